"""torch-facing wrappers of the C-ABI: tensors in, tensors out, autograd where the reference
has it.  PyTorch supplies device memory, streams and torch.distributed; every floating-point
operation of the hot path runs in libuglad_b200.so."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import UgladDims, check


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.UgladError(f"{name} must be a CUDA tensor (uglad_b200 has no CPU path)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def covariance(X: torch.Tensor, return_mean: bool = False):
    """S[b] = (X_b - mean)^T (X_b - mean) / M for X [B,M,D] (prepare_data.py:342-344)."""
    X = _f32c(X, "X")
    if X.dim() == 2:
        X = X.unsqueeze(0)
    B, M, D = X.shape
    S = torch.empty(B, D, D, device=X.device, dtype=torch.float32)
    mean = torch.empty(B, D, device=X.device, dtype=torch.float32)
    lib = _lib.load()
    # the contraction runs on the tensor pipe (tcgen05 3xTF32) from centred, feature-major samples in scratch
    scratch = torch.empty(max(lib.uglad_covariance_scratch_floats(B, M, D), 1), device=X.device, dtype=torch.float32)
    check(lib.uglad_covariance_ws(_ptr(X), B, M, D, _ptr(S), _ptr(mean), _ptr(scratch), _stream(X)), "uglad_covariance_ws")
    return (S, mean) if return_mean else S


def eigh(A: torch.Tensor, indefinite: bool = True) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Batched symmetric eigendecomposition.  Returns (w [B,D], Vt [B,D,D] rows = vectors,
    info [B,4] = sweeps, shift, trace, sum(w))."""
    A = _f32c(A, "A")
    B, D, _ = A.shape
    lib = _lib.load()
    w = torch.empty(B, D, device=A.device, dtype=torch.float32)
    Vt = torch.empty(B, D, D, device=A.device, dtype=torch.float32)
    info = torch.empty(B, 4, device=A.device, dtype=torch.float32)
    ns = lib.uglad_eigh_scratch_floats(B, D)
    scratch = torch.empty(max(ns, 1), device=A.device, dtype=torch.float32)
    check(lib.uglad_eigh(_ptr(A), B, D, 1 if indefinite else 0, _ptr(w), _ptr(Vt), _ptr(info),
                         _ptr(scratch), _stream(A)), "uglad_eigh")
    return w, Vt, info


class ConditionedCovariance:
    """S after the eigenvalue repair of prepare_data.py:345-355 plus its eigendecomposition,
    which every forward reuses for theta_0 = (S + t I)^-1 (glad.py:115-117)."""

    def __init__(self, S: torch.Tensor, offset: float = 0.1, repair: bool = True, warm=None, X=None, mean=None):
        """`warm`: an earlier ConditionedCovariance of the same shape (e.g. the previous batch of a
        stream of similar sample matrices); its eigenvectors seed the solver.  `X` [B,M,D] / `mean`
        [B,D]: the samples S was computed from; with them the repair decision is taken on the
        float64 covariance of the samples, as the reference does (uglad_condition_covariance_x)."""
        S = _f32c(S, "S").clone()
        B, D, _ = S.shape
        lib = _lib.load()
        self.S = S
        M = 0
        if X is not None and mean is not None:
            X, mean = _f32c(X, "X"), _f32c(mean, "mean")
            M = X.shape[1]
        else:
            X = mean = None
        if D > lib.uglad_small_d_max():
            # large-D path: no eigendecomposition is kept (theta_0 comes from a Cholesky inverse)
            self.wS = self.VtS = self.info = None
            if repair:
                scratch = torch.empty(lib.uglad_condition_scratch_floats(B, D), device=S.device, dtype=torch.float32)
                check(lib.uglad_condition_covariance_x(_ptr(S), _ptr(X), _ptr(mean), B, M, D, float(offset), None, None,
                                                       None, _ptr(scratch), None, None, _stream(S)),
                      "uglad_condition_covariance")
            return
        self.wS = torch.empty(B, D, device=S.device, dtype=torch.float32)
        self.VtS = torch.empty(B, D, D, device=S.device, dtype=torch.float32)
        self.info = torch.empty(B, 4, device=S.device, dtype=torch.float32)
        ns = lib.uglad_eigh_scratch_floats(B, D)
        scratch = torch.empty(max(ns, 1), device=S.device, dtype=torch.float32)
        wV = ww = None
        if warm is not None and warm.VtS is not None and warm.VtS.shape == self.VtS.shape:
            wV, ww = warm.VtS, warm.wS
        if repair:
            check(lib.uglad_condition_covariance_x(_ptr(S), _ptr(X), _ptr(mean), B, M, D, float(offset), _ptr(self.wS),
                                                   _ptr(self.VtS), _ptr(self.info), _ptr(scratch), _ptr(wV), _ptr(ww),
                                                   _stream(S)), "uglad_condition_covariance")
        else:
            check(lib.uglad_eigh_warm(_ptr(S), B, D, 1, _ptr(self.wS), _ptr(self.VtS), _ptr(self.info),
                                      _ptr(scratch), _ptr(wV), _ptr(ww), _stream(S)), "uglad_eigh")


def _eig_of(S: torch.Tensor) -> ConditionedCovariance:
    """Eigendecomposition of a covariance batch, cached ON the tensor object (keyed by its
    version counter) so that the epochs of one fit, which pass the same Sb every time
    (main.py:389-399), pay for it once and a recycled allocation can never alias it."""
    hit = getattr(S, "_uglad_eig", None)
    if hit is not None and hit[0] == S._version:
        return hit[1]
    cc = ConditionedCovariance(S, repair=False)
    try:
        S._uglad_eig = (S._version, cc)
    except Exception:  # exotic tensor subclasses without attribute storage
        pass
    return cc


def make_dims(B, D, L, H, init_diag, B_total=None, exact_sqrt=False, lambda_init=1.0) -> UgladDims:
    return UgladDims(int(B), int(D), int(L), int(H), int(init_diag), int(B if B_total is None else B_total),
                     int(bool(exact_sqrt)), float(lambda_init))


# Warm start of the per-layer eigensolver: the workspace of an earlier forward seeds the next one.
# Keyed by problem shape + workspace layout AND by the covariance tensor it was computed for, so
# that alternating inputs of one shape (the train / test covariances of the CV mode, main.py:486-494)
# each continue from their own previous epoch.  A forward on a tensor never seen before falls back
# to the most recent workspace of the same shape (fresh data of a stream: a valid, if weaker, seed --
# a warm start changes the work done, never the converged result).
import collections

_warm: "collections.OrderedDict" = collections.OrderedDict()
_WARM_MAX = 3
warm_start_enabled = True


def reset_warm_start():
    """Forget the eigenvector seeds (the next forward solves from scratch)."""
    _warm.clear()


def _warm_lookup(shape_key, tensor_key):
    hit = _warm.get((shape_key, tensor_key))
    if hit is not None:
        return hit
    for (sk, _), ws in reversed(_warm.items()):
        if sk == shape_key:
            return ws
    return None


def _warm_store(shape_key, tensor_key, ws):
    _warm.pop((shape_key, tensor_key), None)
    _warm[(shape_key, tensor_key)] = ws
    while len(_warm) > _WARM_MAX:
        _warm.popitem(last=False)


def tune(key: str, value: int):
    check(_lib.load().uglad_tune(key.encode(), int(value)), "uglad_tune")
    _warm.clear()   # a knob may change the workspace layout ("small_d_max") or the solver's state


# ---- graph-sharded execution (one process per GPU) ------------------------------------------------
# The graphs of a multitask / consensus batch are independent except for two scalars per epoch
# path: the batch MEAN of ||Z - X||_F^2 that feeds lambda_f after every layer (glad.py:147) and the
# 1/B of the loss (main.py:315).  These helpers are the whole distributed protocol; they are
# backend-agnostic (NCCL on the GPUs, gloo in the CPU tests).
def global_graph_count(B: int, device, group) -> int:
    """Number of graphs over all ranks of `group` (B when not distributed)."""
    if group is None:
        return B
    import torch.distributed as dist
    if dist.get_world_size(group) == 1:
        return B
    cnt = torch.tensor([B], device=device, dtype=torch.int64)
    dist.all_reduce(cnt, group=group)
    return int(cnt.item())


def run_sharded_layers(L: int, layer_fn, normf: torch.Tensor, group) -> None:
    """Drive the L unrolled layers of this rank's shard: layer_fn(k) leaves the LOCAL sum of
    ||Z - X||_F^2 in normf[k]; it is all-reduced before layer k+1 turns it into lambda_{k+1}."""
    import torch.distributed as dist
    for k in range(L):
        layer_fn(k)
        if k + 1 < L:
            dist.all_reduce(normf[k:k + 1], group=group)


def allreduce_shared_gradients(gp: torch.Tensor, group) -> torch.Tensor:
    """The one collective of the data path: the packed rho_l1 / lambda_f / theta_init_offset
    gradient (42 floats for H=3) summed over the shards."""
    import torch.distributed as dist
    dist.all_reduce(gp, group=group)
    return gp


def allreduce_sum(t: torch.Tensor, group) -> torch.Tensor:
    import torch.distributed as dist
    dist.all_reduce(t, group=group)
    return t


def allreduce_min(t: torch.Tensor, group) -> torch.Tensor:
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return t


class PeerExchange:
    """Exchange buffers of the graph-sharded forward (uglad_glad_forward_sharded): one small device
    buffer per rank, mapped into every rank of the box through CUDA IPC, so that the lambda kernels
    exchange the per-layer Frobenius sums themselves (stores over NVLink) and the L layers of a shard
    run as one uninterrupted stream of kernels -- no collective, no host code between the layers.
    Built once per (group, L) with one all-gather of the 64-byte IPC handles."""

    _cache: dict = {}
    enabled = __import__("os").environ.get("UGLAD_PEER_EXCHANGE", "1") != "0"   # 0: per-layer NCCL all-reduce instead

    def __init__(self, group, L: int, device):
        import torch.distributed as dist
        lib = _lib.load()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise _lib.UgladError("peer exchange serves up to 8 ranks of one box")
        nbytes = lib.uglad_peer_slots_bytes(int(L))
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        check(lib.uglad_peer_alloc(nbytes, C.byref(ptr), handle), "uglad_peer_alloc")
        mine = torch.tensor(list(handle.raw), dtype=torch.uint8, device=device)
        every = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(every, mine, group=group)
        self.slots = []
        for r in range(self.world):
            if r == self.rank:
                self.slots.append(ptr.value)
            else:
                p = C.c_void_p()
                check(lib.uglad_peer_open(bytes(every[r].cpu().tolist()), C.byref(p)), "uglad_peer_open")
                self.slots.append(p.value)
        self.tag = 0
        self.tag_dev = torch.zeros(1, dtype=torch.int32, device=device)   # device-side call counter (graph replay)
        dist.barrier(group=group)   # every rank has mapped every buffer before the first store

    @classmethod
    def get(cls, group, L: int, device):
        key = (id(group), int(L), device.index)
        if key not in cls._cache:
            cls._cache[key] = cls(group, L, device)
        return cls._cache[key]

    def next_call(self) -> "_lib.UgladPeers":
        self.tag = (self.tag + 1) & 0xFFFFFFFF or 1
        p = _lib.UgladPeers()
        p.world, p.rank, p.tag = self.world, self.rank, self.tag
        for r, v in enumerate(self.slots):
            p.slots[r] = v
        p.tag_dev = self.tag_dev.data_ptr()
        return p


def _use_peer_exchange(S: torch.Tensor, group) -> bool:
    if not PeerExchange.enabled or group is None or not S.is_cuda:
        return False
    import torch.distributed as dist
    return dist.get_backend(group) == "nccl" and dist.get_world_size(group) <= 8


# explicit (workspace, warm-start workspace) for the next forward: GraphedStep captures the epoch with two static
# workspaces that seed each other instead of the allocate-and-cache scheme of eager calls
_forced_ws = None


class use_workspace:
    def __init__(self, ws: torch.Tensor, warm: Optional[torch.Tensor]):
        self.pair = (ws, warm)

    def __enter__(self):
        global _forced_ws
        _forced_ws = self.pair

    def __exit__(self, *a):
        global _forced_ws
        _forced_ws = None


class GladFunction(torch.autograd.Function):
    """theta_pred = glad(S; params)  (glad.py:74-150) with the hand-written backward."""

    @staticmethod
    def forward(ctx, S, flat_params, L, init_diag, H, lambda_init, exact_sqrt, group, total_graphs=None):
        lib = _lib.load()
        S = _f32c(S, "Sb")
        flat_params = _f32c(flat_params, "params")
        B, D, _ = S.shape
        world = 1
        if group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(group)
        # `total_graphs` (when the caller knows it) avoids an all-reduce + host sync per forward
        B_total = int(total_graphs) if total_graphs is not None else global_graph_count(B, S.device, group)
        dims = make_dims(B, D, L, H, init_diag, B_total, exact_sqrt, lambda_init)
        if flat_params.numel() != lib.uglad_param_count(H):
            raise _lib.UgladError("packed parameter vector has the wrong length")
        n = lib.uglad_workspace_floats(C.byref(dims))
        if n == 0:
            raise _lib.UgladError(lib.uglad_last_error().decode())
        forced = _forced_ws
        if forced is not None:
            ws, warm = forced
            if ws.numel() < n or (warm is not None and warm.numel() < n):
                raise _lib.UgladError("use_workspace: workspace smaller than uglad_workspace_floats")
        else:
            ws = torch.empty(n, device=S.device, dtype=torch.float32)
        # warm start: the previous forward's workspace for the same problem shape (normally the
        # previous epoch of the same fit) seeds the eigensolver; see uglad_glad_forward.
        wkey = (B, D, L, H, init_diag, S.device.index, lib.uglad_eig_path(B, D), n)
        tkey = getattr(S, "_uglad_warm_key", None) or (S.data_ptr(), S._version)
        if forced is None:
            warm = _warm_lookup(wkey, tkey) if warm_start_enabled else None
        eig = _eig_of(S) if (init_diag == 0 and D <= lib.uglad_small_d_max()) else None
        wS, VtS = (eig.wS, eig.VtS) if eig is not None else (None, None)
        st = _stream(S)
        if world == 1:
            check(lib.uglad_glad_forward(C.byref(dims), _ptr(S), _ptr(flat_params), _ptr(wS), _ptr(VtS),
                                         _ptr(ws), _ptr(warm), st), "uglad_glad_forward")
        elif _use_peer_exchange(S, group):
            # one call: the ranks exchange the per-layer Frobenius sums through peer-mapped memory
            peers = PeerExchange.get(group, L, S.device).next_call()
            check(lib.uglad_glad_forward_sharded(C.byref(dims), _ptr(S), _ptr(flat_params), _ptr(wS), _ptr(VtS),
                                                 _ptr(ws), _ptr(warm), C.byref(peers), st), "uglad_glad_forward_sharded")
        else:
            check(lib.uglad_glad_init_forward(C.byref(dims), _ptr(S), _ptr(flat_params), _ptr(wS), _ptr(VtS),
                                              _ptr(ws), st), "uglad_glad_init_forward")
            off = lib.uglad_workspace_offset(C.byref(dims), b"normf")

            def layer(k):
                check(lib.uglad_glad_layer_forward(C.byref(dims), k, _ptr(S), _ptr(flat_params), _ptr(ws),
                                                   _ptr(warm), st), "uglad_glad_layer_forward")

            run_sharded_layers(L, layer, ws[off:off + L], group)
        if warm_start_enabled and forced is None:
            _warm_store(wkey, tkey, ws)
        off = lib.uglad_workspace_offset(C.byref(dims), b"theta")
        theta = ws[off:off + B * D * D].view(B, D, D)
        ctx.dims, ctx.ws, ctx.S, ctx.params, ctx.eig, ctx.group, ctx.world = dims, ws, S, flat_params, eig, group, world
        return theta

    @staticmethod
    def backward(ctx, grad_theta):
        lib = _lib.load()
        dims = ctx.dims
        g = _f32c(grad_theta, "grad_theta")
        gp = torch.empty(ctx.params.numel(), device=g.device, dtype=torch.float32)
        wS, VtS = (ctx.eig.wS, ctx.eig.VtS) if ctx.eig is not None else (None, None)
        check(lib.uglad_glad_backward(C.byref(dims), _ptr(ctx.S), _ptr(ctx.params), _ptr(wS), _ptr(VtS),
                                      _ptr(ctx.ws), _ptr(g), _ptr(gp), _stream(g)), "uglad_glad_backward")
        if ctx.world > 1:
            allreduce_shared_gradients(gp, ctx.group)
        return None, gp, None, None, None, None, None, None, None


class GlassoLossFunction(torch.autograd.Function):
    """loss = sum_b(-logdet theta_b + <S_b, theta_b>) / Bdiv  (main.py:306-315)."""

    @staticmethod
    def forward(ctx, theta, S, Bdiv, struct_theta=None):
        lib = _lib.load()
        theta = _f32c(theta, "theta")
        S = _f32c(S, "S")
        if struct_theta is not None:   # log-cosh structure prior (main.py:325-334), one mask per graph
            struct_theta = _f32c(struct_theta, "struct_theta").expand(theta.shape).contiguous()
        B, D, _ = theta.shape
        sb = S.shape[0]
        loss = torch.empty(1, device=theta.device, dtype=torch.float32)
        need_grad = ctx.needs_input_grad[0]
        grad = torch.empty_like(theta) if need_grad else None
        scratch = torch.empty(lib.uglad_loss_scratch_floats(B, D), device=theta.device, dtype=torch.float32)
        check(lib.uglad_glasso_loss_prior(_ptr(theta), _ptr(S), _ptr(struct_theta), B, D, sb, float(Bdiv), _ptr(loss),
                                          _ptr(grad), _ptr(scratch), _stream(theta)), "uglad_glasso_loss")
        ctx.grad = grad
        return loss[0]

    @staticmethod
    def backward(ctx, gout):
        return ctx.grad * gout, None, None, None


def workspace_floats(B, D, L, H=3, init_diag=0, B_total=None) -> int:
    dims = make_dims(B, D, L, H, init_diag, B_total)
    n = _lib.load().uglad_workspace_floats(C.byref(dims))
    if n == 0:
        raise _lib.UgladError(_lib.load().uglad_last_error().decode())
    return int(n)


class GraphedStep:
    """One training epoch of the reference loop (main.py:389-414: zero_grad, unrolled GLAD forward, glasso
    loss, backward, Adam step) captured as CUDA graphs and replayed: the ~260 kernel launches of an epoch
    cost one graph launch, which is what bounds small batches (one graph at D = 100: ~5 us of launch and
    dependency latency around every small kernel).  Two graphs alternate because every epoch's eigensolver
    is seeded by the previous epoch's workspace: graph 0 works in workspace X seeded by Y, graph 1 in Y
    seeded by X.  The inputs live in static buffers (update_inputs copies new covariances in); the
    optimizer must be capturable (glad.get_optimizers(..., capturable=True)).  With `group` the epoch is
    the graph-sharded one: the layers exchange their Frobenius sums through peer memory (device-side call
    counter) and the gradient all-reduce is captured with the rest."""

    def __init__(self, S, model, optimizer, L=15, INIT_DIAG=0, loss_S=None, struct_theta=None, group=None,
                 total_graphs=None, lazy=False, nan_guard=False):
        """lazy: run nothing here -- the first four step() calls are the eager epochs, the fifth captures (a fit loop
        simply calls step() EPOCHS times; if the capture fails the epochs stay eager).  nan_guard: the NaN stop of
        the reference's direct mode (main.py:405-409: break BEFORE the update): every epoch starts by copying the
        parameters aside, so that step_guarded() can take a replayed update back."""
        from . import main as ug
        if not all(g.get("capturable", False) for g in optimizer.param_groups):
            raise _lib.UgladError("GraphedStep needs a capturable optimizer: glad.get_optimizers(model, lr, capturable=True)")
        dev = S.device
        self.S = S.detach().clone()
        cc = _eig_of(S)
        self.cc = ConditionedCovariance.__new__(ConditionedCovariance)
        self.cc.S = self.S
        self.cc.wS = None if cc.wS is None else cc.wS.clone()
        self.cc.VtS = None if cc.VtS is None else cc.VtS.clone()
        self.cc.info = None
        self.S._uglad_eig = (self.S._version, self.cc)
        self.loss_S = None if loss_S is None else loss_S.detach().clone()
        B, D = S.shape[0], S.shape[1]
        n = workspace_floats(B, D, L, model.H, INIT_DIAG, total_graphs)
        self.ws = [torch.empty(n, device=dev, dtype=torch.float32) for _ in range(2)]
        self.model, self.opt = model, optimizer
        self.calls = 0
        self.dev = dev
        params = [p for p in model.parameters()]
        self.params = params
        self.backup = [torch.empty_like(p) for p in params] if nan_guard else None
        # The epoch closure must not reference `self`: it is stored on the instance, and a reference cycle would keep the
        # captured graphs (with their captured NCCL all-reduce) alive until the cyclic collector runs -- possibly after
        # dist.destroy_process_group(), which then waits for them for ever.
        ws_pair, backup, S_static, loss_static = self.ws, self.backup, self.S, self.loss_S

        def epoch(i, warm, guard=False):
            with use_workspace(ws_pair[i], warm):
                if backup is not None:
                    with torch.no_grad():
                        torch._foreach_copy_(backup, params)
                optimizer.zero_grad(set_to_none=True)
                theta, loss = ug.forward_uGLAD(S_static, model, L=L, INIT_DIAG=INIT_DIAG, loss_Sb=loss_static,
                                               struct_theta=struct_theta, group=group, total_graphs=total_graphs)
                if guard and bool(torch.isnan(loss.detach())):
                    return theta.detach(), loss.detach(), True   # no update (main.py:405-409)
                loss.backward()
                optimizer.step()
            return theta.detach(), loss.detach(), False

        self._epoch = epoch
        self.eager_epochs = 4
        self.graphs, self.out, self.kernels_per_graph = None, [], []
        self.capture_error = None
        self._capture_tried = False
        self.eager_out = []
        if lazy:
            return
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):   # eager epochs: fill both workspaces (cold, then warm) and the optimizer state
            self.eager_out.append(epoch(0, None))
            self.eager_out.append(epoch(1, self.ws[0]))
            self.eager_out.append(epoch(0, self.ws[1]))
            self.eager_out.append(epoch(1, self.ws[0]))
        self.eager_out = [(None, o[1].clone()) for o in self.eager_out]
        cur.wait_stream(side)
        self._capture(strict=True)

    def _capture(self, strict=False):
        """Two alternating captures of the epoch; on failure (strict=False) the epochs simply stay eager."""
        self._capture_tried = True
        torch.cuda.synchronize(self.dev)
        lib = _lib.load()
        graphs, out, kpg = [], [], []
        pool = None
        try:
            for i in range(2):
                g = torch.cuda.CUDAGraph()
                c0 = lib.uglad_launch_count()
                with torch.cuda.graph(g, pool=pool):
                    o = self._epoch(i, self.ws[1 - i])
                out.append((o[0], o[1]))
                kpg.append(int(lib.uglad_launch_count() - c0))   # library kernels replayed per epoch
                pool = g.pool()
                graphs.append(g)
        except Exception as exc:
            if strict:
                raise
            self.capture_error = f"{type(exc).__name__}: {str(exc)[:160]}"
            torch.cuda.synchronize(self.dev)
            return
        self.graphs, self.out, self.kernels_per_graph = graphs, out, kpg
        if self.calls >= self.eager_epochs:
            self.calls -= self.eager_epochs   # replays are indexed from graph 0 (= the epoch after four eager ones)

    def close(self):
        """Drop the captured graphs, their outputs and workspaces NOW (do this before dist.destroy_process_group():
        a graph that captured the gradient all-reduce keeps the NCCL communicator busy until it is destroyed)."""
        self.out = []
        self.graphs = None
        self._capture_tried = True
        self._epoch = None
        self.ws = None
        self.backup = None

    def step_guarded(self, guard=False):
        """One epoch: (theta_pred, loss, stopped).  Eager for the first four calls of a lazy instance, replayed after the
        capture.  guard: stop on a NaN loss without applying the update (needs nan_guard=True for replayed epochs)."""
        if self.graphs is None and not self._capture_tried and self.calls >= self.eager_epochs:
            self._capture()
        i = self.calls & 1
        if self.graphs is not None:
            self.calls += 1
            self.graphs[i].replay()
            theta, loss = self.out[i]
            if guard and bool(torch.isnan(loss)):
                if self.backup is None:
                    raise _lib.UgladError("step_guarded(guard=True) on replayed epochs needs GraphedStep(nan_guard=True)")
                with torch.no_grad():
                    torch._foreach_copy_(self.params, self.backup)
                return theta, loss, True
            return theta, loss, False
        warm = None if self.calls == 0 else self.ws[1 - i]
        theta, loss, stopped = self._epoch(i, warm, guard)
        self.calls += 1
        return theta, loss, stopped

    def update_inputs(self, S, loss_S=None):
        """New covariances of the same shape (and their eigendecomposition, get_covariance attaches it)."""
        self.S.copy_(S)
        cc = _eig_of(S)
        if self.cc.wS is not None:
            self.cc.wS.copy_(cc.wS)
            self.cc.VtS.copy_(cc.VtS)
        if loss_S is not None:
            self.loss_S.copy_(loss_S)

    def step(self):
        """One epoch; returns (theta_pred, loss) -- static tensors when replayed (overwritten two steps later)."""
        theta, loss, _ = self.step_guarded(False)
        return theta, loss


def z_update(X, S, theta_prev, flat_params, H=3):
    """Entrywise soft threshold with the rho_l1 MLP (glad_params.py:56-77); returns
    (Z, sum ||Z-X||_F^2 over the batch)."""
    lib = _lib.load()
    X, S, theta_prev, flat_params = (_f32c(t, "arg") for t in (X, S, theta_prev, flat_params))
    B, D, _ = X.shape
    Z = torch.empty_like(X)
    normf = torch.empty(1, device=X.device, dtype=torch.float32)
    scratch = torch.empty(512 * B + 16, device=X.device, dtype=torch.float32)
    check(lib.uglad_z_update(_ptr(X), _ptr(S), _ptr(theta_prev), _ptr(flat_params), H, B, D, _ptr(Z),
                             _ptr(normf), _ptr(scratch), _stream(X)), "uglad_z_update")
    return Z, normf
