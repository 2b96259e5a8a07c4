"""Parity of the CUDA path against the oracle at BASELINE.json's full sizes, the graph-sharded
execution against the single-call execution, the eigensolver's convergence branches and the
covariance-repair decision on either side of its threshold.  Run on the B200 box: pytest -m gpu"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import uglad_oracle as O  # the checker, never the thing under test

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
THETA_TOL = 1e-4


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def edges_match(theta, ref, margin=1e-5):
    clear = np.abs(ref) > margin
    same = (theta != 0) == (ref != 0)
    return bool(np.all(same | ~clear)) and bool(np.all(np.abs(theta[ref == 0]) <= margin))


@pytest.fixture(autouse=True)
def _cold():
    from uglad_b200 import ops
    assert torch.cuda.is_available()
    ops.reset_warm_start()
    yield
    ops.reset_warm_start()


def _model_from(P):
    from uglad_b200.glad.glad_params import GladParams
    model = GladParams(1.0, 3, 3)
    model.load_state_dict({k: v.detach() for k, v in P.items()})
    return model


def _low_threshold_params(seed, bias=-6.0):
    """rho ~ 0.0025: theta keeps a non-trivial off-diagonal support (as after long training)."""
    P = O.init_params(seed)
    with torch.no_grad():
        P["rho_l1.4.bias"].fill_(bias)
    return P


def test_multitask_full_size_matches_oracle():
    """BASELINE configs[2]: 256 graphs, D=100, M=1000, L=15 -- theta per graph, edge sets, loss and
    all 42 gradients against the oracle on the same inputs (more CTAs than one wave of the grid)."""
    import bench
    from uglad_b200 import main as ug
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    X = bench.synth(256, 100, 1000, 77)
    S = torch.tensor(O.covariance(X), dtype=torch.float32)
    P = _low_threshold_params(77)
    th_o, loss_o = O.forward_loss(S, P, 15, 0)
    loss_o.backward()
    model = _model_from(P)
    for epoch in range(2):   # cold solve, then the warm-started one (same parameters -> same answer)
        model.zero_grad()
        th, loss = ug.forward_uGLAD(S.cuda(), model, L=15, INIT_DIAG=0)
        loss.backward()
        th_n, ref = th.detach().cpu().numpy(), th_o.detach().numpy()
        per_graph = np.array([rel(th_n[b], ref[b]) for b in range(256)])
        assert per_graph.max() < THETA_TOL, (epoch, per_graph.argmax(), per_graph.max())
        assert edges_match(th_n, ref)
        assert (ref != 0).sum() > 256 * 100 * 2          # the support is not just the diagonal
        assert abs(loss.item() - loss_o.item()) < 1e-4 * max(1.0, abs(loss_o.item()))
        for k, p in model.named_parameters():
            assert rel(p.grad.cpu().numpy(), P[k].grad.numpy()) < 1e-3, (epoch, k)


def test_consensus_full_size_matches_oracle():
    """BASELINE configs[3] built exactly like run_uGLAD_missing (main.py:595-622): D=200, M=1000 with
    20 % NaN dropout -> mean imputation -> K=32 row-subsampled covariances + the full-data covariance
    in the loss.  theta, loss, gradients and the consensus against the oracle."""
    import bench
    from uglad_b200 import main as ug
    from uglad_b200.utils import prepare_data
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    rng = np.random.default_rng(78)
    X = bench.synth(1, 200, 1000, 78)[0].astype(np.float64)
    Xm = prepare_data.add_noise_dropout(X[None], dropout=0.2, rng=rng)
    Xi = ug.mean_imputation(Xm)[0]
    folds = O.kfold_blocks(1000, 32)
    S_K = torch.tensor(O.covariance([Xi[tr] for tr, _ in folds]), dtype=torch.float32)
    Sb = torch.tensor(O.covariance(Xi[None]), dtype=torch.float32)
    P = _low_threshold_params(78)
    th_o, loss_o = O.forward_loss(S_K, P, 15, 0, loss_S=Sb)
    loss_o.backward()
    model = _model_from(P)
    # the product's own pipeline on the device builds the same covariances ...
    S_K_dev, Sb_dev = ug.consensus_covariances(prepare_data.convert_to_torch(Xi), 32)
    assert rel(S_K_dev.cpu().numpy(), S_K.numpy()) < 5e-6 and rel(Sb_dev.cpu().numpy(), Sb.numpy()) < 5e-6
    # ... and the hot path on the oracle's inputs gives the oracle's outputs
    th, loss = ug.forward_uGLAD(S_K.cuda(), model, L=15, INIT_DIAG=0, loss_Sb=Sb.cuda())
    loss.backward()
    th_n, ref = th.detach().cpu().numpy(), th_o.detach().numpy()
    per_graph = np.array([rel(th_n[b], ref[b]) for b in range(32)])
    assert per_graph.max() < THETA_TOL, per_graph.max()
    assert edges_match(th_n, ref)
    assert abs(loss.item() - loss_o.item()) < 1e-4 * max(1.0, abs(loss_o.item()))
    for k, p in model.named_parameters():
        assert rel(p.grad.cpu().numpy(), P[k].grad.numpy()) < 1e-3, k
    cons = ug.get_final_precision_from_batch(th.detach(), type="min").cpu().numpy()
    cons_o = O.consensus_min(th_o.detach()).numpy()
    assert rel(cons, cons_o) < THETA_TOL and edges_match(cons, cons_o)


# ---- graph shards == the whole batch ---------------------------------------------------------------
def _run_whole(lib, dims_fn, S, flat, eig, G):
    from uglad_b200 import ops
    B, D = S.shape[0], S.shape[1]
    dims = dims_fn(B, B)
    ws = torch.empty(lib.uglad_workspace_floats(C.byref(dims)), device=S.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
    wS, VtS = (eig.wS, eig.VtS) if eig is not None else (None, None)
    ops.check(lib.uglad_glad_forward(C.byref(dims), P(S), P(flat), P(wS), P(VtS), P(ws), None, st), "forward")
    off = lib.uglad_workspace_offset(C.byref(dims), b"theta")
    theta = ws[off:off + B * D * D].view(B, D, D).clone()
    gp = torch.empty(flat.numel(), device=S.device)
    ops.check(lib.uglad_glad_backward(C.byref(dims), P(S), P(flat), P(wS), P(VtS), P(ws), P(G), P(gp), st), "backward")
    return theta, gp


@pytest.mark.parametrize("D,small_d_max", [(20, 166), (20, 0), (100, 166)], ids=["eigensolver", "large-path", "d100"])
def test_shards_reproduce_the_whole_batch(D, small_d_max):
    """uglad_glad_init_forward / _layer_forward with B_total > B on three shards (3 + 2 + 2 graphs),
    the per-layer Frobenius sums exchanged by hand as the all-reduce would, against uglad_glad_forward
    on all 7 graphs; likewise the summed shard gradients against the whole-batch gradient."""
    from uglad_b200 import _lib, ops
    lib = _lib.load()
    ops.tune("small_d_max", small_d_max)
    try:
        L, H = 6, 3
        rng = np.random.default_rng(D)
        X = rng.random((7, 3 * D, D))
        S = torch.tensor(O.covariance(X), dtype=torch.float32).cuda()
        G = torch.tensor(rng.standard_normal((7, D, D)), dtype=torch.float32).cuda()
        G = (G + G.transpose(1, 2)).contiguous()
        P_ = _low_threshold_params(5)
        flat = torch.cat([P_[k].detach().reshape(-1) for k in O.PARAM_KEYS]).cuda()
        dims_fn = lambda B, Bt: ops.make_dims(B, D, L, H, 0, Bt)
        large = D > lib.uglad_small_d_max()
        eig_all = None if large else ops.ConditionedCovariance(S, repair=False)
        theta_all, gp_all = _run_whole(lib, dims_fn, S, flat, eig_all, G)

        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        Pp = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        shards = [slice(0, 3), slice(3, 5), slice(5, 7)]
        state = []
        for sl in shards:
            Ss = S[sl].contiguous()
            dims = dims_fn(Ss.shape[0], 7)
            ws = torch.empty(lib.uglad_workspace_floats(C.byref(dims)), device=S.device)
            eig = None if large else ops.ConditionedCovariance(Ss, repair=False)
            wS, VtS = (eig.wS, eig.VtS) if eig is not None else (None, None)
            ops.check(lib.uglad_glad_init_forward(C.byref(dims), Pp(Ss), Pp(flat), Pp(wS), Pp(VtS), Pp(ws), st), "init")
            state.append((Ss, dims, ws, eig))
        noff = [lib.uglad_workspace_offset(C.byref(dims), b"normf") for _, dims, _, _ in state]
        for k in range(L):
            for Ss, dims, ws, eig in state:
                ops.check(lib.uglad_glad_layer_forward(C.byref(dims), k, Pp(Ss), Pp(flat), Pp(ws), None, st), "layer")
            total = sum(ws[o + k] for o, (_, _, ws, _) in zip(noff, state))       # what the all-reduce leaves on every rank
            for o, (_, _, ws, _) in zip(noff, state):
                ws[o + k] = total
        gp_sum = torch.zeros_like(gp_all)
        for sl, (Ss, dims, ws, eig) in zip(shards, state):
            off = lib.uglad_workspace_offset(C.byref(dims), b"theta")
            n = Ss.shape[0]
            theta = ws[off:off + n * D * D].view(n, D, D)
            assert rel(theta.cpu().numpy(), theta_all[sl].cpu().numpy()) < 2e-6
            assert edges_match(theta.cpu().numpy(), theta_all[sl].cpu().numpy())
            gp = torch.empty_like(gp_all)
            wS, VtS = (eig.wS, eig.VtS) if eig is not None else (None, None)
            ops.check(lib.uglad_glad_backward(C.byref(dims), Pp(Ss), Pp(flat), Pp(wS), Pp(VtS), Pp(ws),
                                              Pp(G[sl].contiguous()), Pp(gp), st), "backward")
            gp_sum += gp
        assert rel(gp_sum.cpu().numpy(), gp_all.cpu().numpy()) < 2e-5
    finally:
        ops.tune("small_d_max", 166)


# ---- the eigensolver's convergence branches ---------------------------------------------------------
def test_eigensolver_fixup_list_recheck_and_overflow_branches():
    """Cold solves of ill-separated spectra: eigenvalue clusters far tighter than the cosine tolerance
    resolves leave many column pairs above the tolerance after a sweep.  The post-sweep check then
    (a) lists them for the fix-up pass, (b) re-checks after fix-ups whose worst cosine was not tiny and
    (c) overflows its 64-entry list and falls back to a full sweep.  Every branch must have run on
    some matrix of the set (developer knob eig_timing = 2 reports the counts) and every result must
    be a valid eigendecomposition."""
    from uglad_b200 import ops
    rng = np.random.default_rng(11)
    mats = []
    for D in (48, 100, 150):
        for spread in (1e-2, 1e-4, 1e-6):
            Q, _ = np.linalg.qr(rng.standard_normal((D, D)))
            ev = np.concatenate([1.0 + spread * rng.standard_normal(D - 6), np.array([-3.0, -1.0, 0.2, 2.5, 4.0, 9.0])])
            mats.append((Q * ev) @ Q.T)
    for D in (8, 10, 11):   # at most 64 column pairs: a cold sweep leaves a short list of LARGE cosines -> re-check
        A = rng.standard_normal((D, D))
        mats.append((A + A.T) / 2)
    seen = np.zeros(3)
    ops.tune("eig_timing", 2)
    try:
        for A in mats:
            A32 = torch.tensor(A[None], dtype=torch.float32).cuda()
            w, Vt, info = ops.eigh(A32, indefinite=True)
            seen += info[0, 1:4].cpu().numpy() > 0
            w, Vt = w.cpu().numpy().astype(np.float64)[0], Vt.cpu().numpy().astype(np.float64)[0]
            D = A.shape[0]
            assert np.abs(Vt @ Vt.T - np.eye(D)).max() < 2e-5
            assert rel((Vt.T * w) @ Vt, A) < 3e-5
            assert np.abs(np.sort(w) - np.linalg.eigvalsh(A)).max() < 3e-5 * np.abs(A).sum(-1).max()
    finally:
        ops.tune("eig_timing", 0)
    assert seen[0] > 0, "the fix-up pass never ran"
    assert seen[1] > 0, "the list never overflowed into a full sweep"
    assert seen[2] > 0, "no re-check after a fix-up pass"


# ---- covariance repair on either side of the threshold (prepare_data.py:345-355) --------------------
def _samples_with_spectrum(ev, M, rng):
    """Samples whose biased covariance is (up to float64 rounding) R diag(ev) R^T."""
    D = len(ev)
    Z = rng.standard_normal((M, D))
    Z -= Z.mean(0)
    Qz, _ = np.linalg.qr(Z)                      # centred, orthonormal columns: cov = I / M
    R, _ = np.linalg.qr(rng.standard_normal((D, D)))
    return (Qz * np.sqrt(M)) @ (np.sqrt(ev)[:, None] * R.T)


@pytest.mark.parametrize("small_d_max", [166, 0], ids=["eigensolver", "large-path"])
@pytest.mark.parametrize("min_eig,repaired", [(0.0, True), (3e-7, True), (7e-7, True), (1.5e-6, False), (4e-6, False)])
def test_covariance_repair_decision_near_the_threshold(small_d_max, min_eig, repaired):
    """The reference repairs when the smallest float64 eigenvalue is <= 1e-6.  The FP32 solver's own
    eigenvalue error is of that size; the decision is taken in double on the samples' covariance."""
    from uglad_b200 import ops
    from uglad_b200.utils import prepare_data
    ops.tune("small_d_max", small_d_max)
    try:
        rng = np.random.default_rng(int(min_eig * 1e9) + 3)
        D, M = 24, 160
        ev = np.concatenate([[min_eig], np.linspace(0.02, 0.3, D - 1)])
        X = _samples_with_spectrum(ev, M, rng)
        if min_eig == 0.0:
            X[:, -1] = X[:, 0] - 2.0 * X[:, 1]   # an exactly dependent column: one zero eigenvalue
        X32 = X.astype(np.float32).astype(np.float64)   # what the device sees
        want = O.covariance(X32[None], offset=0.1)
        raw = (X32 - X32.mean(0)).T @ (X32 - X32.mean(0)) / M
        assert (np.abs(want[0] - raw).max() > 0.05) == repaired   # the oracle's own decision is the expected one
        S = prepare_data.get_covariance(X32[None], offset=0.1).cpu().numpy()
        assert rel(S, want) < 1e-5, (rel(S, want), np.abs(S[0] - raw).max())
    finally:
        ops.tune("small_d_max", 166)


@pytest.mark.parametrize("D,small_d_max", [(20, 166), (20, 0), (100, 166)], ids=["eigensolver", "large-path", "d100"])
def test_peer_exchange_forward_reproduces_the_whole_batch(D, small_d_max):
    """uglad_glad_forward_sharded: two shards of one batch as two "ranks" on two streams of this GPU,
    their exchange buffers mapped into each other (here: plain device pointers of one process; across
    processes the same pointers come from CUDA IPC).  The lambda kernels exchange the per-layer Frobenius
    sums through those buffers; the shards must reproduce uglad_glad_forward on the whole batch, and two
    consecutive calls (alternating slot parity, new tags) must both be right."""
    from uglad_b200 import _lib, ops
    lib = _lib.load()
    ops.tune("small_d_max", small_d_max)
    try:
        L, H = 6, 3
        rng = np.random.default_rng(D + 1)
        X = rng.random((5, 3 * D, D))
        S = torch.tensor(O.covariance(X), dtype=torch.float32).cuda()
        P_ = _low_threshold_params(6)
        flat = torch.cat([P_[k].detach().reshape(-1) for k in O.PARAM_KEYS]).cuda()
        large = D > lib.uglad_small_d_max()
        G = torch.zeros(5, D, D, device="cuda")
        eig_all = None if large else ops.ConditionedCovariance(S, repair=False)
        theta_all, _ = _run_whole(lib, lambda B, Bt: ops.make_dims(B, D, L, H, 0, Bt), S, flat, eig_all, G)
        Pp = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        shards = [slice(0, 3), slice(3, 5)]
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        slots = [torch.zeros(lib.uglad_peer_slots_bytes(L) // 8, dtype=torch.int64, device="cuda") for _ in shards]
        state = []
        for sl in shards:
            Ss = S[sl].contiguous()
            dims = ops.make_dims(Ss.shape[0], D, L, H, 0, 5)
            ws = torch.empty(lib.uglad_workspace_floats(C.byref(dims)), device="cuda")
            eig = None if large else ops.ConditionedCovariance(Ss, repair=False)
            state.append((Ss, dims, ws, eig))
        torch.cuda.synchronize()
        for tag in (7, 8):
            for r, ((Ss, dims, ws, eig), stream) in enumerate(zip(state, streams)):
                peers = _lib.UgladPeers()
                peers.world, peers.rank, peers.tag = 2, r, tag
                for j, t in enumerate(slots):
                    peers.slots[j] = t.data_ptr()
                wS, VtS = (eig.wS, eig.VtS) if eig is not None else (None, None)
                ops.check(lib.uglad_glad_forward_sharded(C.byref(dims), Pp(Ss), Pp(flat), Pp(wS), Pp(VtS), Pp(ws), None,
                                                         C.byref(peers), C.c_void_p(stream.cuda_stream)), "forward_sharded")
            torch.cuda.synchronize()
            for sl, (Ss, dims, ws, eig) in zip(shards, state):
                off = lib.uglad_workspace_offset(C.byref(dims), b"theta")
                n = Ss.shape[0]
                theta = ws[off:off + n * D * D].view(n, D, D)
                assert rel(theta.cpu().numpy(), theta_all[sl].cpu().numpy()) < 2e-6, tag
                assert edges_match(theta.cpu().numpy(), theta_all[sl].cpu().numpy())
    finally:
        ops.tune("small_d_max", 166)


@pytest.mark.parametrize("B,D", [(2, 256), (4, 204), (1, 333)])
def test_persistent_chain_launch_is_bit_identical_to_separate_launches(B, D):
    """uglad_tune("tc_chain", 1): the ten Newton-Schulz iterations of a layer (forward: 19 stages, backward:
    30 stages incl. the in-kernel antisymmetrisation) as ONE persistent launch with grid barriers.  Same
    tiles, same arithmetic: theta, loss and gradients must equal the separate launches bit for bit --
    also with fewer tiles than CTAs in a stage (idle CTAs must not run ahead of the barrier) and with a D
    whose rows are not 16-byte multiples (direct-store epilogue)."""
    from uglad_b200 import main as ug, ops
    rng = np.random.default_rng(B * 1000 + D)
    X = rng.random((B, 2 * D, D))
    S = torch.tensor(O.covariance(X), dtype=torch.float32).cuda()
    P_ = _low_threshold_params(9)
    out = {}
    for chain in (0, 1):
        ops.tune("tc_chain", chain)
        try:
            model = _model_from(P_)
            th, loss = ug.forward_uGLAD(S, model, L=3, INIT_DIAG=0)
            loss.backward()
            out[chain] = (th.detach().clone(), loss.detach().clone(),
                          torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone())
        finally:
            ops.tune("tc_chain", 0)
    for a, b in zip(out[0], out[1]):
        assert torch.equal(a, b)
