"""Data preparation on either side of the hot path (reference: uglad/utils/prepare_data.py).

get_covariance runs on the GPU (uglad_covariance + uglad_condition_covariance); the table
clean-up and the synthetic Erdos-Renyi generator are host-side numpy."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from .. import ops


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("uglad_b200 needs a CUDA device (there is no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


def convert_to_torch(data, req_grad: bool = False, use_cuda: bool = True) -> torch.Tensor:
    """prepare_data.py:288-307, float32 on the current CUDA device."""
    if not torch.is_tensor(data):
        data = torch.from_numpy(np.asarray(data, dtype=np.float32))
    data = data.to(_device(), dtype=torch.float32)
    data.requires_grad = req_grad
    return data


def get_covariance(Xb, offset: float = 0.1, warm=None) -> torch.Tensor:
    """prepare_data.py:328-356.  Xb: [B,M,D] array/tensor or a list of [M_b,D] arrays with
    different sample counts (multitask mode).  Returns the conditioned covariances as a CUDA
    tensor [B,D,D] with their eigendecomposition attached for glad()'s theta_0.  `warm`: the
    result of an earlier call on similar data (same shapes); it seeds the eigensolver."""
    dev = _device()
    wcc = getattr(warm, "_uglad_eig", (None, None))[1] if warm is not None else None
    if torch.is_tensor(Xb) or (isinstance(Xb, np.ndarray) and Xb.ndim == 3):
        X = convert_to_torch(Xb)
        S, mean = ops.covariance(X, return_mean=True)
        cc = ops.ConditionedCovariance(S, offset=offset, repair=True, warm=wcc, X=X, mean=mean)
    else:  # ragged list (arrays or device tensors): equal shapes are grouped into one batched launch each
        mats = [convert_to_torch(x) for x in Xb]
        shapes = {}
        for i, m in enumerate(mats):
            shapes.setdefault(tuple(m.shape), []).append(i)
        parts = []
        for shp, idx in shapes.items():
            X = torch.stack([mats[i] for i in idx])
            S, mean = ops.covariance(X, return_mean=True)
            parts.append((idx, ops.ConditionedCovariance(S, offset=offset, repair=True, X=X, mean=mean)))
        if len(parts) == 1 and parts[0][0] == list(range(len(mats))):
            cc = parts[0][1]
        else:  # stitch the groups back into input order
            cc = ops.ConditionedCovariance.__new__(ops.ConditionedCovariance)
            order = torch.empty(len(mats), dtype=torch.long, device=dev)
            order[torch.as_tensor([i for idx, _ in parts for i in idx], device=dev)] = torch.arange(len(mats), device=dev)
            cat = lambda name: (None if getattr(parts[0][1], name) is None
                                else torch.cat([getattr(p, name) for _, p in parts])[order].contiguous())
            cc.S, cc.wS, cc.VtS, cc.info = cat("S"), cat("wS"), cat("VtS"), cat("info")
    out = cc.S
    out._uglad_eig = (out._version, cc)
    return out


def bind_host_to_device_numa(device=None) -> Optional[int]:
    """Input-pipeline hygiene for multi-GPU boxes: restrict this process to the CPUs of the NUMA node
    its GPU hangs off, so that pinned staging buffers allocated afterwards (first touch) and the
    threads that fill them are local to the GPU's PCIe root -- with one process per GPU and all of
    them streaming samples at once, cross-socket copies otherwise halve the host->device bandwidth.
    Returns the node, or None when the topology cannot be read (nothing is changed then)."""
    import os
    try:
        dev = _device() if device is None else device
        pr = torch.cuda.get_device_properties(dev)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


class CovariancePrefetcher:
    """Input pipeline for repeated fits on fresh sample batches of one shape [B, M, D]: the
    host->device copy of the NEXT batch and its covariance + conditioning run on a side stream while
    the current batch trains.

        pf = CovariancePrefetcher(); pf.submit(X0_pinned)
        for each step:  S = pf.get(); loss = step(S); pf.submit(X_next_pinned); loss.item()

    Everything is staged in three rotating, preallocated slots (no allocator traffic on the side
    stream, which would otherwise synchronise the device now and then).  The tensor returned by
    get() stays valid until the NEXT get(): at that point the work the consumer has enqueued on it
    is fenced by an event, and the side stream waits for that event before it overwrites the slot
    three submits later -- so the host may run any number of steps ahead of the device.
    Consecutive batches seed each other's eigensolver (a warm start changes the work, never the
    result)."""

    SLOTS = 3

    def __init__(self, device=None, offset: float = 0.1):
        self.device = _device() if device is None else device
        self.stream = torch.cuda.Stream(device=self.device)
        self.offset = float(offset)
        self._slots = None
        self._n = 0
        self._pending = None
        self._handed = None   # the slot the consumer is working on (released at the next get())

    def _make_slots(self, shape):
        import ctypes as C
        from .. import _lib
        lib = _lib.load()
        B, M, D = shape
        large = D > lib.uglad_small_d_max()
        f = dict(device=self.device, dtype=torch.float32)
        slots = []
        for _ in range(self.SLOTS):
            sl = {"X": torch.empty(B, M, D, **f), "S": torch.empty(B, D, D, **f), "mean": torch.empty(B, D, **f),
                  "Xt": torch.empty(max(lib.uglad_covariance_scratch_floats(B, M, D), 1), **f),
                  "scratch": torch.empty(max(lib.uglad_condition_scratch_floats(B, D), 1), **f),
                  "ev": torch.cuda.Event(), "released": None}
            cc = ops.ConditionedCovariance.__new__(ops.ConditionedCovariance)
            cc.S = sl["S"]
            if large:
                cc.wS = cc.VtS = cc.info = None
            else:
                cc.wS, cc.VtS, cc.info = torch.empty(B, D, **f), torch.empty(B, D, D, **f), torch.empty(B, 4, **f)
            sl["cc"] = cc
            slots.append(sl)
        return slots

    def submit(self, X_host: torch.Tensor) -> None:
        import ctypes as C
        from .. import _lib
        lib = _lib.load()
        if X_host.dim() == 2:
            X_host = X_host.unsqueeze(0)
        if self._slots is None or tuple(self._slots[0]["X"].shape) != tuple(X_host.shape):
            self._slots, self._n = self._make_slots(tuple(X_host.shape)), 0
            self._handed = None
            # the slots were just carved out of the caching allocator on the consumer's stream: a recycled
            # block may still be read by kernels queued there
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
        sl = self._slots[self._n % self.SLOTS]
        if sl["released"] is not None:   # the consumer's work on this slot's previous content
            self.stream.wait_event(sl["released"])
            sl["released"] = None
        prev = self._slots[(self._n - 1) % self.SLOTS]["cc"] if self._n > 0 else None
        self._n += 1
        B, M, D = sl["X"].shape
        P = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        with torch.cuda.stream(self.stream):
            st = C.c_void_p(self.stream.cuda_stream)
            sl["X"].copy_(X_host, non_blocking=True)
            ops.check(lib.uglad_covariance_ws(P(sl["X"]), B, M, D, P(sl["S"]), P(sl["mean"]), P(sl["Xt"]), st),
                      "uglad_covariance_ws")
            cc = sl["cc"]
            wV, ww = (prev.VtS, prev.wS) if (prev is not None and prev.VtS is not None) else (None, None)
            ops.check(lib.uglad_condition_covariance_x(P(sl["S"]), P(sl["X"]), P(sl["mean"]), B, M, D, self.offset,
                                                       P(cc.wS), P(cc.VtS), P(cc.info), P(sl["scratch"]), P(wV), P(ww), st),
                      "uglad_condition_covariance")
            sl["ev"].record(self.stream)
        self._pending = sl

    def get(self) -> torch.Tensor:
        sl, self._pending = self._pending, None
        cur = torch.cuda.current_stream(self.device)
        if self._handed is not None:   # everything enqueued so far on the slot handed out before
            ev = torch.cuda.Event()
            ev.record(cur)
            self._handed["released"] = ev
        self._handed = sl
        cur.wait_event(sl["ev"])
        S = sl["S"]
        S._uglad_eig = (S._version, sl["cc"])
        S._uglad_warm_key = ("prefetch", id(self))   # consecutive batches of one stream seed each other (ops._warm)
        return S


# ---- host-side table hygiene (prepare_data.py:361-516 with its default arguments) ---------
def normalize_table(df, typeN: str):
    if typeN == "min_max":
        return (df - df.min()) / (df.max() - df.min())
    if typeN == "mean":
        return (df - df.mean()) / df.std()
    return df


def _cov_spectrum(values: np.ndarray):
    """prepare_data.py:310-325, :568-594: biased covariance of a sample table, its eigenvalues (real
    parts) and the condition number max|eig| / min|eig|."""
    X = np.asarray(values, dtype=np.float64)
    Xc = X - X.mean(axis=0)
    S = Xc.T @ Xc / X.shape[0]
    eig = np.linalg.eigvals(S).real
    with np.errstate(divide="ignore"):
        con = np.abs(eig).max() / np.abs(eig).min()
    return S, eig, con


def get_highly_correlated_features(input_cov: np.ndarray) -> np.ndarray:
    """prepare_data.py:519-549: rank the features by how many strong second-order correlations they
    have.  The covariance matrix's rows are treated as samples; entries of that second covariance
    (diagonal zeroed) at or above the value ranked 10 % from the top of the upper triangle count as
    strong; features are returned by decreasing number of strong partners (ties keep index order)."""
    C = np.asarray(input_cov, dtype=np.float64)
    Cc = C - C.mean(axis=0)
    cov2 = Cc.T @ Cc / C.shape[0]
    np.fill_diagonal(cov2, 0.0)
    mag = np.abs(cov2)
    upper = np.sort(mag[np.triu_indices(mag.shape[0], 1)])[::-1]
    th = upper[int(0.1 * upper.size)]
    rows = np.nonzero(mag >= th)[0]
    feats, counts = np.unique(rows, return_counts=True)
    return feats[np.argsort(-counts, kind="stable")]


def process_table(table, NORM: str = "no", MIN_VARIANCE: float = 0.0, msg: str = "",
                  COND_NUM: float = np.inf, eigval_th: float = 1e-3, VERBOSE: bool = True):
    """prepare_data.py:361-516: drop all-zero rows, mean-impute NaNs, drop constant and duplicate
    columns, normalise, drop columns whose variance is below MIN_VARIANCE; then, while the
    covariance's condition number exceeds COND_NUM, drop the most inter-correlated features -- as
    many as there are eigenvalues below eigval_th (at least one) per pass."""
    import pandas as pd
    table = pd.DataFrame(table).astype(float)
    table = table.loc[~(table == 0).all(axis=1)]
    table = table.fillna(table.mean())
    constant = [c for c in table.columns if table[c].nunique(dropna=False) == 1]
    table = table.drop(columns=constant)
    table = normalize_table(table, NORM)
    table = table.T.drop_duplicates().T
    var = table.var()
    table = table.drop(columns=list(var[var < MIN_VARIANCE].index))
    if COND_NUM != np.inf:
        S, eig, con = _cov_spectrum(table.values)
        itr = 1
        while con > COND_NUM:
            n_small = int(np.sum(eig < eigval_th)) or 1   # still ill-conditioned with no small eigenvalue: drop one
            ranked = get_highly_correlated_features(S)
            drop = table.columns[ranked[: min(n_small, len(ranked))]]
            if VERBOSE:
                print(f"{msg} {itr}: condition number {con}: dropping {len(drop)} highly correlated features {list(drop)}")
            table = table.drop(columns=drop)
            S, eig, con = _cov_spectrum(table.values)
            itr += 1
    if VERBOSE:
        print(f"{msg}: processed table has {table.shape[0]} samples and {table.shape[1]} features")
    return table


# ---- synthetic Erdos-Renyi Gaussian graphical models (prepare_data.py:13-140) --------------
def get_data(num_nodes: int, sparsity: Sequence[float], num_samples: int, batch_size: int = 1,
             w_min: float = 0.5, w_max: float = 1.0, eig_offset: float = 0.1,
             rng: Optional[np.random.Generator] = None):
    """Same construction as the reference (random G(n,p) support, U[w_min,w_max] weights,
    symmetrise, shift the spectrum so the smallest eigenvalue is eig_offset, sample from
    N(0, theta^-1)), driven by a numpy Generator instead of networkx + the global seed."""
    rng = np.random.default_rng() if rng is None else rng
    Xb, thetas = [], []
    for _ in range(batch_size):
        p = rng.uniform(sparsity[0], sparsity[1])
        upper = np.triu(rng.random((num_nodes, num_nodes)) < p, 1)
        adj = (upper | upper.T).astype(np.float64)
        U = rng.random((num_nodes, num_nodes)) * (w_max - w_min) + w_min
        theta = adj * U
        theta = (theta + theta.T) / 2 + np.eye(num_nodes)
        theta = theta + np.eye(num_nodes) * (eig_offset - np.linalg.eigvalsh(theta).min())
        cov = np.linalg.inv(theta)
        Lc = np.linalg.cholesky((cov + cov.T) / 2)
        Xb.append(rng.standard_normal((num_samples, num_nodes)) @ Lc.T)
        thetas.append(theta)
    return np.array(Xb), np.array(thetas)


def add_noise_dropout(Xb: np.ndarray, dropout: float = 0.25, rng: Optional[np.random.Generator] = None):
    """prepare_data.py:143-169: replace a fraction of the entries by NaN."""
    rng = np.random.default_rng() if rng is None else rng
    out = np.array(Xb, dtype=np.float64, copy=True)
    for X in out:
        flat = X.reshape(-1)
        flat[rng.choice(flat.size, size=int(flat.size * dropout), replace=False)] = np.nan
    return out
