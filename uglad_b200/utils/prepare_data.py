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


def get_covariance(Xb, offset: float = 0.1) -> torch.Tensor:
    """prepare_data.py:328-356.  Xb: [B,M,D] array/tensor or a list of [M_b,D] arrays with
    different sample counts (multitask mode).  Returns the conditioned covariances as a CUDA
    tensor [B,D,D] with their eigendecomposition attached for glad()'s theta_0."""
    dev = _device()
    if torch.is_tensor(Xb) or (isinstance(Xb, np.ndarray) and Xb.ndim == 3):
        X = convert_to_torch(Xb)
        S = ops.covariance(X)
    else:  # ragged list: group equal shapes so that each group is one batched launch
        mats = [np.asarray(x, dtype=np.float32) for x in Xb]
        S = torch.empty(len(mats), mats[0].shape[1], mats[0].shape[1], device=dev)
        shapes = {}
        for i, m in enumerate(mats):
            shapes.setdefault(m.shape, []).append(i)
        for shp, idx in shapes.items():
            X = torch.from_numpy(np.stack([mats[i] for i in idx])).to(dev)
            S[idx] = ops.covariance(X)
    cc = ops.ConditionedCovariance(S, offset=offset, repair=True)
    out = cc.S
    out._uglad_eig = (out._version, cc)
    return out


class CovariancePrefetcher:
    """Input pipeline for repeated fits on fresh sample batches: the host->device copy of the NEXT
    batch and its covariance + conditioning run on a side stream while the current batch trains.

        pf = CovariancePrefetcher(); pf.submit(X0_pinned)
        for each step:  S = pf.get(); loss = step(S); pf.submit(X_next_pinned); loss.item()

    Contract: the tensors are allocated on the side stream's pool, so the consumer must have
    synchronised the work that reads S (reading the step's loss does) before S is dropped; the
    prefetcher additionally keeps the two most recent batches referenced."""

    def __init__(self, device=None, offset: float = 0.1):
        self.device = _device() if device is None else device
        self.stream = torch.cuda.Stream(device=self.device)
        self.offset = offset
        self._pending = None
        self._keep = []

    def submit(self, X_host: torch.Tensor) -> None:
        with torch.cuda.stream(self.stream):
            Xd = X_host.to(self.device, non_blocking=True)
            S = get_covariance(Xd, offset=self.offset)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._pending = (S, ev, Xd)

    def get(self) -> torch.Tensor:
        S, ev, Xd = self._pending
        self._pending = None
        torch.cuda.current_stream(self.device).wait_event(ev)
        self._keep = (self._keep + [(S, Xd)])[-2:]
        return S


# ---- host-side table hygiene (prepare_data.py:361-516 with its default arguments) ---------
def normalize_table(df, typeN: str):
    if typeN == "min_max":
        return (df - df.min()) / (df.max() - df.min())
    if typeN == "mean":
        return (df - df.mean()) / df.std()
    return df


def process_table(table, NORM: str = "no", MIN_VARIANCE: float = 0.0, msg: str = "",
                  COND_NUM: float = np.inf, eigval_th: float = 1e-3, VERBOSE: bool = True):
    """Drop all-zero rows, mean-impute NaNs, drop constant and duplicate columns, normalise,
    drop columns whose variance is below MIN_VARIANCE.  (The reference's optional
    condition-number pruning loop only runs for a finite COND_NUM, which fit() never passes.)"""
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
        raise NotImplementedError("condition-number pruning (COND_NUM < inf) is outside the hot path")
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
