"""Training / inference front-end of uGLAD on B200 (reference: uglad/main.py).

Mirrors the reference's public names for the hot path -- uGLAD_GL, uGLAD_multitask,
init_uGLAD, forward_uGLAD, loss_uGLAD, run_uGLAD_direct / _CV / _missing / _multitask,
mean_imputation, get_final_precision_from_batch -- with the same arguments, return values
and training loop (Adam on the glasso loss, main.py:389-414).  Visualisation, pickling and
the MAP-estimate helpers of the reference are outside this path and are not provided.
"""
from __future__ import annotations

import copy
import os
import sys
from time import time
from typing import Optional

import numpy as np
import torch

from . import ops
from .glad import glad
from .glad.glad_params import GladParams
from .utils import prepare_data
from .utils.metrics import report_metrics_all


# ---- model ------------------------------------------------------------------------------
def init_uGLAD(lr: float, theta_init_offset: float = 1.0, nF: int = 3, H: int = 3, capturable: bool = False):
    """main.py:233-249.  `capturable`: an optimizer that ops.GraphedStep can capture in a CUDA graph."""
    model = GladParams(theta_init_offset=theta_init_offset, nF=nF, H=H)
    return model, glad.get_optimizers(model, lr_glad=lr, capturable=capturable)


def loss_uGLAD(theta: torch.Tensor, S: torch.Tensor, struct_theta: Optional[torch.Tensor] = None,
               group=None, total_graphs: Optional[int] = None, replicated_S: bool = False) -> torch.Tensor:
    """main.py:289-335: sum_b(-logdet theta_b + <S_b, theta_b>) / B with B = S.shape[0]
    (the number of graphs over all processes when `group` shards them), plus the optional
    log-cosh structure prior."""
    if replicated_S or (group is None and total_graphs is None):
        B = S.shape[0]   # the reference's divisor (main.py:303); a replicated S (the broadcast full-data covariance of
                         # the consensus mode, main.py:620-622) is not sharded, so its count is not summed over ranks
    else:
        B = int(total_graphs) if total_graphs is not None else ops.global_graph_count(S.shape[0], S.device, group)
    return ops.GlassoLossFunction.apply(theta, S, float(B), struct_theta)


def forward_uGLAD(Sb, model_glad, L: int = 15, INIT_DIAG: int = 0, loss_Sb=None, struct_theta=None,
                  group=None, total_graphs: Optional[int] = None):
    """main.py:252-286: theta = glad(Sb); loss = glasso(theta, loss_Sb or Sb).  With `group` the
    graphs are sharded over ranks; `total_graphs` (their count over all ranks) saves the count
    all-reduce of every call when the caller already knows it."""
    predTheta = glad.glad(Sb, model_glad, L=L, INIT_DIAG=INIT_DIAG, group=group, total_graphs=total_graphs)
    loss = loss_uGLAD(predTheta, Sb if loss_Sb is None else loss_Sb, struct_theta=struct_theta, group=group,
                      total_graphs=total_graphs, replicated_S=loss_Sb is not None)
    return predTheta, loss


def _fit_loop(Sb, model, optimizer, EPOCHS, L, INIT_DIAG, VERBOSE, loss_Sb=None, struct_theta=None,
              tag="", stop_on_nan=False, group=None):
    """The epoch loop shared by the direct / missing / multitask modes (main.py:389-414,
    :616-630, :766-778): zero_grad, forward, backward, Adam step.  The loss is only pulled
    to the host when it is printed (or when the NaN guard of the direct mode needs it)."""
    every = max(int(EPOCHS / 10), 1)
    predTheta, losses = None, []
    total = ops.global_graph_count(Sb.shape[0], Sb.device, group) if group is not None else None
    # The epochs replay from CUDA graphs (ops.GraphedStep: four eager epochs, then two alternating captures) when the
    # fit is long enough to pay for the capture and the optimizer is capturable; same trajectory as eager epochs.
    gs = None
    if (Sb.is_cuda and EPOCHS >= 8 and os.environ.get("UGLAD_EAGER_FIT", "0") != "1"
            and all(g.get("capturable", False) for g in optimizer.param_groups)):
        gs = ops.GraphedStep(Sb, model, optimizer, L=L, INIT_DIAG=INIT_DIAG, loss_S=loss_Sb, struct_theta=struct_theta,
                             group=group, total_graphs=total, lazy=True, nan_guard=stop_on_nan)
    for e in range(EPOCHS):
        if gs is not None:
            # (sharded fits never use the NaN stop: only the direct mode does, main.py:405-409)
            predTheta, loss, stopped = gs.step_guarded(stop_on_nan and group is None)
            shown = loss.detach().clone()
            if stopped:
                print(f"Warning: NaN loss encountered at epoch {e}. Try updating the parameters and train.")
                break
            if VERBOSE and not e % every:
                if group is not None:
                    shown = ops.allreduce_sum(shown.reshape(1), group)[0]
                print(f"{tag}epoch:{e}/{EPOCHS} loss:{shown.item()}")
            losses.append(shown)
            continue
        optimizer.zero_grad()
        predTheta, loss = forward_uGLAD(Sb, model, L=L, INIT_DIAG=INIT_DIAG, loss_Sb=loss_Sb,
                                        struct_theta=struct_theta, group=group, total_graphs=total)
        shown = loss.detach()
        if group is not None and (stop_on_nan or (VERBOSE and not e % every)):
            # each rank holds the loss of its own graphs (already divided by the global count): the
            # reported value and the NaN stop must be the same on every rank
            shown = ops.allreduce_sum(shown.clone().reshape(1), group)[0]
        if stop_on_nan and bool(torch.isnan(shown)):
            print(f"Warning: NaN loss encountered at epoch {e}. Try updating the parameters and train.")
            break
        loss.backward()
        if VERBOSE and not e % every:
            print(f"{tag}epoch:{e}/{EPOCHS} loss:{shown.item()}")
        optimizer.step()
        losses.append(shown)
    if gs is not None:
        if predTheta is not None:
            predTheta = predTheta.clone()   # the replayed epochs write their theta into static buffers
        gs.close()                          # graphs (and a captured gradient all-reduce) are released here, not by the collector
    return predTheta, losses


def _share_model(model, group):
    """One model over all ranks: rank 0's initialisation everywhere."""
    import torch.distributed as dist
    for p in model.parameters():
        dist.broadcast(p.data, src=dist.get_global_rank(group, 0), group=group)


def _compare(trueTheta, predTheta, b=0):
    return report_metrics_all(trueTheta[b].detach().cpu().numpy(), predTheta[b].detach().cpu().numpy())


# ---- modes --------------------------------------------------------------------------------
def run_uGLAD_direct(Xb, trueTheta=None, eval_offset=0.1, EPOCHS=250, lr=0.002, INIT_DIAG=0, L=15,
                     VERBOSE=True):
    """main.py:338-425."""
    Sb = prepare_data.get_covariance(Xb, offset=eval_offset)
    if trueTheta is not None:
        trueTheta = prepare_data.convert_to_torch(trueTheta, req_grad=False)
    model_glad, optimizer_glad = init_uGLAD(lr=lr, theta_init_offset=1.0, nF=3, H=3, capturable=Sb.is_cuda)
    predTheta, losses = _fit_loop(Sb, model_glad, optimizer_glad, EPOCHS, L, INIT_DIAG, VERBOSE,
                                  struct_theta=trueTheta, stop_on_nan=True)
    compare_theta = None
    if trueTheta is not None:
        for b in range(Sb.shape[0]):
            compare_theta = _compare(trueTheta, predTheta, b)
            print(f"Compare - {compare_theta}")
    model_glad.loss_values_ = [float(v) for v in torch.stack(losses).cpu()] if losses else []
    return predTheta, compare_theta, model_glad


def run_uGLAD_CV(Xb, trueTheta=None, eval_offset=0.1, EPOCHS=250, lr=0.002, INIT_DIAG=0, L=15,
                 VERBOSE=True, k_fold=5):
    """main.py:428-550: per fold train on the fold's covariance, keep the parameters with the
    best held-out glasso loss; the best fold's model is rerun on the full covariance."""
    from sklearn.model_selection import KFold
    Xb = np.asarray(Xb)
    Sb = prepare_data.get_covariance(Xb, offset=eval_offset)
    if trueTheta is not None:
        trueTheta = prepare_data.convert_to_torch(trueTheta, req_grad=False)
    results = {}
    for fold, (train, test) in enumerate(KFold(n_splits=k_fold).split(Xb[0])):
        if VERBOSE:
            print(f"Fold num {fold}")
        Sb_train = prepare_data.get_covariance(Xb[:, train], offset=eval_offset)
        Sb_test = prepare_data.get_covariance(Xb[:, test], offset=eval_offset)
        model_glad, optimizer_glad = init_uGLAD(lr=lr, theta_init_offset=1.0, nF=3, H=3)
        best_loss, best_model = np.inf, None
        every = max(int(EPOCHS / 10), 1)
        for e in range(EPOCHS):
            optimizer_glad.zero_grad()
            _, loss_train = forward_uGLAD(Sb_train, model_glad, L=L, INIT_DIAG=INIT_DIAG)
            with torch.no_grad():
                _, loss_test = forward_uGLAD(Sb_test, model_glad, L=L, INIT_DIAG=INIT_DIAG)
            loss_train.backward()
            optimizer_glad.step()
            _loss = loss_test.item()
            if VERBOSE and not e % every:
                print(f"Fold {fold}: epoch:{e}/{EPOCHS} test-loss:{_loss}")
            if _loss < best_loss:
                best_model, best_loss = copy.deepcopy(model_glad), _loss
        results[fold] = {"test_loss": best_loss, "model": best_model}
    model_glad = min(results.values(), key=lambda r: r["test_loss"])["model"]
    with torch.no_grad():
        predTheta, _ = forward_uGLAD(Sb, model_glad, L=L, INIT_DIAG=INIT_DIAG)
    compare_theta = None
    if trueTheta is not None:
        compare_theta = _compare(trueTheta, predTheta, Sb.shape[0] - 1)
        print(f"Comparison - {compare_theta}")
    return predTheta, compare_theta, model_glad


def mean_imputation(Xb: np.ndarray) -> np.ndarray:
    """main.py:647-670: replace NaNs by their column mean (in place, like the reference)."""
    X = Xb[0]
    col_mean = np.nanmean(X, axis=0)
    rows, cols = np.where(np.isnan(X))
    X[rows, cols] = col_mean[cols]
    if np.isnan(X.sum()):
        print("ERROR: One or more columns have all NaNs")
        sys.exit(0)
    return np.expand_dims(X, axis=0)


def get_final_precision_from_batch(predTheta: torch.Tensor, type: str = "min", group=None) -> torch.Tensor:
    """main.py:673-716: consensus over K precision matrices: majority sign (ties -> +) times
    the min (or mean) magnitude.  With `group` the K matrices are sharded over the ranks: the
    local min |theta| / sum sign(theta) are combined by one all-reduce(MIN) and one all-reduce(SUM)
    (both exact in floating point, so N ranks reproduce the single-process result bit for bit)."""
    K, _, D = predTheta.shape
    mag = torch.abs(predTheta)
    votes = torch.sum(torch.sign(predTheta), 0)
    if type == "min":
        value = torch.min(mag, 0)[0]
        if group is not None:
            value = ops.allreduce_min(value.contiguous(), group)
    elif type == "mean":
        if group is not None:
            raise NotImplementedError("sharded consensus implements type='min' (the reference's setting, main.py:632)")
        value = torch.mean(mag, 0)[0]  # sic: the reference indexes the mean too (main.py:705)
    else:
        print(f"Enter valid type min/mean, currently {type}")
        sys.exit(0)
    if group is not None:
        votes = ops.allreduce_sum(votes.contiguous(), group)   # sums of +-1/0: exact
    sign = torch.where(votes >= 0, torch.ones_like(votes), -torch.ones_like(votes))
    return (sign * value).reshape(1, D, D)


def kfold_train_indices(M: int, K: int):
    """Row indices of the K training folds of sklearn's KFold(n_splits=K) without shuffling (the
    reference's row-subsampled batches, main.py:604-606): fold k drops one contiguous block of
    M // K (+1 for the first M % K folds) rows."""
    sizes = np.full(K, M // K, dtype=np.int64)
    sizes[: M % K] += 1
    stops = np.cumsum(sizes)
    rows = np.arange(M)
    return [np.concatenate([rows[: stop - size], rows[stop:]]) for size, stop in zip(sizes, stops)]


def consensus_covariances(X: torch.Tensor, K_batch: int, eval_offset: float = 0.1, group=None, warm=None):
    """main.py:598-610 on the device: X [M, D] (imputed samples, CUDA) -> (S_K, Sb): the covariances
    of this rank's share of the K row-subsampled batches and the full-data covariance [1, D, D].
    The sub-samples are gathered on the device and go through the batched covariance kernels (two
    launches: the folds come in at most two sizes)."""
    M = X.shape[0]
    folds = kfold_train_indices(M, K_batch)
    mine = range(K_batch)
    if group is not None:
        import torch.distributed as dist
        mine = np.array_split(np.arange(K_batch), dist.get_world_size(group))[dist.get_rank(group)]
    X_K = [X[torch.as_tensor(folds[k], device=X.device)] for k in mine]
    w_full, w_K = (warm if warm is not None else (None, None))
    Sb = prepare_data.get_covariance(X.unsqueeze(0), offset=eval_offset, warm=w_full)
    S_K = prepare_data.get_covariance(X_K, offset=eval_offset, warm=w_K)
    return S_K, Sb


def run_uGLAD_missing(Xb, trueTheta=None, eval_offset=0.1, EPOCHS=250, lr=0.002, INIT_DIAG=0, L=15,
                      VERBOSE=True, K_batch=3, group=None):
    """main.py:553-644: mean-impute, build K row-subsampled covariances (the training folds of
    a K-fold split), fit one model on all K with the full-data covariance in the loss, then
    take the consensus.  With `group` the K imputations are sharded over the ranks (every rank
    passes the same Xb): the model is shared exactly as in the multitask mode and the consensus
    is reduced across ranks."""
    if K_batch == 0:
        K_batch = 3
    Xb = mean_imputation(np.array(Xb, dtype=np.float64))
    print(f"Creating K={K_batch} row-subsampled batches")
    S_K, Sb = consensus_covariances(prepare_data.convert_to_torch(Xb[0]), K_batch, eval_offset, group)
    if trueTheta is not None:
        trueTheta = prepare_data.convert_to_torch(trueTheta, req_grad=False)
    model_glad, optimizer_glad = init_uGLAD(lr=lr, theta_init_offset=1.0, nF=3, H=3, capturable=S_K.is_cuda)
    if group is not None:
        _share_model(model_glad, group)
    predTheta, _ = _fit_loop(S_K, model_glad, optimizer_glad, EPOCHS, L, INIT_DIAG, VERBOSE, loss_Sb=Sb, group=group)
    print("Getting the final precision matrix using the consensus strategy")
    predTheta = get_final_precision_from_batch(predTheta.detach(), type="min", group=group)
    compare_theta = None
    if trueTheta is not None:
        compare_theta = _compare(trueTheta, predTheta, 0)
        print(f"Comparison - {compare_theta}")
    return predTheta, compare_theta, model_glad


def run_uGLAD_multitask(Xb, trueTheta=None, eval_offset=0.1, EPOCHS=250, lr=0.002, INIT_DIAG=0, L=15,
                        VERBOSE=True, group=None):
    """main.py:719-789.  With `group`, Xb holds this process's shard of the graphs: the model
    is shared, the Frobenius mean and the loss run over all shards and the MLP gradients are
    all-reduced once per epoch."""
    Sb = prepare_data.get_covariance(Xb, offset=eval_offset)
    if trueTheta is not None:
        trueTheta = prepare_data.convert_to_torch(trueTheta, req_grad=False)
    model_glad, optimizer_glad = init_uGLAD(lr=lr, theta_init_offset=1.0, nF=3, H=3, capturable=Sb.is_cuda)
    if group is not None:
        _share_model(model_glad, group)
    predTheta, _ = _fit_loop(Sb, model_glad, optimizer_glad, EPOCHS, L, INIT_DIAG, VERBOSE, group=group)
    compare_theta = []
    if trueTheta is not None:
        for b in range(len(Xb)):
            rM = _compare(trueTheta, predTheta, b)
            print(f"Metrics for graph {b}: {rM}\n")
            compare_theta.append(rM)
    return predTheta, compare_theta, model_glad


# ---- sklearn-style wrappers -----------------------------------------------------------------
def _clean(X, verbose):
    return np.array(prepare_data.process_table(X, NORM="min_max", VERBOSE=verbose))


class uGLAD_GL(object):
    """GraphicalLassoCV-style wrapper (main.py:34-151): fit(X) fills covariance_, precision_,
    location_, node_names_, model_glad."""

    def __init__(self):
        self.covariance_ = None
        self.precision_ = None
        self.location_ = None
        self.model_glad = None

    def fit(self, X, true_theta=None, eval_offset=0.1, centered=False, epochs=250, lr=0.002, INIT_DIAG=0,
            L=15, verbose=True, k_fold=3, mode="direct", node_names=None):
        print("Running uGLAD")
        start = time()
        X = _clean(X, verbose)
        M, D = X.shape
        Xb = X.reshape(1, M, D)
        true_theta_b = None if true_theta is None else np.asarray(true_theta).reshape(1, D, D)
        common = dict(trueTheta=true_theta_b, eval_offset=eval_offset, EPOCHS=epochs, lr=lr,
                      INIT_DIAG=INIT_DIAG, L=L, VERBOSE=verbose)
        if mode == "missing":
            print("Handling missing data")
            pred_theta, compare_theta, model_glad = run_uGLAD_missing(Xb, K_batch=k_fold, **common)
        elif mode == "cv" and k_fold >= 0:
            print(f"CV mode: {k_fold}-fold")
            pred_theta, compare_theta, model_glad = run_uGLAD_CV(Xb, k_fold=k_fold, **common)
        elif mode == "direct":
            print("Direct Mode")
            pred_theta, compare_theta, model_glad = run_uGLAD_direct(Xb, **common)
        else:
            print(f"ERROR Please enter K-fold value in valid range [0, ), currently entered {k_fold}; "
                  f"Check mode {mode}")
            sys.exit(0)
        Xc = X if centered else X - X.mean(axis=0)
        self.covariance_ = Xc.T @ Xc / M
        self.location_ = X.mean(axis=0)
        self.node_names_ = list(node_names) if node_names is not None else [f"node_{i}" for i in range(D)]
        if pred_theta is not None:
            self.precision_ = pred_theta[0].detach().cpu().numpy()
        if model_glad is not None:
            self.model_glad = model_glad
        print(f"Total runtime: {time() - start} secs\n")
        return compare_theta


class uGLAD_multitask(object):
    """main.py:155-226: one shared model, a list of sample matrices (sample counts may differ)."""

    def __init__(self):
        self.covariance_ = []
        self.precision_ = None
        self.model_glad = None

    def fit(self, Xb, true_theta_b=None, eval_offset=0.1, centered=False, epochs=250, lr=0.002,
            INIT_DIAG=0, L=15, verbose=True, group=None):
        print("Running uGLAD in multi-task mode")
        start = time()
        Xb = [_clean(X, verbose) for X in Xb]
        pred_theta, compare_theta, model_glad = run_uGLAD_multitask(
            Xb, trueTheta=true_theta_b, eval_offset=eval_offset, EPOCHS=epochs, lr=lr,
            INIT_DIAG=INIT_DIAG, L=L, VERBOSE=verbose, group=group)
        cov = []
        for X in Xb:
            Xc = X if centered else X - X.mean(axis=0)
            cov.append(Xc.T @ Xc / X.shape[0])
        self.covariance_ = np.array(cov)
        self.precision_ = pred_theta.detach().cpu().numpy()
        self.model_glad = model_glad
        print(f"Total runtime: {time() - start} secs\n")
        return compare_theta
