"""Edge-recovery metrics (reference: uglad/utils/metrics.py:25-108), numpy only."""
from __future__ import annotations

import numpy as np


def _roc_auc(y: np.ndarray, score: np.ndarray) -> float:
    """Area under the ROC curve via the rank statistic (ties get the average rank)."""
    pos, neg = int(y.sum()), int((1 - y).sum())
    if pos == 0 or neg == 0:
        return float("nan")
    order = np.argsort(score, kind="mergesort")
    s = score[order]
    ranks = np.empty(len(s), dtype=np.float64)
    i = 0
    while i < len(s):
        j = i
        while j + 1 < len(s) and s[j + 1] == s[i]:
            j += 1
        ranks[i:j + 1] = 0.5 * (i + j) + 1.0
        i = j + 1
    r = np.empty_like(ranks)
    r[order] = ranks
    return float((r[y == 1].sum() - pos * (pos + 1) / 2.0) / (pos * neg))


def _average_precision(y: np.ndarray, score: np.ndarray) -> float:
    """sum_n (R_n - R_{n-1}) P_n over distinct thresholds (sklearn's definition)."""
    pos = int(y.sum())
    if pos == 0:
        return float("nan")
    order = np.argsort(-score, kind="mergesort")
    y, s = y[order], score[order]
    tp = np.cumsum(y)
    last = np.r_[np.where(np.diff(s))[0], len(s) - 1]  # last index of every tie group
    prec = tp[last] / (last + 1.0)
    rec = tp[last] / pos
    return float(np.sum(np.diff(np.r_[0.0, rec]) * prec))


def get_auc(y, scores):
    y = np.asarray(y).astype(int)
    scores = np.asarray(scores, dtype=np.float64)
    return _roc_auc(y, scores), _average_precision(y, scores)


def report_metrics_all(trueG: np.ndarray, G: np.ndarray, beta: int = 1) -> dict:
    """FDR, TPR, FPR, SHD, nnz, precision, recall, F-beta, AUPR and AUC of the off-diagonal
    support of G against trueG (upper triangle), rounded to 3 decimals like the reference."""
    trueG, G = np.asarray(trueG).real, np.asarray(G).real
    iu = np.triu_indices(G.shape[-1], 1)
    t = (trueG[iu] != 0).astype(int)
    p = (G[iu] != 0).astype(int)
    auc, aupr = get_auc(t, np.abs(G[iu]))
    TP = int(np.sum(t * p))
    FP = int(np.sum((1 - t) * p))
    FN = int(np.sum(t * (1 - p)))
    P, T = int(p.sum()), int(t.sum())
    F = len(t) - T
    with np.errstate(divide="ignore", invalid="ignore"):
        out = {
            "FDR": np.float64(FP) / P, "TPR": np.float64(TP) / T, "FPR": np.float64(FP) / F,
            "SHD": FP + FN, "nnzTrue": T, "nnzPred": P,
            "precision": np.float64(TP) / (TP + FP), "recall": np.float64(TP) / (TP + FN),
            "Fbeta": np.float64((1 + beta ** 2) * TP) / ((1 + beta ** 2) * TP + beta ** 2 * FN + FP),
            "aupr": aupr, "auc": auc,
        }
    return {k: round(float(v), 3) for k, v in out.items()}
