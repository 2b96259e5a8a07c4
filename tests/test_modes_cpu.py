"""Host-side logic either side of the hot path, and the oracle's mode drivers, pinned to the REAL
reference: tests/golden/modes.npz holds the outputs of the reference's uGLAD_GL.fit (direct with a
structure prior / cv / missing), uGLAD_multitask.fit, metrics.report_metrics_all and process_table
(tests/golden/make_golden.py modes).  Everything here runs on the CPU."""
import os

import numpy as np
import pytest
import torch

from oracle import uglad_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(ROOT, "tests", "golden", "modes", "modes.npz"))


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_report_metrics_all_matches_reference(g):
    """uglad/utils/metrics.py:25-108 (sklearn roc/auc/average precision) vs the numpy rewrite."""
    from uglad_b200.utils.metrics import report_metrics_all
    for i in range(4):
        with np.errstate(all="ignore"):
            got = report_metrics_all(g[f"metrics/{i}/true"], g[f"metrics/{i}/pred"])
        keys, vals = [str(k) for k in g[f"metrics/{i}/keys"]], g[f"metrics/{i}/vals"]
        assert sorted(got) == keys
        for k, v in zip(keys, vals):
            assert (np.isnan(v) and np.isnan(got[k])) or got[k] == pytest.approx(v, abs=1e-12), (i, k, got[k], v)


def test_report_metrics_matches_sklearn_on_ties_and_empty_predictions():
    """Tied scores (many exact zeros after soft thresholding) and the all-zero prediction."""
    from sklearn import metrics as skm
    from uglad_b200.utils.metrics import report_metrics_all
    rng = np.random.default_rng(0)
    d = 15
    tg = np.triu(rng.random((d, d)) < 0.3, 1).astype(float)
    tg = tg + tg.T + np.eye(d)
    pg = np.triu(np.round(rng.standard_normal((d, d)), 0) * (rng.random((d, d)) < 0.5), 1)   # heavy ties
    pg = pg + pg.T + np.eye(d)
    r = report_metrics_all(tg, pg)
    iu = np.triu_indices(d, 1)
    y, sc = (tg[iu] != 0).astype(int), np.abs(pg[iu])
    fpr, tpr, _ = skm.roc_curve(y, sc)
    assert r["auc"] == round(float(skm.auc(fpr, tpr)), 3)
    assert r["aupr"] == round(float(skm.average_precision_score(y, sc)), 3)
    with np.errstate(all="ignore"):
        r0 = report_metrics_all(tg, np.eye(d))
    assert r0["nnzPred"] == 0 and np.isnan(r0["FDR"]) and r0["TPR"] == 0


def test_process_table_matches_reference(g):
    """prepare_data.py:361-516: zero rows, NaNs, constant / duplicated columns, both normalisations."""
    from uglad_b200.utils.prepare_data import process_table
    for norm in ("minmax", "mean"):
        out = np.array(process_table(g["table/raw"].copy(), NORM={"minmax": "min_max", "mean": "mean"}[norm], VERBOSE=False))
        assert out.shape == g[f"table/{norm}"].shape
        assert np.allclose(out, g[f"table/{norm}"], rtol=0, atol=1e-14)
    assert np.allclose(O.clean_table_minmax(g["table/raw"]), g["table/minmax"], rtol=0, atol=1e-14)


def test_process_table_condition_number_pruning_matches_reference(g):
    """prepare_data.py:465-505 (finite COND_NUM): the same columns survive."""
    from uglad_b200.utils.prepare_data import process_table
    out = process_table(g["table/collinear"].copy(), NORM="min_max", COND_NUM=200.0, eigval_th=1e-3, VERBOSE=False)
    assert list(out.columns) == g["table/collinear_kept"].tolist()
    assert np.allclose(np.array(out), g["table/collinear_out"], rtol=0, atol=1e-14)


def test_kfold_indices_match_sklearn():
    from sklearn.model_selection import KFold
    from uglad_b200.main import kfold_train_indices
    for M, K in [(10, 3), (242, 4), (1000, 32), (7, 7), (33, 2)]:
        ours = kfold_train_indices(M, K)
        ref = [tr for tr, _ in KFold(n_splits=K).split(np.zeros((M, 1)))]
        assert len(ours) == K and all(np.array_equal(a, b) for a, b in zip(ours, ref))
        assert all(np.array_equal(a, b[0]) for a, b in zip(ours, O.kfold_blocks(M, K)))


def test_mean_imputation_and_consensus_host_logic(g):
    from uglad_b200.main import get_final_precision_from_batch, mean_imputation
    X = g["missing/X"].copy()
    out = mean_imputation(X[None].copy())[0]
    want = np.where(np.isnan(X), np.nanmean(X, axis=0), X)
    assert np.array_equal(out, want)
    th = torch.tensor(np.random.default_rng(1).standard_normal((5, 6, 6)))
    th[:, 0, 1] = torch.tensor([1.0, -1.0, 0.0, 2.0, -3.0])   # a sign tie -> +
    assert torch.equal(get_final_precision_from_batch(th, "min"), O.consensus_min(th))


def test_initial_parameters_follow_the_reference_seed(golden_dir):
    """Same construction order as glad_params.py:10-31: torch.manual_seed gives the reference's weights."""
    from uglad_b200.glad.glad_params import GladParams
    g0 = np.load(os.path.join(golden_dir, "d10_m500.npz"))
    torch.manual_seed(int(g0["seed"]))
    model = GladParams(1.0, 3, 3, device=torch.device("cpu"))
    for k, v in model.state_dict().items():
        assert np.array_equal(v.numpy(), g0["p0/" + k]), k
    P = O.init_params(int(g0["seed"]))
    for k in O.PARAM_KEYS:
        assert np.array_equal(P[k].detach().numpy(), g0["p0/" + k]), k


# ---- the oracle's mode drivers against the reference's fits -------------------------------------
def test_oracle_direct_mode_with_structure_prior(g):
    torch.set_num_threads(1)
    th = O.fit_direct(g["direct/X"], 31, 10, 0.01, true_theta=g["direct/true_theta"])
    assert rel(th, g["direct/precision"]) < 1e-5


def test_oracle_cv_mode(g):
    torch.set_num_threads(1)
    th = O.fit_cv(g["cv/X"], 32, 10, 0.01, 3)
    assert rel(th, g["cv/precision"]) < 1e-5


def test_oracle_missing_mode(g):
    torch.set_num_threads(1)
    th = O.fit_missing(g["missing/X"], 33, 10, 0.01, 4)
    assert rel(th, g["missing/precision"]) < 1e-5


def test_oracle_multitask_mode(g):
    torch.set_num_threads(1)
    th = O.fit_multitask([g[f"multitask/X{i}"] for i in range(3)], 34, 10, 0.01)
    assert rel(th, g["multitask/precision"]) < 1e-5
