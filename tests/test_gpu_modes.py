"""The callers either side of the hot path, through the drop-in API on the GPU, against the REAL
reference's outputs (tests/golden/modes.npz, produced by tests/golden/make_golden.py modes):
uGLAD_GL.fit in direct (with the log-cosh structure prior), cv and missing (consensus) modes and
uGLAD_multitask.fit.  Same seeds, same data, same epochs: precision_ within 1e-4 relative
Frobenius error, recovered edge set identical up to entries within 1e-5 of the threshold."""
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


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(ROOT, "tests", "golden", "modes", "modes.npz"))


@pytest.fixture(autouse=True)
def _cold():
    from uglad_b200 import ops
    assert torch.cuda.is_available()
    ops.reset_warm_start()
    yield
    ops.reset_warm_start()


def test_fit_direct_with_structure_prior_matches_reference(g):
    """main.py:338-425 + the log-cosh prior of :325-334 (struct_theta = true theta)."""
    from uglad_b200 import main as ug
    torch.manual_seed(31)
    m = ug.uGLAD_GL()
    with np.errstate(all="ignore"):
        cmp_ = m.fit(g["direct/X"].copy(), true_theta=g["direct/true_theta"], epochs=10, lr=0.01, L=15, verbose=False,
                     mode="direct")
    assert rel(m.precision_, g["direct/precision"]) < THETA_TOL
    assert edges_match(m.precision_, g["direct/precision"])
    assert rel(m.covariance_, g["direct/covariance"]) < 1e-12 and rel(m.location_, g["direct/location"]) < 1e-12
    for k, v in zip([str(k) for k in g["direct/metric_keys"]], g["direct/metric_vals"]):
        assert (np.isnan(v) and np.isnan(cmp_[k])) or abs(cmp_[k] - v) <= 2e-3, (k, cmp_[k], v)


def test_struct_prior_loss_and_gradient_match_oracle():
    """uglad_glasso_loss_prior against torch autograd of main.py:306-334, both product back-ends."""
    from uglad_b200 import main as ug, ops
    rng = np.random.default_rng(41)
    for D, B in [(7, 3), (40, 2)]:
        A = rng.standard_normal((B, D, 2 * D))
        theta = torch.tensor(A @ A.transpose(0, 2, 1) / D + 0.5 * np.eye(D), dtype=torch.float32)
        S = torch.tensor(rng.standard_normal((B, D, D)), dtype=torch.float32)
        S = S + S.transpose(1, 2)
        st = torch.tensor((rng.random((B, D, D)) < 0.3) * rng.uniform(0.5, 1.5, (B, D, D)), dtype=torch.float32)
        t_o = theta.clone().requires_grad_(True)
        loss_o = O.glasso_loss(t_o, S, st)
        loss_o.backward()
        for small in (166, 0):
            ops.tune("small_d_max", small)
            try:
                t_g = theta.cuda().requires_grad_(True)
                loss = ug.loss_uGLAD(t_g, S.cuda(), struct_theta=st.cuda())
                loss.backward()
            finally:
                ops.tune("small_d_max", 166)
            assert abs(loss.item() - loss_o.item()) < 1e-5 * max(1.0, abs(loss_o.item()))
            assert rel(t_g.grad.cpu().numpy(), t_o.grad.numpy()) < 2e-5


def test_fit_cv_matches_reference(g):
    """main.py:428-550: 3 folds, best held-out loss per fold, best fold's model on the full covariance."""
    from uglad_b200 import main as ug
    torch.manual_seed(32)
    m = ug.uGLAD_GL()
    m.fit(g["cv/X"].copy(), epochs=10, lr=0.01, L=15, verbose=False, mode="cv", k_fold=3)
    assert rel(m.precision_, g["cv/precision"]) < THETA_TOL
    assert edges_match(m.precision_, g["cv/precision"])


def test_fit_missing_matches_reference(g):
    """main.py:553-644: 20 % NaNs, K = 4 row-subsampled covariances against the full-data covariance,
    consensus by majority sign / minimum magnitude."""
    from uglad_b200 import main as ug
    torch.manual_seed(33)
    m = ug.uGLAD_GL()
    m.fit(g["missing/X"].copy(), epochs=10, lr=0.01, L=15, verbose=False, mode="missing", k_fold=4)
    assert rel(m.precision_, g["missing/precision"]) < THETA_TOL
    assert edges_match(m.precision_, g["missing/precision"])
    assert rel(m.covariance_, g["missing/covariance"]) < 1e-12


def test_fit_multitask_matches_reference(g):
    """main.py:155-226, :719-789: ragged sample counts, one shared model."""
    from uglad_b200 import main as ug
    torch.manual_seed(34)
    mt = ug.uGLAD_multitask()
    mt.fit([g[f"multitask/X{i}"].copy() for i in range(3)], epochs=10, lr=0.01, L=15, verbose=False)
    assert rel(mt.precision_, g["multitask/precision"]) < THETA_TOL
    assert edges_match(mt.precision_, g["multitask/precision"])
    assert rel(mt.covariance_, g["multitask/covariance"]) < 1e-12


def test_consensus_covariances_match_the_reference_pipeline():
    """main.py:598-610 on the device: the K row-subsampled covariances and the full-data covariance."""
    from uglad_b200 import main as ug
    from uglad_b200.utils import prepare_data
    rng = np.random.default_rng(42)
    X = rng.random((203, 14))
    S_K, Sb = ug.consensus_covariances(prepare_data.convert_to_torch(X), 5)
    want = O.covariance([X[tr] for tr, _ in O.kfold_blocks(203, 5)])
    assert rel(S_K.cpu().numpy(), want) < 5e-6
    assert rel(Sb.cpu().numpy(), O.covariance(X[None])) < 5e-6


@pytest.mark.parametrize("B,D", [(3, 20), (1, 100), (2, 200)])
def test_graphed_epochs_equal_eager_epochs(B, D):
    """ops.GraphedStep: the epoch (zero_grad, forward, loss, backward, Adam) captured as two alternating CUDA
    graphs and replayed must walk the same training trajectory as eager calls -- same kernels, same
    warm-start chain (each epoch seeded by the previous one's workspace), capturable Adam on both sides."""
    from uglad_b200 import main as ug, ops
    rng = np.random.default_rng(B * 100 + D)
    X = rng.random((B, 3 * D, D))
    S = torch.tensor(O.covariance(X), dtype=torch.float32).cuda()
    E = 9

    def fresh():
        torch.manual_seed(5)
        model, opt = ug.init_uGLAD(lr=0.01, capturable=True)
        with torch.no_grad():
            model.rho_l1[4].bias.fill_(-6.0)
        ops.reset_warm_start()
        return model, opt

    model, opt = fresh()
    eager = []
    for _ in range(E):
        opt.zero_grad()
        th, loss = ug.forward_uGLAD(S, model, L=15)
        loss.backward()
        opt.step()
        eager.append(loss.item())
    th_e = th.detach().clone()
    model, opt = fresh()
    gs = ops.GraphedStep(S, model, opt, L=15)
    n_eager = gs.eager_epochs
    graphed = []
    for _ in range(E - n_eager):
        th, loss = gs.step()
        graphed.append(loss.item())
    assert np.allclose(graphed, eager[n_eager:], rtol=2e-6, atol=1e-6), (graphed, eager[n_eager:])
    assert rel(th.cpu().numpy(), th_e.cpu().numpy()) < 2e-5


def test_fit_loop_replays_epochs_and_keeps_the_nan_stop():
    """main._fit_loop replays its epochs from CUDA graphs (lazy ops.GraphedStep: four eager epochs, then captures):
    the direct mode's fit must equal the eager fit (UGLAD_EAGER_FIT=1), and the NaN stop of main.py:405-409 must
    still leave the parameters of the epoch BEFORE the NaN loss (the replayed update is taken back)."""
    import os
    from uglad_b200 import main as ug, ops
    rng = np.random.default_rng(3)
    X = rng.random((1, 60, 20))
    out = {}
    for eager in ("1", "0"):
        os.environ["UGLAD_EAGER_FIT"] = eager
        try:
            torch.manual_seed(4)
            ops.reset_warm_start()
            th, _, model = ug.run_uGLAD_direct(X, EPOCHS=14, lr=0.01, L=15, VERBOSE=False)
            out[eager] = (th.detach().cpu().numpy(), np.array(model.loss_values_))
        finally:
            os.environ.pop("UGLAD_EAGER_FIT", None)
    assert rel(out["0"][0], out["1"][0]) < 2e-5
    assert np.allclose(out["0"][1], out["1"][1], rtol=2e-6, atol=1e-6)
    # NaN stop on a replayed epoch: a NaN covariance makes the loss NaN
    S = torch.tensor(O.covariance(X), dtype=torch.float32).cuda()
    torch.manual_seed(4)
    model, opt = ug.init_uGLAD(lr=0.01, capturable=True)
    gs = ops.GraphedStep(S, model, opt, L=15, lazy=True, nan_guard=True)
    for _ in range(6):
        _, loss, stopped = gs.step_guarded(True)
        assert not stopped
    before = [p.detach().clone() for p in model.parameters()]
    gs.S.fill_(float("nan"))   # the graphs' static covariance: <S, theta> and with it the loss is NaN from here on
    _, loss, stopped = gs.step_guarded(True)
    assert stopped and bool(torch.isnan(loss))
    for p, q in zip(model.parameters(), before):
        assert torch.equal(p.detach(), q)
