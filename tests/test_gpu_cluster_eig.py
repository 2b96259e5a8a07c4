"""The cluster eigensolver (csrc/eig_cluster.cu: warm layer solves spread over 1 / 2 / 4 CTAs per graph, odd-even
ordering, boundary columns pushed through distributed shared memory) against the oracle and against the
one-CTA kernel, plus its two rare branches: a warm start that is not positive definite (retry pass of the
one-CTA kernel) and a loose warm start (several sweeps, fix-up lists, full-sweep fallback)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import uglad_oracle as O  # the checker, never the thing under test

pytestmark = pytest.mark.gpu
THETA_TOL = 1e-4


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _cov(B, D, seed, M=None):
    rng = np.random.default_rng(seed)
    X = rng.random((B, M or 3 * D, D))
    return torch.tensor(O.covariance(X), dtype=torch.float32)


def _params(seed):
    P = O.init_params(seed)
    with torch.no_grad():
        P["rho_l1.4.bias"].fill_(-5.0)   # low threshold: a dense, moving support
    return P


def _model_from(P):
    from uglad_b200.glad.glad_params import GladParams
    model = GladParams(1.0, 3, 3)
    model.load_state_dict({k: v.detach().clone() for k, v in P.items()})
    return model.cuda()


def _info(ws, B, D, L=15):
    from uglad_b200 import ops, _lib
    dims = ops.make_dims(B, D, L, 3, 0)
    off = _lib.load().uglad_workspace_offset(C.byref(dims), b"info")
    return ws[off:off + L * B * 4].view(L, B, 4).cpu().numpy()


@pytest.fixture(autouse=True)
def _restore_knobs():
    from uglad_b200 import ops
    yield
    ops.tune("eig_cluster", -1)
    ops.reset_warm_start()


@pytest.mark.parametrize("B,D,nc", [(1, 100, 4), (1, 100, 2), (3, 100, 1), (5, 20, 4), (2, 64, 2), (2, 164, 4),
                                    (1, 16, 2), (40, 100, 2), (150, 36, 1)])
def test_warm_epochs_match_the_oracle(B, D, nc):
    """Three training epochs (the 2nd and 3rd are warm: cluster kernel) against the oracle's trajectory."""
    from uglad_b200 import main as ug, ops
    S = _cov(B, D, 11 * B + D)
    P = _params(3)
    opt_o = torch.optim.Adam(list(P.values()), lr=0.01)
    model = _model_from(P)
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    ops.tune("eig_cluster", nc)
    ops.reset_warm_start()
    Sd = S.cuda()
    for epoch in range(3):
        opt_o.zero_grad()
        th_o, loss_o = O.forward_loss(S, P, 15, 0)
        loss_o.backward()
        opt.zero_grad()
        th, loss = ug.forward_uGLAD(Sd, model, L=15)
        loss.backward()
        assert rel(th.detach().cpu().numpy(), th_o.detach().numpy()) < THETA_TOL, epoch
        assert abs(loss.item() - loss_o.item()) < 1e-4 * max(1.0, abs(loss_o.item())), epoch
        for k, p in model.named_parameters():
            assert rel(p.grad.cpu().numpy(), P[k].grad.numpy()) < 2e-3, (epoch, k)
        opt_o.step()
        opt.step()
    ws = next(reversed(ops._warm.values()))
    sw = _info(ws, B, D)[:, :, 0]
    assert sw.max() < 1000 and sw.min() >= 1   # warm solves, no retry


@pytest.mark.parametrize("nc", [1, 2, 4])
def test_cluster_sizes_agree_with_the_one_cta_kernel(nc):
    from uglad_b200 import main as ug, ops
    S = _cov(6, 100, 5).cuda()
    model = _model_from(_params(8))
    out = {}
    for cfg in (0, nc):
        ops.tune("eig_cluster", cfg)
        ops.reset_warm_start()
        with torch.no_grad():
            ug.glad.glad(S, model, L=15)
            out[cfg] = ug.glad.glad(S, model, L=15).clone()   # warm
    assert rel(out[nc].cpu().numpy(), out[0].cpu().numpy()) < 2e-5


@pytest.mark.parametrize("nc", [1, 4])
def test_indefinite_warm_start_takes_the_retry_pass(nc):
    """The warm workspace belongs to a DIFFERENT problem (theta_init_offset 1.0 instead of 0.01: |b| ~ 1 instead
    of ~ 100), so the shift taken from its eigenvalues is far too small and U0 is not positive definite: the
    check must flag the graphs (info[0] >= 1000), the retry pass solves them from scratch, and theta must equal
    the cold result."""
    from uglad_b200 import main as ug, ops
    B, D = 2, 100
    S1 = _cov(B, D, 21).cuda()
    other = _model_from(_params(4))
    P = _params(4)
    with torch.no_grad():
        P["theta_init_offset"].fill_(0.01)
    model = _model_from(P)
    n = ops.workspace_floats(B, D, 15)
    wsA, wsB, wsC = (torch.empty(n, device="cuda") for _ in range(3))
    ops.tune("eig_cluster", nc)
    with torch.no_grad():
        with ops.use_workspace(wsA, None):
            ug.glad.glad(S1, other, L=15)                 # warm state of another problem
        with ops.use_workspace(wsC, None):
            cold = ug.glad.glad(S1, model, L=15).clone()
        with ops.use_workspace(wsB, wsA):
            warm = ug.glad.glad(S1, model, L=15).clone()
    assert torch.isfinite(warm).all()
    assert rel(warm.cpu().numpy(), cold.cpu().numpy()) < 5e-5
    sw = _info(wsB, B, D)[:, :, 0]
    assert (sw >= 1000).any(), sw.max()   # the branch was exercised


@pytest.mark.parametrize("nc", [1, 2, 4])
def test_loose_warm_start_needs_more_sweeps_and_still_converges(nc):
    """Warm state from a perturbed covariance (5 % relative): several sweeps, long fix-up lists."""
    from uglad_b200 import main as ug, ops
    B, D = 3, 100
    S1 = _cov(B, D, 31)
    rng = np.random.default_rng(32)
    E = rng.standard_normal((B, D, D)).astype(np.float32) * 0.05 * float(S1.abs().mean())
    S2 = (S1 + torch.tensor(E + E.transpose(0, 2, 1)) + 0.2 * torch.eye(D)).cuda()
    S1 = S1.cuda()
    model = _model_from(_params(6))
    n = ops.workspace_floats(B, D, 15)
    wsA, wsB, wsC = (torch.empty(n, device="cuda") for _ in range(3))
    ops.tune("eig_cluster", nc)
    with torch.no_grad():
        with ops.use_workspace(wsA, None):
            ug.glad.glad(S2, model, L=15)
        with ops.use_workspace(wsC, None):
            cold = ug.glad.glad(S1, model, L=15).clone()
        with ops.use_workspace(wsB, wsA):
            warm = ug.glad.glad(S1, model, L=15).clone()
    assert rel(warm.cpu().numpy(), cold.cpu().numpy()) < 5e-5
    sw = _info(wsB, B, D)[:, :, 0]
    assert (sw % 1000).max() >= 2, sw.max()
