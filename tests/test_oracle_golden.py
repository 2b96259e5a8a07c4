"""The oracle pinned against the real reference: tests/golden/*.npz were produced by
tests/golden/make_golden.py importing /root/reference.  The matmul formulation must
reproduce them to float32 round-off and the float64 spectral formulation (the algorithm
the CUDA kernels implement) must agree with the matmul one."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import uglad_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def params(g, tag, dtype=torch.float32):
    return {k: torch.tensor(g[f"{tag}/{k}"], dtype=dtype).requires_grad_(True) for k in O.PARAM_KEYS}


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_matmul_oracle_reproduces_reference(path):
    torch.set_num_threads(1)
    g = np.load(path)
    P = params(g, "p0")
    theta, loss = O.forward_loss(torch.tensor(g["S"]), P, int(g["L"]), int(g["init_diag"]))
    loss.backward()
    assert rel(theta.detach().numpy(), g["theta0"]) < 1e-5
    assert abs(loss.item() - float(g["loss0"])) < 1e-4 * max(1.0, abs(float(g["loss0"])))
    for k in O.PARAM_KEYS:
        assert rel(P[k].grad.numpy(), g["g0/" + k]) < 1e-4, k


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_spectral_formulation_is_the_reference_math(path):
    g = np.load(path)
    P64 = params(g, "p0", torch.float64)
    S = torch.tensor(g["S"], dtype=torch.float64)
    theta, loss = O.forward_loss(S, P64, int(g["L"]), int(g["init_diag"]))
    loss.backward()
    r = O.spectral_forward_backward(g["S"], P64, int(g["L"]), int(g["init_diag"]))
    assert rel(r["theta"], theta.detach().numpy()) < 1e-10
    assert abs(r["loss"] - loss.item()) < 1e-9
    for k in O.PARAM_KEYS:
        assert rel(r["grads"][k], P64[k].grad.numpy()) < 1e-8, k
    # and float64 agrees with the float32 reference to float32 accuracy
    assert rel(r["theta"], g["theta0"]) < 5e-5


def test_covariance_matches_reference(golden_dir):
    for name in ("d10_m500", "d20_b3_multitask"):
        g = np.load(os.path.join(golden_dir, name + ".npz"))
        assert rel(O.covariance(g["X"]), g["S64"]) < 1e-12


def test_training_trajectory_short(golden_dir):
    torch.set_num_threads(1)
    g = np.load(os.path.join(golden_dir, "d10_m500.npz"))
    P = params(g, "p0")
    thT, losses = O.train(torch.tensor(g["S"]), P, int(g["epochs"]), float(g["lr"]), int(g["L"]), int(g["init_diag"]))
    assert np.abs(losses - g["losses"]).max() < 1e-4
    assert rel(thT.numpy(), g["thetaT"]) < 1e-5


def test_consensus(golden_dir):
    g = np.load(os.path.join(golden_dir, "d20_b3_multitask.npz"))
    c = O.consensus_min(torch.tensor(g["thetaT"]))
    assert rel(c.numpy(), g["consensus"]) == 0.0


def test_rank_deficient_covariance_is_repaired():
    rng = np.random.default_rng(3)
    X = rng.standard_normal((1, 5, 12))  # M < D
    S = O.covariance(X, offset=0.1)
    assert abs(np.linalg.eigvalsh(S[0]).min() - 0.1) < 1e-8
