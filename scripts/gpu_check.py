"""Numeric check of every CUDA stage against the oracle / golden vectors; prints the error of
each stage so one GPU run tells the whole story.   python scripts/gpu_check.py [--big]"""
import glob
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import uglad_oracle as O  # noqa: E402  (checker only)
from uglad_b200 import ops  # noqa: E402
from uglad_b200 import main as ug  # noqa: E402
from uglad_b200.glad.glad_params import GladParams  # noqa: E402

dev = torch.device("cuda:0")
rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) /
                         max(np.linalg.norm(np.asarray(b, np.float64)), 1e-300))


def check_eigh():
    rng = np.random.default_rng(0)
    for D, B in [(3, 2), (10, 4), (33, 3), (100, 8), (129, 2), (200, 2), (232, 1)]:
        A = rng.standard_normal((B, D, D))
        A = (A + A.transpose(0, 2, 1)) / 2
        w, Vt, info = ops.eigh(torch.tensor(A, dtype=torch.float32, device=dev), indefinite=True)
        w, Vt, info = w.cpu().numpy().astype(np.float64), Vt.cpu().numpy().astype(np.float64), info.cpu().numpy()
        R = np.einsum("bki,bk,bkj->bij", Vt, w, Vt)
        orth = np.abs(np.einsum("bki,bli->bkl", Vt, Vt) - np.eye(D)).max()
        w64 = np.linalg.eigvalsh(A)
        print(f"eigh indef D={D:4d} B={B}: recon {rel(R, A):.2e} eig {np.abs(np.sort(w, 1) - w64).max() / np.abs(w64).max():.2e} "
              f"orth {orth:.2e} sweeps {info[:, 0].max():.0f} shift {info[0, 1]:.3g}")
        Pd = A @ A.transpose(0, 2, 1) + 0.1 * np.eye(D)
        w, Vt, info = ops.eigh(torch.tensor(Pd, dtype=torch.float32, device=dev), indefinite=False)
        w, Vt = w.cpu().numpy().astype(np.float64), Vt.cpu().numpy().astype(np.float64)
        R = np.einsum("bki,bk,bkj->bij", Vt, w, Vt)
        print(f"eigh   pd  D={D:4d} B={B}: recon {rel(R, Pd):.2e} sweeps {info[:, 0].max().item():.0f}")


def check_cov():
    rng = np.random.default_rng(1)
    for B, M, D in [(2, 50, 7), (3, 500, 20), (1, 1000, 100)]:
        X = rng.random((B, M, D))
        S = ops.covariance(torch.tensor(X, dtype=torch.float32, device=dev)).cpu().numpy()
        print(f"covariance B={B} M={M} D={D}: rel {rel(S, O.covariance(X, offset=0.1)):.2e}")


def load_model(g, tag):
    model = GladParams(1.0, 3, 3)
    sd = {k: torch.tensor(g[f"{tag}/{k}"]) for k in O.PARAM_KEYS}
    model.load_state_dict(sd)
    return model


def check_golden(path):
    g = np.load(path)
    name = os.path.basename(path)[:-4]
    L, idg = int(g["L"]), int(g["init_diag"])
    S = torch.tensor(g["S"], device=dev)
    model = load_model(g, "p0")
    theta, loss = ug.forward_uGLAD(S, model, L=L, INIT_DIAG=idg)
    loss.backward()
    gerr = {k: rel(p.grad.cpu().numpy(), g["g0/" + k]) for k, p in model.named_parameters()}
    print(f"{name}: theta0 rel {rel(theta.detach().cpu().numpy(), g['theta0']):.2e} loss0 abs "
          f"{abs(loss.item() - float(g['loss0'])):.2e} (loss {loss.item():.5f}) grads max rel {max(gerr.values()):.2e}")
    for k, v in gerr.items():
        if v > 1e-3:
            print(f"     grad {k}: rel {v:.2e} ours {p_fmt(dict(model.named_parameters())[k].grad)} ref {g['g0/' + k].ravel()[:4]}")
    # covariance from the stored samples
    Sg = ug.prepare_data.get_covariance(g["X"], offset=0.1)
    print(f"     covariance from X: rel {rel(Sg.cpu().numpy(), g['S']):.2e}")
    # training trajectory
    model = load_model(g, "p0")
    opt = ug.glad.get_optimizers(model, lr_glad=float(g["lr"]))
    t0 = time.time()
    th, losses = ug._fit_loop(S, model, opt, int(g["epochs"]), L, idg, False)
    torch.cuda.synchronize()
    dt = time.time() - t0
    losses = torch.stack(losses).cpu().numpy()
    perr = max(rel(p.detach().cpu().numpy(), g["pT/" + k]) for k, p in model.named_parameters())
    print(f"     train {int(g['epochs'])} epochs: thetaT rel {rel(th.detach().cpu().numpy(), g['thetaT']):.2e} loss traj max abs "
          f"{np.abs(losses - g['losses']).max():.2e} params rel {perr:.2e}  ({dt / int(g['epochs']) * 1e3:.2f} ms/epoch)")
    if "consensus" in g.files:
        c = ug.get_final_precision_from_batch(th.detach(), type="min").cpu().numpy()
        print(f"     consensus rel {rel(c, g['consensus']):.2e}")


def p_fmt(t):
    return t.detach().cpu().numpy().ravel()[:4]


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    check_eigh()
    check_cov()
    for p in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
        check_golden(p)
