"""Parity margins against the reference goldens for the product back-ends: relative Frobenius error of
theta after one forward and after the golden training loop, worst gradient error, worst loss deviation.
python scripts/gpu_parity_margins.py"""
import sys, os, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import uglad_oracle as O
from uglad_b200 import main as ug, ops
from uglad_b200.glad.glad_params import GladParams
dev = torch.device("cuda:0")
for kv in sys.argv[1:]:   # extra knobs, e.g. eig_tol_1e7=40
    k, v = kv.split("="); ops.tune(k, int(v)); print("knob", k, v)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
def load_model(g, tag):
    m = GladParams(1.0, 3, 3)
    m.load_state_dict({k: torch.tensor(g[f"{tag}/{k}"]) for k in O.PARAM_KEYS})
    return m
cases = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))) + sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "large", "*.npz")))
for knobs in ({"eig_raw": 1, "tc_raw": 1}, {"eig_raw": 0, "tc_raw": 0}, {"use_tc": 0})[:(1 if len(sys.argv) > 1 else 3)]:
    for k, v in knobs.items(): ops.tune(k, v)
    for path in cases:
        g = np.load(path)
        if "theta0" not in g.files: continue
        L, idg = int(g["L"]), int(g["init_diag"])
        S = torch.tensor(g["S"], device=dev)
        ops.reset_warm_start()
        model = load_model(g, "p0")
        theta, loss = ug.forward_uGLAD(S, model, L=L, INIT_DIAG=idg)
        loss.backward()
        e_th = rel(theta.detach().cpu().numpy(), g["theta0"])
        e_g = max(rel(p.grad.cpu().numpy(), g["g0/" + k]) for k, p in model.named_parameters())
        msg = f"{str(knobs):34s} {os.path.basename(path):22s} theta {e_th:.1e} grad {e_g:.1e} loss {abs(loss.item() - float(g['loss0'])):.1e}"
        if "thetaT" in g.files and "losses" in g.files:
            ops.reset_warm_start()
            model = load_model(g, "p0")
            opt = ug.glad.get_optimizers(model, lr_glad=float(g["lr"]))
            thT, losses = ug._fit_loop(S, model, opt, int(g["epochs"]), L, idg, False)
            losses = torch.stack(losses).cpu().numpy()
            msg += f" | trained: theta {rel(thT.detach().cpu().numpy(), g['thetaT']):.1e} losses {np.abs(losses - g['losses']).max():.1e}"
        print(msg, flush=True)
    for k in knobs: ops.tune(k, 1)
