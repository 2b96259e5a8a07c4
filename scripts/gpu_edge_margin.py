"""How close the recovered edge sets come to the 1e-5 criterion on the trained low-threshold goldens:
largest |reference entry| among the entries whose support differs, per knob setting and repeat.
python scripts/gpu_edge_margin.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import uglad_oracle as O
from uglad_b200 import main as ug, ops
from uglad_b200.glad.glad_params import GladParams
dev = torch.device("cuda:0")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def load_model(g, tag):
    m = GladParams(1.0, 3, 3)
    m.load_state_dict({k: torch.tensor(g[f"{tag}/{k}"]) for k in O.PARAM_KEYS})
    return m
def worst(theta, ref):
    diff = (theta != 0) != (ref != 0)
    a = np.abs(ref[diff]).max() if diff.any() else 0.0      # reference non-zero, ours zero
    b = np.abs(theta[diff]).max() if diff.any() else 0.0    # ours non-zero, reference zero
    return int(diff.sum()), float(a), float(b), float(np.abs(theta - ref).max())
for knobs in ({}, {"eig_tol_1e7": 10}, {"eig_raw": 0}, {"eig_raw": 0, "eig_tol_1e7": 10}, {"use_tc": 0}):
    for k, v in knobs.items(): ops.tune(k, v)
    for name in ("d100_lowrho.npz", "d20_lowrho.npz"):
        g = np.load(os.path.join(ROOT, "tests", "golden", name))
        L, idg = int(g["L"]), int(g["init_diag"])
        S = torch.tensor(g["S"], device=dev)
        res = []
        for rep in range(3):
            ops.reset_warm_start()
            model = load_model(g, "p0")
            opt = ug.glad.get_optimizers(model, lr_glad=float(g["lr"]))
            thT, _ = ug._fit_loop(S, model, opt, int(g["epochs"]), L, idg, False)
            res.append(worst(thT.detach().cpu().numpy(), g["thetaT"]))
        print(f"{str(knobs):40s} {name:18s} (mismatches, max|ref| there, max|ours| there, max abs err): {res}", flush=True)
    for k in knobs: ops.tune(k, 1 if k != "eig_tol_1e7" else 0)
