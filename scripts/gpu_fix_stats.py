"""Convergence branches of the warm solves at the full multitask batch: fix-up list lengths and full-sweep fallbacks."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from uglad_b200 import main as ug, ops, _lib
from uglad_b200.utils import prepare_data
lib = _lib.load(); dev = torch.device("cuda:0")
B, D = int(sys.argv[1]), int(sys.argv[2])
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1234
S = prepare_data.get_covariance(torch.from_numpy(bench.synth(B, D, 1000, seed)).to(dev))
torch.manual_seed(0)
model, opt = ug.init_uGLAD(lr=0.002)
ops.tune("eig_timing", 2)
for step in range(8):
    opt.zero_grad()
    th, loss = ug.forward_uGLAD(S, model, L=15)
    loss.backward(); opt.step()
    torch.cuda.synchronize()
    ws = next(reversed(ops._warm.values()))
    dims = ops.make_dims(B, D, 15, 3, 0)
    off = lib.uglad_workspace_offset(C.byref(dims), b"info")
    info = ws[off:off + 15 * B * 4].view(15, B, 4).cpu().numpy()
    sw = info[:, :, 0].astype(int)
    ml = info[:, :, 3].astype(int)
    print(f"step {step}: sweeps hist {np.bincount(sw.ravel()).tolist()} layers with a 2-sweep graph {int((sw.max(1) > 1).sum())}/15; "
          f"list length percentiles 50/90/99/max {np.percentile(ml, 50):.0f}/{np.percentile(ml, 90):.0f}/{np.percentile(ml, 99):.0f}/{ml.max()}; "
          f"fallback reasons (1 = list > 64, 100 = third check) {np.unique(info[:, :, 2].astype(int), return_counts=True)}", flush=True)
ops.tune("eig_timing", 0)
