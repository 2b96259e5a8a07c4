"""Per-stage timeline of the persistent chain launch (developer buffer uglad_tc_debug_buffer): B D"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from uglad_b200 import main as ug, ops, _lib
from uglad_b200.utils import prepare_data
lib = _lib.load(); dev = torch.device("cuda:0")
B, D = int(sys.argv[1]), int(sys.argv[2])
bn = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ops.tune("tc_chain_bn", bn)
S = prepare_data.get_covariance(torch.from_numpy(bench.synth(B, D, 1000 if D <= 200 else 4000, 1234)).to(dev))
torch.manual_seed(0)
model, _ = ug.init_uGLAD(lr=0.002)
with torch.no_grad():
    ug.glad.glad(S, model, L=2)
torch.cuda.synchronize()
dbg = torch.zeros(40 * 148 * 8, dtype=torch.int64, device=dev)
lib.uglad_tc_debug_buffer(dbg.data_ptr())
with torch.no_grad():
    ug.glad.glad(S, model, L=1)       # one layer: one forward chain (19 stages)
torch.cuda.synchronize()
lib.uglad_tc_debug_buffer(None)
d = dbg.cpu().numpy().reshape(40, -1, 8)
# grid size: CTAs with a non-zero arrival stamp in stage 0
G = int((d[0, :, 7] > 0).sum())
d = d.reshape(-1)[: 40 * 148 * 8]
d = dbg.cpu().numpy()[: 40 * G * 8].reshape(40, G, 8)
t0 = d[0, :, 0].min()
print(f"B={B} D={D} bn={bn}: grid {G}; per stage (us, relative to chain start): barrier passed (min/max) | first slab ready (median) | "
      f"last MMA (median, max) | epilogue issued (max) | stores complete (max) | arrived (max)")
prev_end = None
for s in range(19):
    x = d[s].astype(np.float64)
    act = x[:, 2] > 0     # CTAs that had a tile
    us = lambda v: (v - t0) / 1e3
    line = (f"  stage {s:2d} tiles/CTA>0: {int(act.sum()):3d}  barrier {us(x[:,1].min()):8.1f} {us(x[:,1].max()):8.1f} | first slab {us(np.median(x[act,2])):8.1f} | "
            f"last MMA {us(np.median(x[act,3])):8.1f} {us(x[act,3].max()):8.1f} | epi issued {us(x[:,5].max()):8.1f} | stores done {us(x[:,6].max()):8.1f} | arrived {us(x[:,7].max()):8.1f}")
    dur = us(x[:, 7].max()) - (prev_end if prev_end is not None else 0.0)
    prev_end = us(x[:, 7].max())
    print(line + f" | stage time {dur:6.1f}")
