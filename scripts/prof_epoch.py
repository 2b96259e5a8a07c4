"""One short training run for profiling: python scripts/prof_epoch.py D B epochs"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uglad_b200 import ops, main as ug
from uglad_b200.utils import prepare_data
D, B, E = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(0)
Xb, _ = prepare_data.get_data(D, [0.05, 0.05], 1000, batch_size=B, eig_offset=1.0, rng=rng)
Xb = (Xb - Xb.min(1, keepdims=True)) / (Xb.max(1, keepdims=True) - Xb.min(1, keepdims=True))
S = prepare_data.get_covariance(Xb)
torch.manual_seed(0)
model, opt = ug.init_uGLAD(lr=0.002)
for e in range(E):
    opt.zero_grad()
    th, loss = ug.forward_uGLAD(S, model, L=15)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
lib = ops._lib.load()
print("loss", loss.item())
