"""Step time of the tcgen05 Newton-Schulz path for each tile width: python scripts/gpu_ab_bn.py B D"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from uglad_b200 import main as ug, ops, _lib
from uglad_b200.utils import prepare_data
B, D = int(sys.argv[1]), int(sys.argv[2])
M = 10 * D if D >= 500 else 1000
lib = _lib.load(); dev = torch.device("cuda:0")
S = prepare_data.get_covariance(torch.from_numpy(bench.synth(B, D, M, 1234)).to(dev))
for bn in (0, 64, 112, 128):
    ops.tune("tc_bn", bn)
    torch.manual_seed(0)
    model, opt = ug.init_uGLAD(lr=0.002)
    def step():
        opt.zero_grad()
        _, loss = ug.forward_uGLAD(S, model, L=15)
        loss.backward(); opt.step()
        return loss
    for _ in range(2): l = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4): l = step()
    e1.record(); torch.cuda.synchronize()
    print(f"B={B} D={D} tc_bn={bn}: step {e0.elapsed_time(e1)/4:.2f} ms loss {l.item():.4f}", flush=True)
ops.tune("tc_bn", 0)
