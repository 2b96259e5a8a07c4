"""A/B of eigensolver knobs on the multitask D=100 step: python scripts/gpu_ab_eig.py key v0 v1 [B D]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from uglad_b200 import main as ug, ops, _lib
from uglad_b200.utils import prepare_data
key, vals = sys.argv[1], [int(v) for v in sys.argv[2:4]]
B = int(sys.argv[4]) if len(sys.argv) > 4 else 256
D = int(sys.argv[5]) if len(sys.argv) > 5 else 100
lib = _lib.load()
dev = torch.device("cuda:0")
S = prepare_data.get_covariance(torch.from_numpy(bench.synth(B, D, 1000, 1234)).to(dev))
for v in vals + vals:
    ops.tune(key, v)
    ops.reset_warm_start()
    torch.manual_seed(0)
    model, opt = ug.init_uGLAD(lr=0.002)
    def step():
        opt.zero_grad()
        _, loss = ug.forward_uGLAD(S, model, L=15)
        loss.backward()
        opt.step()
        return loss
    for _ in range(5): l = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): l = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    lib.uglad_profile(1, None, None)
    for _ in range(3): step()
    torch.cuda.synchronize()
    k_ms, k_n = ctypes.c_double(0), ctypes.c_ulonglong(0)
    lib.uglad_profile_read(0, ctypes.byref(k_ms), ctypes.byref(k_n), None)
    lib.uglad_profile(0, None, None)
    print(f"{key}={v}: step {ms:.3f} ms, eig avg {k_ms.value/max(1,k_n.value)*1e3:.1f} us over {k_n.value} launches, loss {l.item():.5f}", flush=True)
