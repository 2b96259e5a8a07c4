"""A/B of any uglad_tune knob on a workload's step time: python scripts/gpu_ab_knob.py key v0 v1 B D"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, ctypes
from uglad_b200 import _lib
lib = _lib.load()
from uglad_b200 import main as ug, ops
from uglad_b200.utils import prepare_data
key, vals, B, D = sys.argv[1], [int(sys.argv[2]), int(sys.argv[3])], int(sys.argv[4]), int(sys.argv[5])
M = 10 * D if D >= 500 else 1000
dev = torch.device("cuda:0")
S = prepare_data.get_covariance(torch.from_numpy(bench.synth(B, D, M, 1234)).to(dev))
for v in vals + vals:
    ops.tune(key, v)
    ops.reset_warm_start()
    torch.manual_seed(0)
    model, opt = ug.init_uGLAD(lr=0.002)
    def step():
        opt.zero_grad()
        _, loss = ug.forward_uGLAD(S, model, L=15)
        loss.backward(); opt.step()
        return loss
    for _ in range(3): l = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): l = step()
    e1.record(); torch.cuda.synchronize()
    msg = f"B={B} D={D} {key}={v}: step {e0.elapsed_time(e1)/5:.3f} ms loss {l.item():.5f}"
    lib.uglad_profile(1, None, None)   # per-kernel CUDA-event brackets (serialises the launches a little)
    for _ in range(2): step()
    torch.cuda.synchronize()
    for kind, name in ((0, "eig"), (1, "tc_gemm")):
        k_ms, k_n, k_w = ctypes.c_double(0), ctypes.c_ulonglong(0), ctypes.c_double(0)
        lib.uglad_profile_read(kind, ctypes.byref(k_ms), ctypes.byref(k_n), ctypes.byref(k_w))
        if k_n.value: msg += f" | {name}: {k_ms.value/2:.2f} ms/step in {k_n.value//2} launches ({k_ms.value/k_n.value*1e3:.1f} us each)"
    lib.uglad_profile(0, None, None)
    print(msg, flush=True)
