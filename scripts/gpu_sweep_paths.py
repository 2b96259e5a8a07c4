"""Step time of the eigensolver path vs the tcgen05 Newton-Schulz path over (B, D)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from uglad_b200 import main as ug, ops
from uglad_b200.utils import prepare_data
def run(B, D, force):
    rng = np.random.default_rng(0)
    X = rng.random((B, 2 * D, D)).astype(np.float32)
    ops.tune("small_d_max", 0 if force else 232)
    ops.reset_warm_start()
    S = prepare_data.get_covariance(X)
    torch.manual_seed(0)
    model, opt = ug.init_uGLAD(lr=0.002)
    def step():
        opt.zero_grad()
        th, loss = ug.forward_uGLAD(S, model, L=15)
        loss.backward()
        opt.step()
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 3
for B in (1, 8, 32, 148, 256):
    for D in (32, 64, 100, 128, 160, 200, 232):
        a, b = run(B, D, False), run(B, D, True)
        print(f"B={B:4d} D={D:4d}  eig {a:8.2f} ms   tc-ns {b:8.2f} ms   ratio {a/b:5.2f}", flush=True)
ops.tune("small_d_max", 166)
