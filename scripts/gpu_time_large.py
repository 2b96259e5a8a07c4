"""Timing of one fwd+bwd step of the large-D path (D=1000 single graph; D=100 x 256 forced)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from uglad_b200 import main as ug, ops, _lib
from uglad_b200.utils import prepare_data

dev = torch.device("cuda:0")
lib = _lib.load()

def run(B, D, M, steps=3, force=False):
    rng = np.random.default_rng(0)
    Xb, _ = prepare_data.get_data(D, [0.02, 0.02], M, batch_size=B, eig_offset=1.0, rng=rng)
    Xb = (Xb - Xb.min(1, keepdims=True)) / (Xb.max(1, keepdims=True) - Xb.min(1, keepdims=True))
    if force:
        ops.tune("small_d_max", 0)
    ops.reset_warm_start()
    t0 = time.time()
    S = prepare_data.get_covariance(Xb.astype(np.float32))
    torch.cuda.synchronize()
    print(f"B={B} D={D} M={M} covariance+condition {1e3*(time.time()-t0):.1f} ms")
    torch.manual_seed(0)
    model, opt = ug.init_uGLAD(lr=0.002)
    def step():
        opt.zero_grad()
        th, loss = ug.forward_uGLAD(S, model, L=15)
        loss.backward()
        opt.step()
        return loss
    for _ in range(2):
        l = step()
    torch.cuda.synchronize()
    c0 = lib.uglad_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        l = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"   step {ms:.2f} ms  -> {B*15/ms*1e3:.1f} layer-graphs/s, loss {l.item():.4f}, launches/step {(lib.uglad_launch_count()-c0)//steps}")
    ops.tune("small_d_max", 166)

run(1, 1000, 10000)
run(32, 200, 1000)
run(32, 200, 1000, force=True)
run(256, 100, 1000)
run(256, 100, 1000, force=True)
