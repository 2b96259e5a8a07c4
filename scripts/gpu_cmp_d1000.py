"""D=1000 single graph: tcgen05 3xTF32 path vs FP32 SIMT path (theta, loss, grads) + timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from uglad_b200 import main as ug, ops, _lib
from uglad_b200.utils import prepare_data
dev = torch.device("cuda:0")
lib = _lib.load()
D = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
M = 10 * D
rng = np.random.default_rng(0)
Xb, _ = prepare_data.get_data(D, [0.02, 0.02], M, batch_size=1, eig_offset=1.0, rng=rng)
Xb = (Xb - Xb.min(1, keepdims=True)) / (Xb.max(1, keepdims=True) - Xb.min(1, keepdims=True))
S = prepare_data.get_covariance(Xb.astype(np.float32))
res = {}
for use_tc in (0, 1):
    ops.tune("use_tc", use_tc)
    torch.manual_seed(0)
    model, opt = ug.init_uGLAD(lr=0.002)
    th, loss = ug.forward_uGLAD(S, model, L=15)
    loss.backward()
    g = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).cpu().numpy()
    res[use_tc] = (th.detach().cpu().numpy().astype(np.float64), loss.item(), g)
    def step():
        opt.zero_grad()
        th, loss = ug.forward_uGLAD(S, model, L=15)
        loss.backward()
        opt.step()
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): step()
    e1.record(); torch.cuda.synchronize()
    print(f"use_tc={use_tc}: step {e0.elapsed_time(e1)/3:.2f} ms  loss0 {res[use_tc][1]:.5f}", flush=True)
a, b = res[0], res[1]
print("theta rel diff tc vs simt:", np.linalg.norm(a[0] - b[0]) / np.linalg.norm(a[0]))
print("theta asym (tc):", np.linalg.norm(b[0] - b[0].T) / np.linalg.norm(b[0]))
print("loss diff:", a[1] - b[1], "grad rel diff:", np.linalg.norm(a[2] - b[2]) / np.linalg.norm(a[2]))
print("support mismatch:", int(((a[0] != 0) != (b[0] != 0)).sum()), "of", a[0].size)
