"""Phase cycle counts of the Jacobi kernel inside the multitask D=100 step (developer knob eig_timing)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from uglad_b200 import main as ug, ops, _lib
from uglad_b200.utils import prepare_data
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
D = int(sys.argv[2]) if len(sys.argv) > 2 else 100
lib = _lib.load(); dev = torch.device("cuda:0")
S = prepare_data.get_covariance(torch.from_numpy(bench.synth(B, D, 1000, 1234)).to(dev))
torch.manual_seed(0)
model, opt = ug.init_uGLAD(lr=0.002)
ops.tune("eig_timing", 1)
for step in range(6):
    opt.zero_grad()
    th, loss = ug.forward_uGLAD(S, model, L=15)
    loss.backward(); opt.step()
    torch.cuda.synchronize()
    ws = next(reversed(ops._warm.values()))
    dims = ops.make_dims(B, D, 15, 3, 0)
    off = lib.uglad_workspace_offset(C.byref(dims), b"info")
    info = ws[off:off + 15 * B * 4].view(15, B, 4).cpu().numpy()
    print(f"step {step}: sweeps/layer {info[:, :, 0].mean(1).round(2).tolist()}")
    sw = info[:, :, 0].astype(int)
    print(f"   sweep histogram over (layer, graph): {np.bincount(sw.ravel()).tolist()}; max per layer {sw.max(1).tolist()}")
    print(f"   cycles: setup {info[:,:,1].mean():.0f}  sweeps {info[:,:,2].mean():.0f}  tail {info[:,:,3].mean():.0f}; per sweep {info[:,:,2].sum()/info[:,:,0].sum():.0f}", flush=True)
