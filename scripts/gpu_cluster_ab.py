"""A/B of the cluster eigensolver (uglad_tune eig_cluster: 0 one-CTA kernel, 1/2/4 CTAs per graph of the odd-even
kernel): graph-replayed step time, agreement of theta / loss with the one-CTA kernel, sweeps, phase cycles.
usage: gpu_cluster_ab.py [B D]..."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from uglad_b200 import main as ug, ops, _lib
from uglad_b200.utils import prepare_data
lib = _lib.load(); dev = torch.device("cuda:0")
args = [int(x) for x in sys.argv[1:]]
cases = list(zip(args[0::2], args[1::2])) or [(1, 100), (32, 100), (3, 20), (2, 164), (64, 100), (256, 100)]
if os.environ.get("SMALL_D_MAX"): ops.tune("small_d_max", int(os.environ["SMALL_D_MAX"]))
NCS = [int(x) for x in os.environ.get("NCS", "0,1,2,4").split(",")]
for B, D in cases:
    S = prepare_data.get_covariance(torch.from_numpy(bench.synth(B, D, 1000, 1234)).to(dev))
    base = None
    for nc in NCS:
        ops.tune("eig_cluster", nc)
        ops.tune("eig_timing", 0)
        ops.reset_warm_start()
        torch.manual_seed(0)
        model, opt = ug.init_uGLAD(lr=0.002, capturable=True)
        gs = ops.GraphedStep(S, model, opt, L=15)
        for _ in range(4): th, loss = gs.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): th, loss = gs.step()
        e1.record(); torch.cuda.synchronize()
        th = th.detach().clone(); lossv = float(loss)
        # phases: two eager epochs with the timing knob (warm: seeded by the graph's last workspace)
        ops.tune("eig_timing", int(os.environ.get("EIG_TIMING", "1")))
        with ops.use_workspace(gs.ws[0], gs.ws[1]):
            opt.zero_grad(); _, l2 = ug.forward_uGLAD(gs.S, model, L=15)
        torch.cuda.synchronize()
        dims = ops.make_dims(B, D, 15, 3, 0)
        off = lib.uglad_workspace_offset(C.byref(dims), b"info")
        info = gs.ws[0][off:off + 15 * B * 4].view(15, B, 4).cpu().numpy()
        if base is None: base = (th, lossv)
        print(f"B={B} D={D} eig_cluster={nc}: {e0.elapsed_time(e1)/10:.3f} ms/step  loss {lossv:.6f} (d {lossv-base[1]:+.2e})  "
              f"theta rel vs one-CTA {float(torch.linalg.norm(th - base[0]) / torch.linalg.norm(base[0])):.2e}  "
              f"sweeps mean {info[:,:,0].mean():.2f} max {info[:,:,0].max():.0f}  cycles: setup/check {info[:,:,1].mean():.0f} sweeps {info[:,:,2].mean():.0f} tail {info[:,:,3].mean():.0f}", flush=True)
        del gs
ops.tune("eig_cluster", -1); ops.tune("eig_timing", 0)
