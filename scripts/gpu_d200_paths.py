"""32 x D=200 (configs[3] shape): the Newton-Schulz GEMM chain (small_d_max 166) against the eigensolver path with the
cluster kernel (small_d_max 200), graph-replayed step time and agreement of theta / loss."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from uglad_b200 import main as ug, ops
from uglad_b200.utils import prepare_data
dev = torch.device("cuda:0")
B, D = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32, 200)
base = None
for sdm, nc in ((166, -1), (200, 4), (200, 2), (200, 0)):
    ops.tune("small_d_max", sdm); ops.tune("eig_cluster", nc)
    S = prepare_data.get_covariance(torch.from_numpy(bench.synth(B, D, 1000, 1234)).to(dev))
    ops.reset_warm_start()
    torch.manual_seed(0)
    model, opt = ug.init_uGLAD(lr=0.002, capturable=True)
    gs = ops.GraphedStep(S, model, opt, L=15)
    for _ in range(3): th, loss = gs.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): th, loss = gs.step()
    e1.record(); torch.cuda.synchronize()
    th = th.detach().clone(); lv = float(loss)
    if base is None: base = (th, lv)
    print(f"B={B} D={D} small_d_max={sdm} eig_cluster={nc}: {e0.elapsed_time(e1)/10:.3f} ms/step loss {lv:.6f} (d {lv-base[1]:+.2e}) "
          f"theta rel vs NS chain {float(torch.linalg.norm(th-base[0])/torch.linalg.norm(base[0])):.2e}", flush=True)
    del gs
ops.tune("small_d_max", 166); ops.tune("eig_cluster", -1)
