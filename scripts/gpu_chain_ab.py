"""A/B of the persistent chain launch (uglad_tune tc_chain) on the large-D workloads: step time and agreement."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from uglad_b200 import main as ug, ops
from uglad_b200.utils import prepare_data
dev = torch.device("cuda:0")
for name, B, D, M in (("32xD200", 32, 200, 1000), ("D1000", 1, 1000, 10000), ("4xD200", 4, 200, 1000)):
    S = prepare_data.get_covariance(torch.from_numpy(bench.synth(B, D, M, 1234)).to(dev))
    res = {}
    for cfg in (("tc_chain", 0, 0), ("tc_chain", 1, 0), ("tc_chain", 1, 64), ("tc_chain", 1, 128)):
        ops.tune("tc_chain", cfg[1]); ops.tune("tc_chain_bn", cfg[2])
        torch.manual_seed(0)
        model, opt = ug.init_uGLAD(lr=0.002)
        def step():
            opt.zero_grad()
            th, loss = ug.forward_uGLAD(S, model, L=15)
            loss.backward(); opt.step()
            return th, loss
        th, loss = step()
        g = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
        for _ in range(2): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): step()
        e1.record(); torch.cuda.synchronize()
        res[cfg] = (th.detach().clone(), g, float(loss))
        base = res[("tc_chain", 0, 0)]
        print(f"{name} chain={cfg[1]} bn={cfg[2]}: {e0.elapsed_time(e1)/5:.2f} ms/step  loss {float(loss):.6f}  "
              f"theta rel vs launches {float(torch.linalg.norm(th - base[0]) / torch.linalg.norm(base[0])):.2e}  "
              f"grad rel {float(torch.linalg.norm(g - base[1]) / torch.linalg.norm(base[1])):.2e}", flush=True)
ops.tune("tc_chain", 1); ops.tune("tc_chain_bn", 0)
