import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from uglad_b200.utils import prepare_data
dev = torch.device("cuda:0")
X = torch.from_numpy(bench.synth(256, 100, 1000, 1234)).to(dev)
def t(fn, n=5):
    for _ in range(2): r = fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3, r
ms, S0 = t(lambda: prepare_data.get_covariance(X))
print("cold get_covariance ms", ms, "sweeps", S0._uglad_eig[1].info[:, 0].mean().item())
ms, S1 = t(lambda: prepare_data.get_covariance(X, warm=S0))
print("warm get_covariance ms", ms, "sweeps", S1._uglad_eig[1].info[:, 0].mean().item(), "max", S1._uglad_eig[1].info[:, 0].max().item())
print("diff", (S1 - S0).abs().max().item(), (S1._uglad_eig[1].wS.sort(1)[0] - S0._uglad_eig[1].wS.sort(1)[0]).abs().max().item())
