"""One workload, a few steps: the command profiled by ncu (profiles/).  usage: prof_step.py WORKLOAD [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from uglad_b200 import main as ug, ops
from uglad_b200.utils import prepare_data
wl = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
spec = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
S = prepare_data.get_covariance(torch.from_numpy(bench.synth(spec["B"], spec["D"], spec["M"], 1234)).to(dev))
torch.manual_seed(0)
model, opt = ug.init_uGLAD(lr=0.002)
for i in range(steps):
    opt.zero_grad()
    _, loss = ug.forward_uGLAD(S, model, L=15)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("loss", loss.item())
