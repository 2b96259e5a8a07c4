"""Does staging the next step's samples on a side stream overlap with training? (e2e pipeline experiment)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from uglad_b200 import main as ug, ops
from uglad_b200.utils import prepare_data
wl = sys.argv[1] if len(sys.argv) > 1 else "multitask_d100"
spec = bench.WORKLOADS[wl]
B, D, M = spec["B"], spec["D"], spec["M"]
dev = torch.device("cuda:0")
X_host = torch.from_numpy(bench.synth(B, D, M, 1234)).pin_memory()
torch.manual_seed(0)
model, opt = ug.init_uGLAD(lr=0.002)
def step(S):
    opt.zero_grad()
    _, loss = ug.forward_uGLAD(S, model, L=15)
    loss.backward(); opt.step()
    return loss
side = torch.cuda.Stream(device=dev)
WARM = [None, False]
def stage(record):
    with torch.cuda.stream(side):
        Xd = X_host.to(dev, non_blocking=True)
        S = prepare_data.get_covariance(Xd, warm=WARM[0] if WARM[1] else None)
        WARM[0] = S
        ev = torch.cuda.Event(); ev.record(side)
    return S, ev, Xd
def run(mode, n=12):
    ts = []
    if mode == "seq":
        for i in range(n):
            t0 = time.perf_counter()
            S = prepare_data.get_covariance(X_host.to(dev, non_blocking=True))
            l = float(step(S).item())
            ts.append(time.perf_counter() - t0)
    else:
        nxt = stage(False)
        for i in range(n):
            t0 = time.perf_counter()
            S, ev, Xd = nxt
            torch.cuda.current_stream().wait_event(ev)
            loss = step(S)
            nxt = stage(False)
            l = float(loss.item())
            if mode == "pipe_sync":
                side.synchronize()
            ts.append(time.perf_counter() - t0)
            if ts[-1] > 0.03:
                st = torch.cuda.memory_stats()
                print(f"   spike it={i} {ts[-1]*1e3:.0f} ms: device_alloc {st['num_device_alloc']} device_free {st['num_device_free']} retries {st['num_alloc_retries']} reserved {st['reserved_bytes.all.current']/1e9:.2f} GB", flush=True)
            del S, Xd
    print(f"{wl} {mode}: per-iteration ms {[round(t*1e3,1) for t in ts]}  loss {l:.4f}", flush=True)
import sys
if len(sys.argv) > 2:
    ops.tune("tc_pdl", int(sys.argv[2]))
run("seq", 6)
pf = prepare_data.CovariancePrefetcher(dev)
pf.submit(X_host)
ts = []
for i in range(40):
    t0 = time.perf_counter()
    S = pf.get(); loss = step(S); pf.submit(X_host); l = float(loss.item())
    ts.append(time.perf_counter() - t0)
print("slots:", [round(t*1e3,1) for t in ts], l, "device_alloc", torch.cuda.memory_stats()['num_device_alloc'])
