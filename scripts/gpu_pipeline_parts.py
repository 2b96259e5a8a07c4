import sys, os, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from uglad_b200 import _lib, ops
from uglad_b200.utils import prepare_data
lib = _lib.load(); dev = torch.device("cuda:0")
for (B, D, M) in [(256, 100, 1000), (1, 1000, 10000), (32, 200, 1000)]:
    X = torch.from_numpy(bench.synth(B, D, M, 1234)).to(dev)
    S = torch.empty(B, D, D, device=dev); mean = torch.empty(B, D, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
    Xt = torch.empty(lib.uglad_covariance_scratch_floats(B, M, D), device=dev)
    def cov():   # tcgen05 path (uglad_tune("use_tc", 0) selects the FP32 SIMT contraction)
        ops.check(lib.uglad_covariance_ws(P(X), B, M, D, P(S), P(mean), P(Xt), st), "cov")
    for _ in range(3): cov()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): cov()
    e1.record(); torch.cuda.synchronize()
    t_cov = e0.elapsed_time(e1) / 10
    pf = prepare_data.CovariancePrefetcher(device=dev)
    Xh = X.cpu().pin_memory()
    pf.submit(Xh); pf.get(); pf.submit(Xh); pf.get()
    torch.cuda.synchronize()
    e0.record(pf.stream)
    for _ in range(10):
        pf.submit(Xh); pf.get()
    e1.record(pf.stream); torch.cuda.synchronize()
    print(f"B={B} D={D} M={M}: covariance {t_cov*1e3:.0f} us; full pipeline (H2D {X.numel()*4/1e6:.0f} MB + cov + conditioning, alone) {e0.elapsed_time(e1)/10*1e3:.0f} us/step", flush=True)
