"""What programmatic dependent launch buys between dependent small kernels: 100 dependent 100^3 products replayed from
a CUDA graph with tc_pdl on / off, and whole epochs at B = 1 / 32 / 256."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from uglad_b200 import ops, _lib, main as ug
from uglad_b200.utils import prepare_data
lib = _lib.load(); dev = torch.device("cuda:0")
for B, D in ((1, 100), (256, 100)):
    A = torch.randn(B, D, D, device=dev) / D ** 0.5
    X = [torch.randn(B, D, D, device=dev), torch.empty(B, D, D, device=dev)]
    scratch = torch.empty(lib.uglad_tc_gemm_scratch_floats(D, D, D, B), device=dev)
    for pdl in (1, 0):
        ops.tune("tc_pdl", pdl)
        s = torch.cuda.Stream()
        def chain(n):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            for i in range(n):
                lib.uglad_tc_gemm(A.data_ptr(), X[i & 1].data_ptr(), None, X[(i + 1) & 1].data_ptr(), D, D, D, B, C.c_float(1.0), C.c_float(0.0), C.c_float(0.0), scratch.data_ptr(), st)
        with torch.cuda.stream(s):
            chain(4)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            chain(100)
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): g.replay()
        e1.record(); torch.cuda.synchronize()
        print(f"B={B} D={D} tc_pdl={pdl}: {e0.elapsed_time(e1) * 1e3 / 1000:.2f} us per dependent product", flush=True)
for B in (1, 32, 256):
    S = prepare_data.get_covariance(torch.from_numpy(bench.synth(B, 100, 1000, 1234)).to(dev))
    for pdl in (1, 0):
        ops.tune("tc_pdl", pdl)
        ops.reset_warm_start()
        torch.manual_seed(0)
        model, opt = ug.init_uGLAD(lr=0.002, capturable=True)
        gs = ops.GraphedStep(S, model, opt, L=15)
        for _ in range(4): gs.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): gs.step()
        e1.record(); torch.cuda.synchronize()
        print(f"epoch B={B} D=100 tc_pdl={pdl}: {e0.elapsed_time(e1)/10:.3f} ms", flush=True)
        del gs
ops.tune("tc_pdl", 1)
