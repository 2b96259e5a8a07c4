"""Kernel-boundary cost inside a CUDA graph: 200 dependent launches of a trivial torch kernel, of lambda-size library
kernels, and of the 256 x 100^3 tcgen05 product (whose in-kernel life is 5.9 us, scripts/gpu_tc_timeline.py)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uglad_b200 import ops, _lib
lib = _lib.load(); dev = torch.device("cuda:0")
def timed(fn, n, label):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(3)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn(n)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1) * 1e3 / (10 * n):.2f} us per dependent launch", flush=True)
x = torch.zeros(32, device=dev)
timed(lambda n: [x.add_(1.0) for _ in range(n)], 200, "torch add_ on 32 floats")
big = torch.zeros(256, 100, 100, device=dev)
timed(lambda n: [big.add_(1.0) for _ in range(n)], 200, "torch add_ on 256x100x100 floats (10 MB)")
for B in (1, 256):
    D = 100
    A = torch.randn(B, D, D, device=dev) / D ** 0.5
    X = [torch.randn(B, D, D, device=dev), torch.empty(B, D, D, device=dev)]
    scratch = torch.empty(lib.uglad_tc_gemm_scratch_floats(D, D, D, B), device=dev)
    def chain(n):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for i in range(n):
            lib.uglad_tc_gemm(A.data_ptr(), X[i & 1].data_ptr(), None, X[(i + 1) & 1].data_ptr(), D, D, D, B, C.c_float(1.0), C.c_float(0.0), C.c_float(0.0), scratch.data_ptr(), st)
    timed(chain, 200, f"tcgen05 product {B} x 100^3")
    def chain2(n):   # interleaved with a trivial kernel: does a different neighbour change the gap?
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for i in range(n // 2):
            lib.uglad_tc_gemm(A.data_ptr(), X[i & 1].data_ptr(), None, X[(i + 1) & 1].data_ptr(), D, D, D, B, C.c_float(1.0), C.c_float(0.0), C.c_float(0.0), scratch.data_ptr(), st)
            x.add_(1.0)
    timed(chain2, 200, f"tcgen05 product {B} x 100^3 alternating with a trivial kernel (per launch of either)")
