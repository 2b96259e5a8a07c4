"""tcgen05 GEMM: plain operands split in shared memory (tc_raw=1) vs pre-split pairs (tc_raw=0):
back-to-back launch time per product.  python scripts/gpu_ab_raw.py"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from uglad_b200 import _lib, ops
lib = _lib.load(); dev = torch.device("cuda:0")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for (M, N, K, batch) in [(1000, 1000, 1000, 1), (200, 200, 200, 32), (100, 100, 100, 256), (512, 512, 512, 4)]:
    A = torch.randn(batch, M, K, device=dev); B = torch.randn(batch, N, K, device=dev)
    out = torch.empty(batch, M, N, device=dev)
    scratch = torch.empty(lib.uglad_tc_gemm_scratch_floats(M, N, K, batch), device=dev)
    ref = torch.einsum("bmk,bnk->bmn", A.double(), B.double())
    for raw in (0, 1):
        ops.tune("tc_raw", raw)
        rc = lib.uglad_tc_gemm(A.data_ptr(), B.data_ptr(), None, out.data_ptr(), M, N, K, batch, 1.0, 0.0, 0.0, scratch.data_ptr(), st)
        assert rc == 0, lib.uglad_last_error().decode()
        err = (out.double() - ref).abs().max().item()
        for so in (0, 1):
            if raw and so: continue
            lib.uglad_tc_gemm_repeat(A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, batch, 10, so, scratch.data_ptr(), st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.uglad_tc_gemm_repeat(A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, batch, 200, so, scratch.data_ptr(), st)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 200 * 1e3
            print(f"{batch}x{M}x{N}x{K} raw={raw} split_out={so}: {us:.2f} us/launch, {2.0*M*N*K*batch/us*1e-6:.1f} TFLOP/s algorithmic, max err {err:.2e}", flush=True)
ops.tune("tc_raw", 1)
