"""A few launches of one tcgen05 GEMM shape (the command profiled by ncu): gpu_tc_one.py M N K batch [bn] [raw]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uglad_b200 import _lib, ops
lib = _lib.load(); dev = torch.device("cuda:0")
M, N, K, batch = (int(v) for v in sys.argv[1:5])
ops.tune("tc_bn", int(sys.argv[5]) if len(sys.argv) > 5 else 0)
ops.tune("tc_raw", int(sys.argv[6]) if len(sys.argv) > 6 else 1)
A = torch.randn(batch, M, K, device=dev); B = torch.randn(batch, N, K, device=dev)
out = torch.empty(batch, M, N, device=dev)
scratch = torch.empty(lib.uglad_tc_gemm_scratch_floats(M, N, K, batch), device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(6):
    rc = lib.uglad_tc_gemm(A.data_ptr(), B.data_ptr(), None, out.data_ptr(), M, N, K, batch, 1.0, 0.0, 0.0, scratch.data_ptr(), st)
    assert rc == 0, lib.uglad_last_error().decode()
torch.cuda.synchronize()
print("ok", float(out[0, 0, 0]))
