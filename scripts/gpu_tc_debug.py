"""First contact of the tcgen05 GEMM with real hardware: tiny cases, printed errors."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from uglad_b200 import _lib, ops
lib = _lib.load()
dev = torch.device("cuda:0")
def run(M, N, K, batch, bn=0):
    ops.tune("tc_bn", bn)
    rng = np.random.default_rng(1)
    A = rng.standard_normal((batch, M, K)).astype(np.float32)
    B = rng.standard_normal((batch, N, K)).astype(np.float32)
    ref = np.einsum("bmk,bnk->bmn", A.astype(np.float64), B.astype(np.float64))
    dA, dB = torch.tensor(A, device=dev), torch.tensor(B, device=dev)
    out = torch.full((batch, M, N), float("nan"), device=dev)
    scratch = torch.zeros(lib.uglad_tc_gemm_scratch_floats(M, N, K, batch), device=dev)
    rc = lib.uglad_tc_gemm(dA.data_ptr(), dB.data_ptr(), None, out.data_ptr(), M, N, K, batch, 1.0, 0.0, 0.0,
                           scratch.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc: print("rc", rc, lib.uglad_last_error().decode()); return
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)
    err = np.abs(got - ref)
    print(f"M={M} N={N} K={K} batch={batch} bn={bn}: max err {np.nanmax(err):.3e} nan={np.isnan(got).sum()} ref scale {np.abs(ref).max():.2f}", flush=True)
    if np.nanmax(err) > 1e-3:
        print(" got[0,:4,:4]\n", got[0, :4, :4], "\n ref\n", ref[0, :4, :4])
        # single-TF32 reference to tell '3x dropped' from 'layout wrong'
for a in [(128, 128, 32, 1), (128, 128, 8, 1), (128, 64, 64, 1, 64), (100, 100, 100, 2), (256, 256, 256, 1, 128), (1000, 1000, 1000, 1)]:
    run(*a)
