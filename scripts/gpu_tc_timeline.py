"""In-kernel timeline of one tcgen05 GEMM launch: python scripts/gpu_tc_timeline.py M N K batch [bn]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from uglad_b200 import _lib, ops
lib = _lib.load(); dev = torch.device("cuda:0")
M, N, K, batch = (int(v) for v in sys.argv[1:5])
bn = int(sys.argv[5]) if len(sys.argv) > 5 else 0
ops.tune("tc_bn", bn)
ops.tune("tc_raw", int(os.environ.get("TC_RAW", "1")))
ops.tune("tc_atm", int(os.environ.get("TC_ATM", "1")))
A = torch.randn(batch, M, K, device=dev); B = torch.randn(batch, N, K, device=dev)
out = torch.empty(batch, M, N, device=dev)
scratch = torch.empty(lib.uglad_tc_gemm_scratch_floats(M, N, K, batch), device=dev)
dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run():
    rc = lib.uglad_tc_gemm(A.data_ptr(), B.data_ptr(), None, out.data_ptr(), M, N, K, batch, 1.0, 0.0, 0.0, scratch.data_ptr(), st)
    assert rc == 0, lib.uglad_last_error().decode()
for _ in range(3): run()
ops.tune("tc_exp", int(os.environ.get("TC_EXP", "0")))
torch.cuda.synchronize()
lib.uglad_tc_debug_buffer(dbg.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
run(); torch.cuda.synchronize()
d = dbg.cpu().numpy().reshape(148, 16)
lib.uglad_tc_debug_buffer(None)
used = d[d[:, 7] > 0]
gstart, gend = used[:, 7], used[:, 0]
print(f"wall clock (ns): CTA starts spread {gstart.max() - gstart.min()}, first start -> last end {gend.max() - gstart.min()}, "
      f"per-CTA life median {np.median(gend - gstart):.0f} max {(gend - gstart).max()}")
used = used.copy(); used[:, 0] = used[:, 1] - 400
t0 = used[:, 0].min()
names = ["start", "setup", "depwait", "first_full", "last_mma", "acc_ready", "epi_done"]
rel = (used[:, :7] - t0).astype(np.float64)
print(f"M={M} N={N} K={K} batch={batch} bn={bn}: {len(used)} CTAs; cycles relative to first CTA start (min / median / max)")
for i, n in enumerate(names):
    print(f"  {n:10s} {rel[:, i].min():9.0f} {np.median(rel[:, i]):9.0f} {rel[:, i].max():9.0f}")
seg = np.diff(rel[:, :7], axis=1)
print("  MMA thread, first tile (median cycles summed over the K slabs): wait %d, issue %d, commit %d" % tuple(np.median(used[:, 8 + i]) for i in range(3)))
print("  split warps, first tile (raw-operand kernel): wait raw %d, wait split buffer %d, convert %d" % tuple(np.median(used[:, 11 + i]) for i in range(3)))
print("  epilogue warp 2, first tile: tcgen05.ld + wait %d, math + stores %d" % tuple(np.median(used[:, 14 + i]) for i in range(2)))
print("  per-CTA segments (median cycles):", {names[i + 1]: float(np.median(seg[:, i])) for i in range(6)})
# back-to-back launches of the same product (no other kernels in between)
for so in (0, 1):
    for pdl in (0, 1):
        ops.tune("tc_pdl", pdl)
        lib.uglad_tc_gemm_repeat(A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, batch, 10, so, scratch.data_ptr(), st)
        torch.cuda.synchronize(); e0.record()
        lib.uglad_tc_gemm_repeat(A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, batch, 200, so, scratch.data_ptr(), st)
        e1.record(); torch.cuda.synchronize()
        print(f"  split_out={so} pdl={pdl}: {e0.elapsed_time(e1) / 200 * 1e3:.2f} us per launch (200 back-to-back)")
ops.tune("tc_pdl", 1)
