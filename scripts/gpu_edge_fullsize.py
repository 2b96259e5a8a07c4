"""Who is off at full size: GPU theta vs the float32 oracle (the reference's arithmetic) vs the float64
spectral oracle, on graphs of BASELINE configs[2] with a low threshold (non-trivial support); and what
the eigensolver tolerance / the product back-end contribute.  usage: gpu_edge_fullsize.py [B]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from oracle import uglad_oracle as O
from uglad_b200 import main as ug, ops
from uglad_b200.glad.glad_params import GladParams
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
X = bench.synth(B, 100, 1000, 77)
S = torch.tensor(O.covariance(X), dtype=torch.float32)
P = O.init_params(77)
with torch.no_grad():
    P["rho_l1.4.bias"].fill_(-6.0)
torch.set_num_threads(16)
th32, _ = O.forward_loss(S, P, 15, 0)
th32 = th32.detach().numpy().astype(np.float64)
nb64 = min(B, 16)
th64 = O.spectral_forward_backward(S[:nb64].numpy(), P, L=15)["theta"] if B <= 16 else None
model = GladParams(1.0, 3, 3); model.load_state_dict({k: v.detach() for k, v in P.items()})
off = ~np.eye(100, dtype=bool)[None].repeat(B, 0)
def cmp(a, b, na, nb):
    d = np.abs(a - b)
    o = off[: a.shape[0]]
    viol = (((a != 0) != (b != 0)) & (np.abs(b) > 1e-5)) | ((b == 0) & (np.abs(a) > 1e-5))
    print(f"{na} vs {nb}: rel fro {np.linalg.norm(a-b)/np.linalg.norm(b):.2e}  max abs diag {d[~o].max():.2e}  offdiag {d[o].max():.2e}  "
          f"support mismatches {int(((a != 0) != (b != 0)).sum())}  edge-criterion violations {int(viol.sum())}; worst |b| among violations "
          f"{np.abs(b[viol]).max() if viol.any() else 0:.2e} / |a| {np.abs(a[viol]).max() if viol.any() else 0:.2e}", flush=True)
Sd = S.cuda()
for use_tc, tol in ((1, 10), (1, 10), (1, 7), (1, 5), (1, 3), (0, 10)):
    ops.tune("use_tc", use_tc); ops.tune("eig_tol_1e7", tol)
    ops.reset_warm_start()
    with torch.no_grad():
        thg = ug.glad.glad(Sd, model, L=15).cpu().numpy().astype(np.float64)
        thw = ug.glad.glad(Sd, model, L=15).cpu().numpy().astype(np.float64)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): ug.glad.glad(Sd, model, L=15)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 5 * 1e3
    print(f"--- use_tc={use_tc} eig tol={tol}e-7: warm forward {ms:.2f} ms")
    cmp(thg, th32, "gpu-cold", "cpu32")
    cmp(thw, th32, "gpu-warm", "cpu32")
    if th64 is not None:
        cmp(thg[:nb64], th64, "gpu", "f64")
ops.tune("use_tc", 1); ops.tune("eig_tol_1e7", 0)
if th64 is not None:
    cmp(th32[:nb64], th64, "cpu32", "f64")
print("nnz", int((th32 != 0).sum()), "of", th32.size)
