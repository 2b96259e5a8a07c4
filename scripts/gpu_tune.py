"""Sweep eigensolver configurations (lanes per pair, keep-G / warm start) on the GPU."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uglad_b200 import ops, main as ug
from uglad_b200.utils import prepare_data
from scripts.gpu_time import timeit
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)

def eig_sweep():
    for D, B in [(100, 1), (100, 256), (200, 1), (200, 32), (20, 1)]:
        A = rng.standard_normal((B, D, D)).astype(np.float32)
        A = torch.tensor((A + A.transpose(0, 2, 1)) / 2, device=dev)
        for keepg in (0, 1):
            for lp in (4, 8, 16, 32):
                ops.tune("eig_lp", lp); ops.tune("eig_keepg", keepg)
                try:
                    ms = timeit(lambda: ops.eigh(A, indefinite=True), warm=2, rep=5)
                    w, Vt, info = ops.eigh(A, indefinite=True)
                    R = torch.einsum("bki,bk,bkj->bij", Vt.double(), w.double(), Vt.double())
                    err = (torch.linalg.matrix_norm(R - A.double()) / torch.linalg.matrix_norm(A.double())).max().item()
                    print(f"eigh D={D} B={B} keepg={keepg} lp={lp}: {ms*1e3:.0f} us sweeps {info[:,0].max().item():.0f} recon {err:.2e}")
                except Exception as e:
                    print(f"eigh D={D} B={B} keepg={keepg} lp={lp}: FAILED {e}")
    ops.tune("eig_lp", 0); ops.tune("eig_keepg", -1)

def epoch_sweep():
    for D, B, M in [(100, 1, 1000), (100, 256, 1000)]:
        Xb, _ = prepare_data.get_data(D, [0.05, 0.05], M, batch_size=B, eig_offset=1.0, rng=rng)
        Xb = (Xb - Xb.min(1, keepdims=True)) / (Xb.max(1, keepdims=True) - Xb.min(1, keepdims=True))
        S = prepare_data.get_covariance(Xb)
        for warm in (False, True):
            for lp in (4, 8, 16):
                ops.tune("eig_lp", lp); ops.warm_start_enabled = warm; ops.reset_warm_start()
                torch.manual_seed(0)
                model, opt = ug.init_uGLAD(lr=0.002)
                def step():
                    opt.zero_grad()
                    th, loss = ug.forward_uGLAD(S, model, L=15)
                    loss.backward(); opt.step()
                ms = timeit(step, warm=3, rep=8)
                def fwd():
                    with torch.no_grad():
                        ug.glad.glad(S, model, L=15)
                msf = timeit(fwd, warm=2, rep=8)
                print(f"epoch D={D} B={B} warm={warm} lp={lp}: step {ms:.2f} ms (glad fwd only {msf:.2f} ms) -> {B*15/ms*1e3:.0f} layer-graphs/s")
    ops.tune("eig_lp", 0); ops.warm_start_enabled = True

if __name__ == "__main__":
    eig_sweep(); epoch_sweep()
