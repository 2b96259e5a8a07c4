"""Stage timings on the GPU (CUDA events, warm-up, median of repeats)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uglad_b200 import ops  # noqa: E402
from uglad_b200 import main as ug  # noqa: E402
from uglad_b200.utils import prepare_data  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, warm=3, rep=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(rep):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    rng = np.random.default_rng(0)
    for D, B in [(20, 1), (100, 1), (100, 32), (100, 148), (100, 256), (100, 1024), (200, 1), (200, 32)]:
        A = rng.standard_normal((B, D, D)).astype(np.float32)
        A = torch.tensor((A + A.transpose(0, 2, 1)) / 2, device=dev)
        ms = timeit(lambda: ops.eigh(A, indefinite=True))
        _, _, info = ops.eigh(A, indefinite=True)
        print(f"eigh D={D} B={B}: {ms * 1e3:.0f} us  sweeps {info[:, 0].max().item():.0f}  ({B / ms * 1e3:.0f} eig/s)")
    for D, B, M in [(100, 1, 1000), (100, 256, 1000), (200, 32, 1000)]:
        Xb, _ = prepare_data.get_data(D, [0.05, 0.05], M, batch_size=B, eig_offset=1.0, rng=rng)
        Xb = (Xb - Xb.min(1, keepdims=True)) / (Xb.max(1, keepdims=True) - Xb.min(1, keepdims=True))
        S = prepare_data.get_covariance(Xb)
        torch.manual_seed(0)
        model, opt = ug.init_uGLAD(lr=0.002)

        def step():
            opt.zero_grad()
            th, loss = ug.forward_uGLAD(S, model, L=15)
            loss.backward()
            opt.step()

        def fwd():
            with torch.no_grad():
                ug.forward_uGLAD(S, model, L=15)

        ms = timeit(step, warm=3, rep=8)
        msf = timeit(fwd, warm=2, rep=8)
        print(f"epoch D={D} B={B} L=15: fwd+bwd+adam {ms:.2f} ms (fwd+loss only {msf:.2f} ms) -> {B * 15 / ms * 1e3:.0f} layer-graphs/s")


if __name__ == "__main__":
    main()
