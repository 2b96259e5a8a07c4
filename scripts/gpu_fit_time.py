"""Wall time of uGLAD_GL.fit (direct mode, the reference's most common call) with replayed and with eager epochs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from uglad_b200 import main as ug
for D, epochs in ((100, 250), (20, 250)):
    X = bench.synth(1, D, 1000, 7)[0].astype(np.float64)
    for eager in ("0", "1", "0"):
        os.environ["UGLAD_EAGER_FIT"] = eager
        m = ug.uGLAD_GL()
        torch.manual_seed(0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m.fit(X.copy(), epochs=epochs, lr=0.002, L=15, verbose=False, mode="direct")
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"D={D} epochs={epochs} eager_fit={eager}: {dt*1e3:.1f} ms ({dt*1e3/epochs:.3f} ms/epoch) precision[0,:3]={np.round(m.precision_[0,:3],5)}", flush=True)
