"""bench.py contract on the CPU: the reference arm (the vendored reference, or the oracle port when
oracle/_ref is absent, timed on the host cores) prints exactly one JSON line on stdout with the keys
the driver reads, and honours --steps / --warmup on the whole workload."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--workload", "demo_d10"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["value"] > 0 and line["unit"] == "layer-graphs/s"
    from oracle import ref_loader
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"] == "demo_d10"
    assert line["steps"] == 2 and line["warmup"] == 1                      # the arm honours the driver's counts
    assert line["config"]["graphs_per_step"] == line["config"]["graphs_total"]   # and runs the whole workload


def test_vendored_reference_equals_the_oracle_port():
    """oracle/_ref (the real reference, vendored by oracle/build_ref.py) and the oracle port give the
    same theta / loss / gradients on a fresh random problem (skipped where oracle/_ref was not built)."""
    import numpy as np
    import pytest
    import torch
    from oracle import ref_loader, uglad_oracle as O
    ref = ref_loader.load()
    if ref is None:
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py needs /root/reference)")
    torch.set_num_threads(1)
    rng = np.random.default_rng(5)
    X = rng.random((3, 70, 9))
    S64 = ref.prepare_data.get_covariance(X, offset=0.1)
    assert np.abs(S64 - O.covariance(X)).max() < 1e-12
    S = torch.tensor(S64, dtype=torch.float32)
    torch.manual_seed(7)
    model = ref.GladParams(theta_init_offset=1.0, nF=3, H=3)
    th_r, loss_r = ref.main.forward_uGLAD(S, model, L=15, INIT_DIAG=0)
    loss_r.backward()
    P = O.init_params(7)
    th_o, loss_o = O.forward_loss(S, P, 15, 0)
    loss_o.backward()
    assert torch.allclose(th_r, th_o, rtol=1e-5, atol=1e-6) and abs(loss_r.item() - loss_o.item()) < 1e-4
    for k, p in model.named_parameters():
        assert torch.allclose(p.grad, P[k].grad, rtol=1e-3, atol=1e-6), k
