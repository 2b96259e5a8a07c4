#!/usr/bin/env python
"""bench.py -- unrolled-layer-graphs/sec, forward+backward, of the uGLAD hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A step is one training epoch of the reference loop (main.py:389-414): zero_grad, forward of
the L=15 unrolled GLAD layers on every graph, glasso loss, backward, Adam step.  The unit is
one layer of one graph ("layer-graph"): value = graphs * L * steps / time.

Workloads (BASELINE.json configs):
  multitask_d100 (default) configs[2]: 256 graphs, D=100, M=1000, one shared model.  With N
                 GPUs every rank holds 256 graphs (weak scaling); the shards are coupled only
                 by the scalar Frobenius mean per layer and the all-reduce of the 42 MLP
                 gradients per epoch.
  single_d100    configs[1]: one graph, D=100, M=1000 (N=1 only; also reported under "extra"
                 by the default run).
  consensus_d200 configs[3]: 32 imputations at D=200 (tcgen05 Newton-Schulz path).
  single_d1000   configs[4]: one graph, D=1000, M=10000 (tcgen05 Newton-Schulz path; reported
                 under "extra" with its tensor-pipe roofline by the default run).
Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over
ranks.  `value` has the covariances resident in HBM; `e2e` starts from the sample matrices in
pinned host memory every step (H2D copy, covariance, conditioning, fwd+bwd+Adam, loss to
the host; the next step's samples are staged on a side stream while the current step trains).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L_LAYERS = 15
WORKLOADS = {
    "multitask_d100": dict(B=256, D=100, M=1000, config_index=2),
    "single_d100": dict(B=1, D=100, M=1000, config_index=1),
    "consensus_d200": dict(B=32, D=200, M=1000, config_index=3),
    "single_d1000": dict(B=1, D=1000, M=10000, config_index=4),
    "demo_d10": dict(B=1, D=10, M=500, config_index=0),
}


_RESULT_FD = None


def emit(line):
    """The one JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def synth(B, D, M, seed):
    """Synthetic Erdos-Renyi Gaussian-graphical-model samples, min-max normalised like
    uGLAD_GL.fit does (process_table NORM='min_max')."""
    from uglad_b200.utils import prepare_data
    rng = np.random.default_rng(seed)
    p = 0.05 if D <= 200 else 0.01
    Xb, _ = prepare_data.get_data(D, [p, p], M, batch_size=B, eig_offset=1.0, rng=rng)
    Xb = (Xb - Xb.min(1, keepdims=True)) / (Xb.max(1, keepdims=True) - Xb.min(1, keepdims=True))
    return Xb.astype(np.float32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.th.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
def cpu_oracle_rate(B, D, M, seed, steps, warmup, max_graphs):
    """Reference arm / cpu_baseline: the oracle port (torch CPU restatement of the reference,
    oracle/uglad_oracle.py) on the host cores, bounded sample of the same workload."""
    import torch
    from oracle import uglad_oracle as O
    try:  # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it may
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        pass
    nb = min(B, max_graphs)
    X = synth(nb, D, M, seed)
    S = torch.tensor(O.covariance(X), dtype=torch.float32)
    P = O.init_params(seed)
    opt = torch.optim.Adam(list(P.values()), lr=0.002)

    def step():
        opt.zero_grad()
        _, loss = O.forward_loss(S, P, L_LAYERS, 0)
        loss.backward()
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return nb * L_LAYERS * steps / dt, dt / steps * 1e3, nb, torch.get_num_threads()


def run_reference(args, wl, rank, world):
    """--impl reference: the reference's CPU path (the oracle port; the reference itself is pure
    Python/torch and cannot be pip-installed here) on the box's host cores."""
    if rank != 0:
        return
    spec = WORKLOADS[wl]
    big = spec["D"] >= 500
    steps, warmup = (1, 0) if big else (max(1, min(args.steps, 3)), 1)
    rate, ms, nb, cores = cpu_oracle_rate(spec["B"], spec["D"], spec["M"], 1234, steps, warmup, max_graphs=32)
    sample = f"{nb} of {spec['B']} graphs per step, {steps} steps after {warmup} warm-up (oracle port, torch CPU, batched)"
    line = {
        "impl": "reference", "metric": "unrolled-layer-graphs/sec fwd+bwd", "value": rate, "unit": "layer-graphs/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl, "baseline_config_index": spec["config_index"], "D": spec["D"], "M": spec["M"],
                   "L": L_LAYERS, "graphs_per_step": nb},
        "cpu_baseline": {"value": rate, "unit": "layer-graphs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "layer-graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
KERNELS = ((0, "eig_jacobi_small_kernel"), (1, "tc_gemm_kernel"))
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures
# (profiles/r01_ncu_full_*.txt), keyed by (kernel, graphs, D); bytes
NCU_TRAFFIC = {("eig", 256, 100): 30.8e6, ("tc", 1, 1000): 10.7e6}   # tc: r01_ncu_full_tc_gemm_d1000_v3.txt (8-12 MB)


def time_gpu_workload(name, steps, warmup, rank, world, group, dev, profile=False):
    """One workload on this rank's GPU.  `value`: covariances resident in HBM.  `e2e`: every step
    starts from the sample matrices in pinned host memory (H2D, covariance, conditioning,
    fwd + bwd + Adam, loss back to the host).  `prof`: a separate pass with CUDA-event brackets
    around every launch of the two dominant kernels."""
    import ctypes
    import torch
    import torch.distributed as dist
    from uglad_b200 import _lib, main as ug, ops
    from uglad_b200.utils import prepare_data
    spec = WORKLOADS[name]
    B, D, M = spec["B"], spec["D"], spec["M"]
    lib = _lib.load()
    X_host = torch.from_numpy(synth(B, D, M, 1234 + rank)).pin_memory()
    S = prepare_data.get_covariance(X_host.to(dev))
    torch.manual_seed(0)
    model, opt = ug.init_uGLAD(lr=0.002, theta_init_offset=1.0, nF=3, H=3)
    ops.reset_warm_start()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(group=group)
            torch.cuda.synchronize(dev)

    def step(Sb):
        opt.zero_grad()
        _, loss = ug.forward_uGLAD(Sb, model, L=L_LAYERS, INIT_DIAG=0, group=group if world > 1 else None,
                                   total_graphs=B * world)
        loss.backward()
        opt.step()
        return loss

    # e2e: every step consumes a fresh copy of the samples from pinned host memory (H2D,
    # covariance, conditioning) and returns its loss to the host.  As a data loader would, the
    # input pipeline (prepare_data.CovariancePrefetcher) stages the NEXT step's samples on a side
    # stream while the current step trains; one H2D copy and one D2H read per step stay inside the
    # timed region.
    pf = prepare_data.CovariancePrefetcher(dev)
    pf.submit(X_host)

    def e2e_step():
        Sb = pf.get()                                    # this step's covariance (staged during the last step)
        loss = step(Sb)                                  # enqueue fwd + bwd + Adam
        pf.submit(X_host)                                # H2D + covariance + conditioning of the next samples
        return float(loss.item())                        # D2H of the loss

    def timed(fn, n):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX, group=group)
        return float(ms.item())

    for _ in range(warmup):
        step(S)
    c0 = lib.uglad_launch_count()
    with ClockSampler(dev.index) as clk:
        ms_total = timed(lambda: step(S), steps)
    launches = lib.uglad_launch_count() - c0
    for _ in range(min(warmup, 2)):
        e2e_step()
    ms_e2e = timed(e2e_step, steps)
    prof = None
    if profile:
        psteps = max(1, min(steps, 5))
        lib.uglad_profile(1, None, None)
        ms_prof = timed(lambda: step(S), psteps)
        prof = {"steps": psteps, "ms_per_step": ms_prof / psteps}
        for kind, kname in KERNELS:
            k_ms, k_n, k_work = ctypes.c_double(0), ctypes.c_ulonglong(0), ctypes.c_double(0)
            lib.uglad_profile_read(kind, ctypes.byref(k_ms), ctypes.byref(k_n), ctypes.byref(k_work))
            prof[kname] = {"ms": k_ms.value, "launches": int(k_n.value), "work": k_work.value}
        lib.uglad_profile(0, None, None)
    units = B * world * L_LAYERS * steps
    return dict(value=units / ms_total * 1e3, ms_per_step=ms_total / steps, e2e_value=units / ms_e2e * 1e3,
                e2e_ms_per_step=ms_e2e / steps, h2d=int(X_host.numel() * 4), d2h=4, launches=int(launches),
                prof=prof, clocks=clk.summary(), B=B, D=D, M=M)


def roofline_of(r, peaks):
    """Roofline object of the workload's dominant kernel (largest share of the profiled step).
    achieved = algorithmic work of the launches / their CUDA-event durations."""
    prof = r["prof"]
    if not prof:
        return None
    eig, tcg = prof["eig_jacobi_small_kernel"], prof["tc_gemm_kernel"]
    step_ms = prof["ms_per_step"] * prof["steps"]
    if tcg["launches"] and tcg["ms"] >= eig["ms"]:
        key = "bf16_tflops_sustained"
        peak = peaks.get(key)
        src = f"measured (MEASURED_PEAKS.json {key}: cuBLAS dense bf16 inside a long step)"
        if not peak:
            peak, src = 1400.0, "fallback (B200_PROFILING.md sustained bf16)"
        achieved = tcg["work"] / (tcg["ms"] * 1e-3) / 1e12
        return {"bound": "tensor", "kernel": "tc_gemm_kernel (tcgen05.mma kind::tf32, 3xTF32: hi/lo operands, two MMAs per K granule)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": NCU_TRAFFIC.get(("tc", r["B"], r["D"])),
                "peak_source": src, "avg_launch_ms": tcg["ms"] / tcg["launches"],
                "launches_per_step": tcg["launches"] / prof["steps"], "kernel_share_of_step": tcg["ms"] / step_ms,
                "tf32_pipe_frac": 3.0 * achieved / (peak / 2.0),
                "note": "achieved counts the algorithmic 2MNK flops per product; every product costs three "
                        "TF32 passes (hi*hi, hi*lo, lo*hi) and the TF32 pipe peaks at half the bf16 rate, so the "
                        "tensor pipe itself runs at tf32_pipe_frac of its own peak"}
    peak = peaks.get("hbm_gbs")
    src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    if not peak:
        peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = eig["work"] / (eig["ms"] * 1e-3) / 1e9 if eig["launches"] else None
    return {"bound": "hbm", "kernel": "eig_jacobi_small_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": (achieved / peak) if achieved else None, "traffic": NCU_TRAFFIC.get(("eig", r["B"], r["D"])),
            "peak_source": src,
            "avg_launch_ms": (eig["ms"] / eig["launches"]) if eig["launches"] else None,
            "launches_per_step": eig["launches"] / prof["steps"],
            "kernel_share_of_step": eig["ms"] / step_ms,
            "note": "shared-memory-resident Jacobi solver: the binding limit is SM issue/smem latency, not HBM "
                    "(see DESIGN.md)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="multitask_d100", choices=sorted(WORKLOADS))
    ap.add_argument("--no-extra", action="store_true", help="skip the extra per-config measurements")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: everything else that libraries print there
    # (NCCL's "NCCL version ..." banner under torchrun, for one) is sent to stderr
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, args.workload, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; uglad_b200 has no CPU path (use --impl reference for the CPU arm)")
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    wl = args.workload
    r = time_gpu_workload(wl, args.steps, args.warmup, rank, world, group, dev, profile=True)
    spec = WORKLOADS[wl]
    extra = {}
    if world == 1 and not args.no_extra and wl == "multitask_d100":
        for name, st, wu in (("single_d100", args.steps, args.warmup), ("consensus_d200", min(args.steps, 10), 3),
                             ("single_d1000", min(args.steps, 5), 3)):
            x = time_gpu_workload(name, st, wu, rank, world, group, dev, profile=True)
            extra[name] = {"baseline_config_index": WORKLOADS[name]["config_index"], "value": x["value"],
                           "unit": "layer-graphs/s", "ms_per_step": x["ms_per_step"], "e2e_value": x["e2e_value"],
                           "gpu_launches": x["launches"], "roofline": roofline_of(x, peaks)}

    if rank == 0:
        D, B = r["D"], r["B"]
        big = D >= 500
        cpu = None
        if world == 1:  # the CPU baseline is reported at N=1 only
            cpu_rate, cpu_ms, cpu_nb, cores = cpu_oracle_rate(spec["B"], spec["D"], spec["M"], 1234, 1 if big else 2,
                                                              0 if big else 1, max_graphs=32)
            cpu = {"value": cpu_rate, "unit": "layer-graphs/s", "cores": cores, "kind": "port",
                   "sample": f"{cpu_nb} of {spec['B']} graphs per step (oracle port, torch CPU)"}
        line = {
            "metric": "unrolled-layer-graphs/sec fwd+bwd", "value": r["value"], "unit": "layer-graphs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl, "baseline_config_index": spec["config_index"], "graphs_per_gpu": B,
                       "graphs_total": B * world, "D": D, "M": r["M"], "L": L_LAYERS, "H": 3,
                       "parallelism": f"graph-sharded x{world}", "l2_policy": "working set per step exceeds L2 "
                       "(saved theta / theta_k1 / eigenvectors of 15 layers: %.0f MB)" % (B * D * D * 4 * 3 * L_LAYERS / 1e6)},
            "clocks": r["clocks"],
            "e2e": {"value": r["e2e_value"], "unit": "layer-graphs/s", "ms_per_step": r["e2e_ms_per_step"],
                    "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
            "gpu_launches": r["launches"],
            "roofline": roofline_of(r, peaks),
            "cpu_baseline": cpu,
            "extra": extra,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
