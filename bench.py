#!/usr/bin/env python
"""bench.py -- unrolled-layer-graphs/sec, forward+backward, of the uGLAD hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--scaling weak|strong]
                    [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A step is one training epoch of the reference loop (main.py:389-414): zero_grad, forward of
the L=15 unrolled GLAD layers on every graph, glasso loss, backward, Adam step.  The unit is
one layer of one graph ("layer-graph"): value = graphs * L * steps / time.

Workloads (BASELINE.json configs):
  multitask_d100 (default) configs[2]: 256 graphs, D=100, M=1000, one shared model.  With N
                 GPUs every rank holds 256 graphs ("weak", the default) or 256 / N ("strong",
                 configs[2] read literally; also reported under "extra" by every N > 1 run).
  single_d100    configs[1]: one graph, D=100, M=1000 (N=1; reported under "extra").
  consensus_d200 configs[3]: the consensus pipeline of main.py:553-644 at D=200: one sample matrix
                 with 20 % missing values, mean imputation, K=32 row-subsampled covariances trained
                 against the full-data covariance (sharded by imputation over the ranks; extra).
  single_d1000   configs[4]: one graph, D=1000, M=10000 (tcgen05 Newton-Schulz path; extra, N=1).
Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over
ranks.  `value` has the covariances resident in HBM; `e2e` starts from the sample matrices in
pinned host memory every step (H2D copy, covariance, conditioning, fwd+bwd+Adam, loss to
the host; the next step's samples are staged on a side stream while the current step trains).

--impl reference: the REAL reference (oracle/_ref, vendored by oracle/build_ref.py; the oracle port
when that is absent) on the host cores, same workload, same --steps / --warmup.
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
    "consensus_d200": dict(B=32, D=200, M=1000, config_index=3, consensus=True, dropout=0.2),
    "single_d1000": dict(B=1, D=1000, M=10000, config_index=4),
    "demo_d10": dict(B=1, D=10, M=500, config_index=0),
}
METRIC = "unrolled-layer-graphs/sec fwd+bwd"
UNIT = "layer-graphs/s"

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


def synth_consensus(D, M, dropout, seed):
    """configs[3]: ONE sample matrix with `dropout` of its entries missing, mean-imputed
    (main.py:595-597).  Returns the imputed [M, D] float32 matrix."""
    X = synth(1, D, M, seed)[0].astype(np.float64)
    rng = np.random.default_rng(seed + 1)
    flat = X.reshape(-1)
    flat[rng.choice(flat.size, size=int(flat.size * dropout), replace=False)] = np.nan
    col_mean = np.nanmean(X, axis=0)
    r, c = np.where(np.isnan(X))
    X[r, c] = col_mean[c]
    return X.astype(np.float32)


def kfold_train(M, K):
    sizes = np.full(K, M // K)
    sizes[: M % K] += 1
    stops = np.cumsum(sizes)
    rows = np.arange(M)
    return [np.concatenate([rows[: e - s], rows[e:]]) for s, e in zip(sizes, stops)]


class ClockSampler:
    """SM clocks / throttle reasons sampled while the timed region runs: NVML in a thread (one query per 50 ms;
    eight ranks each forking nvidia-smi loops starved one another and returned no samples), nvidia-smi as the
    fallback when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.sm, self.mx, self.reasons, self.stop, self.th, self.nvml = [], [], set(), False, None, None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        self.nvml = pynvml
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            return pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            return pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def _poll(self, h):
        n = self.nvml
        while not self.stop:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)))
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for name, bit in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        try:
            h = self._nvml_handle()
            self.th = threading.Thread(target=self._poll, args=(h,), daemon=True)
            self.th.start()
            return self
        except Exception:
            self.nvml = None
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
        if self.nvml is not None:
            time.sleep(0.06)
            self.stop = True
            self.th.join(timeout=2)
        elif self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.th.join(timeout=2)

    def summary(self):
        sm, mx, reasons = list(self.sm), list(self.mx), set(self.reasons)
        names = [n for n, _ in self.REASONS]
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
# CPU arm: the reference itself on the host cores
# ------------------------------------------------------------------------------------------
def cpu_reference_rate(name, steps, warmup, max_graphs=None):
    """One workload through the reference's own CPU code path (oracle/_ref: uglad.main.forward_uGLAD
    + Adam, exactly the loop of main.py:389-414 / :616-630), or through the oracle port when the
    vendored copy is absent.  `max_graphs` bounds the sample (None: the whole workload)."""
    import torch
    from oracle import ref_loader
    spec = WORKLOADS[name]
    B, D, M = spec["B"], spec["D"], spec["M"]
    nb = B if max_graphs is None else min(B, max_graphs)
    threads_note = ""
    try:  # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it may
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        pass
    if nb > 1 and D > 128:
        # torch 2.11's batched CPU LU (the reference's torch.inverse / torch.logdet on [B, D, D]) is broken in
        # this image on more than one thread for D > 128 (oneMKL SLASWP parameter error, then garbage or a hang:
        # profiles/r02_torch_cpu_batched_lu_bug.txt).  The unmodified reference can only run this shape on 1 thread.
        torch.set_num_threads(1)
        threads_note = "; 1 thread: torch's batched CPU LU is broken multi-threaded for D > 128 in this image"
    ref = ref_loader.load()
    kind = "reference" if ref is not None else "port"
    if ref is not None:
        cov = lambda Xs: ref.prepare_data.convert_to_torch(ref.prepare_data.get_covariance(Xs, offset=0.1), req_grad=False)
    else:
        from oracle import uglad_oracle as O
        cov = lambda Xs: torch.tensor(O.covariance(Xs), dtype=torch.float32)
    loss_S = None
    if spec.get("consensus"):
        X = synth_consensus(D, M, spec["dropout"], 1234).astype(np.float64)
        folds = kfold_train(M, B)[:nb]
        S = cov([X[tr] for tr in folds])
        loss_S = cov(X[None])
    else:
        S = cov(synth(nb, D, M, 1234).astype(np.float64))
    torch.manual_seed(0)
    if ref is not None:
        import contextlib
        import io
        import warnings
        model, opt = ref.main.init_uGLAD(lr=0.002, theta_init_offset=1.0, nF=3, H=3)

        def step():
            opt.zero_grad()
            with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
                warnings.simplefilter("ignore")
                _, loss = ref.main.forward_uGLAD(S, model, L=L_LAYERS, INIT_DIAG=0, loss_Sb=loss_S)
            loss.backward()
            opt.step()
            return loss
    else:
        P = O.init_params(0)
        opt = torch.optim.Adam(list(P.values()), lr=0.002)

        def step():
            opt.zero_grad()
            _, loss = O.forward_loss(S, P, L_LAYERS, 0, loss_S=loss_S)
            loss.backward()
            opt.step()
            return loss

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    what = ("the reference itself (oracle/_ref: uglad.main.forward_uGLAD + Adam, per-matrix Newton-Schulz loop)"
            if kind == "reference" else "oracle port (torch CPU restatement, batched)")
    sample = f"{nb} of {B} graphs per step, {steps} timed steps after {warmup} warm-up; {what}{threads_note}"
    return dict(value=nb * L_LAYERS * steps / dt, ms_per_step=dt / steps * 1e3, graphs=nb,
                cores=torch.get_num_threads(), kind=kind, sample=sample)


def cpu_baseline_object(name):
    """Bounded sample for the `cpu_baseline` key of our own arm: ~10-30 s of CPU work in total over
    the four workloads."""
    plan = {"multitask_d100": (2, 1, 32), "single_d100": (4, 1, None), "consensus_d200": (1, 1, 4),
            "single_d1000": (1, 0, None), "demo_d10": (5, 1, None)}
    steps, warmup, mg = plan[name]
    r = cpu_reference_rate(name, steps, warmup, mg)
    return {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
            "ms_per_step": r["ms_per_step"]}


def run_reference(args, wl, rank, world):
    """--impl reference: the reference's CPU path on the box's host cores, the same workload and the
    same --steps / --warmup as our arm.  At N > 1 the workload of ONE rank is timed (the reference
    is a single process; its rate does not depend on how many graphs follow)."""
    if rank != 0:
        return
    spec = WORKLOADS[wl]
    r = cpu_reference_rate(wl, args.steps, args.warmup, None)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl, "baseline_config_index": spec["config_index"], "graphs_per_gpu": spec["B"],
                   "graphs_total": spec["B"], "D": spec["D"], "M": spec["M"], "L": L_LAYERS, "H": 3,
                   "graphs_per_step": r["graphs"]},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
KERNELS = ((0, "eig_jacobi_oe_kernel"), (1, "tc_gemm_kernel"))   # kind 0: the eigensolver launches (warm: eig_cluster.cu)
# dram__bytes_read.sum + dram__bytes_write.sum per launch, read from the committed `ncu --set full`
# captures under profiles/ (the file is named in the roofline object): keyed by (kernel, graphs, D)
NCU_TRAFFIC = {
    ("eig", 256, 100): (10.35e6, "profiles/r02_ncu_full_eig_oe_multitask_d100.txt"),
    ("eig", 1, 100): (0.0965e6, "profiles/r02_ncu_full_eig_oe_single_d100.txt"),
    ("tc", 1, 1000): (10.7e6, "profiles/r01_ncu_full_tc_gemm_d1000_v4.txt"),
}


def shard_range(total, rank, world):
    idx = np.array_split(np.arange(total), world)[rank]
    return int(idx[0]), int(idx[-1]) + 1


def time_gpu_workload(name, steps, warmup, rank, world, group, dev, scaling="weak", profile=False, use_graph=True):
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
    D, M = spec["D"], spec["M"]
    consensus = bool(spec.get("consensus"))
    lib = _lib.load()
    grp = group if world > 1 else None
    if consensus:
        # one sample matrix, K imputations sharded over the ranks; every rank stages the matrix
        B_total = spec["B"]
        lo, hi = shard_range(B_total, rank, world)
        B = hi - lo
        X_host = torch.from_numpy(synth_consensus(D, M, spec["dropout"], 1234)).pin_memory()
        folds = [torch.as_tensor(f, device=dev) for f in kfold_train(M, B_total)[lo:hi]]

        def build(Xd, warm=None):
            S_K = prepare_data.get_covariance([Xd[f] for f in folds])
            Sb = prepare_data.get_covariance(Xd.unsqueeze(0))
            return S_K, Sb
        S, loss_S = build(X_host.to(dev))
    else:
        if scaling == "strong":
            B_total = spec["B"]
            lo, hi = shard_range(B_total, rank, world)
            X_all = synth(B_total, D, M, 1234)          # the same 256 graphs whatever N is
            X_host = torch.from_numpy(np.ascontiguousarray(X_all[lo:hi])).pin_memory()
        else:
            X_host = torch.from_numpy(synth(spec["B"], D, M, 1234 + rank)).pin_memory()
            B_total = spec["B"] * world
        B = X_host.shape[0]
        S, loss_S = prepare_data.get_covariance(X_host.to(dev)), None
    torch.manual_seed(0)
    model, opt = ug.init_uGLAD(lr=0.002, theta_init_offset=1.0, nF=3, H=3, capturable=True)
    ops.reset_warm_start()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(group=group)
            torch.cuda.synchronize(dev)

    def step_eager(Sb, lS=None):
        opt.zero_grad()
        _, loss = ug.forward_uGLAD(Sb, model, L=L_LAYERS, INIT_DIAG=0, loss_Sb=lS, group=grp, total_graphs=B_total)
        loss.backward()
        opt.step()
        return loss

    # The epoch is replayed from CUDA graphs (ops.GraphedStep: two alternating captures, each seeded by the other's
    # workspace); the eager path stays for the per-kernel profiling pass and as the fallback if capture fails.
    gs, graph_note = None, None
    if use_graph:
        try:
            gs = ops.GraphedStep(S, model, opt, L=L_LAYERS, INIT_DIAG=0, loss_S=loss_S, group=grp, total_graphs=B_total)
        except Exception as exc:   # e.g. a collective that cannot be captured on this stack
            graph_note = f"capture failed, eager launches: {type(exc).__name__}: {str(exc)[:120]}"
            torch.cuda.synchronize(dev)
            ops.reset_warm_start()
    replayed = [0]

    def step(Sb, lS=None):
        if gs is None:
            return step_eager(Sb, lS)
        if Sb is not S:
            gs.update_inputs(Sb, lS)
        replayed[0] += gs.kernels_per_graph[gs.calls & 1]
        return gs.step()[1]

    # e2e: every step consumes a fresh copy of the samples from pinned host memory (H2D,
    # covariance, conditioning) and returns its loss to the host.  As a data loader would, the
    # input pipeline (prepare_data.CovariancePrefetcher) stages the NEXT step's samples on a side
    # stream while the current step trains; one H2D copy and one D2H read per step stay inside the
    # timed region.
    if consensus:
        side = torch.cuda.Stream(device=dev)
        X_dev = [torch.empty_like(X_host, device=dev) for _ in range(2)]
        side.wait_stream(torch.cuda.current_stream(dev))
        staged = {}

        def stage(i):
            # runs on the side stream, concurrently with the step enqueued on the main stream; the tensors of
            # the previous generation are released only after that step's loss.item() has synchronised
            with torch.cuda.stream(side):
                X_dev[i].copy_(X_host, non_blocking=True)
                staged["S"] = build(X_dev[i])
                staged["ev"] = torch.cuda.Event()
                staged["ev"].record(side)
        stage(0)
        flip = [1]

        def e2e_step():
            torch.cuda.current_stream(dev).wait_event(staged["ev"])
            S_K, Sb = staged["S"]
            loss = step(S_K, Sb)
            stage(flip[0])
            flip[0] ^= 1
            return float(loss.item())
    else:
        pf = prepare_data.CovariancePrefetcher(dev)
        pf.submit(X_host)

        def e2e_step():
            Sb = pf.get()                                    # this step's covariance (staged during the last step)
            loss = step(Sb)                                  # enqueue fwd + bwd + Adam
            pf.submit(X_host)                                # H2D + covariance + conditioning of the next samples
            return float(loss.item())                        # D2H of the loss

    def timed(fn, n, stream=None):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream if stream is not None else torch.cuda.current_stream(dev)):
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
        step(S, loss_S)
    c0, r0 = lib.uglad_launch_count(), replayed[0]
    with ClockSampler(dev.index) as clk:
        ms_total = timed(lambda: step(S, loss_S), steps)
    launches = (lib.uglad_launch_count() - c0) + (replayed[0] - r0)   # eager launches + kernels replayed from the graphs
    # e2e: the training step runs on a high-priority stream, so that the staging of the NEXT batch (H2D,
    # covariance, conditioning on the prefetcher's ordinary-priority side stream) fills idle SM slots instead
    # of delaying the kernels on the step's critical path
    hp = torch.cuda.Stream(device=dev, priority=-1)
    hp.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(hp):
        for _ in range(min(warmup, 2)):
            e2e_step()
    ms_e2e = timed(e2e_step, steps, stream=hp)
    torch.cuda.current_stream(dev).wait_stream(hp)
    prof = None
    if profile:
        psteps = max(1, min(steps, 5))
        for _ in range(2):   # the eager path keeps its own warm-start chain: seed it before the brackets go on
            step_eager(S, loss_S)
        lib.uglad_profile(1, None, None)
        ms_prof = timed(lambda: step_eager(S, loss_S), psteps)
        prof = {"steps": psteps, "ms_per_step": ms_prof / psteps}
        for kind, kname in KERNELS:
            k_ms, k_n, k_work = ctypes.c_double(0), ctypes.c_ulonglong(0), ctypes.c_double(0)
            lib.uglad_profile_read(kind, ctypes.byref(k_ms), ctypes.byref(k_n), ctypes.byref(k_work))
            prof[kname] = {"ms": k_ms.value, "launches": int(k_n.value), "work": k_work.value}
        k_ms, k_n, k_work = ctypes.c_double(0), ctypes.c_ulonglong(0), ctypes.c_double(0)
        lib.uglad_profile_read(2, ctypes.byref(k_ms), ctypes.byref(k_n), ctypes.byref(k_work))
        prof["eig_fp32_flops"] = k_work.value
        lib.uglad_profile(0, None, None)
    units = B_total * L_LAYERS * steps
    del S, loss_S
    ops.reset_warm_start()
    if gs is not None:
        gs.close()   # the captured graphs (with the captured gradient all-reduce) go now, not when the collector runs
    del gs
    return dict(cuda_graph=(graph_note or ("replayed" if use_graph else "off")),
                value=units / ms_total * 1e3, ms_per_step=ms_total / steps, e2e_value=units / ms_e2e * 1e3,
                e2e_ms_per_step=ms_e2e / steps, h2d=int(X_host.numel() * 4), d2h=4, launches=int(launches),
                prof=prof, clocks=clk.summary(), B=B, B_total=B_total, D=D, M=M)


def roofline_of(r, peaks):
    """Roofline object of the workload's dominant kernel (largest share of the profiled step).
    achieved = algorithmic work of the launches / their CUDA-event durations."""
    prof = r["prof"]
    if not prof:
        return None
    eig, tcg = prof["eig_jacobi_oe_kernel"], prof["tc_gemm_kernel"]
    step_ms = prof["ms_per_step"] * prof["steps"]
    if tcg["launches"] and tcg["ms"] >= eig["ms"]:
        key = "bf16_tflops_sustained"
        peak = peaks.get(key)
        src = f"measured (MEASURED_PEAKS.json {key}: cuBLAS dense bf16 inside a long step)"
        if not peak:
            peak, src = 1400.0, "fallback (B200_PROFILING.md sustained bf16)"
        achieved = tcg["work"] / (tcg["ms"] * 1e-3) / 1e12
        traffic = NCU_TRAFFIC.get(("tc", r["B"], r["D"]), (None, None))
        return {"bound": "tensor", "kernel": "tc_gemm_kernel (tcgen05.mma kind::tf32, 3xTF32: hi/lo operands, two MMAs per K granule)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic[0], "traffic_source": traffic[1],
                "peak_source": src, "avg_launch_ms": tcg["ms"] / tcg["launches"],
                "launches_per_step": tcg["launches"] / prof["steps"], "kernel_share_of_step": tcg["ms"] / step_ms,
                "tf32_pipe_frac": 3.0 * achieved / (peak / 2.0),
                "note": "achieved counts the algorithmic 2MNK flops per product; every product costs three "
                        "TF32 passes (hi*hi, hi*lo, lo*hi) and the TF32 pipe peaks at half the bf16 rate, so the "
                        "tensor pipe itself runs at tf32_pipe_frac of its own peak"}
    if not eig["launches"]:
        return None
    # The Jacobi solver keeps its matrix in shared memory: HBM sees one read of the inputs and one
    # write of the eigenvectors per launch and is nowhere near binding.  Its roof is the SM's FP32
    # SIMT pipe: `achieved` = the FP32 flops the kernel itself counted (2D per column-pair dot
    # product, 8D per applied rotation) over the event-timed launch durations, `peak` = SMs x 128
    # FMA lanes x 2 x the SM clock.  The HBM view stays in `hbm` for reference.
    sm_mhz = peaks.get("sm_max_mhz") or 1965.0
    peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    achieved = prof.get("eig_fp32_flops", 0.0) / (eig["ms"] * 1e-3) / 1e12
    hbm_peak = peaks.get("hbm_gbs") or 6650.0
    hbm_ach = eig["work"] / (eig["ms"] * 1e-3) / 1e9
    traffic = NCU_TRAFFIC.get(("eig", r["B"], r["D"]), (None, None))
    return {"bound": "fp32", "kernel": "eig_jacobi_oe_kernel (warm layer solves; eig_jacobi_small_kernel: the loss solve)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": traffic[0], "traffic_source": traffic[1],
            "peak_source": f"148 SMs x 128 FP32 lanes x 2 x {sm_mhz:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz)",
            "avg_launch_ms": eig["ms"] / eig["launches"], "launches_per_step": eig["launches"] / prof["steps"],
            "kernel_share_of_step": eig["ms"] / step_ms,
            "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak},
            "note": "shared-memory-resident one-sided Jacobi: bound by SM issue slots and the shared-memory / "
                    "MUFU latency of a barrier-separated dependent chain, not by HBM (DESIGN.md); the flops are "
                    "counted by the kernel (dot products and applied rotations), not assumed"}


def shard_check(rank, world, group, dev):
    """N ranks == 1 rank on the GPUs: a small multitask batch (8 graphs per rank, D=100) through the
    sharded forward + backward, compared on rank 0 with the single-process run over ALL graphs."""
    import torch
    import torch.distributed as dist
    from uglad_b200 import main as ug, ops
    from uglad_b200.utils import prepare_data
    nb, D, M = 8, 100, 400
    X_all = synth(nb * world, D, M, 4321)
    S_all = prepare_data.get_covariance(torch.from_numpy(X_all).to(dev))
    mine = S_all[rank * nb:(rank + 1) * nb].contiguous()
    torch.manual_seed(1)
    model, _ = ug.init_uGLAD(lr=0.002)
    ops.reset_warm_start()
    th, loss = ug.forward_uGLAD(mine, model, L=L_LAYERS, group=group, total_graphs=nb * world)
    loss.backward()
    g_sh = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
    tot = loss.detach().clone().reshape(1)
    dist.all_reduce(tot, group=group)
    out = None
    if rank == 0:
        model.zero_grad()
        ops.reset_warm_start()
        th1, loss1 = ug.forward_uGLAD(S_all, model, L=L_LAYERS)
        loss1.backward()
        g1 = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        rt = float(torch.linalg.norm(th - th1[:nb]) / torch.linalg.norm(th1[:nb]))
        rg = float(torch.linalg.norm(g_sh - g1) / torch.linalg.norm(g1))
        rl = abs(float(tot) - float(loss1)) / max(1.0, abs(float(loss1)))
        out = {"graphs": nb * world, "theta_rel": rt, "grad_rel": rg, "loss_rel": rl,
               "ok": bool(rt < 1e-5 and rg < 1e-4 and rl < 1e-5)}
    ops.reset_warm_start()
    dist.barrier(group=group)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="multitask_d100", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-extra", action="store_true", help="skip the extra per-config measurements")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay of the epoch")
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
    numa = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
        # N processes stream their samples from one host at once: keep each one's pinned staging memory and
        # threads on the NUMA node of its GPU
        from uglad_b200.utils import prepare_data as _pd
        numa = _pd.bind_host_to_device_numa(dev)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    wl = args.workload
    spec = WORKLOADS[wl]
    scaling = "strong" if spec.get("consensus") else args.scaling
    r = time_gpu_workload(wl, args.steps, args.warmup, rank, world, group, dev, scaling=scaling, profile=True,
                          use_graph=not args.no_graph)
    extra = {}

    def extra_entry(name, x, with_cpu):
        e = {"baseline_config_index": WORKLOADS[name]["config_index"], "value": x["value"], "unit": UNIT,
             "ms_per_step": x["ms_per_step"], "e2e_value": x["e2e_value"], "e2e_ms_per_step": x["e2e_ms_per_step"],
             "h2d_bytes_per_step": x["h2d"], "graphs_total": x["B_total"], "graphs_per_gpu": x["B"],
             "cuda_graph": x["cuda_graph"],
             "gpu_launches": x["launches"], "roofline": roofline_of(x, peaks)}
        if with_cpu and rank == 0:
            e["cpu_baseline"] = cpu_baseline_object(name)
        return e

    if not args.no_extra and wl == "multitask_d100":
        if world == 1:
            for name, st, wu in (("single_d100", args.steps, args.warmup), ("consensus_d200", min(args.steps, 10), 3),
                                 ("single_d1000", min(args.steps, 5), 3)):
                x = time_gpu_workload(name, st, wu, rank, world, group, dev, scaling="strong", profile=True,
                                      use_graph=not args.no_graph)
                extra[name] = extra_entry(name, x, with_cpu=True)
        else:
            # the literal configs[2] / configs[3]: a FIXED 256-graph batch / 32 imputations over the N GPUs
            if scaling == "weak":
                x = time_gpu_workload("multitask_d100", args.steps, args.warmup, rank, world, group, dev,
                                      scaling="strong", profile=False, use_graph=not args.no_graph)
                extra["multitask_d100_strong"] = dict(extra_entry("multitask_d100", x, with_cpu=False), scaling="strong")
            x = time_gpu_workload("consensus_d200", min(args.steps, 10), 3, rank, world, group, dev, scaling="strong",
                                  profile=False, use_graph=not args.no_graph)
            extra["consensus_d200"] = dict(extra_entry("consensus_d200", x, with_cpu=False), scaling="strong")
            extra["shard_check"] = shard_check(rank, world, group, dev)

    if rank == 0:
        D, B = r["D"], r["B"]
        cpu = cpu_baseline_object(wl) if world == 1 else None   # the CPU baseline is reported at N=1 only
        line = {
            "metric": METRIC, "value": r["value"], "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl, "baseline_config_index": spec["config_index"], "graphs_per_gpu": B,
                       "graphs_total": r["B_total"], "D": D, "M": r["M"], "L": L_LAYERS, "H": 3,
                       "parallelism": f"graph-sharded x{world}", "host_numa_node": numa, "cuda_graph": r["cuda_graph"],
                       "l2_policy": "working set per step exceeds L2 "
                       "(saved theta / theta_k1 / eigenvectors of 15 layers: %.0f MB)" % (B * D * D * 4 * 3 * L_LAYERS / 1e6)},
            "clocks": r["clocks"],
            "e2e": {"value": r["e2e_value"], "unit": UNIT, "ms_per_step": r["e2e_ms_per_step"],
                    "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
            "gpu_launches": r["launches"],
            "roofline": roofline_of(r, peaks),
            "cpu_baseline": cpu,
            "extra": extra,
        }
        emit(line)
    if world > 1:
        # The result line is out.  Should the teardown of some rank ever block (a captured collective that outlives the
        # communicator did exactly that once: 6 minutes until the driver's limit), leave anyway after half a minute.
        wd = threading.Timer(30.0, lambda: os._exit(0))
        wd.daemon = True
        wd.start()
        # nothing that captured a collective may outlive the communicator: collect, drain, meet, then tear down
        import gc
        gc.collect()
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
