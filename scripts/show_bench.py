"""Print the key numbers of bench.py JSON lines: python scripts/show_bench.py file.json ..."""
import json, sys
for f in sys.argv[1:]:
    for line in open(f):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        e = d.get("e2e", {})
        print(f"{f}: N={d.get('n_gpus')} value {d['value']:.0f} ({d['ms_per_step']:.3f} ms)  e2e {e.get('value', 0):.0f} ({e.get('ms_per_step', 0):.3f} ms)"
              f" launches {d.get('gpu_launches')} graph={d['config'].get('cuda_graph')} clocks={d.get('clocks')}")
        r = d.get("roofline") or {}
        print("   roofline", {k: r.get(k) for k in ("kernel", "bound", "achieved", "peak", "frac", "avg_launch_ms", "kernel_share_of_step", "tf32_pipe_frac")})
        cb = d.get("cpu_baseline")
        if cb:
            print("   cpu", cb)
        for k, v in d.get("extra", {}).items():
            if isinstance(v, dict) and "ms_per_step" in v:
                r = v.get("roofline") or {}
                cb = v.get("cpu_baseline") or {}
                print(f"   {k}: {v['value']:.0f} ({v['ms_per_step']:.3f} ms) e2e {v['e2e_ms_per_step']:.3f} ms graph={v.get('cuda_graph')} "
                      f"frac={r.get('frac')} tf32={r.get('tf32_pipe_frac')} share={r.get('kernel_share_of_step')} cpu={cb.get('value')} ({cb.get('kind')})")
            else:
                print("   ", k, v)
