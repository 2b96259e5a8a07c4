"""Key metrics of an `ncu --set full` report: ncu -i X.ncu-rep --page raw --csv | python scripts/ncu_summary.py [title]"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
units = rows[1]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__cluster_dim_x", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__cycles_active.min", "sm__cycles_active.max", "sm__cycles_active.avg"]
stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
idx = {h: i for i, h in enumerate(hdr)}
if len(sys.argv) > 1:
    print("#", " ".join(sys.argv[1:]))
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    print(f"launch {r[idx['ID']]}: {r[idx['Kernel Name']][:70]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
    for k in keys:
        if k in idx:
            print(f"   {k:78s} {r[idx[k]]:>16s} {units[idx[k]]}")
    st = sorted(((float(r[idx[h]].replace(',', '')) if r[idx[h]] not in ('', 'n/a') else 0.0, h) for h in stall), reverse=True)[:7]
    for v, h in st:
        print(f"   stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:8.2f} warps per issue")
