#!/usr/bin/env python
"""Digest of one kernel of an ncu --set full report: python tools/ncu_digest.py <rep> <kernel regex> > profiles/ncu_<name>.txt
(the metrics DESIGN.md quotes, the stall reasons per issued instruction, then tools/ncu_src.py's opcode / hot-line table)."""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
print("==", d.get("Kernel Name", ("?", ""))[0][:120])
for k in want:
    if k in d:
        print(f"   {k:75s} {d[k][0]} {d[k][1]}")
print("   -- stall reasons, warps per issued instruction --")
for h in sorted(hdr):
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
        try:
            v = float(d[h][0].replace(",", ""))
        except ValueError:
            continue
        if v >= 0.1:
            print(f"   {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:40s} {v:.2f}")
print("   -- executed warp instructions by opcode, stall samples, hottest lines (tools/ncu_src.py) --")
sys.stdout.flush()
out = subprocess.run([sys.executable, __file__.replace("ncu_digest.py", "ncu_src.py"), rep, pat, "30"], capture_output=True, text=True).stdout
print("\n".join("   " + l for l in out.splitlines()))
