#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full) and a launch list CSV into the text summaries committed under profiles/.

    python profiles/summarize.py gpurun_out/prof_r1b.ncu-rep gpurun_out/launches_r1b.csv r1b
"""
import csv
import json
import subprocess
import sys
from collections import Counter, defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                d[k] = f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip()
        out.append(d)
    return out


def source_mix(rep, pattern):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pattern}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        return None
    hdr = rows[1]
    isrc, iex, ism = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    c, smp = Counter(), Counter()
    tot = 0
    for r in rows[2:]:
        if len(r) <= iex:
            continue
        s = r[isrc].strip().split()
        op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
        c[op] += int(r[iex]); smp[op] += int(r[ism]); tot += int(r[iex])
    return tot, c, smp


def main():
    rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    lines = []
    for d in raw(rep):
        lines.append(f"== {d['kernel'][:110]}")
        for k in KEYS:
            if k in d:
                lines.append(f"   {k:75s} {d[k]}")
    for pat in ("stft_tc_kernel", "stft_main", "frame_chain"):
        m = source_mix(rep, pat)
        if m:
            tot, c, smp = m
            lines.append(f"== executed warp instructions by opcode: {pat} (total {tot})")
            for op, v in c.most_common(18):
                lines.append(f"   {op:10s} {v:13d} {100 * v / tot:5.1f}%   stall samples {smp[op]}")
    open(f"profiles/ncu_full_{tag}.txt", "w").write("\n".join(lines) + "\n")
    # launch list: per-kernel totals and shares
    rows = list(csv.reader(l for l in open(launches) if l.startswith('"')))
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg, cnt = defaultdict(float), Counter()
    for r in rows[1:]:
        name = r[ik].split("(")[0][:60]
        agg[name] += float(r[iv].replace(",", "")); cnt[name] += 1
    tot = sum(agg.values())
    with open(f"profiles/launches_{tag}.txt", "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
            f.write(f"{k:62s} launches {cnt[k]:3d}  total {v / 1e3:10.1f} us  share {100 * v / tot:5.1f}%\n")
    print(open(f"profiles/ncu_full_{tag}.txt").read())
    print(open(f"profiles/launches_{tag}.txt").read())


if __name__ == "__main__":
    main()
