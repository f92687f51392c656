#!/usr/bin/env python
"""Summarise an .ncu-rep: per-kernel headline metrics + stall-reason shares (reads `ncu -i ... --csv`)."""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
names = []
for r in data:
    print("====", r[idx["Kernel Name"]][:110])
    names.append(r[idx["Kernel Name"]])
    for w in want:
        if w in idx:
            print(f"  {w:78s} {units[idx[w]]:16s} {r[idx[w]]}")
for k, name in enumerate(names):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    d = [r for r in rows[2:] if r and r[0].startswith("0x")]
    ix = {c: i for i, c in enumerate(h)}
    tot = sum(int(r[ix["# Samples"]]) for r in d)
    agg = collections.Counter()
    for c in h:
        if c.startswith("stall_") and "Not Issued" not in c:
            agg[c] = sum(int(r[ix[c]] or 0) for r in d)
    ops = collections.Counter()
    for r in d:
        t = r[ix["Source"]].split()
        op = t[1] if t and t[0].startswith("@") else (t[0] if t else "")
        ops[op.split(".")[0]] += int(r[ix["Instructions Executed"]])
    print("==== stalls", name[:80], "samples", tot, "SASS lines", len(d))
    print("   ", ", ".join(f"{k2[6:]} {100 * v / tot:.1f}%" for k2, v in agg.most_common(9)))
    ti = sum(ops.values())
    print("    instr mix:", ", ".join(f"{k2} {100 * v / ti:.1f}%" for k2, v in ops.most_common(12)), f"(total warp instr {ti:.3g})")
