#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics, SASS opcode mix, per-source-line mix.
usage: python tools/ncu_summary.py X.ncu-rep n_warps > profiles/<round>/X.summary.txt"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nw = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "launch__registers_per_thread ",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum ", "smsp__issue_active.avg.pct",
        "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_fma.avg.pct", "sm__pipe_fmaheavy_cycles_active.avg",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued",
        "smsp__average_warps_issue_stalled", "gpu__dram_throughput.avg", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg", "sm__cycles_elapsed.max ", "smsp__cycles_active.avg ", "sass__inst_executed_local",
        "sm__throughput.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "smsp__inst_executed_pipe_fma", "sm__inst_executed_pipe_fmaheavy"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print("== raw metrics (", rep, ") ==")
for k in range(2, len(rows)):
    print("-- kernel:", rows[k][rows[0].index("Kernel Name")] if "Kernel Name" in rows[0] else k)
    for h, u, v in zip(rows[0], rows[1], rows[k]):
        if any((h + " ").startswith(key) or key in h + " " for key in KEYS) and v not in ("", "0"):
            print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ia = hdr.index("Instructions Executed"); isrc = hdr.index("Source"); ismp = hdr.index("# Samples")
cnt = collections.Counter(); smp = collections.Counter(); tot = 0
for r in rows[2:]:
    if len(r) <= ia: continue
    t = r[isrc].split()
    if t and t[0].startswith('@'): t = t[1:]
    op = '.'.join((t[0] if t else '?').split('.')[:3])
    try: n = int(r[ia]); s = int(r[ismp])
    except ValueError: continue
    cnt[op] += n; tot += n; smp[op] += s
st = max(sum(smp.values()), 1)
print(f"\n== SASS opcode mix: {tot} warp instructions, {tot/nw:.0f} per warp ==")
for op, n in cnt.most_common(45): print(f"{op:26s} {n/tot*100:6.2f}%  {n/nw:12.0f} per warp   stall samples {smp[op]/st*100:5.1f}%")
