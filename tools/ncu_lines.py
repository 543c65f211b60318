#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source file / line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv; python tools/ncu_lines.py x.csv [n_warps]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
nw = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
cur = None
agg = collections.Counter(); smp = collections.Counter(); src = {}
hdr = None
def num(x):
    try: return int(x)
    except ValueError: return 0
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No": hdr = r; ia = r.index("Instructions Executed"); ismp = r.index("# Samples"); continue
    if r and r[0] == "Function Name": continue
    if hdr is None or len(r) <= ia: continue
    if r[0] != "":
        key = (cur, num(r[0])); agg[key] += num(r[ia]); smp[key] += num(r[ismp]); src[key] = r[1].strip()[:100]
tot = sum(agg.values()); st = sum(smp.values())
print("warp instructions:", tot, "per warp:", tot / nw)
byfile = collections.Counter(); sf = collections.Counter()
for k, v in agg.items(): byfile[k[0]] += v; sf[k[0]] += smp[k]
for f, v in byfile.most_common(): print(f"{f}: {v/tot*100:.1f}% inst, {sf[f]/st*100:.1f}% samples")
for k, v in agg.most_common(70): print(f"{k[0]}:{k[1]:4d} {v/tot*100:5.2f}% smp {smp[k]/st*100:5.2f}%  {src[k]}")
