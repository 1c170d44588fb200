#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv --print-source sass,cuda` export by CUDA source line:
usage: ncu_lines.py file.csv [kernel_index] -> top lines by stall samples and by instructions executed"""
import csv, sys
from collections import defaultdict

def num(x):
    try:
        return int(x)
    except ValueError:
        return 0

path = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
blocks, cur = [], None
for row in csv.reader(open(path)):
    if row and row[0] == "File Path":
        pass
    if row and row[0] == "Function Name":
        cur = {"fn": row[1], "rows": [], "hdr": None}; blocks.append(cur); continue
    if cur is None: continue
    if row and row[0] == "Line No": cur["hdr"] = row; continue
    if cur["hdr"] and len(row) == len(cur["hdr"]): cur["rows"].append(row)
# kernels appear in order; group blocks by launch: a new launch starts when the function name repeats the first one
print(len(blocks), "blocks:", [(i, bb["fn"][10:40], len(bb["rows"])) for i, bb in enumerate(blocks)])
b = blocks[which]
hdr = b["hdr"]; iL = 0; iS = 1; iA = hdr.index("# Samples"); iE = hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = defaultdict(lambda: [0, 0, defaultdict(int)]); src = {}
for r in b["rows"]:
    key = r[iL]; src[key] = r[iS]
    agg[key][0] += num(r[iA]); agg[key][1] += num(r[iE])
    for i in stalls: agg[key][2][hdr[i]] += num(r[i])
ts = sum(v[0] for v in agg.values()); te = sum(v[1] for v in agg.values())
print(b["fn"][:80], "samples", ts, "inst", te)
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    top = sorted(v[2].items(), key=lambda kv: -kv[1])[:3]
    print("%5s %5.1f%% smp %5.1f%% inst  %-70s %s" % (key, 100 * v[0] / ts, 100 * v[1] / te, src[key].strip()[:70], ", ".join("%s=%d" % (k.replace("stall_", ""), n) for k, n in top)))
