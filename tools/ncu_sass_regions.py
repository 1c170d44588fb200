#!/usr/bin/env python3
"""Per-address view of an `ncu --page source --csv --print-source sass,cuda` export: python tools/ncu_sass_regions.py file.csv [block]
Prints, for the chosen kernel block, every backward-branch loop (address range, instruction mix, samples, executed instructions)
and the share of stall samples outside loops, so that the time of a fused kernel can be attributed to its phases."""
import csv, re, sys
from collections import Counter, defaultdict

path = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
blocks, cur = [], None
fpath = ""
for row in csv.reader(open(path)):
    if row and row[0] == "File Path":
        fpath = row[1]; continue
    if row and row[0] == "Function Name":
        cur = {"fn": row[1], "rows": [], "file": fpath}; blocks.append(cur); continue
    if cur is None or len(row) < 8: continue
    if row[2].startswith("0x"): cur["rows"].append(row)
# a launch = consecutive blocks (one per source file) of the same function; merge blocks until the function name changes or addresses repeat
launches = []
for b in blocks:
    # blocks of one launch (one per source file) follow each other; a block for the file that opened the launch starts the next one
    if launches and launches[-1]["fn"] == b["fn"] and b["file"] not in launches[-1]["files"]:
        launches[-1]["files"].add(b["file"])
        for r in b["rows"]:
            if r[2] not in launches[-1]["seen"]:
                launches[-1]["seen"].add(r[2]); launches[-1]["rows"].append(r)
    else:
        launches.append({"fn": b["fn"], "rows": list(b["rows"]), "seen": set(r[2] for r in b["rows"]), "files": {b["file"]}})
print(len(launches), "launches:", [(i, l["fn"][10:48], len(l["rows"])) for i, l in enumerate(launches)])
L = launches[which]
ins = sorted({int(r[2], 16): (int(r[2], 16), r[3].strip(), int(r[6] or 0), int(r[7] or 0), r[0]) for r in L["rows"]}.values())
base = ins[0][0]
ins = [(a - base, s, smp, ex, ln) for a, s, smp, ex, ln in ins]
tot_s = sum(i[2] for i in ins); tot_e = sum(i[3] for i in ins)
print(L["fn"][:90], "instructions", len(ins), "samples", tot_s, "executed", tot_e)
def op(s): return re.sub(r'^@!?U?P\d+\s+', '', s).split()[0].split('.')[0]
loops = []
for a, s, smp, ex, ln in ins:
    m = re.search(r'BRA\S*\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?0x([0-9a-f]+)', s) or re.search(r'BRA.*?0x([0-9a-f]+)', s)
    if m:
        t = int(m.group(1), 16)
        t = t - base if t >= base else t
        if t < a: loops.append((t, a))
loops.sort(key=lambda x: (x[0], -x[1]))
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
for t, a in loops:
    body = [i for i in ins if t <= i[0] <= a]
    c = Counter(op(i[1]) for i in body)
    s = sum(i[2] for i in body); e = sum(i[3] for i in body)
    print("loop %05x-%05x n=%4d fp64=%3d dmma=%2d | samples %5.1f%% executed %5.1f%% | %s" % (
        t, a, len(body), sum(c[k] for k in FP64), c["DMMA"], 100.0 * s / tot_s, 100.0 * e / tot_e,
        " ".join("%s:%d" % kv for kv in c.most_common(8))))
if len(sys.argv) > 3:
    lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    for i in ins:
        if lo <= i[0] <= hi: print("%05x %5d %9d  %s" % (i[0], i[2], i[3], i[1]))
if len(sys.argv) == 3 or (len(sys.argv) > 5 and sys.argv[5] == "buckets"):
    # coarse address buckets with the dominant source lines
    B = 0x400
    bk = defaultdict(lambda: [0, 0, Counter()])
    for a, s, smp, ex, ln in ins:
        k = a // B; bk[k][0] += smp; bk[k][1] += ex
    print("bucket  samples%  executed%")
    for k in sorted(bk):
        if bk[k][0] * 200 > tot_s: print("%05x  %5.1f  %5.1f" % (k * B, 100.0 * bk[k][0] / tot_s, 100.0 * bk[k][1] / tot_e))
