"""Static resource / instruction-mix table of every kernel in libmle_cuda.so: registers, stack, spills, barriers from the
ptxas log written by the build (-Xptxas -v), instruction counts from `cuobjdump -sass` (DMMA = FP64 tensor instruction,
UBLKCP = TMA bulk copy, MUFU = special function unit, BAR = CTA barrier).  Runs without a GPU.

    python tools/sass_summary.py > profiles/r01_sass_summary.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multilevelestimators.jl_b200", "lib")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = {}
    for n, d in zip(names, out):
        d = re.sub(r"^void ", "", d)
        d = re.sub(r"\(.*$", "", d)            # drop the argument list
        short[n] = d.replace("mle::", "")
    return short


def ptxas_table():
    rows, cur = {}, None
    for line in open(os.path.join(LIB, "ptxas.log")):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = m.group(1)
            rows[cur] = {}
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and "stack" not in rows[cur]:
            rows[cur].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
        m = re.search(r"Used (\d+) registers, used (\d+) barriers", line)
        if m:
            rows[cur].update(regs=int(m.group(1)), bars=int(m.group(2)))
            sm = re.search(r"(\d+) bytes smem", line)
            rows[cur]["smem"] = int(sm.group(1)) if sm else 0
    return rows


OPS = ("DMMA", "DFMA", "DADD", "DMUL", "DSETP", "MUFU", "BAR", "LDS", "STS", "LDG", "STG", "UBLKCP", "SHFL", "ATOM", "RED")


def sass_table():
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(LIB, "libmle_cuda.so")], capture_output=True, text=True).stdout
    rows, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            rows[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            rows[cur]["total"] += 1
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    rows[cur][o] += 1
    return rows


def main():
    pt, ss = ptxas_table(), sass_table()
    names = sorted(set(pt) & set(ss))
    short = demangle(names)
    print("# static summary of libmle_cuda.so (sm_100a): ptxas -v resources and SASS instruction counts per kernel")
    print("# %-58s %4s %6s %6s %6s %7s | %s" % ("kernel", "regs", "stack", "spl.st", "spl.ld", "st.smem",
                                                  " ".join("%6s" % o for o in ("total",) + OPS)))
    for n in sorted(names, key=lambda k: short[k]):
        p, c = pt[n], ss[n]
        print("%-60s %4d %6d %6d %6d %7d | %s" % (short[n][:60], p.get("regs", 0), p.get("stack", 0), p.get("spill_st", 0),
                                                p.get("spill_ld", 0), p.get("smem", 0),
                                                " ".join("%6d" % c[o] for o in ("total",) + OPS)))


if __name__ == "__main__":
    main()
