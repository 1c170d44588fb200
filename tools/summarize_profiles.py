#!/usr/bin/env python3
"""Turn the scratch ncu output under gpurun_out/ into the tracked summaries under profiles/.

usage: python tools/summarize_profiles.py <round tag> <launches.csv> <full .ncu-rep | its --page raw --csv export>
  profiles/<tag>_launches.csv        per-kernel totals of the `--metrics gpu__time_duration.sum` launch list
  profiles/<tag>_kernels.csv         one row per profiled launch of the `--set full` capture (key metrics)
"""
import collections
import csv
import os
import subprocess
import sys

tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles")
os.makedirs(root, exist_ok=True)

rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
hdr = rows[0]
ik, iv, iu, ip = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Process Name")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    key = (r[ip], r[ik].split("(")[0])
    tot[key] += float(r[iv].replace(",", "")) * scale
    cnt[key] += 1
ours = sum(v for (p, k), v in tot.items() if "mle::" in k)
with open(os.path.join(root, tag + "_launches.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["process", "kernel", "launches", "total_ms", "share_of_libmle_cuda_time_pct"])
    for (p, k), v in tot.most_common():
        w.writerow([p, k, cnt[(p, k)], "%.4f" % v, "%.2f" % (100 * v / ours) if "mle::" in k else ""])

# the full capture: a .ncu-rep, or its `ncu -i <rep> --page raw --csv` export (made on the GPU box when the report itself is too
# large to bring back)
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "gpu__time_duration.sum",
        "smsp__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
cols = []
for wname in want:
    for i, h in enumerate(hdr):
        if h.endswith(wname):
            cols.append((wname + (" [" + units[i] + "]" if units[i] else ""), i))
            break
stall = [i for i, h in enumerate(hdr) if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def col(name):
    for i, h in enumerate(hdr):
        if h.endswith(name):
            return i
    return None


def num(r, i, scale_units=False):
    if i is None or r[i] in ("", "n/a"):
        return float("nan")
    v = float(r[i].replace(",", ""))
    return v * SCALE.get(units[i], 1.0) if scale_units else v


i_rd, i_wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
i_fp64 = col("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")
i_tens = col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
with open(os.path.join(root, tag + "_kernels.csv"), "w", newline="") as f:
    w = csv.writer(f)
    # the last four columns are unit-free figures bench.py reads: DRAM bytes of the launch, FP64-pipe and DMMA-pipe busy shares
    w.writerow([c for c, _ in cols] + ["top warp stall reasons (% of samples)", "dram_bytes", "pipe_fp64_pct", "pipe_dmma_pct",
                                         "pipe_fp64_plus_dmma_pct"])
    for r in data:
        tot_s = sum(float(r[i] or 0) for i in stall) or 1.0
        top = sorted(((float(r[i] or 0) / tot_s * 100, hdr[i].split("issue_stalled_")[-1]) for i in stall), reverse=True)[:5]
        fp, te = num(r, i_fp64), num(r, i_tens)
        w.writerow([r[i][:90] for _, i in cols] + [" ".join("%s=%.1f" % (n, v) for v, n in top),
                                                   "%.0f" % (num(r, i_rd, True) + num(r, i_wr, True)), "%.2f" % fp, "%.2f" % te,
                                                   "%.2f" % (fp + te)])
print("wrote profiles/%s_launches.csv and profiles/%s_kernels.csv" % (tag, tag))
