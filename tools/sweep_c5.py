"""Config C5, the raw sweep: N = 2^20 .. 2^30 lattice points x s = 100 / 250 / 1000 dims, 16 random shifts,
all dims Normal(0,1), integrand Q = exp(sum_j a_j xi_j), a_j = 0.5/j, shift-averaged moments.

Two arms per (N, s):
  fused    mle_sample_batch_dev("expsum"): generate -> integrand -> moments, nothing written per sample (0 HBM bytes/coordinate)
  points   mle_points_dev: generate -> store [m x n x R] (8 B per coordinate, the HBM-bound form), N capped so the output
           fits in 32 GiB

Every row also checks the estimate against the closed form E[Q] = exp(sum a_j^2 / 2) in units of the shift standard error.
Usage: python tools/sweep_c5.py [--max-log2 30] [--out gpurun_out/c5_sweep.json]"""
import argparse
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import mle_b200

ap = argparse.ArgumentParser()
ap.add_argument("--max-log2", type=int, default=30)
ap.add_argument("--min-log2", type=int, default=20)
ap.add_argument("--dims", type=int, nargs="*", default=[100, 250, 1000])
ap.add_argument("--out", default=None)
args = ap.parse_args()

eng = mle_b200.Engine(0)
stream = torch.cuda.ExternalStream(eng.stream)
R = 16
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm_peak = float(peaks.get("hbm_copy_gbs", peaks.get("hbm_gbs", 6547.5))) if isinstance(peaks, dict) else 6547.5


def splitmix(seed, n):
    out = np.empty(n)
    x = seed
    for i in range(n):
        x = (x + 0x9E3779B97F4A7C15) & (2**64 - 1)
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
        z ^= z >> 31
        out[i] = (z >> 11) * 2.0**-53
    return out


def timed(fn):
    fn()
    eng.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    fn()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3


rows = []
for s in args.dims:
    w = [0.5 / (j + 1) for j in range(s)]
    exact = math.exp(0.5 * sum(x * x for x in w))
    lat = eng.lattice(None, s)
    d = eng.dists([(mle_b200.NORMAL, [0.0, 1.0])] * s)
    gm = eng.model("expsum", 1, [float(s)] + w)
    shifts = torch.from_numpy(splitmix(2024, R * s).reshape(R, s)).cuda()
    mom = torch.zeros(R * 2 * 3, dtype=torch.float64, device="cuda")
    for lg in range(args.min_log2, args.max_log2 + 1, 2):
        n = (1 << lg) // R
        t = timed(lambda: eng.sample_batch_dev(gm, lat, d, [0], shifts.data_ptr(), R, 0, n, s, mom.data_ptr()))
        mh = mom.cpu().numpy().reshape(R, 2, 3)
        means = mh[:, 1, 1]
        est, se = means.mean(), math.sqrt(means.var(ddof=1) / R)
        row = {"s": s, "log2_points": lg, "points_per_shift": n, "fused_s": t, "fused_coords_per_s": R * n * s / t,
               "fused_points_per_s": R * n / t, "estimate": est, "exact": exact, "std_err": se, "err_in_se": (est - exact) / se}
        # the store-bound form, while the output fits
        nbytes = 8 * R * n * s
        if nbytes <= 32 * 2**30:
            out = torch.empty(R * n * s, dtype=torch.float64, device="cuda")
            tp = timed(lambda: eng.points_dev(lat, d, shifts.data_ptr(), R, 0, n, s, out.data_ptr()))
            row.update(points_s=tp, points_gbs=nbytes / tp * 1e-9, points_frac_of_hbm_peak=nbytes / tp * 1e-9 / hbm_peak)
            del out
        rows.append(row)
        print("s=%4d N=2^%d  fused %9.3f ms  %7.1f G coord/s  est %.9f exact %.9f (%+.2f se)%s" % (
            s, lg, t * 1e3, row["fused_coords_per_s"] * 1e-9, est, exact, row["err_in_se"],
            "   points %9.3f ms %7.1f GB/s (%.0f%% of %.0f)" % (row["points_s"] * 1e3, row["points_gbs"], 100 * row["points_frac_of_hbm_peak"], hbm_peak)
            if "points_s" in row else ""), flush=True)
if args.out:
    json.dump({"device": torch.cuda.get_device_name(0), "shifts": R, "hbm_peak_gbs": hbm_peak, "rows": rows}, open(args.out, "w"), indent=1)
