"""Per-level timing of the lognormal1d kernel at the C2 bench sizes: python tools/ln_levels.py  (env knobs: MLE_CUDA_LIB, MLE_LEVELS=0,1)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mle_b200, mle_b200.kl as kl
import bench as B
if os.environ.get("MLE_LEVELS"): B.LEVELS = [int(x) for x in os.environ["MLE_LEVELS"].split(",")]

eng = mle_b200.Engine(0)
m, R = B.M_KL, B.R_SHIFTS
z = eng.genvec("K_3600_32")[:m]
lat = eng.lattice(z, m); d = eng.dists([(1, [0.0, 1.0])] * m)
gm = eng.model("lognormal1d", 1, [16])
for lev in B.LEVELS: gm.set_basis(lev, kl.kl_basis_1d(lev, m=m))
sh = torch.from_numpy(B.level_shifts()).cuda()
mom = torch.zeros(R * 6, dtype=torch.float64, device="cuda")
stream = torch.cuda.ExternalStream(eng.stream, device=0)
ms = []
for li, lev in enumerate(B.LEVELS):
    n = B.points_per_shift(lev)
    for _ in range(2): eng.sample_batch_dev(gm, lat, d, [lev], sh[li].data_ptr(), R, 0, n, m, mom.data_ptr())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(stream)
    for _ in range(5): eng.sample_batch_dev(gm, lat, d, [lev], sh[li].data_ptr(), R, 0, n, m, mom.data_ptr())
    b.record(stream); torch.cuda.synchronize()
    ms.append(a.elapsed_time(b) / 5)
tf = [B.algorithmic_flops_per_sample(lev) * B.points_per_shift(lev) * R / (ms[i] * 1e-3) * 1e-12 for i, lev in enumerate(B.LEVELS)]
print(os.environ.get("MLE_LN_WS", "-"), os.environ.get("MLE_CUDA_LIB", "-"), "ms", " ".join("%.3f" % x for x in ms), "sum %.3f" % sum(ms), "| TF", " ".join("%.1f" % x for x in tf))
