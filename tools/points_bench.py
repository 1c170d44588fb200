"""points_kernel throughput (stage 1-2 store-bound form): 2^18 points x 100 dims x 16 shifts, Normal inverse CDF and raw."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mle_b200
eng = mle_b200.Engine(0)
m, R, npts = 100, 16, 1 << 18
lat = eng.lattice(None, m); d = eng.dists([(1, [0.0, 1.0])] * m)
sh = torch.from_numpy(np.random.default_rng(0).random((R, m))).cuda()
out = torch.empty(R * npts * m, dtype=torch.float64, device="cuda")
stream = torch.cuda.ExternalStream(eng.stream, device=0)
for name, dd in (("normal", d), ("raw", None)):
    for _ in range(2): eng.points_dev(lat, dd, sh.data_ptr(), R, 0, npts, m, out.data_ptr())
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(5): eng.points_dev(lat, dd, sh.data_ptr(), R, 0, npts, m, out.data_ptr())
    b.record(stream); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print("%s: %.3f ms, %.1f GB/s, %.1f G coordinates/s" % (name, ms, out.numel() * 8 / ms * 1e-6, out.numel() / ms * 1e-6), flush=True)
