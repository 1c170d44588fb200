"""Small fixed case for ncu: the latency variant of the lognormal1d kernel at levels 0, 3 (one CTA per sample tile) and 5, 6
(thread-block clusters), 16 shifts x 4 points, run twice (warm + profiled)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mle_b200, mle_b200.kl as kl
eng = mle_b200.Engine(0)
m, R = 100, 16
lat = eng.lattice(None, m); d = eng.dists([(1, [0.0, 1.0])] * m); gm = eng.model("lognormal1d", 1, [16])
levels = [0, 3, 5, 6]
for lev in levels: gm.set_basis(lev, kl.kl_basis_1d(lev, m=m))
sh = torch.from_numpy(np.random.default_rng(0).random((R, m))).cuda()
mom = torch.zeros(R * 6, dtype=torch.float64, device="cuda")
for rep in range(2):
    for lev in levels:
        eng.sample_batch_dev(gm, lat, d, [lev], sh.data_ptr(), R, 0, 4, m, mom.data_ptr())
    eng.sync()
print("ok", eng.launch_count)
