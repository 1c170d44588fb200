"""Small fixed case for ncu: one points launch and lognormal1d launches at levels 0, 3 and 6 (run twice: warm + profiled)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mle_b200, mle_b200.kl as kl

eng = mle_b200.Engine(0)
m, R = 100, 16
z = eng.genvec("K_3600_32")[:m]
lat = eng.lattice(z, m)
d = eng.dists([(1, [0.0, 1.0])] * m)
gm = eng.model("lognormal1d", 1, [16])
levels = [0, 3, 6]
for lev in levels:
    gm.set_basis(lev, kl.kl_basis_1d(lev, m=m))
sh = torch.from_numpy(np.random.default_rng(0).random((R, m))).cuda()
mom = torch.zeros(R * 2 * 3, dtype=torch.float64, device="cuda")
npts = 1 << 16
out = torch.empty(R * npts * m, dtype=torch.float64, device="cuda")
ns = {0: 1 << 16, 3: 1 << 12, 6: 1 << 9}
if len(sys.argv) > 1:
    ns = {0: 1 << 20, 3: 46341, 6: 2048}
for rep in range(2):
    eng.points_dev(lat, d, sh.data_ptr(), R, 0, npts, m, out.data_ptr())
    for lev in levels:
        eng.sample_batch_dev(gm, lat, d, [lev], sh.data_ptr(), R, 0, ns[lev], m, mom.data_ptr())
    eng.sync()
print("ok", eng.launch_count)
