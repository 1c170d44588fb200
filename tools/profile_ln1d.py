"""ncu target: lognormal1d launches at the levels given on the command line (default 0 1 2), C2 shapes, n = 2^17 / 2^(1.5 l) per shift."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mle_b200, mle_b200.kl as kl

levels = [int(a) for a in sys.argv[1:]] or [0, 1, 2]
eng = mle_b200.Engine(0)
m, R = 100, 16
z = eng.genvec("K_3600_32")[:m]
lat = eng.lattice(z, m)
d = eng.dists([(1, [0.0, 1.0])] * m)
gm = eng.model("lognormal1d", 1, [16])
for lev in levels:
    gm.set_basis(lev, kl.kl_basis_1d(lev, m=m))
sh = torch.from_numpy(np.random.default_rng(0).random((R, m))).cuda()
mom = torch.zeros(R * 2 * 3, dtype=torch.float64, device="cuda")
for rep in range(2):
    for lev in levels:
        n = max(64, int((1 << 17) * 2.0 ** (-1.5 * lev)))
        eng.sample_batch_dev(gm, lat, d, [lev], sh.data_ptr(), R, 0, n, m, mom.data_ptr())
    eng.sync()
print("ok", eng.launch_count)
