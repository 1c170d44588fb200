import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import mle_b200, mle_b200.kl as kl
eng = mle_b200.Engine(0)
m, R = 100, 16
lat = eng.lattice(None, m); d = eng.dists([(1, [0.0, 1.0])] * m); gm = eng.model("lognormal1d", 1, [16])
sh = np.random.default_rng(0).random((R, m))
for lev in (0, 3, 6):
    gm.set_basis(lev, kl.kl_basis_1d(lev, m=m))
    for n in (1, 100):
        for i in range(3): eng.sample_batch(gm, lat, d, [lev], sh, i * n, n)
