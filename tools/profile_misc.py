"""One launch each of the kernels other than the headline one, for an ncu capture (profiles/r01_other_kernels.csv):
points_kernel with the Normal inverse CDF and raw, the fused expsum kernel (config C5 shape), the 2-D register-window
kernel at 32x32 and 128x128 (config C3 shape)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mle_b200, mle_b200.kl as kl

eng = mle_b200.Engine(0)
R, s = 16, 100
lat = eng.lattice(None, s)
d = eng.dists([(mle_b200.NORMAL, [0.0, 1.0])] * s)
sh = torch.rand(R, s, dtype=torch.float64, device="cuda")
n = 1 << 18
out = torch.empty(R * n * s, dtype=torch.float64, device="cuda")
mom = torch.zeros(R * 6, dtype=torch.float64, device="cuda")
w = [0.5 / (j + 1) for j in range(s)]
gm = eng.model("expsum", 1, [float(s)] + w)
g2 = eng.model("lognormal2d", 1, [4])
m2 = 400
d2 = eng.dists([(mle_b200.NORMAL, [0.0, 1.0])] * m2)
for lev in (3, 5):
    g2.set_basis(lev, kl.kl_basis_2d(lev, m=m2))
for rep in range(2):  # the second round is the one to profile (--launch-skip 5 --launch-count 5)
    eng.points_dev(lat, d, sh.data_ptr(), R, 0, n, s, out.data_ptr())
    eng.points_dev(lat, None, sh.data_ptr(), R, 0, n, s, out.data_ptr())
    eng.sample_batch_dev(gm, lat, d, [0], sh.data_ptr(), R, 0, 1 << 20, s, mom.data_ptr())
    eng.sample_batch_dev(g2, None, d2, [3], None, 1, 0, 8192, m2, mom.data_ptr(), mc_seed=1)
    eng.sample_batch_dev(g2, None, d2, [5], None, 1, 0, 1184, m2, mom.data_ptr(), mc_seed=1)
    eng.sync()
print("ok")
