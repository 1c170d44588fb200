"""ncu target: one lognormal2d launch (C3 shape, 400 KL terms, MC inputs) at the level given on the command line (default 5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mle_b200, mle_b200.kl as kl
lev = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
eng = mle_b200.Engine(0)
m = 400
d = eng.dists([(1, [0.0, 1.0])] * m)
gm = eng.model("lognormal2d", 1, [4])
gm.set_basis(lev, kl.kl_basis_2d(lev, m=m))
mom = torch.zeros(6, dtype=torch.float64, device="cuda")
for rep in range(2):
    eng.sample_batch_dev(gm, None, d, [lev], None, 1, 0, n, m, mom.data_ptr(), mc_seed=1)
    eng.sync()
print("ok")
