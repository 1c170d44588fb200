import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mle_b200, mle_b200.kl as kl
eng = mle_b200.Engine(0)
m = 400
d = eng.dists([(1, [0.0, 1.0])] * m)
gm = eng.model("lognormal2d", 1, [4])
mom = torch.zeros(6, dtype=torch.float64, device="cuda")
levels = [int(x) for x in os.environ.get('P2D_LEVELS', '0,2').split(',')]
for lev, n in [(l, max(296 * 8, (1 << 17) >> (3 * l))) for l in levels]:
    gm.set_basis(lev, kl.kl_basis_2d(lev, m=m))
    for rep in range(2):
        eng.sample_batch_dev(gm, None, d, [lev], None, 1, 0, n, m, mom.data_ptr(), mc_seed=1)
    eng.sync()
print("ok")
