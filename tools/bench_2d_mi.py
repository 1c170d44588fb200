"""Throughput of the 2-D multi-index lognormal model per index (config C4 shape: AD(2)/QMC, 100 KL terms, 16 shifts)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mle_b200, mle_b200.kl as kl
eng = mle_b200.Engine(0)
m, R = 100, 16
lat = eng.lattice(None, m)
d = eng.dists([(1, [0.0, 1.0])] * m)
gm = eng.model("lognormal2d", 2, [4])
sh = torch.rand(R, m, dtype=torch.float64, device="cuda")
mom = torch.zeros(R * 6, dtype=torch.float64, device="cuda")
stream = torch.cuda.ExternalStream(eng.stream)
for idx in ((0, 0), (1, 0), (0, 1), (1, 1), (2, 0), (0, 2), (2, 1), (1, 2), (2, 2), (3, 1), (1, 3), (3, 3), (4, 0), (0, 4), (4, 2), (2, 4), (4, 4)):
    Nx, Ny = 4 << idx[0], 4 << idx[1]
    n = max(16, (1 << 22) // (Nx * Ny * max(Nx, 8)))
    gm.set_basis(idx, kl.kl_basis_2d(idx, m=m))
    f = lambda: eng.sample_batch_dev(gm, lat, d, list(idx), sh.data_ptr(), R, 0, n, m, mom.data_ptr())
    f(); eng.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); f(); b.record(stream); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print("index %s grid %3dx%-3d n=%7d x %d shifts  %8.3f ms  %10.0f samples/s" % (idx, Nx, Ny, n, R, ms, R * n / ms * 1e3), flush=True)
