"""Throughput of the 2-D lognormal model per level (config C3 shape: 400 KL terms, MC inputs, grid 4*2^l)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mle_b200, mle_b200.kl as kl
eng = mle_b200.Engine(0)
m = 400
d = eng.dists([(1, [0.0, 1.0])] * m)
gm = eng.model("lognormal2d", 1, [4])
mom = torch.zeros(6, dtype=torch.float64, device="cuda")
stream = torch.cuda.ExternalStream(eng.stream)
for lev, n in ((0, 1 << 18), (1, 1 << 17), (2, 1 << 16), (3, 1 << 14), (4, 1 << 11), (5, 1184)):
    gm.set_basis(lev, kl.kl_basis_2d(lev, m=m))
    eng.sample_batch_dev(gm, None, d, [lev], None, 1, 0, n, m, mom.data_ptr(), mc_seed=1); eng.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    eng.sample_batch_dev(gm, None, d, [lev], None, 1, 0, n, m, mom.data_ptr(), mc_seed=1)
    b.record(stream); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    N = 4 << lev
    flops = 2.0 * (N + 1) ** 2 * m + 2.0 * ((N - 1) ** 3 * (N / 2 - 1) * 2 + (N - 1) ** 3) * (1 + (0.125 if lev else 0))
    print("level %d grid %3dx%-3d n=%7d  %8.3f ms  %10.0f samples/s  ~%.2f TFLOP/s (contraction + full-window updates)" % (lev, N, N, n, ms, n / ms * 1e3, flops * n / ms * 1e-9))
