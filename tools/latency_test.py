"""Latency of one mle_sample_batch call (host buffers) for the tiny batches of the QMC loop (src/run.jl:96-113)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mle_b200, mle_b200.kl as kl
eng = mle_b200.Engine(0)
m, R = 100, 16
lat = eng.lattice(None, m); d = eng.dists([(1, [0.0, 1.0])] * m); gm = eng.model("lognormal1d", 1, [16])
sh = np.random.default_rng(0).random((R, m))
for lev in (0, 3, 6):
    gm.set_basis(lev, kl.kl_basis_1d(lev, m=m))
    for n in (1, 2, 12, 100, 1000):
        eng.sample_batch(gm, lat, d, [lev], sh, 0, n)
        t0 = time.perf_counter(); reps = 50
        for i in range(reps): mom = eng.sample_batch(gm, lat, d, [lev], sh, i * n, n)
        dt = (time.perf_counter() - t0) / reps
        mom, tg = eng.sample_batch(gm, lat, d, [lev], sh, 0, n, want_time=True)   # device time of the two kernels (CUDA events)
        print("level %d n=%4d per shift: %.1f us per call (GPU %.1f us), %.0f samples/s" % (lev, n, dt * 1e6, tg * 1e6, R * n / dt))
