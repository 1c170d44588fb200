import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mle_b200, mle_b200.kl as kl
eng = mle_b200.Engine(0)
m, R = 100, 16
lat = eng.lattice(None, m); d = eng.dists([(1, [0.0, 1.0])] * m); gm = eng.model("lognormal1d", 1, [16])
sh = np.random.default_rng(0).random((R, m))
gm.set_basis(0, kl.kl_basis_1d(0, m=m))
dp = C.POINTER(C.c_double)
idx = np.array([0], dtype=np.int32); mom = np.empty((R, 1, 2, 3)); t = C.c_double(0)
args = (eng.ctx, gm.handle, lat.handle, d.handle, idx.ctypes.data_as(C.POINTER(C.c_int32)), 1, sh.ctypes.data_as(dp), R, 0, 1, m, 0, mom.ctypes.data_as(dp), None, C.byref(t))
for rep in range(3):
    t0 = time.perf_counter()
    for i in range(200): eng.lib.mle_sample_batch(*args)
    dt = (time.perf_counter() - t0) / 200
    t0 = time.perf_counter()
    for i in range(200): eng.sample_batch(gm, lat, d, [0], sh, 0, 1, want_time=True)
    dt2 = (time.perf_counter() - t0) / 200
    args2 = args[:-1] + (None,)
    t0 = time.perf_counter()
    for i in range(200): eng.lib.mle_sample_batch(*args2)
    dt3 = (time.perf_counter() - t0) / 200
    print("raw ctypes %.1f us, Engine.sample_batch %.1f us, raw without gpu_seconds %.1f us (GPU %.1f us)" % (dt * 1e6, dt2 * 1e6, dt3 * 1e6, t.value * 1e6))
