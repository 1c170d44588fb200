"""Latency of one small host-buffer mle_sample_batch call on the 2-D models (C4 shape: multi-index, 100 KL terms, 16 shifts;
C3 shape: multilevel MC, 400 KL terms): raw ctypes call, median of 100."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mle_b200, mle_b200.kl as kl
eng = mle_b200.Engine(0)
dp = C.POINTER(C.c_double)
def call(gm, lat, d, idx, sh, R, n, m, seed=0):
    ia = np.array(idx, dtype=np.int32); mom = np.empty((R, 1, 2, 3))
    args = (eng.ctx, gm.handle, lat.handle if lat else None, d.handle, ia.ctypes.data_as(C.POINTER(C.c_int32)), len(idx),
            sh.ctypes.data_as(dp) if sh is not None else None, R, 0, n, m, seed, mom.ctypes.data_as(dp), None, None)
    for i in range(10): eng.lib.mle_sample_batch(*args)
    ts = []
    for i in range(100):
        t0 = time.perf_counter(); rc = eng.lib.mle_sample_batch(*args); ts.append(time.perf_counter() - t0)
    assert rc == 0
    return float(np.median(ts)) * 1e6
m, R = 100, 16
lat = eng.lattice(None, m); d = eng.dists([(1, [0.0, 1.0])] * m); gm = eng.model("lognormal2d", 2, [4])
sh = np.random.default_rng(0).random((R, m))
for idx in ((0, 0), (1, 1), (2, 1), (2, 2), (3, 1), (3, 3), (4, 2), (4, 4)):
    gm.set_basis(idx, kl.kl_basis_2d(idx, m=m))
    print("C4 index %s grid %3dx%-3d: " % (idx, 4 << idx[0], 4 << idx[1]) + ", ".join("n=%d: %.1f us" % (n, call(gm, lat, d, idx, sh, R, n, m)) for n in (1, 8, 32)), flush=True)
m3 = 400
d3 = eng.dists([(1, [0.0, 1.0])] * m3); g3 = eng.model("lognormal2d", 1, [4])
for lev in range(0, 6):
    g3.set_basis(lev, kl.kl_basis_2d(lev, m=m3))
    print("C3 level %d grid %3dx%-3d (MC): " % (lev, 4 << lev, 4 << lev) + ", ".join("n=%d: %.1f us" % (n, call(g3, None, d3, [lev], None, 1, n, m3, seed=5)) for n in (4, 16, 64)), flush=True)
