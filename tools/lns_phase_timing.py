"""Debug: SM-clock cycles between the phase boundaries of the latency variant (CTA (0,0) of a 16 x n launch).
Build first: tools/build_variant.sh lnst -DMLE_LNS_TIMING ; run with MLE_CUDA_LIB=tools/bin/libmle_lnst.so"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mle_b200, mle_b200.kl as kl
eng = mle_b200.Engine(0)
m, R = 100, 16
lat = eng.lattice(None, m); d = eng.dists([(1, [0.0, 1.0])] * m); gm = eng.model("lognormal1d", 1, [16])
sh = np.random.default_rng(0).random((R, m))
names = ["prologue", "generator", "contraction", "exp", "reciprocals", "flux sums + combine", "partials", "finalize"]
for lev in (0, 3, 6):
    gm.set_basis(lev, kl.kl_basis_1d(lev, m=m))
    for n in (1, 8):
        for i in range(5): eng.sample_batch(gm, lat, d, [lev], sh, 0, n)
        t = (C.c_longlong * 16)()
        eng.lib.mle_debug_lns_timing(t)
        d_ = [t[i + 1] - t[i] for i in range(8)]
        print("level %d n=%d: " % (lev, n) + ", ".join("%s %d" % (names[i], d_[i]) for i in range(8)) + " | total %d cycles = %.2f us at 1.965 GHz" % (t[8] - t[0], (t[8] - t[0]) / 1965.0))
