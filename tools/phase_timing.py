"""Debug: per-phase wall cycles of lognormal1d_kernel (build with `make EXTRA=-DMLE_PHASE_TIMING`)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mle_b200, mle_b200.kl as kl
eng = mle_b200.Engine(0)
lib = eng.lib
m, R = 100, 16
lat = eng.lattice(None, m); d = eng.dists([(1, [0.0, 1.0])] * m); gm = eng.model("lognormal1d", 1, [16])
sh = torch.from_numpy(np.random.default_rng(0).random((R, m))).cuda()
mom = torch.zeros(R * 6, dtype=torch.float64, device="cuda")
names = ["gen", "gemm", "wait_sweep_done", "exp+barrier", "sweep", "tile_overhead"]
for lev, n in ((0, 1 << 20), (1, 370728), (2, 131072), (3, 46341), (4, 16384), (6, 2048)):
    gm.set_basis(lev, kl.kl_basis_1d(lev, m=m))
    out = (C.c_ulonglong * 8)()
    eng.sample_batch_dev(gm, lat, d, [lev], sh.data_ptr(), R, 0, n, m, mom.data_ptr()); eng.sync()
    lib.mle_debug_phase_cycles(out, 1)
    eng.sample_batch_dev(gm, lat, d, [lev], sh.data_ptr(), R, 0, n, m, mom.data_ptr()); eng.sync()
    lib.mle_debug_phase_cycles(out, 1)
    v = np.array(list(out)[:6], dtype=float); tot = v.sum()
    print("level", lev, "  ".join("%s %.1f%%" % (nm, 100 * x / tot) for nm, x in zip(names, v)), " cycles/CTA %.3g" % (tot / 736))
