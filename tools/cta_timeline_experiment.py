"""Per-CTA start / end times and SM ids of one lognormal1d launch (build with `make EXTRA=-DMLE_PHASE_TIMING`), alone and with a
16 MiB fill kernel started right after the launch on another stream."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mle_b200, mle_b200.kl as kl
eng = mle_b200.Engine(0)
lib = eng.lib
m, R = 100, 16
lat = eng.lattice(None, m); d = eng.dists([(1, [0.0, 1.0])] * m); gm = eng.model("lognormal1d", 1, [16])
sh = torch.rand(R, m, dtype=torch.float64, device="cuda")
mom = torch.zeros(R * 6, dtype=torch.float64, device="cuda")
flush = torch.empty(16 << 20, dtype=torch.uint8, device="cuda")
gm.set_basis(0, kl.kl_basis_1d(0, m=m))
n = 1 << 20
f = lambda: eng.sample_batch_dev(gm, lat, d, [0], sh.data_ptr(), R, 0, n, m, mom.data_ptr())
f(); torch.cuda.synchronize()
NC = 736
buf = (C.c_ulonglong * (3 * NC))()
for mode in ("alone", "fill", "fill", "alone"):
    f()
    if mode == "fill":
        flush.fill_(1)
    torch.cuda.synchronize()
    lib.mle_debug_cta_times(buf, NC)
    a = np.array(list(buf), dtype=np.float64).reshape(NC, 3)
    t0 = a[:, 0].min()
    start, end, sm = (a[:, 0] - t0) * 1e-6, (a[:, 1] - t0) * 1e-6, a[:, 2].astype(int)
    life = end - start
    per_sm = np.bincount(sm, minlength=148)
    print("%-5s  kernel span %.3f ms | CTA start: max %.3f ms | lifetime mean %.3f min %.3f max %.3f ms | end: min %.3f max %.3f | CTAs per SM: %s" % (
        mode, end.max(), start.max(), life.mean(), life.min(), life.max(), end.min(), end.max(),
        dict(zip(*np.unique(per_sm, return_counts=True)))))
    late = np.argsort(end)[-5:]
    print("       slowest CTAs: ", [(int(i), int(sm[i]), round(float(life[i]), 2), int(per_sm[sm[i]])) for i in late], "(cta, sm, lifetime ms, CTAs on that SM)")
    sm_end = np.array([end[sm == q].max() if (sm == q).any() else np.nan for q in range(148)])
    sm_mean = np.array([life[sm == q].mean() if (sm == q).any() else np.nan for q in range(148)])
    print("       per-SM finish time: min %.2f max %.2f ms; mean CTA lifetime on SMs 0-73: %.3f, on SMs 74-147: %.3f ms; spread WITHIN an SM (max-min lifetime), mean over SMs: %.3f ms" % (
        np.nanmin(sm_end), np.nanmax(sm_end), np.nanmean(sm_mean[:74]), np.nanmean(sm_mean[74:]),
        np.mean([life[sm == q].max() - life[sm == q].min() for q in range(148) if (sm == q).any()])))
    order = np.argsort(sm_end)
    print("       earliest-finishing SMs:", [(int(q), round(float(sm_end[q]), 2)) for q in order[:6]], " latest:", [(int(q), round(float(sm_end[q]), 2)) for q in order[-6:]])
    for k in np.unique(per_sm):
        sel = per_sm[sm] == k
        print("       CTAs on SMs holding %d CTAs: mean lifetime %.3f ms (n=%d)" % (k, life[sel].mean(), sel.sum()))
