"""Debug: marginal cost of each phase of lognormal1d_kernel (MLE_SKIP bitmask, honoured by debug builds only --
`make EXTRA=-DMLE_PHASE_TIMING`; results are invalid, only the times matter)."""
import os, sys, subprocess, json
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys; sys.path.insert(0, %r)
import numpy as np, torch, mle_b200, mle_b200.kl as kl
eng = mle_b200.Engine(0); m, R = 100, 16
lat = eng.lattice(None, m); d = eng.dists([(1, [0.0, 1.0])] * m); gm = eng.model("lognormal1d", 1, [16])
sh = torch.from_numpy(np.random.default_rng(0).random((R, m))).cuda(); mom = torch.zeros(R * 6, dtype=torch.float64, device="cuda")
st = torch.cuda.ExternalStream(eng.stream); out = []
for lev, n in ((0, 1 << 20), (2, 131072), (4, 16384), (6, 2048)):
    gm.set_basis(lev, kl.kl_basis_1d(lev, m=m))
    eng.sample_batch_dev(gm, lat, d, [lev], sh.data_ptr(), R, 0, n, m, mom.data_ptr()); eng.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(3): eng.sample_batch_dev(gm, lat, d, [lev], sh.data_ptr(), R, 0, n, m, mom.data_ptr())
    b.record(st); torch.cuda.synchronize(); out.append(round(a.elapsed_time(b) / 3, 3))
print(out)
''' % root
for name, mask in (("full", 0), ("no generator", 1), ("no contraction", 2), ("no exp", 4), ("no sweeps", 8), ("generator only", 14), ("no gen, no sweeps", 9)):
    env = dict(os.environ, MLE_SKIP=str(mask))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("%-20s levels 0/2/4/6 ms: %s" % (name, r.stdout.strip() or r.stderr[-200:]))
