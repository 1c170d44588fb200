"""Why did the C2 step run 4 % faster when the L2-flush memset of the NEXT bench iteration overlapped it?  Level launches alone
vs with a 256 MiB fill started right after them on another stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mle_b200, mle_b200.kl as kl
eng = mle_b200.Engine(0)
m, R = 100, 16
lat = eng.lattice(None, m); d = eng.dists([(1, [0.0, 1.0])] * m); gm = eng.model("lognormal1d", 1, [16])
sh = torch.rand(R, m, dtype=torch.float64, device="cuda")
mom = torch.zeros(R * 6, dtype=torch.float64, device="cuda")
stream = torch.cuda.ExternalStream(eng.stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
sizes = [int(x) for x in os.environ.get("FILL_MB", "256").split(",")]
for lev, n in ((0, 1 << 20), (1, 370728), (3, 46341)):
    gm.set_basis(lev, kl.kl_basis_1d(lev, m=m))
    f = lambda: eng.sample_batch_dev(gm, lat, d, [lev], sh.data_ptr(), R, 0, n, m, mom.data_ptr())
    f(); torch.cuda.synchronize()
    for fmb in sizes:
      fl = flush[:fmb << 20]
      for mode in ("alone", "fill after launch"):
        ts = []
        for rep in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); f(); b.record(stream)
            if mode == "fill after launch":
                fl.fill_(rep)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print("level %d  fill %4d MiB %-18s %s ms" % (lev, fmb, mode, [round(t, 3) for t in ts]), flush=True)
