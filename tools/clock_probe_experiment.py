"""Is the concurrent-fill speed-up a clock effect?  Build with `make EXTRA=-DMLE_PHASE_TIMING`: the kernel then adds up its
clock64() cycles per CTA; the same work costing the same cycles in less wall time means a higher SM clock."""
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
stream = torch.cuda.ExternalStream(eng.stream)
flush = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
gm.set_basis(0, kl.kl_basis_1d(0, m=m))
n = 1 << 20
f = lambda: eng.sample_batch_dev(gm, lat, d, [0], sh.data_ptr(), R, 0, n, m, mom.data_ptr())
f(); torch.cuda.synchronize()
out = (C.c_ulonglong * 8)()
for mode in ("alone", "fill after launch", "alone", "fill after launch"):
    for rep in range(4):
        lib.mle_debug_phase_cycles(out, 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); f(); b.record(stream)
        if mode == "fill after launch":
            flush.fill_(rep)
        torch.cuda.synchronize()
        lib.mle_debug_phase_cycles(out, 1)
        cyc = sum(list(out)[:6]) / 740.0
        ms = a.elapsed_time(b)
        print("%-18s %.3f ms  %.4g cycles per CTA  -> %.0f MHz effective" % (mode, ms, cyc, cyc / (ms * 1e-3) * 1e-6), flush=True)
