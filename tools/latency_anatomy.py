"""Where one small host-buffer mle_sample_batch call spends its time (the tiny batches of the QMC loop, src/run.jl:96-113):
   wall  = raw ctypes call, host buffers, median over 200 calls (no gpu_seconds: no event records)
   kern  = the same launch issued 200 times back to back through mle_sample_batch_dev, CUDA events on the library's stream:
           the kernel's own duration once it exceeds the launch rate
   So wall - kern = launch latency + completion flag over PCIe + host bookkeeping."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mle_b200, mle_b200.kl as kl
eng = mle_b200.Engine(0)
m, R = 100, 16
lat = eng.lattice(None, m); d = eng.dists([(1, [0.0, 1.0])] * m); gm = eng.model("lognormal1d", 1, [16])
sh = np.random.default_rng(0).random((R, m))
stream = torch.cuda.ExternalStream(eng.stream, device=0)
sh_dev = torch.from_numpy(sh).cuda(); mom_dev = torch.zeros(R * 6, dtype=torch.float64, device="cuda")
dp = C.POINTER(C.c_double)
levels = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 3, 6]
for lev in levels:
    gm.set_basis(lev, kl.kl_basis_1d(lev, m=m))
    for n in (1, 8, 32, 100):
        idx = np.array([lev], dtype=np.int32); mom = np.empty((R, 1, 2, 3))
        args = (eng.ctx, gm.handle, lat.handle, d.handle, idx.ctypes.data_as(C.POINTER(C.c_int32)), 1, sh.ctypes.data_as(dp), R, 0, n, m, 0, mom.ctypes.data_as(dp), None, None)
        for i in range(20): eng.lib.mle_sample_batch(*args)
        ts = []
        for i in range(200):
            t0 = time.perf_counter(); eng.lib.mle_sample_batch(*args); ts.append(time.perf_counter() - t0)
        wall = float(np.median(ts)) * 1e6
        with torch.cuda.stream(stream):
            for i in range(20): eng.sample_batch_dev(gm, lat, d, [lev], sh_dev.data_ptr(), R, 0, n, m, mom_dev.data_ptr())
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for i in range(200): eng.sample_batch_dev(gm, lat, d, [lev], sh_dev.data_ptr(), R, 0, n, m, mom_dev.data_ptr())
            b.record(stream)
        torch.cuda.synchronize()
        kern = a.elapsed_time(b) / 200 * 1e3
        print("level %d n=%3d per shift: wall %.1f us, kernel %.1f us back to back, overhead %.1f us" % (lev, n, wall, kern, wall - kern), flush=True)
