#!/usr/bin/env python3
"""bench.py -- fine+coarse samples/sec of the sampling hot path on N B200s, with roofline and CPU baseline.

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): MLQMC on the 1-D lognormal diffusion problem -- rank-1 lattice
(first 100 entries of K_3600_32), R = 16 random shifts, 100 Normal(0,1) KL terms, levels 0..6 (grid 16 * 2^l cells).
One STEP = one `parallel_sample!`-sized batch on every level l = 0..6 with n_l = 2^20 * 2^(-1.5 l) points per shift
(the N_l ~ sqrt(V_l / C_l) allocation of an MLQMC run with beta = 2, gamma = 1), i.e. 25.9 M fine+coarse samples.
Every sample runs the whole path: lattice point -> shift -> Phi^-1 -> KL field -> exp -> fine and coarse tridiagonal
solves -> (dQ, Qf) -> per-(level, shift) moments.  Multi-GPU: each rank owns its own contiguous slice of point numbers
for every shift (weak scaling: n_l per rank fixed) and one NCCL all-gather per step combines the moment triplets,
merged in rank order with Chan's formula.

Extra keys of the JSON line: `strong` (the same step with every level's points SPLIT over the ranks: fixed total
work), `extra.c3` (config C3: MLMC on the 2-D lognormal model, 400 KL terms, levels 0-5, point slices per rank + merge),
`extra.c4` (config C4: multi-index QMC batches on the 2-D anisotropic model over the indices of TD(2) up to size 3),
`extra.c5` (config C5: 2^24 lattice points x 1000 dims, fused generate -> integrand -> moments), `run_e2e` (N = 1: wall time of
run(estimator, tol) on the GPU backend vs the CPU oracle backend, tools/run_e2e.py; `run_e2e.two_d`: the same for whole runs of
the 2-D configurations C4 and C3).
Launched WITHOUT torchrun and --gpus N > 1 the bench runs single-process on ONE multi-GPU context (mle_init_multi).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_KL = 100
R_SHIFTS = 16
LEVELS = list(range(7))
N0_POINTS = 1 << 20
SHIFT_SEED = 2024


_JSON_FD = None


def emit(line):
    """the one JSON line, on the process's real stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def points_per_shift(level, scale=1.0):
    return max(1, int(round(N0_POINTS * scale * 2.0 ** (-1.5 * level))))


def splitmix_uniforms(seed, count):
    M = (1 << 64) - 1
    s = int(seed)
    out = np.empty(count)
    for i in range(count):
        s = (s + 0x9E3779B97F4A7C15) & M
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        z ^= z >> 31
        out[i] = (z >> 11) * 2.0 ** -53
    return out


def level_shifts():
    """independent shift vectors per (level, shift) like the reference (src/internals.jl:138)"""
    u = splitmix_uniforms(SHIFT_SEED, len(LEVELS) * R_SHIFTS * M_KL)
    return u.reshape(len(LEVELS), R_SHIFTS, M_KL)


def algorithmic_flops_per_sample(level):
    """FP64 operations of the reference algorithm per sample (add/mul/div/sqrt = 1, fused multiply-add = 2), DESIGN.md"""
    N = 16 << level
    gen = M_KL * (3 + 0.75 * 31 + 0.1875 * 35 + 0.0625 * 72)
    kl = 2.0 * (N + 1) * M_KL
    ex = 34.0 * (N + 1)
    solve = 6.0 * N + (6.0 * (N // 2) if level > 0 else 0.0) + 8  # per edge: mean, 1/kappa, A += r, D = fma(w, r, D)
    return gen + kl + ex + solve


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.lines, self.index = None, [], str(index)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", self.index], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def fp64_peak_tflops():
    """live DFMA / DMMA pipe peak from tools/fp64_microbench (MEASURED_PEAKS.json carries no FP64 figure)"""
    exe = os.path.join(ROOT, "tools", "bin", "fp64_microbench")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout.strip().splitlines()[-1]
        return json.loads(out)
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        return None


# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_run(scale, threads_note=True):
    """the oracle run the reference's way (scalar per-sample path, OpenMP over the (shift, k) product) on every level;
    returns (samples, seconds, cores)"""
    from oracle import oracle as O
    import mle_b200.kl as kl
    O.build()
    z = O.genvec("K_3600_32", M_KL)
    dl = [(O.NORMAL, [0.0, 1.0])] * M_KL
    om = O.Model(O.MODEL_LOGNORMAL1D, 1, [16.0])
    for lev in LEVELS:
        om.set_basis(lev, kl.kl_basis_1d(lev, m=M_KL))
    sh = level_shifts()
    O.sample_batch(om, z, dl, [0], sh[0], 0, 256)  # warm the thread pool
    total, t0 = 0, time.perf_counter()
    for lev in LEVELS:
        n = points_per_shift(lev, scale)
        O.sample_batch(om, z, dl, [lev], sh[lev], 0, n)
        total += n * R_SHIFTS
    return total, time.perf_counter() - t0, O.num_threads()


def run_reference(args):
    """--impl reference: the reference's CPU path.  Julia is not installed here or on the GPU box and the reference is pure
    Julia (no C sources to compile into oracle/_ref), so this arm times the C oracle port on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the reference arm is entitled to every host core
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    scale = 1.0 / 16
    for _ in range(args.warmup):
        cpu_reference_run(scale / 8)
    times, samples, cores = [], 0, 1
    for _ in range(args.steps):
        samples, dt, cores = cpu_reference_run(scale)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = samples / (ms * 1e-3)
    sample = "every level 0..6 at 1/16 of the bench step's points per shift (%d samples per step)" % samples
    line = {"impl": "reference", "metric": "fine+coarse samples/sec", "value": value, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(),
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config():
    return {"workload": "C2: MLQMC 1-D lognormal diffusion, rank-1 lattice K_3600_32[:100], 16 random shifts, 100-term KL, "
                        "levels 0-6 (grid 16*2^l), n_l = 2^20 * 2^(-1.5 l) points per shift per GPU",
            "levels": LEVELS, "shifts": R_SHIFTS, "kl_terms": M_KL,
            "points_per_shift": [points_per_shift(lev) for lev in LEVELS],
            "samples_per_step_per_gpu": sum(points_per_shift(lev) for lev in LEVELS) * R_SHIFTS,
            "parallelism": "point ranges sharded over ranks, one NCCL all-gather of the moment triplets per step",
            "l2": "inputs are a few hundred KB by construction (the KL basis is re-read from L2 by design); a 256 MiB "
                  "buffer is rewritten between steps, outside the per-step event pairs"}



def _ncu_rows():
    """rows of the committed ncu summary of this command (tools/summarize_profiles.py -> profiles/r02_final_kernels.csv)"""
    import csv
    path = os.path.join(ROOT, "profiles", "r02_final_kernels.csv")
    try:
        rows = list(csv.DictReader(open(path)))
    except Exception:  # noqa: BLE001
        return None
    return [r for r in rows if "lognormal1d" in r.get("Kernel Name", "")][:len(LEVELS)] or None


def ncu_pipe_busy():
    """FP64-pipe + DMMA-pipe busy share per level launch (one physical pipe on B200), from the committed ncu summary"""
    rows = _ncu_rows()
    if not rows:
        return None
    try:
        return [round(float(r["pipe_fp64_plus_dmma_pct"]) / 100.0, 3) for r in rows]
    except Exception:  # noqa: BLE001
        return None


def ncu_traffic():
    rows = _ncu_rows()
    if not rows:
        return None
    try:
        return sum(float(r["dram_bytes"]) for r in rows)
    except Exception:  # noqa: BLE001
        return None


# ---------------------------------------------------------------------------------------------------------------
# configs C3 and C5 (BASELINE.json configs[2], configs[4]) at the --gpus the driver passes
# ---------------------------------------------------------------------------------------------------------------
C3_M, C3_LEVELS = 400, list(range(6))
C3_POINTS = [1 << 20, 1 << 18, 1 << 16, 1 << 13, 1 << 11, 1184]   # MC samples per level and rank (1184 = 148 SMs x 8 samples per CTA)
C5_DIMS, C5_LOG2_POINTS = 1000, 24


def c3_flops_per_sample(level):
    """2-D model: field 2 (N+1)^2 m + exp 34 (N+1)^2 + frontal elimination 2 N^4 (full-window rank-1 updates), coarse solve 1/16"""
    N = 4 << level
    nodes = (N + 1) ** 2
    return 2.0 * nodes * C3_M + 34.0 * nodes + 2.0 * float(N) ** 4 * (1.0 + (1.0 / 16 if level else 0.0))


def bench_extras(eng, torch, dist, stream, world, rank):
    import mle_b200
    import mle_b200.kl as kl
    out = {}
    def tmax(ms):
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    # ---- C3: ML()/MC() on the 2-D lognormal diffusion, 400 KL terms, levels 0-5; rank r takes points [r n, (r+1) n) ---------
    try:
        d3 = eng.dists([(mle_b200.NORMAL, [0.0, 1.0])] * C3_M)
        g3 = eng.model("lognormal2d", 1, [4])
        for lev in C3_LEVELS:
            g3.set_basis(lev, kl.kl_basis_2d(lev, m=C3_M))
        mom = torch.zeros(len(C3_LEVELS), 6, dtype=torch.float64, device="cuda")
        gat = torch.zeros(world, mom.numel(), dtype=torch.float64, device="cuda") if world > 1 else None

        def step_c3():
            for li, lev in enumerate(C3_LEVELS):
                eng.sample_batch_dev(g3, None, d3, [lev], None, 1, rank * C3_POINTS[li], C3_POINTS[li], C3_M, mom[li].data_ptr(),
                                     mc_seed=1000 + lev)
            if world > 1:
                with torch.cuda.stream(stream):
                    dist.all_gather_into_tensor(gat.view(-1), mom.view(-1))
        step_c3(); torch.cuda.synchronize()
        lvl = []
        for li, lev in enumerate(C3_LEVELS):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            eng.sample_batch_dev(g3, None, d3, [lev], None, 1, rank * C3_POINTS[li], C3_POINTS[li], C3_M, mom[li].data_ptr(), mc_seed=1000 + lev)
            b.record(stream); torch.cuda.synchronize()
            lvl.append(a.elapsed_time(b))
        if world > 1:
            dist.barrier()
        reps = 3
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(stream)
        for _ in range(reps):
            step_c3()
        b.record(stream); torch.cuda.synchronize()
        ms = tmax(a.elapsed_time(b) / reps)
        per_rank = sum(C3_POINTS)
        flops = sum(c3_flops_per_sample(lev) * C3_POINTS[li] for li, lev in enumerate(C3_LEVELS))
        out["c3"] = {"workload": "C3: ML()/MC() 2-D lognormal diffusion, 400-term KL, levels 0-5 (grid 4*2^l squared), MC points per level "
                                 "and rank %s, ranks take disjoint point slices, one all-gather of the moments" % C3_POINTS,
                     "value": per_rank * world / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "scaling": "weak",
                     "per_level_ms_rank0": lvl,
                     "per_level_samples_per_s_rank0": [C3_POINTS[i] / (lvl[i] * 1e-3) for i in range(len(lvl))],
                     "roofline": {"bound": "tensor", "unit": "TFLOP/s", "achieved": flops / (sum(lvl) * 1e-3) * 1e-12,
                                  "note": "algorithmic FP64 operations (field 2(N+1)^2 m, exp, frontal elimination 2N^4) / kernel "
                                          "time on rank 0; peak = the FP64 roof of `roofline.peak`"}}
        g3.close()
    except Exception as e:  # noqa: BLE001
        out["c3"] = {"error": repr(e)}
    # ---- C4: multi-index QMC batches on the 2-D anisotropic model (BASELINE configs[3]: AD(2)/QMC), the indices of TD(2) up to ----
    # size 3 as an adaptive run visits them, 100 KL terms, 16 shifts; rank r takes points [r n, (r+1) n) of every index
    try:
        C4_M, C4_R = 100, 16
        idxs = [(lx, ly) for s_ in range(4) for lx in range(s_ + 1) for ly in [s_ - lx]]
        z4 = eng.genvec("K_3600_32")[:C4_M]
        lat4 = eng.lattice(z4, C4_M)
        d4 = eng.dists([(mle_b200.NORMAL, [0.0, 1.0])] * C4_M)
        g4 = eng.model("lognormal2d", 2, [4])
        for idx in idxs:
            g4.set_basis(idx, kl.kl_basis_2d(idx, m=C4_M))
        n4 = [max(16, int(8192 * 2.0 ** (-1.5 * (lx + ly)))) for lx, ly in idxs]   # points per shift: N ~ 2^(-1.5 |index|)
        sh4 = torch.from_numpy(splitmix_uniforms(4242, len(idxs) * C4_R * C4_M).reshape(len(idxs), C4_R, C4_M)).cuda()
        mom4 = torch.zeros(len(idxs), C4_R * 6, dtype=torch.float64, device="cuda")

        def step_c4():
            for i, idx in enumerate(idxs):
                eng.sample_batch_dev(g4, lat4, d4, list(idx), sh4[i].data_ptr(), C4_R, rank * n4[i], n4[i], C4_M, mom4[i].data_ptr())
        step_c4(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        a.record(stream)
        for _ in range(reps):
            step_c4()
        b.record(stream); torch.cuda.synchronize()
        ms4 = tmax(a.elapsed_time(b) / reps)
        tot4 = sum(n4) * C4_R
        out["c4"] = {"workload": "C4: multi-index QMC batches on the 2-D anisotropic lognormal model (mixed differences over diff(index)), "
                                 "indices of TD(2) up to size 3 (grids 4*2^lx x 4*2^ly), 100 KL terms, 16 shifts, points per shift %s per rank" % n4,
                     "value": tot4 * world / (ms4 * 1e-3), "unit": "samples/s", "ms_per_step": ms4, "scaling": "weak",
                     "indices": [list(i) for i in idxs]}
        g4.close(); lat4.close(); d4.close()
    except Exception as e:  # noqa: BLE001
        out["c4"] = {"error": repr(e)}
    # ---- C5: 2^24 lattice points x 1000 dims, 16 shifts, fused generate -> exp(sum a_j xi_j) -> moments ------------------------
    try:
        s5 = C5_DIMS
        z5 = eng.genvec("K_3600_32")[:s5]
        l5 = eng.lattice(z5, s5)
        d5 = eng.dists([(mle_b200.NORMAL, [0.0, 1.0])] * s5)
        g5 = eng.model("expsum", 1, [float(s5)] + [0.5 / (j + 1) for j in range(s5)])
        sh5 = torch.from_numpy(splitmix_uniforms(2024, R_SHIFTS * s5).reshape(R_SHIFTS, s5)).cuda()
        n5 = (1 << C5_LOG2_POINTS) // R_SHIFTS   # points per shift and rank
        mom5 = torch.zeros(R_SHIFTS * 6, dtype=torch.float64, device="cuda")
        for _ in range(2):
            eng.sample_batch_dev(g5, l5, d5, [0], sh5.data_ptr(), R_SHIFTS, rank * n5, n5, s5, mom5.data_ptr())
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(3):
            eng.sample_batch_dev(g5, l5, d5, [0], sh5.data_ptr(), R_SHIFTS, rank * n5, n5, s5, mom5.data_ptr())
        b.record(stream); torch.cuda.synchronize()
        ms = tmax(a.elapsed_time(b) / 3)
        coords = float(n5) * R_SHIFTS * s5
        est = mom5.cpu().numpy().reshape(R_SHIFTS, 2, 3)[:, 1, 1]
        exact = float(np.exp(0.5 * sum((0.5 / (j + 1)) ** 2 for j in range(s5))))
        fp64_per_coord = 40.0   # FP64-pipe instructions per coordinate of the generator (SASS count)
        ceiling = 148 * 64 * 1.965e9 / fp64_per_coord
        out["c5"] = {"workload": "C5: 2^%d lattice points x %d dims per rank, 16 shifts, Normal inverse CDF, integrand exp(sum a_j xi_j), "
                                 "shift-averaged moments; fused kernel, 0 HBM bytes per coordinate" % (C5_LOG2_POINTS, s5),
                     "value": coords * world / (ms * 1e-3), "unit": "coordinates/s", "ms": ms, "scaling": "weak",
                     "estimate": float(est.mean()), "closed_form": exact,
                     "std_errors_off": float(abs(est.mean() - exact) / (est.std(ddof=1) / np.sqrt(R_SHIFTS))),
                     "roofline": {"bound": "min(hbm, fp64)", "achieved": coords / (ms * 1e-3), "peak": ceiling, "unit": "coordinates/s",
                                  "frac": coords / (ms * 1e-3) / ceiling,
                                  "note": "the fused sweep writes nothing per coordinate, so HBM does not bound it; the FP64 pipe does: "
                                          "148 SMs x 64 lanes x 1.965 GHz / 40 FP64 instructions per coordinate"}}
        g5.close()
    except Exception as e:  # noqa: BLE001
        out["c5"] = {"error": repr(e)}
    return out

# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import mle_b200
    import mle_b200.kl as kl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmle_cuda has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = mle_b200.Engine(local)
    stream = torch.cuda.ExternalStream(eng.stream, device=local)
    z = eng.genvec("K_3600_32")[:M_KL]
    lat = eng.lattice(z, M_KL)
    dists = eng.dists([(mle_b200.NORMAL, [0.0, 1.0])] * M_KL)
    model = eng.model("lognormal1d", 1, [16])
    for lev in LEVELS:
        model.set_basis(lev, kl.kl_basis_1d(lev, m=M_KL))
    sh_host = level_shifts()
    n_l = [points_per_shift(lev) for lev in LEVELS]
    samples_per_step = sum(n_l) * R_SHIFTS  # per rank
    nmom = R_SHIFTS * 2 * 3

    # ---- device-resident arm ("value"): shifts already in HBM, moments stay in HBM -------------------------------------
    sh_dev = torch.from_numpy(sh_host).cuda()
    mom_dev = torch.zeros(len(LEVELS), R_SHIFTS, 1, 2, 3, dtype=torch.float64, device="cuda")
    gathered = torch.zeros(world, mom_dev.numel(), dtype=torch.float64, device="cuda") if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step_device():
        for li, lev in enumerate(LEVELS):
            k0 = rank * n_l[li]  # this rank's slice of point numbers
            eng.sample_batch_dev(model, lat, dists, [lev], sh_dev[li].data_ptr(), R_SHIFTS, k0, n_l[li], M_KL,
                                 mom_dev[li].data_ptr())
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(gathered.view(-1), mom_dev.view(-1))

    def merged_moments():
        if world == 1:
            return mom_dev.cpu().numpy()
        g = gathered.cpu().numpy().reshape(world, len(LEVELS), R_SHIFTS, 1, 2, 3)
        acc = g[0].copy()
        for r in range(1, world):
            acc = mle_b200.moments_merge(acc, g[r])
        return acc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        # the previous step has finished before the flush starts: a fill kernel that overlaps a step (it did in an earlier
        # version of this loop: the next iteration's flush was enqueued while the step's kernels were still running) makes the
        # FP64-bound kernels 7-11 % FASTER -- tools/concurrent_fill_experiment.py, profiles/r01_concurrent_fill_experiment.txt --
        # which is not the steady state of back-to-back sample batches
        torch.cuda.synchronize()
        flush.fill_(i & 0xFF)     # evict L2 between steps (torch's stream)
        torch.cuda.synchronize()
        ev[i][0].record(stream)
        step_device()
        ev[i][1].record(stream)
    barrier()
    launches = eng.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    ms_local = sum(ms_steps) / len(ms_steps)
    ms_t = torch.tensor([ms_local], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_step = float(ms_t.item())
    value = samples_per_step * world / (ms_step * 1e-3)
    final = merged_moments()

    # ---- per-kernel timing of the dominant kernel (level launches), CUDA events on the launching stream ----------------
    lvl_ms = []
    for li, lev in enumerate(LEVELS):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        a.record(stream)
        for _ in range(reps):
            eng.sample_batch_dev(model, lat, dists, [lev], sh_dev[li].data_ptr(), R_SHIFTS, rank * n_l[li], n_l[li], M_KL,
                                 mom_dev[li].data_ptr())
        b.record(stream)
        torch.cuda.synchronize()
        lvl_ms.append(a.elapsed_time(b) / reps)

    # ---- end-to-end arm through the host-buffer C-ABI call (H2D of the shifts + D2H of the moments every call) -----------
    # two shift sets used in turn: the library skips the upload of shifts it already holds (an adaptive run keeps one shift
    # vector per (index, shift) for its whole life, src/internals.jl:138), and this arm must pay the H2D copy every step
    sh_sets = [torch.from_numpy(sh_host.copy()).pin_memory().numpy(),
               torch.from_numpy(np.ascontiguousarray(1.0 - sh_host)).pin_memory().numpy()]
    e2e_calls = [0]

    def step_e2e():
        # one `update_samples` call (src/sample.jl:8-13): the seven levels' batches, host buffers in and out, one synchronisation
        sh_np = sh_sets[e2e_calls[0] & 1]
        e2e_calls[0] += 1
        out = eng.update_samples(model, lat, dists, [([lev], sh_np[li], rank * n_l[li], n_l[li], M_KL)
                                                     for li, lev in enumerate(LEVELS)])
        mom = np.stack(out)
        if world > 1:
            t = torch.from_numpy(mom).cuda()
            g = torch.empty(world, t.numel(), dtype=torch.float64, device="cuda")
            dist.all_gather_into_tensor(g.view(-1), t.view(-1))
            g = g.cpu().numpy().reshape((world,) + mom.shape)
            acc = g[0].copy()
            for r in range(1, world):
                acc = mle_b200.moments_merge(acc, g[r])
            mom = acc
        return mom

    step_e2e()
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = 2 * max(2, args.steps // 4)  # even: ends on the shift set the device-resident arm used
    for _ in range(e2e_steps):
        mom_e2e = step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = samples_per_step * world / float(e2e_t.item())
    del mom_e2e
    mom_chk = eng.update_samples(model, lat, dists, [([lev], sh_sets[0][li], rank * n_l[li], n_l[li], M_KL) for li, lev in enumerate(LEVELS)])
    mom_chk = np.stack(mom_chk)
    if world == 1:
        assert np.allclose(mom_chk[..., 1], final[..., 1], rtol=1e-12), "device-resident and host-buffer arms disagree"

    # ---- strong scaling: the same step with every level's n_l points SPLIT over the ranks (fixed total work) ------------------
    strong = None
    if world > 1:
        def step_strong():
            for li, lev in enumerate(LEVELS):
                base, rem = divmod(n_l[li], world)
                cnt = base + (1 if rank < rem else 0)
                k0 = rank * base + min(rank, rem)
                eng.sample_batch_dev(model, lat, dists, [lev], sh_dev[li].data_ptr(), R_SHIFTS, k0, max(cnt, 1), M_KL, mom_dev[li].data_ptr())
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(gathered.view(-1), mom_dev.view(-1))
        for _ in range(3):
            step_strong()
        barrier()
        sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i in range(args.steps):
            torch.cuda.synchronize()
            flush.fill_(i & 0xFF)
            torch.cuda.synchronize()
            sev[i][0].record(stream)
            step_strong()
            sev[i][1].record(stream)
        barrier()
        st = torch.tensor([sum(a.elapsed_time(b) for a, b in sev) / len(sev)], dtype=torch.float64, device="cuda")
        dist.all_reduce(st, op=dist.ReduceOp.MAX)
        strong = {"scaling": "strong", "value": samples_per_step / (float(st.item()) * 1e-3), "unit": "samples/s",
                  "ms_per_step": float(st.item()), "samples_per_step_total": samples_per_step,
                  "note": "every level's n_l points per shift split over the ranks (contiguous slices), one all-gather of the moments"}

    extra = bench_extras(eng, torch, dist, stream, world, rank)

    if rank == 0:
        peaks = measured_peaks()
        fp64 = fp64_peak_tflops()  # a subprocess on device 0 = rank 0's GPU, after the timed regions
        flops_step = sum(algorithmic_flops_per_sample(lev) * n_l[li] * R_SHIFTS for li, lev in enumerate(LEVELS))
        kernel_ms = sum(lvl_ms)
        achieved_tf = flops_step / (kernel_ms * 1e-3) * 1e-12
        peak_tf = fp64.get("dmma_m8n8k4_tflops")
        roofline = {"kernel": "lognormal1d_kernel (7 launches per step, one per level)", "bound": "tensor",
                    "pipe": "FP64: on B200 DMMA (mma.sync f64) and DFMA share one pipe (measured: 37.1 / 33.9 / 36.7 TFLOP/s "
                            "alone / alone / mixed), so the FP64 tensor figure is the roof of the whole kernel",
                    "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": (achieved_tf / peak_tf) if peak_tf else None,
                    "peak_source": "measured live by tools/fp64_microbench (MEASURED_PEAKS.json has no FP64 figure)",
                    # dram__bytes_read.sum + dram__bytes_write.sum of the 7 launches of one step, from the ncu --set full capture
                    # of this command (profiles/r02_final_kernels.csv)
                    "traffic": ncu_traffic(), "traffic_unit": "bytes per step (all 7 launches); algorithmic HBM bytes per sample: 0",
                    # FP64 + DMMA pipe-busy fractions of the 7 launches, read from the committed ncu summary of this command
                    # (profiles/r02_final_kernels.csv: sm__inst_executed_pipe_fp64 + sm__pipe_tensor_cycles_active); None when absent
                    "pipe_busy_ncu_per_level": ncu_pipe_busy(),
                    # the same, weighted by this run's measured launch times: the share of the step during which the FP64 pipe
                    # (the roof of these kernels) was busy -- `frac` counts algorithmic FLOPs against a roof only a pure FMA
                    # stream reaches (at 100 % pipe utilisation `frac` would be 0.65)
                    "pipe_busy_time_weighted": (lambda pb: None if not pb or len(pb) != len(lvl_ms) else
                                                round(sum(p * t for p, t in zip(pb, lvl_ms)) / sum(lvl_ms), 4))(ncu_pipe_busy()),
                    "per_level_ms": lvl_ms,
                    "per_level_samples_per_s": [n_l[li] * R_SHIFTS / (lvl_ms[li] * 1e-3) for li in range(len(LEVELS))],
                    "per_level_tflops": [algorithmic_flops_per_sample(lev) * n_l[li] * R_SHIFTS / (lvl_ms[li] * 1e-3) * 1e-12
                                         for li, lev in enumerate(LEVELS)]}
        # the HBM-bound stage on its own: points kernel (lattice -> shift -> Phi^-1 -> 8 B/coordinate stored)
        hbm = None
        if world == 1:
            npts = 1 << 18
            out = torch.empty(R_SHIFTS * npts * M_KL, dtype=torch.float64, device="cuda")
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(2):
                eng.points_dev(lat, dists, sh_dev[0].data_ptr(), R_SHIFTS, 0, npts, M_KL, out.data_ptr())
            torch.cuda.synchronize()
            a.record(stream)
            for _ in range(5):
                eng.points_dev(lat, dists, sh_dev[0].data_ptr(), R_SHIFTS, 0, npts, M_KL, out.data_ptr())
            b.record(stream)
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            gbs = out.numel() * 8 / (ms * 1e-3) * 1e-9
            hbm_peak = peaks["hbm_gbs"] if peaks else 6650.0
            hbm = {"kernel": "points_kernel (2^18 points x 100 dims x 16 shifts = 3.36 GB, larger than L2)", "bound": "hbm",
                   "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                   "peak_source": "MEASURED_PEAKS.json (measured copy)" if peaks else "fallback 6650 GB/s",
                   "algorithmic_bytes_per_coordinate": 8, "traffic": None,
                   "coordinates_per_s": out.numel() / (ms * 1e-3),
                   "note": "with the Normal inverse CDF (about 45 FP64 instructions per coordinate) this stage is bound by the "
                           "FP64 pipe, not by HBM: 8 B x (FP64 issue rate / 45) = 3.3 TB/s is its ceiling; `raw` below is "
                           "the same kernel without the transform (lattice point + shift only), which IS store-bound"}
            # the same kernel without the inverse CDF: shifted lattice points only (stage 1 alone)
            for _ in range(2):
                eng.points_dev(lat, None, sh_dev[0].data_ptr(), R_SHIFTS, 0, npts, M_KL, out.data_ptr())
            torch.cuda.synchronize()
            a.record(stream)
            for _ in range(5):
                eng.points_dev(lat, None, sh_dev[0].data_ptr(), R_SHIFTS, 0, npts, M_KL, out.data_ptr())
            b.record(stream)
            torch.cuda.synchronize()
            ms_raw = a.elapsed_time(b) / 5
            gbs_raw = out.numel() * 8 / (ms_raw * 1e-3) * 1e-9
            hbm["raw"] = {"achieved": gbs_raw, "frac": gbs_raw / hbm_peak, "unit": "GB/s",
                          "coordinates_per_s": out.numel() / (ms_raw * 1e-3)}
            del out
        run_e2e = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import run_e2e as RE
                run_e2e = RE.measure(1, tols=(1e-4, 2e-5, 5e-6, 1e-6), reps=3)
                # whole runs of the 2-D configurations (C4: AD(2)/QMC multi-index, C3: ML/MC), same two arms
                run_e2e["two_d"] = [RE.measure_2d("c4", 1), RE.measure_2d("c3", 1)]
            except Exception as e:  # noqa: BLE001
                run_e2e = {"error": repr(e)}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            total, dt, cores = cpu_reference_run(0.5)
            cpu = {"value": total / dt, "unit": "samples/s", "cores": cores, "kind": "port",
                   "sample": "one pass over levels 0..6 at 1/2 of the step's points per shift (%d samples, %.1f s)" % (total, dt)}
        line = {"metric": "fine+coarse samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(sh_host.nbytes),
                        "d2h_bytes_per_step": int(len(LEVELS) * nmom * 8), "ms_per_step": float(e2e_t.item()) * 1e3},
                "gpu_launches": int(launches), "strong": strong, "extra": extra, "run_e2e": run_e2e, "roofline": roofline, "roofline_hbm_stage": hbm, "cpu_baseline": cpu,
                "fp64_microbench": fp64 or None,
                "estimate": {"mean_dQ_per_level": [float(final[li, :, 0, 0, 1].mean()) for li in range(len(LEVELS))],
                             "sum": float(sum(final[li, :, 0, 0, 1].mean() for li in range(len(LEVELS))))}}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


def run_single_process_multi(args):
    """`python bench.py --gpus N` WITHOUT torchrun: one process, ONE multi-GPU context (mle_init_multi, the library's
    nb_of_workers analogue).  Every level's request carries N x n_l points per shift (weak scaling: the same per-GPU work as the
    N = 1 line); the library splits each request over its devices (whole shifts per device when R >= N), issues all members and
    merges on the host.  There is no device-resident arm across several devices, so `value` and `e2e` are the same host-buffer
    measurement (wall clock around the blocking `mle_update_samples` call)."""
    import torch
    import mle_b200
    import mle_b200.kl as kl
    N = args.gpus
    if torch.cuda.device_count() < N:
        raise SystemExit("bench.py --gpus %d: only %d CUDA devices visible" % (N, torch.cuda.device_count()))
    eng = mle_b200.Engine(list(range(N)))
    z = eng.genvec("K_3600_32")[:M_KL]
    lat = eng.lattice(z, M_KL)
    dists = eng.dists([(mle_b200.NORMAL, [0.0, 1.0])] * M_KL)
    model = eng.model("lognormal1d", 1, [16])
    for lev in LEVELS:
        model.set_basis(lev, kl.kl_basis_1d(lev, m=M_KL))
    sh_host = level_shifts()
    sh_sets = [sh_host.copy(), np.ascontiguousarray(1.0 - sh_host)]
    n_l = [points_per_shift(lev) * N for lev in LEVELS]
    samples_per_step = sum(n_l) * R_SHIFTS
    calls = [0]

    def step():
        sh_np = sh_sets[calls[0] & 1]
        calls[0] += 1
        return eng.update_samples(model, lat, dists, [([lev], sh_np[li], 0, n_l[li], M_KL) for li, lev in enumerate(LEVELS)])

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = eng.launch_count
    t0 = time.perf_counter()
    for _ in range(args.steps):
        mom = step()
    dt = (time.perf_counter() - t0) / args.steps
    launches = eng.launch_count - launches0
    clocks = sampler.stop()
    mom = np.stack(mom)
    value = samples_per_step / dt
    cfg = workload_config()
    cfg["parallelism"] = "single process, one multi-GPU context (mle_init_multi): every request split over %d devices inside the library" % N
    emit({"metric": "fine+coarse samples/sec", "value": value, "unit": "samples/s", "n_gpus": N, "steps": args.steps,
          "warmup": max(args.warmup, 3), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "f64", "data": "synthetic", "config": cfg, "clocks": clocks,
          "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": int(sh_host.nbytes),
                  "d2h_bytes_per_step": int(len(LEVELS) * R_SHIFTS * 6 * 8), "ms_per_step": dt * 1e3},
          "gpu_launches": int(launches), "launcher": "single process (no torchrun)",
          "estimate": {"mean_dQ_per_level": [float(mom[li, :, 0, 0, 1].mean()) for li in range(len(LEVELS))],
                       "sum": float(sum(mom[li, :, 0, 0, 1].mean() for li in range(len(LEVELS))))}})
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: native libraries (the NCCL version banner, for one) write to file descriptor 1
    # directly, so fd 1 is pointed at stderr for the run and the JSON line goes to a private duplicate of the real stdout
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    elif args.gpus > 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        run_single_process_multi(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
