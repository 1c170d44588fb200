"""Run-level end-to-end timing: wall seconds of run(estimator, tol) on config C2 (MLQMC, 1-D lognormal diffusion, 16 shifts,
100-term KL, levels <= 6) for several tolerances, GPU backend (one context on --gpus devices) vs the CPU oracle backend on all
host cores.  Prints one JSON object; bench.py embeds it as `run_e2e`.   python tools/run_e2e.py [--gpus N] [--tols 1e-4,2e-5]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np


class Counting:
    """backend wrapper: counts the calls that reach the device / the oracle and the wall time spent inside them (the rest of a
    run is the Python twin of the reference's host loop, identical on both arms)"""
    def __init__(self, inner):
        self.inner, self.calls, self.samples, self.seconds = inner, 0, 0, 0.0
    def lattice(self, *a, **k): return self.inner.lattice(*a, **k)
    def dists(self, *a, **k): return self.inner.dists(*a, **k)
    def model(self, spec): return self.inner.model(spec)
    def sample_batch(self, model, lattice, dists, index, shifts, k0, n, m, mc_seed, want_samples=False):
        self.calls += 1; self.samples += n * (1 if shifts is None else len(shifts))
        t0 = time.perf_counter()
        out = self.inner.sample_batch(model, lattice, dists, index, shifts, k0, n, m, mc_seed, want_samples=want_samples)
        self.seconds += time.perf_counter() - t0
        return out
    def update_samples(self, model, lattice, dists, entries, mc_seed):
        self.calls += 1; self.samples += sum(e[3] * (1 if e[1] is None else len(e[1])) for e in entries)
        t0 = time.perf_counter()
        out = self.inner.update_samples(model, lattice, dists, entries, mc_seed)
        self.seconds += time.perf_counter() - t0
        return out


def measure(gpus=1, tols=(1e-4, 2e-5, 5e-6), reps=3, with_cpu=True):
    import mle_b200, mle_b200.host as H, mle_b200.kl as kl
    from oracle_backend import OracleBackend
    m, L = 100, 6
    spec = H.ModelSpec("lognormal1d", 1, [16], {lev: kl.kl_basis_1d(lev, m=m) for lev in range(L + 1)})

    def make(backend):
        return H.Estimator(H.ML(), H.QMC(), spec, [H.Normal()] * m, backend, nb_of_shifts=16, max_index_set_param=L,
                           cost_model=lambda i: 2.0 ** i[0], shift_seed=7)

    eng = mle_b200.Engine(list(range(gpus)) if gpus > 1 else 0)
    out = {"config": "C2: ML()/QMC(), lognormal1d, 16 shifts, 100 KL terms, levels <= 6, cost model 2^l, run(estimator, tol)",
           "gpus": gpus, "cpu_cores": os.cpu_count(), "tols": []}
    for tol in tols:
        row = {"tol": tol}
        best = None
        for _ in range(reps):
            be = Counting(H.GpuBackend(eng))
            est = make(be)
            t0 = time.perf_counter(); h = H.run(est, tol); dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, be.calls, be.samples, h, be.seconds)
        dt, calls, samples, hg, inside = best
        row.update(gpu_seconds=dt, gpu_calls=calls, gpu_calls_per_s=calls / dt, gpu_seconds_in_library=inside,
                   gpu_us_per_call_in_library=inside / calls * 1e6, samples=samples, mean=hg["mean"], rmse=hg["rmse"],
                   levels=len(hg["current_index_set"]))
        if with_cpu:
            bestc = None
            for _ in range(max(1, reps - 1)):
                be = Counting(OracleBackend())
                est = make(be)
                t0 = time.perf_counter(); ho = H.run(est, tol); dtc = time.perf_counter() - t0
                if bestc is None or dtc < bestc[0]:
                    bestc = (dtc, ho, be.seconds)
            row.update(cpu_seconds=bestc[0], cpu_seconds_in_oracle=bestc[2], speedup=bestc[0] / dt, speedup_in_library=bestc[2] / inside, same_decisions=bestc[1]["nb_of_samples"] == hg["nb_of_samples"],
                       mean_rel_diff=abs(bestc[1]["mean"] - hg["mean"]) / abs(hg["mean"]))
        out["tols"].append(row)
    eng.close()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--tols", default="1e-4,2e-5,5e-6")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    print(json.dumps(measure(a.gpus, tuple(float(t) for t in a.tols.split(",")), with_cpu=not a.no_cpu)))
