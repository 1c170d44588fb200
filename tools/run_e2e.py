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


def measure_2d(config="c4", gpus=1, tol=None, reps=2, with_cpu=True):
    """the same for the 2-D models: c4 = AD(2)/QMC on the anisotropic multi-index model (100 KL terms, 16 shifts, indices up to size
    4), c3 = ML()/MC() on the multilevel model (400 KL terms, levels <= 4)"""
    import mle_b200, mle_b200.host as H, mle_b200.kl as kl
    from oracle_backend import OracleBackend
    if config == "c4":
        m, L = 100, 4
        bases = {i: kl.kl_basis_2d(i, m=m) for i in H.TD(2).get_index_set(L)}
        spec = H.ModelSpec("lognormal2d", 2, [4], bases)
        tol = 1e-5 if tol is None else tol
        make = lambda be: H.Estimator(H.AD(2), H.QMC(), spec, [H.Normal()] * m, be, max_index_set_param=L, nb_of_shifts=16,
                                      nb_of_warm_up_samples=8, cost_model=lambda i: 2.0 ** sum(i), nb_of_tols=4, min_index_set_param=1, shift_seed=7)
        name = "C4: AD(2)/QMC(), lognormal2d multi-index, 16 shifts, 100 KL terms, indices of size <= 4, cost model 2^|index|"
    else:
        m, L = 400, 4
        spec = H.ModelSpec("lognormal2d", 1, [4], {lev: kl.kl_basis_2d(lev, m=m) for lev in range(L + 1)})
        tol = 2e-5 if tol is None else tol
        make = lambda be: H.Estimator(H.ML(), H.MC(), spec, [H.Normal()] * m, be, max_index_set_param=L, cost_model=lambda i: 4.0 ** i[0],
                                      nb_of_tols=4, mc_seed=11)
        name = "C3: ML()/MC(), lognormal2d, 400 KL terms, levels <= 4, cost model 4^l"
    eng = mle_b200.Engine(list(range(gpus)) if gpus > 1 else 0)
    row = {"config": name + ", run(estimator, %g)" % tol, "gpus": gpus, "cpu_cores": os.cpu_count(), "tol": tol}
    best = None
    for _ in range(reps):
        be = Counting(H.GpuBackend(eng)); est = make(be)
        t0 = time.perf_counter(); h = H.run(est, tol); dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, be.calls, be.samples, h, be.seconds)
    dt, calls, samples, hg, inside = best
    row.update(gpu_seconds=dt, gpu_calls=calls, gpu_seconds_in_library=inside, samples=samples, mean=hg["mean"], rmse=hg["rmse"],
               indices=len(hg["current_index_set"]))
    if with_cpu:
        be = Counting(OracleBackend()); est = make(be)
        t0 = time.perf_counter(); ho = H.run(est, tol); dtc = time.perf_counter() - t0
        row.update(cpu_seconds=dtc, cpu_seconds_in_oracle=be.seconds, speedup=dtc / dt, speedup_in_library=be.seconds / inside,
                   same_decisions=ho["nb_of_samples"] == hg["nb_of_samples"], mean_rel_diff=abs(ho["mean"] - hg["mean"]) / abs(hg["mean"]))
    eng.close()
    return row


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--tols", default="1e-4,2e-5,5e-6")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4"])
    a = ap.parse_args()
    if a.config != "c2":
        print(json.dumps(measure_2d(a.config, a.gpus, with_cpu=not a.no_cpu)))
        sys.exit(0)
    print(json.dumps(measure(a.gpus, tuple(float(t) for t in a.tols.split(",")), with_cpu=not a.no_cpu)))
