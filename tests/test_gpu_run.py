"""GPU end-to-end: run(estimator, tol) through the host mirror on libmle_cuda against the same loop on the CPU oracle.
Identical shifts and a fixed cost model make the adaptive loop deterministic, so both must take the same
decisions (sample counts per level) and agree on estimates and per-level variances to 1e-12 relative (north star)."""
import numpy as np
import pytest

from oracle_backend import OracleBackend

pytestmark = pytest.mark.gpu


def both(engine, make):
    import mle_b200.host as H
    g = make(H, H.GpuBackend(engine))
    o = make(H, OracleBackend())
    return H, g, o


def compare(hg, ho, rtol=1e-12):
    assert len(hg) == len(ho)
    for i in range(len(hg)):
        a, b = hg[i], ho[i]
        assert a["nb_of_samples"] == b["nb_of_samples"], (i, a["nb_of_samples"], b["nb_of_samples"])
        assert a["current_index_set"] == b["current_index_set"]
        assert abs(a["mean"] - b["mean"]) <= rtol * abs(b["mean"])
        assert abs(a["varest"] - b["varest"]) <= 1e-9 * abs(b["varest"])  # variance of 16 nearly equal shift means
        for key in ("E", "dE", "V", "dV"):
            for idx in b[key]:
                x, y = a[key][idx], b[key][idx]
                if np.isnan(y):
                    assert np.isnan(x)
                else:
                    assert abs(x - y) <= rtol * abs(y), (key, idx, x, y)


def test_mlqmc_lognormal1d_run_matches_oracle(engine):
    """config C2 at test size: ML()/QMC(), 16 shifts, 100-term KL, levels up to 4"""
    import mle_b200.kl as kl

    def make(H, backend):
        spec = H.ModelSpec("lognormal1d", 1, [16], {lev: kl.kl_basis_1d(lev, m=100) for lev in range(5)})
        return H.Estimator(H.ML(), H.QMC(), spec, [H.Normal()] * 100, backend, nb_of_shifts=16, max_index_set_param=4,
                           cost_model=lambda i: 2.0 ** i[0], shift_seed=7, nb_of_tols=4)

    H, g, o = both(engine, make)
    hg, ho = H.run(g, 2e-4), H.run(o, 2e-4)
    compare(hg, ho)
    assert hg["rmse"] <= 2e-4 * 1.5 or hg["current_index_set"][-1] == (4,)
    assert len(hg["current_index_set"]) >= 3
    assert 0.05 < hg["mean"] < 0.3


def test_mlqmc_lognormal1d_run_to_level_6(engine):
    """config C2 at its stated depth: ML()/QMC(), 16 shifts, 100-term KL, max_index_set_param = 6; the tolerance is small
    enough that the loop adds every level up to 6 (grid 1024), and both backends take the same decisions all the way"""
    import mle_b200.kl as kl

    def make(H, backend):
        spec = H.ModelSpec("lognormal1d", 1, [16], {lev: kl.kl_basis_1d(lev, m=100) for lev in range(7)})
        return H.Estimator(H.ML(), H.QMC(), spec, [H.Normal()] * 100, backend, nb_of_shifts=16, max_index_set_param=6,
                           cost_model=lambda i: 2.0 ** i[0], shift_seed=7, nb_of_tols=6)

    H, g, o = both(engine, make)
    hg, ho = H.run(g, 1e-6), H.run(o, 1e-6)
    compare(hg, ho)
    assert hg["current_index_set"][-1] == (6,)
    assert 0.119 < hg["mean"] < 0.120


def test_mlmc_expsum_run_matches_oracle(engine):
    def make(H, backend):
        w = [0.5 / (j + 1) for j in range(64)]
        return H.Estimator(H.ML(), H.MC(), H.ModelSpec("expsum", 1, [4] + w), [H.Normal()] * 64, backend,
                           max_index_set_param=4, cost_model=lambda i: 2.0 ** i[0], nb_of_tols=3)

    H, g, o = both(engine, make)
    compare(H.run(g, 5e-3), H.run(o, 5e-3))


def test_problem18_regression_suite_on_gpu(engine):
    """test/regression.jl on the device: ML x {MC, QMC} x {max, sum, norm} and TD(2)/FT(2) x QMC"""
    for method in ("MC", "QMC"):
        for agg, tol in (("max", 1e-2), ("sum", 1e-1), ("norm", 1e-1)):
            def make(H, backend):
                return H.Estimator(H.ML(), getattr(H, method)(), H.ModelSpec("problem18", 1), [H.Uniform(-0.5, 0.5)], backend,
                                   nb_of_qoi=13, max_index_set_param=3, aggregation=agg, cost_model=lambda i: 2.0 ** i[0])
            H, g, o = both(engine, make)
            compare(H.run(g, tol), H.run(o, tol))
    for iset in ("TD", "FT"):
        def make(H, backend):
            return H.Estimator(getattr(H, iset)(2), H.QMC(), H.ModelSpec("problem18", 2), [H.Uniform(-0.5, 0.5)] * 2, backend,
                               nb_of_qoi=13, max_index_set_param=3, cost_model=lambda i: 2.0 ** sum(i), nb_of_shifts=8)
        H, g, o = both(engine, make)
        compare(H.run(g, 1e-2), H.run(o, 1e-2))


def test_mlmc_lognormal2d_run_matches_oracle(engine):
    """config C3 at test size: ML()/MC() on the 2-D lognormal diffusion problem, levels up to 3 (32 x 32 cells)"""
    import mle_b200.kl as kl

    def make(H, backend):
        spec = H.ModelSpec("lognormal2d", 1, [4], {lev: kl.kl_basis_2d(lev, m=60) for lev in range(4)})
        return H.Estimator(H.ML(), H.MC(), spec, [H.Normal()] * 60, backend, max_index_set_param=3,
                           cost_model=lambda i: 4.0 ** i[0], nb_of_tols=3, mc_seed=5)

    H, g, o = both(engine, make)
    hg, ho = H.run(g, 2e-3), H.run(o, 2e-3)
    compare(hg, ho)
    assert 0.05 < hg["mean"] < 0.09


def test_adaptive_multi_index_lognormal2d_run_matches_oracle(engine):
    """config C4 at test size: AD(2)/QMC, 2-D lognormal diffusion with anisotropic multi-index differences; the adaptive
    index-set construction on the device follows the oracle's step for step (same profits -> same logbook)"""
    import mle_b200.kl as kl

    def make(H, backend):
        bases = {i: kl.kl_basis_2d(i, m=40) for i in H.TD(2).get_index_set(3)}
        spec = H.ModelSpec("lognormal2d", 2, [4], bases)
        return H.Estimator(H.AD(2), H.QMC(), spec, [H.Normal()] * 40, backend, max_index_set_param=3, nb_of_shifts=8,
                           nb_of_warm_up_samples=8, cost_model=lambda i: 2.0 ** sum(i), nb_of_tols=3, min_index_set_param=1)

    H, g, o = both(engine, make)
    hg, ho = H.run(g, 2e-3), H.run(o, 2e-3)
    compare(hg, ho)
    assert hg["logbook"] == ho["logbook"]
    assert 0.05 < hg["mean"] < 0.09


def test_single_level_mc_c1_run_matches_oracle(engine):
    """config C1 on the device: SL()/MC(), 16 Normal() inputs, run(estimator, 1e-3); closed form E = exp(sum a_j^2 / 2)"""
    import math
    w = [0.5 / (j + 1) for j in range(16)]

    def make(H, backend):
        return H.Estimator(H.SL(), H.MC(), H.ModelSpec("expsum", 1, [16] + w), [H.Normal()] * 16, backend,
                           cost_model=lambda i: 1.0, nb_of_tols=4)

    H, g, o = both(engine, make)
    hg, ho = H.run(g, 1e-3), H.run(o, 1e-3)
    compare(hg, ho)
    exact = math.exp(0.5 * sum(a * a for a in w))
    assert abs(hg["mean"] - exact) < 4e-3 and hg["rmse"] <= 1e-3 * 1.01
    assert hg["nb_of_samples"][(0,)] > 100000     # theta = 0.5: N = 2 V / eps^2


def test_unbiased_runs_match_oracle(engine):
    """U(1) x {MC, QMC} and U(2) x QMC on the reference's own test function: the device evaluates every idx <= index on the
    same points; index draws, pmf updates and estimates agree with the oracle-backed loop"""
    for method, d in (("MC", 1), ("QMC", 1), ("QMC", 2)):
        def make(H, backend):
            kw = dict(max_search_space=H.ML()) if d == 1 else {}
            if method == "QMC":
                kw["nb_of_shifts"] = 8   # a QMC-only key, src/options.jl:41
            return H.Estimator(H.U(d), getattr(H, method)(), H.ModelSpec("problem18", d), [H.Uniform(-0.5, 0.5)] * d, backend,
                               nb_of_qoi=13, max_index_set_param=3, cost_model=lambda i: 2.0 ** sum(i), nb_of_tols=3, **kw)
        H, g, o = both(engine, make)
        hg, ho = H.run(g, 5e-3), H.run(o, 5e-3)
        compare(hg, ho)
        assert hg["bias"] == 0.0 and abs(hg["rmse"] - ho["rmse"]) <= 1e-9 * ho["rmse"]
        assert all(abs(g.pmf[k] - o.pmf[k]) <= 1e-9 * o.pmf[k] for k in o.pmf)
