"""CPU tests of the host-side mirror of the reference interface (Estimator / run / History) on the oracle backend, and of
the multi-rank sharding logic with the gloo backend (world_size 2)."""
import math
import os

import numpy as np
import pytest

from oracle_backend import OracleBackend


def host():
    import mle_b200.host as H
    return H


def test_index_sets_match_reference_cardinalities():
    """test/index_set.jl:21-52: 3-d sets at sz = 7 and weighted 2-d sets (1, 3/5) at sz = 6"""
    H = host()
    for name, n3, n2 in (("FT", 512, 28), ("TD", 120, 17), ("HC", 38, 11), ("ZC", 98, 15)):
        T = getattr(H, name)
        s3 = T(3).get_index_set(7)
        assert len(s3) == n3 and s3[0] == (0, 0, 0) and (1, 1, 1) in s3
        assert (8, 0, 0) not in s3 and (0, 8, 0) not in s3 and (0, 0, 8) not in s3
        assert H.is_admissable(s3, (8, 0, 0)) and not H.is_admissable(s3, (0, 0, 0)) and not H.is_admissable(s3, (0, 9, 0))
        s2 = T(1 / 1, 3 / 5).get_index_set(6)
        assert len(s2) == n2 and s2[0] == (0, 0) and (1, 1) in s2 and (0, 4) not in s2
    assert H.TD(2).get_index_set(2) == [(0, 0), (1, 0), (2, 0), (0, 1), (1, 1), (0, 2)]   # src/index_set.jl:246-255
    assert len(H.ML().get_index_set(4)) == 5 and H.SL().get_index_set(3) == [(0,)]
    with pytest.raises(ValueError):
        H.TD(1)
    with pytest.raises(ValueError):
        H.HC(1.0, 1.5)


def test_level_and_index_like_the_reference_suite():
    """test/index.jl:6-29"""
    H = host()
    level_0, level_2 = H.Level(0), H.Level(2)
    assert tuple(a + b for a, b in zip(level_0, level_2)) == level_2
    with pytest.raises(TypeError):
        H.Level(0.)
    index_0_0, index_2_3 = H.Index(0, 0), H.Index(2, 3)
    assert len(index_0_0) == len(index_2_3) == 2
    assert tuple(a + b for a, b in zip(index_0_0, index_2_3)) == index_2_3
    with pytest.raises(TypeError):
        H.Index(0., 2)
    d_index = H.diff(H.Index(1))
    assert len(d_index) == 1 and d_index[H.Index(0)] == -1
    # src/index.jl:49-58 on multi-indices: the mixed difference
    assert H.diff((0,)) == {} and H.diff((0, 0)) == {}
    assert H.diff((2, 3)) == {(1, 2): 1, (2, 2): -1, (1, 3): -1}
    assert H.diff((2, 0)) == {(1, 0): -1} and H.diff((0, 4)) == {(0, 3): -1}
    assert H.diff((1, 1, 1)) == {(0, 0, 0): -1, (1, 0, 0): 1, (0, 1, 0): 1, (0, 0, 1): 1, (1, 1, 0): -1, (1, 0, 1): -1, (0, 1, 1): -1}


@pytest.mark.parametrize("name,index", [("problem18", (2, 1)), ("problem18", (0, 3)), ("lognormal2d", (1, 2)), ("lognormal2d", (2, 0))])
def test_device_model_differences_follow_diff_signs(name, index):
    """docs/src/example.md:233-241 / test/regression.jl:93-101: dQ(index) = Qf(index) + sum over diff(index) of sign * Qf(kappa)
    on the same random inputs -- checked sample by sample with the raw samples of every index involved"""
    import mle_b200.kl as kl
    H = host()
    be = OracleBackend()
    if name == "problem18":
        m, dist = 2, H.Uniform(-0.5, 0.5)
        spec = H.ModelSpec("problem18", 2)
    else:
        m, dist = 30, H.Normal()
        keys = [index] + list(H.diff(index))
        spec = H.ModelSpec("lognormal2d", 2, [4], {k: kl.kl_basis_2d(k, m=m) for k in keys})
    model, dists, lat = be.model(spec), be.dists([dist] * m), be.lattice(None, m)
    shifts = np.random.default_rng(3).random((3, m))

    def raw(idx):
        return be.sample_batch(model, lat, dists, list(idx), shifts, 5, 40, m, 0, want_samples=True)[2]   # [R, q, X, n]

    top = raw(index)
    want = top[:, :, 1, :].copy()
    for kappa, sign in H.diff(index).items():
        want += sign * raw(kappa)[:, :, 1, :]
    scale = np.abs(top[:, :, 1, :]).max()
    assert np.abs(top[:, :, 0, :] - want).max() <= 4e-15 * scale


def test_index_set_types_like_the_reference_suite(capsys):
    """the rest of test/index_set.jl:8-20, :54-86: SL / ML, AD, U, sizes 0 and -1, show and print"""
    H = host()
    for T, n, first in ((H.SL, 1, 0), (H.ML, 6, 0)):
        t = T()
        idxset = t.get_index_set(5)
        assert len(idxset) == n and idxset[0] == H.Level(first) and t.length(5) == n
        if T is H.ML:
            assert idxset[-1] == H.Level(5)
        with pytest.raises(TypeError):
            T(2)
        assert "%s" % t == T.__name__
    for T, n3 in ((H.FT, 512), (H.TD, 120), (H.HC, 38), (H.ZC, 98)):
        assert T(3).length(7) == n3 and "%s" % T(3) == "%s{3}" % T.__name__ and "%s" % T(1 / 1, 3 / 5) == "%s{2}" % T.__name__
        assert (7, 0) not in T(1 / 1, 3 / 5).get_index_set(6)
        with pytest.raises(ValueError):
            T(1, 0)
    ad = H.AD(2)
    assert ad.ndims == 2 and "%s" % ad == "AD{2}"
    with pytest.raises(ValueError):
        ad.get_index_set(4)
    with pytest.raises(ValueError):
        H.AD(1)
    with pytest.raises(TypeError):
        H.AD()
    u = H.U(1)
    assert u.ndims == 1 and "%s" % u == "U{1}" and "%s" % H.U(2) == "U{2}"
    with pytest.raises(ValueError):
        u.get_index_set(1)
    with pytest.raises(TypeError):
        H.U()
    with pytest.raises(TypeError):
        H.U(1, 2)
    with pytest.raises(ValueError):
        H.U(2).get_index_set(4)
    assert len(H.TD(2).get_index_set(0)) == 1 and H.TD(2).get_index_set(-1) == []
    with pytest.raises(ValueError):
        H.TD(3).print(4)                    # only for d = 2
    H.TD(2).print(3)
    out = capsys.readouterr().out            # the docstring example, src/index_set.jl:286-291
    assert [row.rstrip() for row in out.splitlines()] == ["  \u25FC", "  \u25FC \u25FC", "  \u25FC \u25FC \u25FC", "  \u25FC \u25FC \u25FC \u25FC"]
    pic = H.format_index_set([(0, 0), (1, 0)], [(0, 1)])
    assert pic == "  \u25A3 \n  \u25FC \u25FC \n"


def test_distribution_constructors_like_the_reference_suite(oracle):
    """test/distribution.jl: the constructor rules of all four distributions (the quantile values are pinned in
    tests/test_oracle_golden.py and, on the device, in tests/test_gpu_points.py)"""
    from fractions import Fraction
    H = host()
    inf, nan = math.inf, math.nan
    assert H.Uniform() == H.Uniform(0., 1.) == H.Uniform(0, 1.)                                  # :7-9
    for bad in ((inf, 1.), (0., nan), (1., 0.)):                                                 # :13-15
        with pytest.raises(ValueError):
            H.Uniform(*bad)
    assert H.Normal() == H.Normal(0., 1.) == H.Normal(0, 1.)                                     # :25-27
    for bad in ((inf, 1.), (0., nan), (0., -1.)):                                                # :31-33
        with pytest.raises(ValueError):
            H.Normal(*bad)
    assert H.TruncatedNormal() == H.TruncatedNormal(0., 1.) == H.TruncatedNormal(0., 1., -2., 2.)  # :45-46
    assert H.TruncatedNormal(0., 1., -1., 1.)[1][2:] == [-1.0, 1.0]
    assert H.TruncatedNormal(0., 1., Fraction(-2, 3), 10)[1][2:] == [-2 / 3, 10.0]               # :48
    for bad in ((inf, 1.), (0., nan), (0., 1., inf, 2.), (0., 1., -2., inf), (0., -1.), (0., 1., 1., 0.)):   # :53-58
        with pytest.raises(ValueError):
            H.TruncatedNormal(*bad)
    assert H.Weibull() == H.Weibull(2., math.sqrt(2.)) and H.Weibull(2, 1.) == H.Weibull(2., 1.)  # :70-72
    for bad in ((inf, 1.), (0., nan), (1., -1), (-1., 1)):                                       # :76-79
        with pytest.raises(ValueError):
            H.Weibull(*bad)
    # transform(Uniform(), x) == x, test/distribution.jl:18-20, through the same (kind, parameters) pairs the device receives
    x = np.arange(1, 10) / 10.0
    z = np.array([1], dtype=np.uint32)
    pts = oracle.points(z, [H.Uniform()], x.reshape(9, 1), 0, 1)    # point 0 of nine "shifts" = the shifts themselves
    assert np.array_equal(pts[:, 0, 0], x)


def test_option_validation_mirrors_reference():
    H = host()
    spec = H.ModelSpec("problem18", 1)
    be = OracleBackend()
    with pytest.raises(ValueError):
        H.Estimator(H.ML(), H.MC(), spec, [H.Uniform(-0.5, 0.5)], be, nb_of_warm_up_samples=1)   # test/estimator.jl
    with pytest.raises(ValueError):
        H.Estimator(H.ML(), H.QMC(), spec, [H.Uniform(-0.5, 0.5)], be, nb_of_shifts=1)
    with pytest.raises(ValueError):
        H.Estimator(H.ML(), H.MC(), spec, [H.Uniform(-0.5, 0.5)], be, aggregation="median")
    with pytest.raises(ValueError):
        H.Estimator(H.ML(), H.MC(), spec, [H.Uniform(-0.5, 0.5)], be, no_such_option=3)
    with pytest.raises(TypeError):
        H.Estimator(H.ML(), H.MC(), lambda l, x: (0, 0), [H.Uniform()], be)
    with pytest.raises(ValueError):
        H.Normal(0.0, -1.0)
    with pytest.raises(ValueError):
        H.Uniform(1.0, 0.0)
    e = H.Estimator(H.ML(), H.QMC(), spec, [H.Uniform(-0.5, 0.5)], be)
    assert e["nb_of_warm_up_samples"] == 1 and e["nb_of_shifts"]((0,)) == 10 and e["max_index_set_param"] == 10
    assert e["continuation_mul_factor"] == 1.3 and e["sample_mul_factor"] == 1.2  # src/parse.jl:80, :216


def test_option_scoping_and_rules_follow_reference(tmp_path):
    """src/options.jl:10-43 (which keys an estimator accepts) and the per-key rules of src/parse.jl:35-307"""
    H = host()
    spec, spec2 = H.ModelSpec("problem18", 1), H.ModelSpec("problem18", 2)
    be, U1, U2 = OracleBackend(), [H.Uniform(-0.5, 0.5)], [H.Uniform(-0.5, 0.5)] * 2

    def bad(iset, method, sp, dists, **kw):
        with pytest.raises(ValueError) as e:
            H.Estimator(iset, method, sp, dists, be, **kw)
        return str(e.value)

    # scoping: QMC-only, multilevel-only, AD-only and U-only keys
    assert "invalid option nb_of_shifts" in bad(H.ML(), H.MC(), spec, U1, nb_of_shifts=4)
    assert "invalid option sample_mul_factor" in bad(H.ML(), H.MC(), spec, U1, sample_mul_factor=2)
    assert "invalid option point_generator" in bad(H.ML(), H.MC(), spec, U1, point_generator=[1, 3])
    assert "invalid option do_regression" in bad(H.SL(), H.MC(), spec, U1, do_regression=False)
    assert "invalid option min_splitting" in bad(H.U(1), H.MC(), spec, U1, max_search_space=H.ML(), min_splitting=0.6)
    assert "invalid option penalization" in bad(H.U(2), H.QMC(), spec2, U2, penalization=0.3)
    assert "invalid option max_search_space" in bad(H.TD(2), H.QMC(), spec2, U2, max_search_space=H.FT(2))
    H.Estimator(H.U(1), H.MC(), spec, U1, be, max_search_space=H.ML(), sample_mul_factor=2)        # valid for U with MC
    H.Estimator(H.AD(2), H.QMC(), spec2, U2, be, do_regression=False, penalization=0.0, acceptance_rate=1)
    assert set(H.valid_options(H.SL(), H.MC())) - {"shifts", "shift_seed", "mc_seed"} == {
        "nb_of_warm_up_samples", "nb_of_qoi", "qoi_with_max_var", "aggregation", "nb_of_tols", "continuation_mul_factor",
        "continuate", "save", "save_samples", "verbose", "folder", "name", "cost_model", "nb_of_workers",
        "nb_of_uncertainties", "max_index_set_param", "min_index_set_param"}
    # types
    assert "must be of type Signed" in bad(H.ML(), H.MC(), spec, U1, nb_of_warm_up_samples=2.5)
    assert "must be of type Signed" in bad(H.ML(), H.MC(), spec, U1, nb_of_qoi=True)
    assert "must be of type String" in bad(H.ML(), H.MC(), spec, U1, aggregation=3)
    assert "must be of type Bool" in bad(H.ML(), H.MC(), spec, U1, continuate=1)
    assert "must be of type Bool" in bad(H.ML(), H.MC(), spec, U1, save_samples="yes")
    assert "must be of type Function" in bad(H.ML(), H.MC(), spec, U1, cost_model=2.0)
    assert "must be of type Function" in bad(H.ML(), H.MC(), spec, U1, nb_of_uncertainties=1)
    assert "must be of type Real" in bad(H.ML(), H.MC(), spec, U1, continuation_mul_factor="2")
    assert "AbstractIndexSet" in bad(H.AD(2), H.MC(), spec2, U2, max_search_space=3)
    # ranges
    assert "nb_of_qoi must be larger than 0" in bad(H.ML(), H.MC(), spec, U1, nb_of_qoi=0)
    assert "nb_of_tols must be larger than 0" in bad(H.ML(), H.MC(), spec, U1, nb_of_tols=0)
    assert "continuation_mul_factor must be larger than 1" in bad(H.ML(), H.MC(), spec, U1, continuation_mul_factor=1.0)
    assert "must be finite" in bad(H.ML(), H.MC(), spec, U1, continuation_mul_factor=math.inf)
    assert "min_splitting must be larger than 0" in bad(H.ML(), H.MC(), spec, U1, min_splitting=0.0)
    assert "max_splitting must be smaller than or equal to 1" in bad(H.ML(), H.MC(), spec, U1, max_splitting=1.5)
    assert "min_splitting must be smaller than or equal to optional key max_splitting" in bad(
        H.ML(), H.MC(), spec, U1, min_splitting=0.9, max_splitting=0.6)
    assert "min_index_set_param must be smaller than or equal to" in bad(H.ML(), H.MC(), spec, U1, max_index_set_param=1)
    assert "max_index_set_param must be larger than 0" in bad(H.ML(), H.MC(), spec, U1, max_index_set_param=0, min_index_set_param=0)
    assert "must be finite" in bad(H.ML(), H.QMC(), spec, U1, sample_mul_factor=math.nan)
    assert "nb_of_workers must be larger than 0" in bad(H.ML(), H.MC(), spec, U1, nb_of_workers=0)
    assert "cannot be a Function" in bad(H.U(1), H.QMC(), spec, U1, max_search_space=H.ML(), nb_of_shifts=lambda i: 4)
    assert "penalization must be larger than or equal to 0" in bad(H.AD(2), H.MC(), spec2, U2, penalization=-0.1)
    assert "qoi_with_max_var must be larger than 0" in bad(H.ML(), H.MC(), spec, U1, qoi_with_max_var=0)
    assert "qoi_with_max_var must be smaller than or equal to 13" in bad(H.ML(), H.MC(), spec, U1, nb_of_qoi=13, qoi_with_max_var=14)
    assert "must not contain a ." in bad(H.ML(), H.MC(), spec, U1, name="my.estimator")
    assert "is not a directory" in bad(H.ML(), H.MC(), spec, U1, save=True, folder=str(tmp_path / "missing"))
    # defaults and normalisation
    e = H.Estimator(H.ML(), H.QMC(), spec, U1, be, name="run7", nb_of_workers=3, nb_of_shifts=lambda i: 4 + i[0])
    assert e["name"] == "run7.jld2" and e["folder"] == os.getcwd() and e["nb_of_workers"]((2,)) == 3   # src/parse.jl:116-140
    assert e["nb_of_shifts"]((2,)) == 6 and e["save"] is False
    assert H.Estimator(H.ML(), H.MC(), spec, U1, be, name="a.jld2")["name"] == "a.jld2"
    assert H.Estimator(H.U(1), H.QMC(), spec, U1, be, max_search_space=H.ML())["nb_of_warm_up_samples"] == 2   # src/parse.jl:36
    assert H.Estimator(H.ML(), H.MC(), spec, U1, be)["nb_of_warm_up_samples"] == 20


class _NullBackend:
    """accepts any model / distributions / lattice: for tests of the constructor alone"""

    def model(self, spec):
        return None

    def dists(self, d):
        return list(d)

    def lattice(self, z, s):
        return None


def test_estimator_constructor_like_the_reference_suite():
    """test/estimator.jl:6-99, line by line (ArgumentError -> ValueError; `qoi` is a ModelSpec instead of a closure)"""
    H = host()
    ml, ad, mc, qmc = H.ML(), H.AD(3), H.MC(), H.QMC()
    qoi, distr, distrs, be = H.ModelSpec("expsum", 1, [4.0]), H.Normal(), [H.Normal() for _ in range(1000)], _NullBackend()

    def ok(iset, method, d=distr, **kw):
        e = H.Estimator(iset, method, qoi, d, be, **kw)
        assert isinstance(e.index_set, type(iset)) and isinstance(e.sample_method, type(method))
        return e

    def throws(iset, method, **kw):
        with pytest.raises(ValueError):
            H.Estimator(iset, method, qoi, distr, be, **kw)

    assert len(ok(ml, mc, distrs).distributions) == 1000           # :15
    assert len(ok(ml, qmc).distributions) == 1                     # :16
    throws(ml, mc, nb_of_warm_up_samples=False)                    # :18-20
    throws(ml, mc, nb_of_warm_up_samples=0)
    ok(ml, mc, nb_of_warm_up_samples=10)
    throws(ml, mc, nb_of_qoi=False)                                # :22-24
    throws(ml, mc, nb_of_qoi=0)
    ok(ml, mc, nb_of_qoi=10)
    throws(ml, mc, continuate="blabla")                            # :26-27
    ok(ml, mc, continuate=False)
    throws(ml, mc, nb_of_tols="blabla")                            # :29-31
    throws(ml, mc, nb_of_tols=0)
    ok(ml, mc, nb_of_tols=20)
    throws(ml, mc, continuation_mul_factor="blabla")               # :33-35
    throws(ml, mc, continuation_mul_factor=math.nan)
    ok(ml, qmc, continuation_mul_factor=2)
    throws(ml, mc, folder=1.0)                                     # :37-38
    ok(ml, mc, folder=".")
    throws(ml, mc, name=1.0)                                       # :40-41
    assert ok(ml, mc, name="MyEstimator")["name"] == "MyEstimator.jld2"
    throws(ml, mc, save_samples=1)                                 # :43-44
    ok(ml, mc, save_samples=True)
    throws(ml, mc, verbose=1)                                      # :46-47
    ok(ml, mc, verbose=True)
    throws(ml, mc, cost_model="blabla")                            # :49-50
    ok(ml, mc, cost_model=lambda i: 2 ** sum(i))
    throws(ml, mc, robustify_bias_estimate="blabla")               # :52-53
    ok(ml, mc, robustify_bias_estimate=False)
    throws(ml, mc, do_mse_splitting=1)                             # :55-56
    ok(ml, mc, do_mse_splitting=False)
    throws(ml, mc, min_splitting=2)                                # :58-59
    ok(ml, mc, min_splitting=0.6)
    throws(ml, mc, max_splitting=0.4)                              # :61-63
    throws(ml, mc, max_splitting=0.6, min_splitting=0.7)
    ok(ml, mc, max_splitting=0.8)
    throws(ml, mc, max_index_set_param=1.0)                        # :65-67
    throws(ml, mc, max_index_set_param=-1)
    ok(ml, mc, max_index_set_param=6)
    throws(ml, mc, sample_mul_factor="blabla")                     # :69-72
    throws(ml, mc, sample_mul_factor=math.inf)
    ok(ml, qmc, sample_mul_factor=1.5)
    ok(ml, qmc, sample_mul_factor=2)
    throws(ml, mc, nb_of_workers=math.inf)                         # :74-76
    ok(ml, mc, nb_of_workers=4)
    assert ok(ml, mc, nb_of_workers=lambda i: 2)["nb_of_workers"]((0,)) == 2
    throws(ml, qmc, nb_of_shifts="blabla")                         # :78-80
    assert ok(ml, qmc, nb_of_shifts=16)["nb_of_shifts"]((1,)) == 16
    assert ok(ml, qmc, nb_of_shifts=lambda i: 32)["nb_of_shifts"]((1,)) == 32
    throws(ml, qmc, point_generator=3)                             # :82
    throws(ml, mc, do_regression="blabla")                         # :84-85
    throws(ml, mc, aggregation="blabla")
    throws(H.AD(2), mc, max_search_space=True)                     # :87-89
    throws(ad, mc, max_search_space=H.TD(2))
    assert isinstance(ok(ad, qmc, max_search_space=H.FT(3))["max_search_space"], H.FT)
    throws(ad, mc, penalization=-0.1)                              # :91-93
    throws(ad, mc, penalization=1.1)
    ok(ad, mc, penalization=0.5)
    throws(ad, mc, acceptance_rate=-0.1)                           # :95-97
    throws(ad, mc, acceptance_rate=1.1)
    ok(ad, mc, acceptance_rate=0.8)


def test_history_is_saved_after_every_tolerance_and_reloads(tmp_path):
    """src/history.jl:99-107: with save = true the history file under folder/name is rewritten after each tolerance; untitled
    estimators get UntitledEstimator, UntitledEstimator1, ... (src/parse.jl:142-153).  The mirror writes JSON."""
    H = host()
    kw = dict(nb_of_qoi=13, max_index_set_param=3, cost_model=lambda i: 2.0 ** sum(i), nb_of_tols=3, save=True,
              folder=str(tmp_path))
    est = H.Estimator(H.TD(2), H.QMC(), H.ModelSpec("problem18", 2), [H.Uniform(-0.5, 0.5)] * 2, OracleBackend(),
                      nb_of_shifts=4, save_samples=True, **kw)
    assert est["name"] == "UntitledEstimator.jld2"
    h = H.run(est, 2e-2)
    assert h["folder"] == str(tmp_path) and h.path == str(tmp_path / "UntitledEstimator.json") and os.path.isfile(h.path)
    assert repr(h) == "MultilevelEstimators.jl history file"
    g = H.History.load(h.path)
    assert len(g) == len(h) == 3
    for i in range(3):
        a, b = h[i], g[i]
        assert set(a) == set(b)
        assert a["current_index_set"] == b["current_index_set"] and a["index_set"] == b["index_set"]
        assert a["nb_of_samples"] == b["nb_of_samples"] and a["tol"] == b["tol"]
        for key in ("mean", "varest", "mse", "rmse", "bias", "alpha", "beta", "gamma"):
            assert np.array_equal(np.asarray(a[key], dtype=float), np.asarray(b[key], dtype=float), equal_nan=True), key
        for key in ("E", "V", "dE", "dV", "W"):
            assert set(a[key]) == set(b[key])
            for idx in a[key]:
                assert np.array_equal(np.asarray(a[key][idx], dtype=float), np.asarray(b[key][idx], dtype=float), equal_nan=True)
        for idx in a["samples_diff"]:
            assert np.array_equal(a["samples_diff"][idx], np.asarray(b["samples_diff"][idx]))
    # the next untitled estimators in the same folder do not overwrite the first one
    e2 = H.Estimator(H.ML(), H.MC(), H.ModelSpec("problem18", 1), [H.Uniform(-0.5, 0.5)], OracleBackend(), save=True,
                     folder=str(tmp_path))
    assert e2["name"] == "UntitledEstimator1.jld2"
    H.run(e2, [5e-2])
    e3 = H.Estimator(H.ML(), H.MC(), H.ModelSpec("problem18", 1), [H.Uniform(-0.5, 0.5)], OracleBackend(), save=True,
                     folder=str(tmp_path))
    assert e3["name"] == "UntitledEstimator2.jld2"
    # nothing is written without save
    e4 = H.Estimator(H.ML(), H.MC(), H.ModelSpec("problem18", 1), [H.Uniform(-0.5, 0.5)], OracleBackend(), name="quiet",
                     folder=str(tmp_path))
    H.run(e4, [5e-2])
    assert not os.path.exists(tmp_path / "quiet.json")


@pytest.mark.parametrize("method", ["MC", "QMC"])
@pytest.mark.parametrize("aggregation,tol", [("max", 1e-2), ("sum", 1e-1), ("norm", 1e-1)])
def test_regression_problem18_multilevel(method, aggregation, tol):
    """test/regression.jl:49-66: run(Estimator(ML(), MC()/QMC(), problem_18, Uniform(-0.5,0.5); nb_of_qoi=13,
    max_index_set_param=3, aggregation), tol) completes; here we also check the History"""
    H = host()
    est = H.Estimator(H.ML(), getattr(H, method)(), H.ModelSpec("problem18", 1), [H.Uniform(-0.5, 0.5)], OracleBackend(),
                      nb_of_qoi=13, max_index_set_param=3, aggregation=aggregation, cost_model=lambda i: 2.0 ** i[0])
    h = H.run(est, tol)
    assert len(h) == 10                                   # 10 tolerances, factor 1.3 apart (src/parse.jl:70-86)
    assert abs(h[0]["tol"] / h[1]["tol"] - 1.3) < 1e-12 and abs(h["tol"] - tol) < 1e-15
    for key in ("type", "ndims", "name", "elapsed", "tol", "current_index_set", "index_set", "mse", "rmse", "mean", "varest",
                "bias", "E", "V", "dE", "dV", "T", "alpha", "beta", "gamma", "nb_of_samples", "W"):
        assert key in h                                   # src/history.jl:57-84
    assert h["rmse"] <= 2 * tol or h["current_index_set"][-1] == (3,)
    ns = [h[i]["nb_of_samples"][(0,)] for i in range(len(h))]
    assert ns == sorted(ns)                               # samples persist across tolerances
    # exact value of the level-3 model: f_j + E[xi^3] = f_j (odd moment vanishes); QoI 1 is (0-2)^2 = 4
    if aggregation == "max":
        assert abs(h["mean"] - 4.0) < 5 * tol


def test_multi_index_problem18_td():
    H = host()
    est = H.Estimator(H.TD(2), H.QMC(), H.ModelSpec("problem18", 2), [H.Uniform(-0.5, 0.5)] * 2, OracleBackend(),
                      nb_of_qoi=13, max_index_set_param=3, cost_model=lambda i: 2.0 ** sum(i), nb_of_shifts=8)
    h = H.run(est, 1e-2)
    assert h["ndims"] == 2 and (1, 1) in h["index_set"]
    assert abs(h["mean"] - 4.0) < 0.05


def test_single_level_mc_expsum_closed_form():
    """config C1: SL()/MC() on 16 Normal() inputs, E[exp(sum a_j xi_j)] = exp(sum a_j^2 / 2)"""
    H = host()
    w = [0.5 / (j + 1) for j in range(16)]
    est = H.Estimator(H.SL(), H.MC(), H.ModelSpec("expsum", 1, [16] + w), [H.Normal()] * 16, OracleBackend(),
                      cost_model=lambda i: 1.0, nb_of_tols=4)
    h = H.run(est, 4e-3)
    exact = math.exp(0.5 * sum(a * a for a in w))
    assert abs(h["mean"] - exact) < 4 * 4e-3
    n = h["nb_of_samples"][(0,)]
    assert abs(n - 2 * h["V"][(0,)] / 4e-3**2) / n < 0.2   # theta = 1/2 for SL: N = V / (theta eps^2), src/run.jl:50


def test_qmc_growth_rule_and_nan_variance_with_one_sample():
    H = host()
    est = H.Estimator(H.ML(), H.QMC(), H.ModelSpec("problem18", 1), [H.Uniform(-0.5, 0.5)], OracleBackend(), nb_of_qoi=13,
                      cost_model=lambda i: 2.0 ** i[0])
    est.sample((0,), 1)
    assert math.isnan(est.var((0,)))                     # var of a 1-vector is NaN
    assert est.varest((0,)) > 0                          # ... while the across-shift variance drives the loop
    est.current_index_set.add((0,))
    assert est.next_number_of_samples((0,)) == {(0,): 2}  # ceil(1 * 1.2)
    est.sample((0,), 4)
    assert est.nb_of_samples[(0,)] == 5 and est.next_number_of_samples((0,)) == {(0,): 6}
    assert est.moments[(0,)].shape == (10, 13, 2, 3) and (est.moments[(0,)][..., 0] == 5).all()


def test_split_range_covers_everything():
    H = host()
    for n, world in [(1, 2), (7, 2), (8, 8), (3, 8), (1000, 3)]:
        got = []
        for r in range(world):
            s, c = H.split_range(5, n, world, r)
            got += list(range(s, s + c))
        assert got == list(range(5, 5 + n))


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import mle_b200.host as H
    be = H.ShardedBackend(OracleBackend())
    est = H.Estimator(H.ML(), H.QMC(), H.ModelSpec("problem18", 1), [H.Uniform(-0.5, 0.5)], be, nb_of_qoi=13,
                      max_index_set_param=3, cost_model=lambda i: 2.0 ** i[0], nb_of_shifts=4, nb_of_tols=3)
    h = H.run(est, 2e-2)
    if rank == 0:
        np.save(out, np.array([h["mean"], h["varest"], float(sum(h["nb_of_samples"].values()))]))
    dist.barrier()
    dist.destroy_process_group()


def _make_unbiased(H, backend, method):
    return H.Estimator(H.U(2), getattr(H, method)(), H.ModelSpec("problem18", 2), [H.Uniform(-0.5, 0.5)] * 2, backend,
                       nb_of_qoi=13, max_index_set_param=3, cost_model=lambda i: 2.0 ** sum(i), nb_of_tols=3, save_samples=True)


def _worker_unbiased(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import mle_b200.host as H
    res = []
    for method in ("MC", "QMC"):
        est = _make_unbiased(H, H.ShardedBackend(OracleBackend()), method)
        h = H.run(est, 5e-3)
        first = h["samples_diff"][(0, 0)]
        res += [h["mean"], h["varest"], float(sum(h["nb_of_samples"].values())), float(first.shape[-1]), float(first.sum())]
    if rank == 0:
        np.save(out, np.array(res))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_unbiased_run_matches_single_rank(tmp_path):
    """U(2) x {MC, QMC} with save_samples on two gloo ranks: each rank evaluates its slice of the points for every idx <= index,
    a second all-gather brings the raw samples back in point order, and the per-sample sums, pmf updates and index draws are
    identical on all ranks and equal to the single-rank run"""
    import torch.multiprocessing as mp
    H = host()
    want = []
    for method in ("MC", "QMC"):
        h = H.run(_make_unbiased(H, OracleBackend(), method), 5e-3)
        first = h["samples_diff"][(0, 0)]
        want += [h["mean"], h["varest"], float(sum(h["nb_of_samples"].values())), float(first.shape[-1]), float(first.sum())]
    out = str(tmp_path / "u.npy")
    mp.spawn(_worker_unbiased, args=(2, 29531, out), nprocs=2, join=True)
    got = np.load(out)
    for g, w, tol in zip(got, want, [1e-12, 1e-9, 0, 0, 1e-12] * 2):
        assert abs(g - w) <= tol * abs(w), (got, want)


def test_two_rank_sharding_matches_single_rank(tmp_path):
    """world_size-2 gloo run: every sample_batch is split over the ranks, one all-gather merges the moments"""
    import torch.multiprocessing as mp
    H = host()
    est = H.Estimator(H.ML(), H.QMC(), H.ModelSpec("problem18", 1), [H.Uniform(-0.5, 0.5)], OracleBackend(), nb_of_qoi=13,
                      max_index_set_param=3, cost_model=lambda i: 2.0 ** i[0], nb_of_shifts=4, nb_of_tols=3)
    h = H.run(est, 2e-2)
    out = str(tmp_path / "r0.npy")
    mp.spawn(_worker, args=(2, 29517, out), nprocs=2, join=True)
    got = np.load(out)
    assert got[2] == float(sum(h["nb_of_samples"].values()))      # identical decisions
    assert abs(got[0] - h["mean"]) <= 1e-12 * abs(h["mean"])
    assert abs(got[1] - h["varest"]) <= 1e-9 * abs(h["varest"])


def _worker_default_cost(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import mle_b200.host as H
    res = []
    for index_set, method, spec, dl in ((H.ML(), H.QMC(), H.ModelSpec("problem18", 1), [H.Uniform(-0.5, 0.5)]),
                                        (H.AD(2), H.MC(), H.ModelSpec("problem18", 2), [H.Uniform(-0.5, 0.5)] * 2)):
        est = H.Estimator(index_set, method, spec, dl, H.ShardedBackend(OracleBackend()), nb_of_qoi=13, max_index_set_param=3,
                          nb_of_tols=3, **({"nb_of_shifts": 4} if isinstance(method, H.QMC) else {}))   # cost_model: default (time)
        h = H.run(est, 2e-2)
        res += [h["mean"], float(sum(h["nb_of_samples"].values())), float(sum(est.total_time.values()))]
    np.save(out % rank, np.array(res))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_run_with_the_default_time_cost_model(tmp_path):
    """the reference's default cost model is the measured time of each sample! call (src/sample.jl:27, src/internals.jl:106).
    Sharded, every rank must charge the SAME time or the ranks request different sample counts and deadlock in the all-gather:
    the backend returns the maximum of the ranks' wall times and the estimator uses it.  Both ranks finish with identical
    sample counts, estimates and accumulated times."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "dc%d.npy")
    mp.spawn(_worker_default_cost, args=(2, 29557, out), nprocs=2, join=True)
    a, b = np.load(out % 0), np.load(out % 1)
    assert np.array_equal(a, b), (a, b)
    assert a[2] > 0 and a[5] > 0


def test_adaptive_index_set_mechanics():
    """src/methods.jl:223-251 and test/index_set.jl:54-60: AD starts from the root index, moves the index with the largest
    profit into the old set, activates its admissable forward neighbours inside max_search_space, and logs every step"""
    H = host()
    with pytest.raises(ValueError):
        H.AD(1)
    with pytest.raises(NotImplementedError):
        H.AD(2).get_index_set(3)
    spec = H.ModelSpec("problem18", 2)
    be = OracleBackend()
    with pytest.raises(ValueError):
        H.Estimator(H.TD(2), H.MC(), spec, [H.Uniform(-0.5, 0.5)] * 2, be, penalization=0.5)     # AD-only key, src/options.jl:36
    with pytest.raises(ValueError):
        H.Estimator(H.AD(2), H.MC(), spec, [H.Uniform(-0.5, 0.5)] * 2, be, max_search_space=H.TD(3))
    with pytest.raises(ValueError):
        H.Estimator(H.AD(2), H.MC(), spec, [H.Uniform(-0.5, 0.5)] * 2, be, acceptance_rate=1.5)
    est = H.Estimator(H.AD(2), H.MC(), spec, [H.Uniform(-0.5, 0.5)] * 2, be, nb_of_qoi=13, max_index_set_param=3,
                      cost_model=lambda i: 2.0 ** sum(i))
    assert isinstance(est["max_search_space"], H.TD) and est["penalization"] == 0.5 and est["acceptance_rate"] == 1.0
    first = est.new_index_set(0)
    assert first == [(0, 0)] and est.active_set == {(0, 0)} and not est.old_set
    est.sample((0, 0), 20)
    est.current_index_set.update(first)
    second = est.new_index_set(1)
    assert second[0] == (0, 0) and set(second[1:]) == {(1, 0), (0, 1)}
    assert est.old_set == {(0, 0)} and est.active_set == {(1, 0), (0, 1)}
    assert est.logbook[-1][2] == (0, 0) and len(est.logbook) == 2
    for i in second[1:]:
        est.sample(i, 20)
    est.current_index_set.update(second)
    third = est.new_index_set(2)
    # (1, 1) is not admissable until both (1, 0) and (0, 1) are old
    assert (1, 1) not in third and len(est.old_set) == 2


@pytest.mark.parametrize("method", ["MC", "QMC"])
def test_regression_problem18_adaptive(method):
    """test/regression.jl:105-121 shape: run(Estimator(AD(2), MC()/QMC(), problem_18, ...; max_index_set_param = 3)) completes
    with a downward-closed index set inside the TD search space and a logbook in the History"""
    H = host()
    est = H.Estimator(H.AD(2), getattr(H, method)(), H.ModelSpec("problem18", 2), [H.Uniform(-0.5, 0.5)] * 2, OracleBackend(),
                      nb_of_qoi=13, max_index_set_param=3, cost_model=lambda i: 2.0 ** sum(i), nb_of_tols=4)
    h = H.run(est, 1e-2)
    assert len(h) == 4 and "logbook" in h and h["type"] == "Estimator{AD, %s}" % method
    cur = set(h["current_index_set"])
    space = set(H.TD(2).get_index_set(3))
    assert (0, 0) in cur and cur <= space
    for idx in cur:                                       # downward closed
        for k in range(2):
            if idx[k] > 0:
                assert tuple(v - (t == k) for t, v in enumerate(idx)) in cur
    old, active, _ = h["logbook"][-1]
    assert old | active == cur and not (old & active)
    assert h["rmse"] <= 2e-2 or cur == space


def test_adaptive_lognormal2d_multi_index_run():
    """config C4 at test size: AD(2)/QMC on the anisotropic 2-D lognormal problem (multi-index differences), oracle backend"""
    import mle_b200.kl as kl
    H = host()
    m = 20
    bases = {i: kl.kl_basis_2d(i, m=m) for i in H.TD(2).get_index_set(2)}
    spec = H.ModelSpec("lognormal2d", 2, [4], bases)
    est = H.Estimator(H.AD(2), H.QMC(), spec, [H.Normal()] * m, OracleBackend(), max_index_set_param=2, nb_of_shifts=4,
                      nb_of_warm_up_samples=4, cost_model=lambda i: 2.0 ** sum(i), nb_of_tols=2, min_index_set_param=1)
    h = H.run(est, 5e-3)
    assert 0.05 < h["mean"] < 0.09 and h["ndims"] == 2
    assert all(abs(h["dE"][i]) < abs(h["dE"][(0, 0)]) for i in h["index_set"] if i != (0, 0))


def test_save_samples_history_keys():
    """src/history.jl:92-95: with save_samples the History carries the raw samples; mean / var recomputed from them the
    reference's way (src/methods.jl:48-57) agree with the moment-based statistics"""
    H = host()
    est = H.Estimator(H.ML(), H.QMC(), H.ModelSpec("problem18", 1), [H.Uniform(-0.5, 0.5)], OracleBackend(), nb_of_qoi=13,
                      max_index_set_param=2, cost_model=lambda i: 2.0 ** i[0], nb_of_tols=2, nb_of_shifts=4, save_samples=True)
    h = H.run(est, 5e-2)
    assert "samples" in h and "samples_diff" in h
    for idx in h["index_set"]:
        sd = h["samples_diff"][idx]                        # [R, q, n]
        assert sd.shape[0] == 4 and sd.shape[1] == 13 and sd.shape[2] == h["nb_of_samples"][idx]
        assert abs(sd[:, 0, :].mean(axis=1).mean() - h["dE"][idx]) <= 1e-12 * max(1.0, abs(h["dE"][idx]))
        assert abs(h["samples"][idx][:, 0, :].mean(axis=1).mean() - h["E"][idx]) <= 1e-12 * max(1.0, abs(h["E"][idx]))


@pytest.mark.parametrize("method", ["MC", "QMC"])
def test_regression_problem18_unbiased(method):
    """test/regression.jl:128-160: run(Estimator(U(1), MC()/QMC(), test_function2, Uniform(-0.5, 0.5); nb_of_qoi = 13,
    max_index_set_param = 3, max_search_space = ML()), tol): every idx <= the drawn index is evaluated on the same inputs, the
    per-sample sums of dQ_idx / Prob(idx) are the unbiased estimator; bias = 0 in the History"""
    H = host()
    with pytest.raises(ValueError):
        H.U(0)
    with pytest.raises(NotImplementedError):
        H.U(1).get_index_set(2)
    est = H.Estimator(H.U(1), getattr(H, method)(), H.ModelSpec("problem18", 1), [H.Uniform(-0.5, 0.5)], OracleBackend(),
                      nb_of_qoi=13, max_index_set_param=3, max_search_space=H.ML(), cost_model=lambda i: 2.0 ** i[0],
                      nb_of_tols=5, save_samples=True)
    assert est["nb_of_warm_up_samples"] == (20 if method == "MC" else 2)          # src/parse.jl:36
    assert abs(sum(est.pmf.values()) - 1.0) < 1e-15 and abs(est.Prob((0,)) - 1.0) < 1e-15
    p = [est.pmf[(l,)] for l in range(4)]
    assert all(abs(p[l + 1] / p[l] - math.exp(-1.5)) < 1e-12 for l in range(3))    # Geometric(1 - exp(-1.5)), src/internals.jl:221
    h = H.run(est, 2e-3)
    assert len(h) == 5 and h["bias"] == 0.0 and h["type"] == "Estimator{U, %s}" % method
    assert h["rmse"] <= 2e-3 and abs(h["mean"] - 4.0) < 6 * 2e-3
    assert abs(sum(est.pmf.values()) - 1.0) < 1e-12
    # the samples of idx come from every call at an index >= idx (src/sample.jl:65-74)
    n_at_or_above = {i: sum(n for j, n in h["nb_of_samples"].items() if j[0] >= i[0]) for i in h["index_set"]}
    assert all(h["samples_diff"][i].shape[-1] == n_at_or_above[i] for i in h["index_set"])


def test_unbiased_multi_index_run():
    """test/regression.jl:163-181: U(2) with the default TD(2) search space"""
    H = host()
    est = H.Estimator(H.U(2), H.QMC(), H.ModelSpec("problem18", 2), [H.Uniform(-0.5, 0.5)] * 2, OracleBackend(), nb_of_qoi=13,
                      max_index_set_param=3, cost_model=lambda i: 2.0 ** sum(i), nb_of_tols=3, nb_of_shifts=8)
    assert set(est.pmf) == set(H.TD(2).get_index_set(3))
    # Prob(index) sums the pmf over the indices at or after `index` in column-major order (CartesianIndex `>=`, src/methods.jl:333)
    assert abs(est.Prob((0, 1)) - sum(v for k, v in est.pmf.items() if (k[1], k[0]) >= (1, 0))) < 1e-15
    h = H.run(est, 1e-2)
    assert h["rmse"] <= 1e-2 and h["ndims"] == 2 and h["bias"] == 0.0
    with pytest.raises(ValueError):
        H.Estimator(H.U(1), H.MC(), H.ModelSpec("problem18", 1), [H.Uniform(-0.5, 0.5)], OracleBackend())   # TD(1) does not exist


_AGG = [("max", 1e-2), ("sum", 1e-1), ("norm", 1e-1)]   # test/regression.jl:12-14


@pytest.mark.parametrize("method", ["MC", "QMC"])
@pytest.mark.parametrize("iset", ["TD", "FT", "HC", "ZC", "AD"])
@pytest.mark.parametrize("aggregation,tol", _AGG)
def test_reference_regression_matrix_multi_index(iset, method, aggregation, tol, tmp_path):
    """test/regression.jl:106-125, every cell: run(Estimator(T(2), MC()/QMC(), problem_18, 2 x Uniform(-0.5, 0.5); nb_of_qoi = 13,
    max_index_set_param = 3, save_samples = true, folder = tempdir(), aggregation, qoi_with_max_var = 3), tol) with the default
    (timing) cost model; the reference asserts completion only, here the History is checked as well"""
    H = host()
    est = H.Estimator(getattr(H, iset)(2), getattr(H, method)(), H.ModelSpec("problem18", 2), [H.Uniform(-0.5, 0.5)] * 2,
                      OracleBackend(), nb_of_qoi=13, max_index_set_param=3, save_samples=True, folder=str(tmp_path),
                      aggregation=aggregation, qoi_with_max_var=3)
    h = H.run(est, tol)
    assert len(h) == 10 and h["ndims"] == 2 and h["type"] == "Estimator{%s, %s}" % (iset, method)
    assert h["folder"] == str(tmp_path) and not os.listdir(tmp_path)          # save defaults to false in this mirror
    assert math.isfinite(h["rmse"]) and math.isfinite(h["mean"]) and h["varest"] >= 0
    full = set(est.search_space if iset == "AD" else getattr(H, iset)(2).get_index_set(3))
    assert set(h["current_index_set"]) <= full and (0, 0) in h["current_index_set"]
    assert h["rmse"] <= 2 * tol or est.max_level_exceeded() or (iset == "AD" and len(h["current_index_set"]) > 1)
    for idx in h["current_index_set"]:
        n = h["nb_of_samples"][idx]
        assert n >= 1 and h["samples"][idx].shape == h["samples_diff"][idx].shape
        assert h["samples"][idx].shape[1:] == (13, n)
    if iset == "AD":
        assert h["logbook"]


@pytest.mark.parametrize("method", ["MC", "QMC"])
@pytest.mark.parametrize("d", [1, 2])
@pytest.mark.parametrize("aggregation,tol", _AGG)
def test_reference_regression_matrix_unbiased(d, method, aggregation, tol):
    """test/regression.jl:140-178, every cell: U(1) (max_search_space = ML()) and U(2) x MC/QMC x the three aggregations with
    test_function2 (all idx <= index on the same inputs)"""
    H = host()
    kw = dict(max_search_space=H.ML()) if d == 1 else {}
    est = H.Estimator(H.U(d), getattr(H, method)(), H.ModelSpec("problem18", d), [H.Uniform(-0.5, 0.5)] * d, OracleBackend(),
                      nb_of_qoi=13, max_index_set_param=3, save_samples=True, aggregation=aggregation, **kw)
    h = H.run(est, tol)
    assert len(h) == 10 and h["bias"] == 0.0 and h["rmse"] <= tol * (1 + 1e-12)
    # an index whose differences have zero sample variance gets probability 0 (f = sqrt(V/C) = 0, src/methods.jl:311-318)
    assert abs(sum(est.pmf.values()) - 1) < 1e-12 and all(p >= 0 for p in est.pmf.values()) and est.pmf[(0,) * d] > 0
    # first batch: ceil(max(nb_of_warm_up_samples, 0) * (sample_mul_factor - 1)) draws, src/methods.jl:287-290
    assert math.isfinite(h["mean"]) and sum(h["nb_of_samples"].values()) >= (1 if method == "QMC" else 4)


def test_mc_streams_are_independent_across_indices():
    """the reference draws every MC input from one global stream (src/sample.jl:42): samples on different levels are
    independent.  The device keys its uniforms by (seed, point number, coordinate), so the host gives every index its own seed"""
    H = host()
    seeds = {H.index_seed(1, i) for i in [(0,), (1,), (2,), (0, 0), (1, 0), (0, 1), (1, 1), (2, 1)]}
    assert len(seeds) == 8 and all(0 <= x < 2**64 for x in seeds)
    assert H.index_seed(1, (3,)) != H.index_seed(2, (3,)) and H.index_seed(7, (2, 1)) == H.index_seed(7, (2, 1))
    w = [0.5 / (j + 1) for j in range(16)]
    est = H.Estimator(H.ML(), H.MC(), H.ModelSpec("expsum", 1, [16] + w), [H.Normal()] * 16, OracleBackend(),
                      save_samples=True, max_index_set_param=3)
    est.sample((0,), 4000)
    est.sample((1,), 4000)
    q0 = np.concatenate(est.samples[(0,)], axis=-1)[0, 0, 1, :]
    q1 = np.concatenate(est.samples[(1,)], axis=-1)[0, 0, 1, :]
    # m_l = min(16, 16 * 2^l) = 16 on both levels: with a shared stream the fine values would be IDENTICAL sample by sample
    assert not np.array_equal(q0, q1) and abs(np.corrcoef(q0, q1)[0, 1]) < 0.06
    # continuing an index continues its own stream (point numbers go on, no repeats)
    est.sample((0,), 4000)
    q0b = np.concatenate(est.samples[(0,)], axis=-1)[0, 0, 1, :]
    assert len(np.unique(q0b)) == 8000


def test_unbiased_point_numbers_count_all_indices():
    """src/sample.jl:25: for U estimators Istart = sum of the samples taken on ALL indices + 1"""
    H = host()

    class Spy(OracleBackend):
        calls = []

        def sample_batch(self, model, lattice, dists, index, shifts, k0, n, m, mc_seed, want_samples=False):
            Spy.calls.append((tuple(index), k0, n))
            return OracleBackend.sample_batch(self, model, lattice, dists, index, shifts, k0, n, m, mc_seed, want_samples)

    est = H.Estimator(H.U(1), H.QMC(), H.ModelSpec("problem18", 1), [H.Uniform(-0.5, 0.5)], Spy(), nb_of_qoi=13,
                      max_search_space=H.ML(), max_index_set_param=3, nb_of_shifts=4)
    est.sample((0,), 5)
    est.sample((2,), 3)
    est.sample((0,), 2)
    assert Spy.calls == [((0,), 0, 5), ((0,), 5, 3), ((1,), 5, 3), ((2,), 5, 3), ((0,), 8, 2)]
    assert est.nb_of_samples[(0,)] == 7 and est.nb_of_samples[(2,)] == 3


def test_batched_update_samples_equals_index_by_index():
    """with a deterministic cost model all due indices of one update_samples (src/sample.jl:8-13) go to the backend in one
    batched call (QMC only: MC needs one stream per index); the run is the same as index by index"""
    H = host()

    class OneByOne(OracleBackend):
        update_samples = property()   # hasattr(...) is False: the host falls back to sample! per index

    def make(backend, method="QMC"):
        return H.Estimator(H.ML(), getattr(H, method)(), H.ModelSpec("problem18", 1), [H.Uniform(-0.5, 0.5)], backend, nb_of_qoi=13,
                           max_index_set_param=3, cost_model=lambda i: 2.0 ** i[0], nb_of_tols=4,
                           **(dict(nb_of_shifts=4) if method == "QMC" else {}))

    assert not hasattr(OneByOne(), "update_samples")
    b1, b2 = OracleBackend(), OneByOne()
    e1, e2 = make(b1), make(b2)
    # the QMC loop grows one index at a time, so drive update_samples directly with several due indices
    for e in (e1, e2):
        e.current_index_set.update([(0,), (1,), (2,)])
        e.update_samples({(0,): 8, (1,): 5, (2,): 3})
        e.update_samples({(0,): 8, (1,): 9, (2,): 4})      # (0,) is not due
    assert getattr(b1, "batched_calls", 0) == 2
    assert e1.nb_of_samples == e2.nb_of_samples and e1.total_work == e2.total_work
    for idx in ((0,), (1,), (2,)):
        assert np.array_equal(e1.moments[idx], e2.moments[idx])
    assert e1.total_time[(2,)] >= 0 and e1.mean() == e2.mean() and e1.varest() == e2.varest()
    # MC never takes the batched path
    b3 = OracleBackend()
    H.run(make(b3, "MC"), 2e-2)
    assert getattr(b3, "batched_calls", 0) == 0


def _drive_batched(H, backend):
    est = H.Estimator(H.ML(), H.QMC(), H.ModelSpec("problem18", 1), [H.Uniform(-0.5, 0.5)], backend, nb_of_qoi=13,
                      max_index_set_param=3, cost_model=lambda i: 2.0 ** i[0], nb_of_shifts=4)
    est.current_index_set.update([(0,), (1,), (2,), (3,)])
    est.update_samples({(0,): 9, (1,): 5, (2,): 1, (3,): 2})     # (2,): one point -> rank 1 has nothing to do for it
    est.update_samples({(0,): 20, (1,): 5, (2,): 4, (3,): 3})
    return est


def _worker_batched(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import mle_b200.host as H
    inner = OracleBackend()
    est = _drive_batched(H, H.ShardedBackend(inner))
    assert inner.batched_calls == 2
    if rank == 0:
        np.save(out, np.stack([est.moments[(l,)] for l in range(4)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_batched_update_samples(tmp_path):
    """ShardedBackend.update_samples: every rank runs its slice of every due index through the inner backend's batched call and
    one all-gather merges the moment tables of all entries"""
    import torch.multiprocessing as mp
    H = host()
    want = _drive_batched(H, OracleBackend())
    out = str(tmp_path / "b.npy")
    mp.spawn(_worker_batched, args=(2, 29543, out), nprocs=2, join=True)
    got = np.load(out)
    for l in range(4):
        w = want.moments[(l,)]
        assert np.array_equal(got[l][..., 0], w[..., 0])
        assert np.allclose(got[l][..., 1], w[..., 1], rtol=1e-13, atol=0)
        assert np.allclose(got[l][..., 2], w[..., 2], rtol=1e-10, atol=1e-28)


def test_warnings_of_the_reference_are_kept():
    """src/print.jl:241-247: `Maximum level L = .. reached, no convergence yet.` (verbose only, src/run.jl:126) and `Trying to add
    index .. outside search space` (always, src/methods.jl:242)"""
    import warnings
    H = host()
    est = H.Estimator(H.ML(), H.MC(), H.ModelSpec("problem18", 1), [H.Uniform(-0.5, 0.5)], OracleBackend(), nb_of_qoi=13,
                      max_index_set_param=1, min_index_set_param=1, verbose=True)
    with pytest.warns(UserWarning, match="Maximum level L = 1 reached, no convergence yet."):
        H.run(est, [1e-2])   # convergence is only tested for L > min_index_set_param (src/run.jl:124)
    est = H.Estimator(H.ML(), H.MC(), H.ModelSpec("problem18", 1), [H.Uniform(-0.5, 0.5)], OracleBackend(), nb_of_qoi=13,
                      max_index_set_param=1, min_index_set_param=1)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        H.run(est, [1e-2])                     # silent without verbose
    ad = H.Estimator(H.AD(2), H.MC(), H.ModelSpec("problem18", 2), [H.Uniform(-0.5, 0.5)] * 2, OracleBackend(), nb_of_qoi=13,
                     max_index_set_param=1, min_index_set_param=1, max_search_space=H.TD(2))
    for idx in ((0, 0), (1, 0), (0, 1)):
        ad.sample(idx, 4)
    ad.old_set, ad.active_set = {(0, 0)}, {(1, 0), (0, 1)}
    with pytest.warns(UserWarning, match="outside search space"):
        new = ad.new_index_set(2)      # the most profitable active index moves to the old set; its forward neighbour (2, 0)
    assert len(new) == 1 and ad.max_index_set and ad.max_index_set <= {(2, 0), (0, 2), (1, 1)}   # or (0, 2) is outside TD(2) at size 1


def test_verbose_report_follows_print_jl(capsys):
    """verbose = true: the tables and messages of src/print.jl (header :34-45, status table :59-83, sample lines :214-222,
    optimal numbers :97-121, QMC / MSE checks :186-202, AD profits :132-151, U pmf :156-175, footer :47-54); the run itself is
    unchanged by it"""
    H = host()

    def run(iset, method, d, tol, **kw):
        def make(verbose):
            return H.Estimator(iset, getattr(H, method)(), H.ModelSpec("problem18", d), [H.Uniform(-0.5, 0.5)] * d, OracleBackend(),
                               nb_of_qoi=13, max_index_set_param=3, cost_model=lambda i: 2.0 ** sum(i), name="demo",
                               verbose=verbose, **kw)
        quiet = H.run(make(False), [tol])
        assert capsys.readouterr().out == ""
        loud = H.run(make(True), [tol])
        out = capsys.readouterr().out
        assert loud["nb_of_samples"] == quiet["nb_of_samples"] and loud["mean"] == quiet["mean"]
        return out, loud

    out, h = run(H.ML(), "MC", 1, 1e-2)
    lines = out.splitlines()
    assert lines[0] == "+" + "-" * 79 + "+" and all(len(ln) == 81 for ln in lines[:6])
    assert lines[1].startswith("| *** MultilevelEstimators.jl @") and lines[1].endswith("|")
    assert lines[2].startswith("| *** This is a Estimator{ML, MC}") and lines[3].startswith("| *** Simulating demo ")
    assert lines[4].startswith("| *** Tolerance on RMSE ϵ = 1.000e-02")
    assert "Currently running on level 0." in lines and "Taking 20 warm-up samples at level (0)... done" in lines
    head = "| level         | E             | dE            | V             | dV            | N             | W             | "
    assert head in lines and "+" + ("-" * 15 + "+") * 7 in lines
    row = [ln for ln in lines if ln.startswith("| (0)")][-1]
    assert row == "| (0)           |%s   |%s   |%s   |%s   | %-13s |%s   |" % (
        "%12.5e" % h["E"][(0,)], "%12.5e" % h["dE"][(0,)], "%12.5e" % h["V"][(0,)], "%12.5e" % h["dV"][(0,)],
        h["nb_of_samples"][(0,)], "%12.5e" % 1.0)
    assert "Samples will be updated to" in lines and "| level         | N             | " in lines
    assert any(ln.startswith("Taking ") and " additional sample" in ln for ln in lines)
    assert "Checking convergence..." in lines and any(ln.startswith("  ==> Rates: α ≈ ") for ln in lines)
    assert any(ln.startswith("  ==> MSE splitting parameter ≈ 0.") for ln in lines)
    assert lines[-4] == "+" + "-" * 79 + "+" and lines[-2].startswith("| *** Successfull termination.") and len(lines[-2]) == 81
    assert any(ln.startswith("Convergence reached. RMSE ≈") for ln in lines) or h["rmse"] > 1e-2

    out, h = run(H.TD(2), "QMC", 2, 2e-2, nb_of_shifts=4)
    assert "Currently running with L = 1." in out and "Shape of the index set:" in out and "\u25FC" in out
    assert "Taking 4 x 1 warm-up sample at index (1, 0)... done" in out
    assert "Checking if variance of estimator is larger than θ*ϵ^2..." in out and "  ==> " in out
    assert "| index         | E " in out and "Estimator{TD{2}, QMC}" in out

    out, h = run(H.AD(2), "MC", 2, 2e-2)
    assert "Shape of the index set (\u25A3 = active set):" in out and "Profit indicators:" in out
    assert "| index         | profit        | " in out and "Index with max profit is (" in out

    out, h = run(H.U(1), "MC", 1, 2e-2, max_search_space=H.ML())
    assert "Using approximate probability mass function:" in out and "| level         | P             | " in out
    assert "Checking if variance of estimator is larger than ϵ^2..." in out and "Convergence reached. RMSE ≈" in out

    out, h = run(H.SL(), "MC", 1, 5e-2)
    assert "Currently running on finest level." in out and "Checking convergence..." not in out


@pytest.mark.parametrize("case", ["ml_qmc", "td_qmc", "ad_qmc", "ml_mc"])
def test_memoisation_does_not_change_a_run(case):
    """the host loop remembers statistics between queries (per-index moments, rates, bias estimates -- the latter invalidated only
    by new samples on the indices they were computed from, sorted key lists); a run with every cache switched off must take the
    same decisions and return the same History"""
    H = host()

    def make():
        if case == "ml_qmc":
            return H.Estimator(H.ML(), H.QMC(), H.ModelSpec("problem18", 1), [H.Uniform(-0.5, 0.5)], OracleBackend(), nb_of_qoi=13,
                               nb_of_shifts=8, cost_model=lambda i: 2.0 ** i[0], shift_seed=3, max_index_set_param=5)
        if case == "td_qmc":
            return H.Estimator(H.TD(2), H.QMC(), H.ModelSpec("problem18", 2), [H.Uniform(-0.5, 0.5)] * 2, OracleBackend(), nb_of_qoi=13,
                               nb_of_shifts=8, cost_model=lambda i: 2.0 ** sum(i), shift_seed=3, max_index_set_param=4)
        if case == "ad_qmc":
            return H.Estimator(H.AD(2), H.QMC(), H.ModelSpec("problem18", 2), [H.Uniform(-0.5, 0.5)] * 2, OracleBackend(), nb_of_qoi=13,
                               nb_of_shifts=8, cost_model=lambda i: 2.0 ** sum(i), shift_seed=3, max_index_set_param=4)
        return H.Estimator(H.ML(), H.MC(), H.ModelSpec("problem18", 1), [H.Uniform(-0.5, 0.5)], OracleBackend(), nb_of_qoi=13,
                           cost_model=lambda i: 2.0 ** i[0], mc_seed=5, max_index_set_param=5)
    a, b = make(), make()
    b._no_memo = True
    ha, hb = H.run(a, 3e-4), H.run(b, 3e-4)   # ten tolerances, ~70 sample! calls on four levels (ml_qmc)
    assert len(ha) == len(hb)
    for i in range(len(ha)):
        for key in ("nb_of_samples", "current_index_set", "mean", "varest", "bias", "rmse", "alpha", "beta", "gamma", "dE", "dV"):
            assert repr(ha[i][key]) == repr(hb[i][key]), (i, key)
    if case == "ad_qmc":
        assert ha["logbook"] == hb["logbook"]
