"""GPU parity, stages 1-2 (through the C ABI): lattice points bit-exact against the oracle and the reference's golden
vectors; transformed inputs within 2 ulp (north star) -- in practice bit-exact for every distribution: the elementary functions
the reference takes from Julia's runtime (log1p in the erfinv tail, log and ^ in Weibull) are the same table algorithms on
both sides (include/mle_math_tables.inc)."""
import numpy as np
import pytest

from conftest import splitmix_uniforms, ulp_diff

pytestmark = pytest.mark.gpu

NORMAL01 = (1, [0.0, 1.0])


def f(strs):
    return np.array([float(s) for s in strs])


def test_golden_G1_G2_G3(engine, goldens):
    g = goldens["G1"]
    lat = engine.lattice(g["z"], g["s"], g["nmax"])
    assert np.array_equal(engine.points(lat, None, None, g["k"], 1)[0, 0], f(g["point"]))
    g = goldens["G2"]
    lat = engine.lattice(None, g["s"])  # default rule = first 16 entries of K_3600_32
    assert np.array_equal(engine.points(lat, None, None, g["k"], 1)[0, 0], f(g["point"]))
    g = goldens["G3"]
    shift = f(g["point_equals_shift"])
    assert np.array_equal(engine.points(lat, None, shift[None, :], 0, 1)[0, 0], shift)


def test_golden_G4_normal_transform(engine, goldens):
    """transform(Normal(), x) for the doctest's 10 inputs, fed through point 0 of a lattice shifted by x"""
    g = goldens["G4"]
    x, y = f(g["x"]), f(g["y"])
    lat = engine.lattice(None, 10)
    d = engine.dists([NORMAL01] * 10)
    got = engine.points(lat, d, x[None, :], 0, 1)[0, 0]
    assert ulp_diff(got, y).max() <= 2   # north-star tolerance
    assert np.array_equal(got, y)        # in fact bit-exact, all three erfinv branches


@pytest.mark.parametrize("key", ["G5", "G6", "G7"])
def test_golden_G5_G7(engine, goldens, key):
    g = goldens[key]
    x, y = f(g["x"]), f(g["y"])
    lat = engine.lattice(None, 9)
    d = engine.dists([(g["kind"], g["pars"])] * 9)
    got = engine.points(lat, d, x[None, :], 0, 1)[0, 0]
    assert np.allclose(got, y, rtol=1.5e-8, atol=1e-15)  # the reference's own tolerance (isapprox)
    assert np.max(np.abs(got - y)) < 2e-15


@pytest.mark.parametrize("s,R,k0,n", [(1, 1, 0, 1), (2, 3, 5, 7), (16, 4, 0, 1000), (100, 16, 12345, 4099),
                                      (101, 3, 1, 41), (250, 2, 2**31, 300), (3600, 2, 2**32 - 40, 40)])
def test_shifted_points_bit_exact(engine, oracle, s, R, k0, n):
    z = oracle.genvec("K_3600_32", s)
    lat = engine.lattice(z, s)
    sh = splitmix_uniforms(100 + s, R * s).reshape(R, s)
    got = engine.points(lat, None, sh, k0, n)
    want = oracle.points(z, None, sh, k0, n)
    assert got.shape == (R, n, s)
    assert np.array_equal(got, want)
    assert (got >= 0).all() and (got < 1).all()


def test_unshifted_points_bit_exact_and_ckn(engine, oracle):
    z = oracle.genvec("CKN_250_20", 250)
    lat = engine.lattice(z, 250, 2**20)
    got = engine.points(lat, None, None, 0, 2000)
    assert np.array_equal(got, oracle.points(z, None, None, 0, 2000, nmax=2**20))
    assert (got[0, 0] == 0).all()


def test_large_generating_vector_entries(engine, oracle):
    """z >= 2^21: the product phi*2^-32*z is no longer exact, so the separately rounded multiply/add order matters"""
    rng = np.random.default_rng(5)
    z = rng.integers(2**28, 2**32, size=33, dtype=np.uint64).astype(np.uint32)
    lat = engine.lattice(z, 33)
    sh = splitmix_uniforms(9, 2 * 33).reshape(2, 33)
    assert np.array_equal(engine.points(lat, None, sh, 777, 513), oracle.points(z, None, sh, 777, 513))


def test_beyond_n_and_padding_are_refused(engine, goldens):
    import mle_b200
    g = goldens["G1"]
    lat = engine.lattice(g["z"], g["s"], g["nmax"])
    with pytest.raises(mle_b200.MLEArgumentError):
        engine.points(lat, None, None, 56, 1)       # src/lattice.jl:41-44: rand(s) beyond n
    with pytest.raises(mle_b200.MLEArgumentError):
        engine.points(lat, None, None, 0, 1, m=3)   # m > s: the reference pads with rand(), src/sample.jl:85
    with pytest.raises(mle_b200.MLEArgumentError):
        engine.lattice([1, 2], 0)                   # s must be > 0, src/lattice.jl:182
    with pytest.raises(mle_b200.MLEArgumentError):
        engine.lattice(None, 3601)                  # src/lattice.jl:187
    with pytest.raises(mle_b200.MLEArgumentError):
        engine.dists([(1, [0.0, -1.0])])            # Normal sigma <= 0, test/distribution.jl:33
    with pytest.raises(mle_b200.MLEArgumentError):
        engine.dists([(0, [1.0, 0.0])])             # Uniform a >= b


def test_empty_batch(engine):
    lat = engine.lattice(None, 4)
    assert engine.points(lat, None, None, 0, 0).shape == (1, 0, 4)


@pytest.mark.parametrize("s,R,n", [(16, 2, 1000), (100, 16, 2048), (7, 1, 5000)])
def test_normal_transform_within_2ulp(engine, oracle, s, R, n):
    z = oracle.genvec("K_3600_32", s)
    lat = engine.lattice(z, s)
    dl = [NORMAL01] * s
    d = engine.dists(dl)
    sh = splitmix_uniforms(2024, R * s).reshape(R, s)
    got = engine.points(lat, d, sh, 3, n)
    want = oracle.points(z, dl, sh, 3, n)
    assert np.isfinite(got).all()
    ulps = ulp_diff(got, want)
    assert ulps.max() <= 2               # north-star tolerance
    assert np.array_equal(got, want)     # in fact bit-exact (same Horner order, IEEE division, shared log(1-|u|))
    u = 2 * oracle.points(z, None, sh, 3, n) - 1
    assert (np.abs(u) > 0.9375).any() and (np.abs(u) <= 0.75).any()


def test_point_zero_unshifted_is_minus_inf(engine):
    """x = 0 -> Phi^-1(0) = -Inf like the reference"""
    lat = engine.lattice(None, 5)
    d = engine.dists([NORMAL01] * 5)
    got = engine.points(lat, d, None, 0, 2)
    assert np.isneginf(got[0, 0]).all() and np.isfinite(got[0, 1]).all()


def test_general_distribution_mix(engine, oracle):
    s = 12
    z = oracle.genvec("K_3600_32", s)
    lat = engine.lattice(z, s)
    dl = [(0, [-0.5, 0.5]), (1, [1.5, 2.0]), (2, [0.0, 1.0, -2.0, 2.0]), (3, [2.0, 2.0**0.5])] * 3
    d = engine.dists(dl)
    sh = splitmix_uniforms(77, 3 * s).reshape(3, s)
    got = engine.points(lat, d, sh, 0, 1500)
    want = oracle.points(z, dl, sh, 0, 1500)
    kinds = np.array([k for k, _ in dl])
    assert np.array_equal(got[..., kinds == 0], want[..., kinds == 0])        # Uniform: bit-exact
    assert np.array_equal(got[..., kinds == 1], want[..., kinds == 1])        # Normal(mu, sigma): bit-exact
    assert np.array_equal(got[..., kinds == 2], want[..., kinds == 2])        # TruncatedNormal: bit-exact
    assert np.array_equal(got[..., kinds == 3], want[..., kinds == 3])        # Weibull: shared log / pow algorithms, bit-exact
    assert (np.abs(got[..., kinds == 2]) <= 2).all()


def test_weibull_shapes_and_small_arguments(engine, oracle):
    """Weibull(k, lambda) for shape parameters below and above 1 (1/k amplifies or damps the logarithm's rounding), on lattice
    points and on tiny / near-1 uniforms through the MC stream; bit-exact against the oracle, whose log / pow are checked against
    exact values in tests/test_oracle_golden.py"""
    s = 10
    z = oracle.genvec("K_3600_32", s)
    lat = engine.lattice(z, s)
    dl = [(3, [k, lam]) for k, lam in ((0.5, 1.0), (0.75, 2.0), (1.0, 1.0), (1.5, 0.3), (2.0, 2.0 ** 0.5), (3.7, 1.0), (5.0, 10.0),
                                       (10.0, 1.0), (0.3, 1.0), (2.5, 1.5))]
    d = engine.dists(dl)
    sh = splitmix_uniforms(5, 2 * s).reshape(2, s)
    got = engine.points(lat, d, sh, 0, 4000)
    want = oracle.points(z, dl, sh, 0, 4000)
    assert np.array_equal(got, want)
    assert (got >= 0).all() and np.isfinite(got).all()
    got0 = engine.points(lat, d, None, 0, 3)      # unshifted point 0 is x = 0: lambda * 0^(1/k) = 0
    assert (got0[0, 0] == 0).all()


def test_m_smaller_than_s(engine, oracle):
    """nb_of_uncertainties(index) < s: only the leading m coordinates are used (src/sample.jl:86)"""
    z = oracle.genvec("K_3600_32", 50)
    lat = engine.lattice(z, 50)
    dl = [NORMAL01] * 50
    d = engine.dists(dl)
    sh = splitmix_uniforms(5, 2 * 50).reshape(2, 50)
    got = engine.points(lat, d, sh, 0, 333, m=20)
    want = oracle.points(z, dl, sh, 0, 333)[..., :20]
    assert got.shape == (2, 333, 20) and np.array_equal(got, want)


def test_mc_stream_matches_oracle(engine, oracle):
    dl = [NORMAL01] * 16
    d = engine.dists(dl)
    raw = engine.points(None, None, None, 10, 4000, m=16, mc_seed=1)
    assert np.array_equal(raw, oracle.points(None, None, None, 10, 4000, m=16, mc_seed=1))
    assert (raw > 0).all() and (raw < 1).all()
    got = engine.points(None, d, None, 10, 4000, m=16, mc_seed=1)
    assert np.array_equal(got, oracle.points(None, dl, None, 10, 4000, m=16, mc_seed=1))
    assert abs(got.mean()) < 0.02 and abs(got.std() - 1) < 0.02


def test_full_size_property_mean_and_range(engine):
    """a BASELINE-size sweep slice (2^20 points x 100 dims x 1 shift kept on the device is covered by bench.py);
    here: 2^16 x 100 x 2 on the host -- every coordinate in [0,1), lattice mean = 1/2 - 2^-17 per coordinate up to the shift"""
    lat = engine.lattice(None, 100)
    sh = splitmix_uniforms(1, 200).reshape(2, 100)
    pts = engine.points(lat, None, sh, 0, 2**16)
    assert (pts >= 0).all() and (pts < 1).all()
    # the first 2^16 points of a shifted lattice sequence are equispaced (spacing 2^-16) in every coordinate with odd z
    srt = np.sort(pts[0, :, 0])
    assert np.allclose(np.diff(srt), 2.0**-16, atol=1e-10)
