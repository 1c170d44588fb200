"""GPU parity, stages 3-4 (through the C ABI): per-(qoi, shift) moments of one `parallel_sample!` call against the
oracle run on identical points.  North star: level means and variances within 1e-12 relative."""
import numpy as np
import pytest

from conftest import splitmix_uniforms

pytestmark = pytest.mark.gpu

NORMAL01 = (1, [0.0, 1.0])
RTOL = 1e-12  # north star: means and variances within 1e-12 relative for identical points


def check_moments(got, want, smp_got=None, smp_want=None, rtol=RTOL):
    assert got.shape == want.shape
    assert np.array_equal(got[..., 0], want[..., 0])
    scale = np.sqrt(want[..., 2] / np.maximum(want[..., 0], 1)) + np.abs(want[..., 1])  # mean error relative to |mean| + std
    assert np.all(np.abs(got[..., 1] - want[..., 1]) <= rtol * scale)
    assert np.all(np.abs(got[..., 2] - want[..., 2]) <= rtol * np.abs(want[..., 2]) + 1e-300)
    if smp_got is not None:
        # per-sample (dQ, Qf): same operations in the same order on both sides (including log/exp) -> bit-identical
        assert np.array_equal(smp_got, smp_want)


def expsum_setup(engine, oracle, s, m0=4):
    z = oracle.genvec("K_3600_32", s)
    w = [0.5 / (j + 1) for j in range(s)]
    return (z, engine.lattice(z, s), [NORMAL01] * s, engine.dists([NORMAL01] * s),
            engine.model("expsum", 1, [m0] + w), oracle.Model(oracle.MODEL_EXPSUM, 1, [m0] + w))


@pytest.mark.parametrize("s,R,level,k0,n", [(16, 1, 0, 0, 1), (16, 4, 0, 0, 2), (16, 16, 2, 0, 127), (100, 16, 3, 0, 128),
                                            (100, 3, 6, 1000, 129), (100, 16, 1, 7, 5000), (130, 2, 9, 0, 777)])
def test_expsum_qmc(engine, oracle, s, R, level, k0, n):
    z, lat, dl, d, gm, om = expsum_setup(engine, oracle, s)
    sh = splitmix_uniforms(2024, R * s).reshape(R, s)
    got, smp = engine.sample_batch(gm, lat, d, [level], sh, k0, n, want_samples=True)
    want, smp_w = oracle.sample_batch(om, z, dl, [level], sh, k0, n, want_samples=True)
    if n > 1:
        check_moments(got, want, smp, smp_w)
    else:
        assert np.allclose(got[..., 1], want[..., 1], rtol=1e-14) and (got[..., 2] == 0).all()


def test_expsum_mc_and_closed_form(engine, oracle):
    """config C1: SL/MC on 16 Normal() inputs, Q = exp(sum a_j xi_j), E[Q] = exp(sum a_j^2 / 2)"""
    s = 16
    w = [0.5 / (j + 1) for j in range(s)]
    d = engine.dists([NORMAL01] * s)
    gm = engine.model("expsum", 1, [s] + w)
    om = oracle.Model(oracle.MODEL_EXPSUM, 1, [s] + w)
    n = 200000
    got = engine.sample_batch(gm, None, d, [0], None, 0, n, mc_seed=1)
    want = oracle.sample_batch(om, None, [NORMAL01] * s, [0], None, 0, n, mc_seed=1)
    check_moments(got, want)
    exact = np.exp(0.5 * np.sum(np.square(w)))
    se = np.sqrt(got[0, 0, 1, 2] / (n - 1) / n)
    assert abs(got[0, 0, 1, 1] - exact) < 5 * se


def test_batches_merge_like_one_call(engine, oracle):
    """samples persist across calls in the reference: merging the moments of consecutive k-ranges
    must equal the moments of the union"""
    import mle_b200
    z, lat, dl, d, gm, om = expsum_setup(engine, oracle, 32)
    sh = splitmix_uniforms(4, 5 * 32).reshape(5, 32)
    whole = engine.sample_batch(gm, lat, d, [2], sh, 0, 3000)
    acc = engine.sample_batch(gm, lat, d, [2], sh, 0, 1)
    for k0, n in [(1, 9), (10, 990), (1000, 2000)]:
        acc = mle_b200.moments_merge(acc, engine.sample_batch(gm, lat, d, [2], sh, k0, n))
    check_moments(acc, whole)
    check_moments(acc, oracle.sample_batch(om, z, dl, [2], sh, 0, 3000))


@pytest.mark.parametrize("ndims,index,R,n", [(1, [0], 4, 300), (1, [3], 16, 1000), (2, [0, 0], 2, 257), (2, [2, 1], 4, 640),
                                             (2, [3, 3], 3, 100), (2, [0, 2], 1, 50)])
def test_problem18_reference_test_function(engine, oracle, ndims, index, R, n):
    """test/regression.jl: problem_18 with Uniform(-0.5, 0.5) inputs, 13 QoI, multi-index differences"""
    z = oracle.genvec("K_3600_32", ndims)
    lat = engine.lattice(z, ndims)
    dl = [(0, [-0.5, 0.5])] * ndims
    d = engine.dists(dl)
    gm = engine.model("problem18", ndims)
    om = oracle.Model(oracle.MODEL_PROBLEM18, ndims)
    assert gm.nqoi == 13
    sh = splitmix_uniforms(31, R * ndims).reshape(R, ndims)
    got, smp = engine.sample_batch(gm, lat, d, index, sh, 0, n, want_samples=True)
    want, smp_w = oracle.sample_batch(om, z, dl, index, sh, 0, n, want_samples=True)
    assert got.shape == (R, 13, 2, 3)
    assert np.array_equal(smp, smp_w)
    check_moments(got, want)


def test_problem18_mc(engine, oracle):
    d = engine.dists([(0, [-0.5, 0.5])])
    gm = engine.model("problem18", 1)
    om = oracle.Model(oracle.MODEL_PROBLEM18, 1)
    got = engine.sample_batch(gm, None, d, [1], None, 0, 5000, mc_seed=3)
    want = oracle.sample_batch(om, None, [(0, [-0.5, 0.5])], [1], None, 0, 5000, mc_seed=3)
    check_moments(got, want)


def test_sample_unbiased_reduces_the_accumulator_on_the_device(engine, oracle):
    """mle_sample_unbiased (src/sample.jl:65-74, :113-124): every idx in zero(index):index on the same points in ONE call; the
    per-idx moment tables equal those of separate mle_sample_batch calls bit for bit, and (n, mean, M2) of the per-sample sums
    Z = sum_idx (1/Prob(idx)) dQ_idx equal the host reduction of the raw samples"""
    import itertools
    rng = np.random.default_rng(4)
    cases = []
    d1 = engine.dists([(0, [-0.5, 0.5])])
    cases.append((engine.model("problem18", 1), None, d1, [3], None, 1, 7, 500, 1, 5))
    z = oracle.genvec("K_3600_32", 2)
    lat2 = engine.lattice(z, 2)
    d2 = engine.dists([(0, [-0.5, 0.5])] * 2)
    sh2 = splitmix_uniforms(9, 4 * 2).reshape(4, 2)
    cases.append((engine.model("problem18", 2), lat2, d2, [2, 1], sh2, 4, 3, 64, 2, 0))
    zl, latl, dl, dd, gm, om = lognormal_setup(engine, oracle, 20, [0, 1, 2])
    shl = splitmix_uniforms(10, 3 * 20).reshape(3, 20)
    cases.append((gm, latl, dd, [2], shl, 3, 0, 77, 20, 0))
    for model, lat, dists, index, sh, R, k0, n, m, seed in cases:
        box = sorted(itertools.product(*[range(i + 1) for i in index]), key=lambda t: t[::-1])
        w = 1.0 / rng.uniform(0.05, 1.0, len(box))
        moms, acc = engine.sample_unbiased(model, lat, dists, index, sh, k0, n, w, m=m, mc_seed=seed)
        Z = None
        for i, idx in enumerate(box):
            mom, smp = engine.sample_batch(model, lat, dists, list(idx), sh, k0, n, m=m, mc_seed=seed, want_samples=True)
            assert np.array_equal(moms[i], mom)
            term = w[i] * smp[:, :, 0, :]
            Z = term if Z is None else Z + term
        zm = Z.mean(axis=-1)
        assert np.array_equal(acc[..., 0], np.full(zm.shape, float(n)))
        assert np.allclose(acc[..., 1], zm, rtol=1e-14, atol=0)
        m2 = ((Z - zm[..., None]) ** 2).sum(axis=-1)
        assert np.allclose(acc[..., 2], m2, rtol=1e-12, atol=1e-300)
    import mle_b200
    with pytest.raises(mle_b200.MLEArgumentError):
        engine.sample_unbiased(cases[0][0], None, d1, [3], None, 0, 10, [1.0, 1.0], m=1)   # one weight per idx


def lognormal_setup(engine, oracle, m, levels, n0=16):
    import mle_b200.kl as kl
    z = oracle.genvec("K_3600_32", m)
    gm = engine.model("lognormal1d", 1, [n0])
    om = oracle.Model(oracle.MODEL_LOGNORMAL1D, 1, [float(n0)])
    for lev in levels:
        phi = kl.kl_basis_1d(lev, m=m, n0=n0)
        gm.set_basis(lev, phi)
        om.set_basis(lev, phi)
    return z, engine.lattice(z, m), [NORMAL01] * m, engine.dists([NORMAL01] * m), gm, om


@pytest.mark.parametrize("level,R,k0,n", [(0, 16, 0, 1), (0, 2, 0, 64), (1, 3, 5, 65), (2, 16, 0, 200), (3, 2, 0, 127),
                                          (4, 16, 100, 32), (5, 2, 0, 70), (6, 2, 0, 40)])
def test_lognormal1d_levels(engine, oracle, level, R, k0, n):
    """config C2 shape: 100 KL terms, grid 2^(l+4), levels 0-6, randomly shifted lattice"""
    z, lat, dl, d, gm, om = lognormal_setup(engine, oracle, 100, [level])
    sh = splitmix_uniforms(2024, R * 100).reshape(R, 100)
    got, smp = engine.sample_batch(gm, lat, d, [level], sh, k0, n, want_samples=True)
    want, smp_w = oracle.sample_batch(om, z, dl, [level], sh, k0, n, want_samples=True)
    # per-sample values: identical KL accumulation order, identical exp, identical elimination order -> bit-identical
    assert np.array_equal(smp, smp_w)
    if n > 1:
        check_moments(got, want)


def test_lognormal1d_level6_bench_slice(engine, oracle):
    """the finest level of the C2 bench step at its bench size: level 6 (1024 cells), 16 shifts x 2048 points, point numbers
    starting where a second rank's slice would start; per-sample values bit-identical, moments to 1e-12"""
    z, lat, dl, d, gm, om = lognormal_setup(engine, oracle, 100, [6])
    sh = splitmix_uniforms(2024, 16 * 100).reshape(16, 100)
    got, smp = engine.sample_batch(gm, lat, d, [6], sh, 2048, 2048, want_samples=True)
    want, smp_w = oracle.sample_batch(om, z, dl, [6], sh, 2048, 2048, want_samples=True)
    assert np.array_equal(smp, smp_w)
    check_moments(got, want)


def test_lognormal1d_small_kl_and_errors(engine, oracle):
    import mle_b200
    z, lat, dl, d, gm, om = lognormal_setup(engine, oracle, 10, [0, 1])
    sh = splitmix_uniforms(8, 2 * 10).reshape(2, 10)
    got = engine.sample_batch(gm, lat, d, [1], sh, 0, 500)
    check_moments(got, oracle.sample_batch(om, z, dl, [1], sh, 0, 500))
    with pytest.raises(mle_b200.MLEError):
        engine.sample_batch(gm, lat, d, [2], sh, 0, 10)   # no basis for level 2
    with pytest.raises(mle_b200.MLEArgumentError):
        engine.sample_batch(gm, lat, d, [1, 0], sh, 0, 10)  # wrong index dimension
    with pytest.raises(mle_b200.MLEArgumentError):
        engine.sample_batch(gm, lat, d, [1], sh, 0, 0)    # n must be positive


def test_moments_are_deterministic(engine, oracle):
    z, lat, dl, d, gm, om = lognormal_setup(engine, oracle, 100, [2])
    sh = splitmix_uniforms(3, 16 * 100).reshape(16, 100)
    a = engine.sample_batch(gm, lat, d, [2], sh, 0, 3000)
    b = engine.sample_batch(gm, lat, d, [2], sh, 0, 3000)
    assert np.array_equal(a, b)


def test_full_size_sweep_properties(engine):
    """BASELINE-size batch (config C5 shape: 2^20 lattice points x 100 Normal dims x 16 shifts, expsum integrand) checked
    through size-independent properties: closed-form mean, additivity of the moments over point ranges, determinism."""
    import mle_b200
    s, R, n = 100, 16, 1 << 20
    w = [0.5 / (j + 1) for j in range(s)]
    lat = engine.lattice(None, s)
    d = engine.dists([NORMAL01] * s)
    gm = engine.model("expsum", 1, [s] + w)
    sh = splitmix_uniforms(2024, R * s).reshape(R, s)
    whole = engine.sample_batch(gm, lat, d, [0], sh, 0, n)
    assert (whole[..., 0] == n).all()
    exact = np.exp(0.5 * np.sum(np.square(w)))
    shift_means = whole[:, 0, 1, 1]
    est, se = shift_means.mean(), shift_means.std(ddof=1) / np.sqrt(R)
    assert abs(est - exact) < 6 * se and se < 1e-4          # 16.8 M lattice points: QMC error well below MC's 3e-4
    half_a = engine.sample_batch(gm, lat, d, [0], sh, 0, n // 2)
    half_b = engine.sample_batch(gm, lat, d, [0], sh, n // 2, n // 2)
    merged = mle_b200.moments_merge(half_a, half_b)
    check_moments(merged, whole)
    assert np.array_equal(whole, engine.sample_batch(gm, lat, d, [0], sh, 0, n))
    # the second half of a 2^20-point lattice sequence refines the first: its per-shift means differ only slightly
    assert np.max(np.abs(half_a[:, 0, 1, 1] - half_b[:, 0, 1, 1])) < 5e-3


def test_full_size_lognormal_level_properties(engine, oracle):
    """C2 level 0 at bench size (2^20 points x 16 shifts): moments are additive over point ranges and the first 4096 points
    of every shift agree with the oracle bit for bit"""
    import mle_b200
    z, lat, dl, d, gm, om = lognormal_setup(engine, oracle, 100, [0])
    sh = splitmix_uniforms(2024, 16 * 100).reshape(16, 100)
    n = 1 << 20
    whole = engine.sample_batch(gm, lat, d, [0], sh, 0, n)
    parts = [engine.sample_batch(gm, lat, d, [0], sh, k0, cnt) for k0, cnt in ((0, 4096), (4096, n // 2 - 4096), (n // 2, n // 2))]
    acc = parts[0]
    for p in parts[1:]:
        acc = mle_b200.moments_merge(acc, p)
    check_moments(acc, whole)
    check_moments(parts[0], oracle.sample_batch(om, z, dl, [0], sh, 0, 4096))
    assert 0.1 < whole[:, 0, 1, 1].mean() < 0.14   # E[u(1/2)] a little below the a == 1 value 1/8 (Jensen)


def lognormal2d_setup(engine, oracle, m, levels, n0=4):
    import mle_b200.kl as kl
    z = oracle.genvec("K_3600_32", m)
    gm = engine.model("lognormal2d", 1, [n0])
    om = oracle.Model(oracle.MODEL_LOGNORMAL2D, 1, [float(n0)])
    for lev in levels:
        phi = kl.kl_basis_2d(lev, m=m, n0=n0)
        gm.set_basis(lev, phi)
        om.set_basis(lev, phi)
    return z, engine.lattice(z, m), [NORMAL01] * m, engine.dists([NORMAL01] * m), gm, om


@pytest.mark.parametrize("level,R,k0,n,m", [(0, 3, 0, 1, 40), (0, 2, 5, 130, 40), (1, 4, 0, 37, 40), (2, 2, 0, 50, 40), (3, 2, 0, 17, 40),
                                            (4, 2, 0, 5, 40), (5, 1, 0, 2, 24), (1, 2, 0, 20, 400)])
def test_lognormal2d_levels(engine, oracle, level, R, k0, n, m):
    """configs C3/C4 shape: 2-D lognormal diffusion, grid 4*2^l per direction (warp-per-sample path up to 32x32, CTA-per-sample
    path for 64x64 and 128x128), lattice inputs; per-sample (dQ, Qf) bit-identical to the oracle's frontal elimination"""
    z, lat, dl, d, gm, om = lognormal2d_setup(engine, oracle, m, [level])
    sh = splitmix_uniforms(2024, R * m).reshape(R, m)
    got, smp = engine.sample_batch(gm, lat, d, [level], sh, k0, n, want_samples=True)
    want, smp_w = oracle.sample_batch(om, z, dl, [level], sh, k0, n, want_samples=True)
    assert np.array_equal(smp, smp_w)
    if n > 1:
        check_moments(got, want)


@pytest.mark.parametrize("level,n", [(0, 64), (2, 24), (4, 6), (5, 3)])
def test_lognormal2d_c3_shape_400_terms_mc(engine, oracle, level, n):
    """config C3 at its stated size: ML()/MC(), 400-term KL field, levels up to 5 (128 x 128 cells), MC inputs; per-sample
    (dQ, Qf) bit-identical to the oracle"""
    z, lat, dl, d, gm, om = lognormal2d_setup(engine, oracle, 400, [level])
    got, smp = engine.sample_batch(gm, None, d, [level], None, 0, n, mc_seed=11, want_samples=True)
    want, smp_w = oracle.sample_batch(om, None, dl, [level], None, 0, n, mc_seed=11, want_samples=True)
    assert np.array_equal(smp, smp_w)
    check_moments(got, want)


def test_lognormal2d_mc(engine, oracle):
    """config C3: ML()/MC() on the 2-D problem (counter-based uniforms instead of lattice points)"""
    z, lat, dl, d, gm, om = lognormal2d_setup(engine, oracle, 40, [2])
    got = engine.sample_batch(gm, None, d, [2], None, 0, 300, mc_seed=9)
    want = oracle.sample_batch(om, None, dl, [2], None, 0, 300, mc_seed=9)
    check_moments(got, want)


def lognormal2d_mi_setup(engine, oracle, m, index, n0=4):
    """multi-index 2-D model: bases for the index and the coarser grids of diff(index)"""
    import mle_b200.kl as kl
    z = oracle.genvec("K_3600_32", m)
    gm = engine.model("lognormal2d", 2, [n0])
    om = oracle.Model(oracle.MODEL_LOGNORMAL2D, 2, [float(n0)])
    phi = kl.kl_basis_2d(tuple(index), m=m, n0=n0)
    gm.set_basis(tuple(index), phi)
    om.set_basis(tuple(index), phi)
    return z, engine.lattice(z, m), [NORMAL01] * m, engine.dists([NORMAL01] * m), gm, om


@pytest.mark.parametrize("index,R,n,m", [((0, 0), 2, 33, 40), ((1, 0), 2, 40, 40), ((0, 1), 3, 40, 40), ((1, 1), 2, 40, 40), ((2, 1), 2, 21, 40),
                                         ((0, 3), 2, 9, 40), ((3, 0), 2, 9, 40), ((2, 3), 1, 7, 32), ((4, 1), 1, 3, 24), ((1, 4), 1, 3, 24),
                                         ((5, 0), 1, 2, 16), ((0, 6), 1, 2, 16)])
def test_lognormal2d_multi_index(engine, oracle, index, R, n, m):
    """config C4 shape: anisotropic grids 4*2^lx x 4*2^ly, mixed differences over diff(index) (src/index.jl:49-58) from
    strided views of the one fine field; per-sample (dQ, Qf) bit-identical to the oracle"""
    z, lat, dl, d, gm, om = lognormal2d_mi_setup(engine, oracle, m, index)
    sh = splitmix_uniforms(404, R * m).reshape(R, m)
    got, smp = engine.sample_batch(gm, lat, d, list(index), sh, 3, n, want_samples=True)
    want, smp_w = oracle.sample_batch(om, z, dl, list(index), sh, 3, n, want_samples=True)
    assert np.array_equal(smp, smp_w)
    check_moments(got, want)


def test_lognormal2d_multi_index_telescopes(engine, oracle):
    """sum of dQ over the full tensor index set {0..2}^2 equals Qf of the finest index, sample by sample (same inputs, nested
    grids): the defining property of the multi-index difference"""
    import mle_b200.kl as kl
    m, n = 24, 16
    z = oracle.genvec("K_3600_32", m)
    gm = engine.model("lognormal2d", 2, [4])
    lat, d = engine.lattice(z, m), engine.dists([NORMAL01] * m)
    sh = splitmix_uniforms(5, m).reshape(1, m)
    fine = kl.kl_basis_2d((2, 2), m=m, n0=4).reshape(17, 17, m)
    total = np.zeros(n)
    for lx in range(3):
        for ly in range(3):
            # nested bases: the coarse grid's nodes are a strided subset of the finest grid's
            phi = fine[:: 2 ** (2 - ly), :: 2 ** (2 - lx), :].reshape(-1, m)
            gm.set_basis((lx, ly), phi)
            _, smp = engine.sample_batch(gm, lat, d, [lx, ly], sh, 0, n, want_samples=True)
            total += smp[0, 0, 0]
            if (lx, ly) == (2, 2):
                qf = smp[0, 0, 1].copy()
    assert np.allclose(total, qf, rtol=1e-12, atol=0)


def test_lognormal2d_argument_errors(engine):
    import mle_b200
    import mle_b200.kl as kl
    gm = engine.model("lognormal2d", 2, [4])
    with pytest.raises(mle_b200.MLEArgumentError):
        gm.set_basis((1, 0), np.zeros((5 * 5, 8)))       # rows must be 9 x 5 nodes
    with pytest.raises(mle_b200.MLEArgumentError):
        engine.model("lognormal2d", 3, [4])
    gm.set_basis((1, 0), kl.kl_basis_2d((1, 0), m=8))
    lat, d = engine.lattice(None, 8), engine.dists([NORMAL01] * 8)
    with pytest.raises(mle_b200.MLEError):
        engine.sample_batch(gm, lat, d, [0, 1], np.zeros((1, 8)), 0, 4)   # no basis for (0, 1)
    with pytest.raises(mle_b200.MLEArgumentError):
        engine.sample_batch(gm, lat, d, [1], np.zeros((1, 8)), 0, 4)      # index dimension mismatch


@pytest.mark.parametrize("n0,ndims,index,n", [(6, 1, (0,), 40), (6, 1, (1,), 24), (6, 1, (2,), 9), (6, 2, (1, 0), 20), (6, 2, (0, 2), 12),
                                              (2, 1, (0,), 50), (2, 1, (1,), 50), (2, 2, (0, 1), 30), (2, 2, (1, 1), 30), (10, 1, (1,), 7)])
def test_lognormal2d_other_base_grids(engine, oracle, n0, ndims, index, n):
    """base grids that are not 4 * 2^l: Nx = 6, 12, 24, 20 (not a power of two) and Nx = 2 run on the shared-memory-window
    kernel, Nx = 4 with a 2-cell coarse grid on the register-window kernel; all bit-identical to the oracle"""
    import mle_b200.kl as kl
    m = 24
    z = oracle.genvec("K_3600_32", m)
    gm = engine.model("lognormal2d", ndims, [n0])
    om = oracle.Model(oracle.MODEL_LOGNORMAL2D, ndims, [float(n0)])
    key = index[0] if ndims == 1 else tuple(index)
    phi = kl.kl_basis_2d(key, m=m, n0=n0)
    gm.set_basis(key, phi)
    om.set_basis(key, phi)
    lat, d = engine.lattice(z, m), engine.dists([NORMAL01] * m)
    sh = splitmix_uniforms(99, 2 * m).reshape(2, m)
    got, smp = engine.sample_batch(gm, lat, d, list(index), sh, 1, n, want_samples=True)
    want, smp_w = oracle.sample_batch(om, z, [NORMAL01] * m, list(index), sh, 1, n, want_samples=True)
    assert np.array_equal(smp, smp_w)
    check_moments(got, want)


def test_lognormal2d_both_kernels_agree(engine, oracle, monkeypatch):
    """MLE_2D_WINDOW_SMEM=1 forces the shared-memory-window kernel: same bits as the register-window kernel"""
    import mle_b200
    import mle_b200.kl as kl
    m = 32
    phi = kl.kl_basis_2d(2, m=m)
    sh = splitmix_uniforms(3, 2 * m).reshape(2, m)
    out = []
    for force in ("0", "1"):
        monkeypatch.setenv("MLE_2D_WINDOW_SMEM", force)
        eng = mle_b200.Engine(0)
        gm = eng.model("lognormal2d", 1, [4])
        gm.set_basis(2, phi)
        _, smp = eng.sample_batch(gm, eng.lattice(None, m), eng.dists([NORMAL01] * m), [2], sh, 0, 40, want_samples=True)
        out.append(smp)
        gm.close()
        eng.close()
    assert np.array_equal(out[0], out[1])


@pytest.mark.parametrize("level,n,m", [(0, 70000, 24), (2, 12000, 24), (3, 6000, 24), (4, 1500, 16), (0, 300, 400)])
def test_lognormal2d_many_tiles_per_cta(engine, oracle, level, n, m):
    """more samples than the persistent grid holds at once (several CTA iterations, a ragged last tile, the L2 scratch field from
    64x64 up, 400 KL terms generated in several chunks): every per-sample value still bit-identical"""
    z, lat, dl, d, gm, om = lognormal2d_setup(engine, oracle, m, [level])
    got, smp = engine.sample_batch(gm, None, d, [level], None, 5, n, mc_seed=3, want_samples=True)
    want, smp_w = oracle.sample_batch(om, None, dl, [level], None, 5, n, mc_seed=3, want_samples=True)
    assert np.array_equal(smp, smp_w)
    check_moments(got, want)


@pytest.mark.parametrize("model", ["expsum", "problem18", "lognormal1d", "lognormal2d"])
def test_empty_and_single_sample_batches(engine, oracle, model):
    """n = 0 is refused (update_samples never asks for it: `n_due > 0 && sample!`, src/sample.jl:11) and n = 1 per shift works
    (the QMC warm-up, src/parse.jl:36: the variance of one sample is NaN in the reference, M2 = 0 here)"""
    import mle_b200
    import mle_b200.kl as kl
    m = 16
    if model == "expsum":
        gm = engine.model("expsum", 1, [4] + [0.1] * m)
    elif model == "problem18":
        gm, m = engine.model("problem18", 1), 1
    else:
        gm = engine.model(model, 1, [16 if model == "lognormal1d" else 4])
        gm.set_basis(0, kl.kl_basis_1d(0, m=m) if model == "lognormal1d" else kl.kl_basis_2d(0, m=m))
    dist = (0, [-0.5, 0.5]) if model == "problem18" else NORMAL01
    lat, d = engine.lattice(None, m), engine.dists([dist] * m)
    sh = splitmix_uniforms(4, 3 * m).reshape(3, m)
    with pytest.raises(mle_b200.MLEArgumentError):
        engine.sample_batch(gm, lat, d, [0], sh, 0, 0)
    mom1, smp = engine.sample_batch(gm, lat, d, [0], sh, 7, 1, want_samples=True)
    assert np.all(mom1[..., 0] == 1) and np.all(mom1[..., 2] == 0) and np.array_equal(mom1[..., 1], smp[..., 0])


MIXED = [(2, [0.0, 1.0, -2.0, 2.0]), (0, [-1.0, 1.0]), (1, [0.1, 0.7])]   # TruncatedNormal, Uniform, Normal(mu, sigma)


@pytest.mark.parametrize("model,ndims,index,n0,mc", [("expsum", 1, (1,), 0, False), ("lognormal1d", 1, (2,), 16, False),
                                                     ("lognormal1d", 1, (1,), 16, True), ("lognormal2d", 1, (2,), 4, False),
                                                     ("lognormal2d", 2, (1, 2), 4, True), ("lognormal2d", 1, (1,), 6, False)])
def test_general_distribution_mix_through_every_kernel(engine, oracle, model, ndims, index, n0, mc):
    """heterogeneous `distributions` (src/estimator.jl:116: one object per dimension): the general inverse-CDF path
    (TruncatedNormal, Uniform, Normal(mu, sigma) interleaved; Weibull's log/pow come from the platform libm and are covered
    to 4 ulp in test_gpu_points.py) inside every fused kernel, lattice and MC inputs"""
    import mle_b200.kl as kl
    m = 36
    dl = [MIXED[j % 3] for j in range(m)]
    z = oracle.genvec("K_3600_32", m)
    if model == "expsum":
        w = [0.3 / (j + 1) for j in range(m)]
        gm, om = engine.model("expsum", 1, [4] + w), oracle.Model(oracle.MODEL_EXPSUM, 1, [4] + w)
    else:
        kind = oracle.MODEL_LOGNORMAL1D if model == "lognormal1d" else oracle.MODEL_LOGNORMAL2D
        gm, om = engine.model(model, ndims, [n0]), oracle.Model(kind, ndims, [float(n0)])
        key = index[0] if ndims == 1 else tuple(index)
        phi = 0.3 * (kl.kl_basis_1d(key, m=m, n0=n0) if model == "lognormal1d" else kl.kl_basis_2d(key, m=m, n0=n0))
        gm.set_basis(key, phi)
        om.set_basis(key, phi)
    d = engine.dists(dl)
    if mc:
        got, smp = engine.sample_batch(gm, None, d, list(index), None, 0, 90, m=m, mc_seed=17, want_samples=True)
        want, smp_w = oracle.sample_batch(om, None, dl, list(index), None, 0, 90, m=m, mc_seed=17, want_samples=True)
    else:
        sh = splitmix_uniforms(8, 3 * m).reshape(3, m)
        got, smp = engine.sample_batch(gm, engine.lattice(z, m), d, list(index), sh, 2, 70, want_samples=True)
        want, smp_w = oracle.sample_batch(om, z, dl, list(index), sh, 2, 70, want_samples=True)
    assert np.array_equal(smp, smp_w)
    check_moments(got, want)


@pytest.mark.parametrize("m,level", [(200, 1), (300, 0), (129, 2)])
def test_lognormal1d_many_kl_terms(engine, oracle, m, level):
    """more than 128 KL terms: the generator switches to 64-bit class masks and multiple passes over the dimensions"""
    z, lat, dl, d, gm, om = lognormal_setup(engine, oracle, m, [level])
    sh = splitmix_uniforms(77, 3 * m).reshape(3, m)
    got, smp = engine.sample_batch(gm, lat, d, [level], sh, 11, 150, want_samples=True)
    want, smp_w = oracle.sample_batch(om, z, dl, [level], sh, 11, 150, want_samples=True)
    assert np.array_equal(smp, smp_w)
    check_moments(got, want)


def test_update_samples_matches_individual_calls(engine, oracle):
    """mle_update_samples: several sample! requests (different levels, shift counts, point ranges) in one call with one
    synchronisation return exactly what the individual mle_sample_batch calls return"""
    z, lat, dl, d, gm, om = lognormal_setup(engine, oracle, 100, [0, 1, 3])
    entries = []
    for i, (lev, R, k0, n) in enumerate([(0, 16, 0, 700), (1, 16, 5, 300), (3, 4, 0, 33), (0, 16, 700, 1)]):
        entries.append(([lev], splitmix_uniforms(50 + i, R * 100).reshape(R, 100), k0, n, 100))
    got = engine.update_samples(gm, lat, d, entries)
    for (idx, sh, k0, n, m), g in zip(entries, got):
        assert np.array_equal(g, engine.sample_batch(gm, lat, d, idx, sh, k0, n))
        check_moments(g, oracle.sample_batch(om, z, dl, idx, sh, k0, n))
    # MC entries (no shifts) and the argument checks
    w = [0.1] * 8
    ge = engine.model("expsum", 1, [2] + w)
    d8 = engine.dists([NORMAL01] * 8)
    mc = engine.update_samples(ge, None, d8, [([0], None, 0, 50, 8), ([2], None, 50, 20, 8)], mc_seed=4)
    assert np.array_equal(mc[1], engine.sample_batch(ge, None, d8, [2], None, 50, 20, m=8, mc_seed=4))
    import mle_b200
    with pytest.raises(mle_b200.MLEArgumentError):
        engine.update_samples(gm, lat, d, [([0], entries[0][1], 0, 0, 100)])    # n = 0 is refused like in mle_sample_batch


def _engine_with_env(**env):
    """a second context whose library switches (read in mle_init) differ from the session engine's"""
    import os
    import mle_b200
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return mle_b200.Engine(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("level,n", [(1, 40), (3, 33), (6, 5)])
def test_lognormal1d_small_batch_variant_is_bit_identical(oracle, level, n):
    """the TMA-staged small-batch variant of the phase-synchronous kernel (basis panels by cp.async.bulk + mbarrier) against the
    regular kernel and the oracle: same per-sample values bit for bit, with a partial last chunk (level 1) and many chunks (6)"""
    outs = []
    for staged in (0, 1):
        eng = _engine_with_env(MLE_LN_STAGED=staged, MLE_LN_BLOCK=1)
        z, lat, dl, d, gm, om = lognormal_setup(eng, oracle, 100, [level])
        sh = splitmix_uniforms(31, 3 * 100).reshape(3, 100)
        got, smp = eng.sample_batch(gm, lat, d, [level], sh, 7, n, want_samples=True)
        outs.append((got, smp))
        eng.close()
    want, smp_w = oracle.sample_batch(om, z, dl, [level], sh, 7, n, want_samples=True)
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[1][1], smp_w)
    assert np.array_equal(outs[0][0], outs[1][0])
    check_moments(outs[1][0], want)


@pytest.mark.parametrize("level,n,k0", [(0, 1, 0), (0, 9, 5), (1, 8, 7), (2, 40, 3), (3, 33, 1), (4, 8, 0), (4, 30, 9), (5, 3, 11), (6, 5, 7), (6, 17, 2 ** 20 + 3)])
def test_lognormal1d_latency_variant_is_bit_identical(oracle, level, n, k0):
    """the latency variant for tiny batches (mle_kernel_ln1d_small.cuh: 8-sample tiles, whole field in shared memory, basis panels
    by cp.async.bulk, reciprocals in parallel, flux sums in sequence) against the throughput kernels and the oracle: the same
    per-sample values bit for bit -- one tile, partial tiles, several tiles per CTA, single and many panels"""
    outs = []
    # throughput kernels | latency variant, one CTA per sample tile | latency variant with thread-block clusters on the fine grids
    for env in (dict(MLE_LN_SMALL=0), dict(MLE_LN_SMALL=1, MLE_LN_CLUSTER=0), dict(MLE_LN_SMALL=1)):
        eng = _engine_with_env(**env)
        z, lat, dl, d, gm, om = lognormal_setup(eng, oracle, 100, [level])
        sh = splitmix_uniforms(31 + level, 3 * 100).reshape(3, 100)
        got, smp = eng.sample_batch(gm, lat, d, [level], sh, k0, n, want_samples=True)
        mom_only = eng.sample_batch(gm, lat, d, [level], sh, k0, n)   # zero-copy result path, one CTA per shift when n <= 8
        outs.append((got, smp, mom_only))
        eng.close()
    want, smp_w = oracle.sample_batch(om, z, dl, [level], sh, k0, n, want_samples=True)
    for got, smp, mom_only in outs:
        assert np.array_equal(smp, smp_w)
        check_moments(got, want)
        check_moments(mom_only, want)
    assert np.array_equal(outs[1][0], outs[1][2]) and np.array_equal(outs[2][0], outs[2][2]) and np.array_equal(outs[1][0], outs[2][0])


def test_lognormal1d_latency_variant_mc_and_other_shapes(oracle):
    """MC inputs (no lattice), a KL basis that is not a multiple of four terms, a grid with n0 = 4 (a single partial row tile)"""
    for small in (0, 1):
        eng = _engine_with_env(MLE_LN_SMALL=small)
        z, lat, dl, d, gm, om = lognormal_setup(eng, oracle, 37, [0, 2], n0=4)
        for level in (0, 2):
            got, smp = eng.sample_batch(gm, None, d, [level], None, 3, 21, mc_seed=99 + level, want_samples=True)
            want, smp_w = oracle.sample_batch(om, None, dl, [level], None, 3, 21, mc_seed=99 + level, want_samples=True)
            assert np.array_equal(smp, smp_w)
            check_moments(got, want)
            sh = splitmix_uniforms(5, 2 * 37).reshape(2, 37)
            got, smp = eng.sample_batch(gm, lat, d, [level], sh, 0, 13, want_samples=True)
            want, smp_w = oracle.sample_batch(om, z, dl, [level], sh, 0, 13, want_samples=True)
            assert np.array_equal(smp, smp_w)
            check_moments(got, want)
        eng.close()


@pytest.mark.parametrize("level,n", [(0, 5), (0, 3000), (2, 7), (2, 2500), (4, 3)])
def test_lognormal1d_reciprocal_of_all_ones_mantissa(engine, oracle, level, n):
    """a field that is 1 - 2^-53 at every node (one KL term phi = -2^-53 everywhere, input Uniform(1, 1 + 2^-52)): every interface
    coefficient has an all-ones mantissa, the case in which a Newton reciprocal started from a seed with a zero low word sticks on
    a tie and returns 1 instead of 1 + 2^-52 (found by tools/fuzz_parity.py, fixed by rcp_seed in csrc/mle_device.cuh, isolated in
    tools/ddiv_seed_test.cu) -- latency variant and throughput kernels, bit for bit against the oracle's IEEE division"""
    n0, N = 16, 16 << level
    gm = engine.model("lognormal1d", 1, [n0])
    om = oracle.Model(oracle.MODEL_LOGNORMAL1D, 1, [float(n0)])
    phi = np.full((N + 1, 1), -2.0 ** -53)
    gm.set_basis(level, phi)
    om.set_basis(level, phi)
    dl = [(0, [1.0, 1.0 + 2.0 ** -52])]   # xi is 1 or 1 + 2^-52: exp(phi xi) rounds to 1 - 2^-53 either way
    z = oracle.genvec("K_3600_32", 1)
    sh = splitmix_uniforms(3, 2).reshape(2, 1)
    got, smp = engine.sample_batch(gm, engine.lattice(z, 1), engine.dists(dl), [level], sh, 0, n, want_samples=True)
    want, smp_w = oracle.sample_batch(om, z, dl, [level], sh, 0, n, want_samples=True)
    assert np.array_equal(smp, smp_w)
    assert smp_w[0, 0, 1, 0] != 0.125   # the oracle's value carries the rounded-up reciprocals: not the exact 1/8 of a == 1
    check_moments(got, want)


@pytest.mark.parametrize("n", [64, 4000])
def test_overflowing_samples_leave_non_finite_statistics(engine, oracle, n):
    """exp(sum a_j xi_j) with a_j = 300 overflows for part of the inputs: the samples are Inf on both sides (one exp algorithm), and
    the variance of a shift that contains them is non-finite like var() of the reference -- the M2 clamp must not turn the NaN
    into 0 (it did, through fmax; found by tools/fuzz_parity.py with random distribution parameters)"""
    s_ = 4
    z = oracle.genvec("K_3600_32", s_)
    w = [300.0] * s_
    gm, om = engine.model("expsum", 1, [4.0] + w), oracle.Model(oracle.MODEL_EXPSUM, 1, [4.0] + w)
    sh = splitmix_uniforms(11, 2 * s_).reshape(2, s_)
    dl = [NORMAL01] * s_
    got, smp = engine.sample_batch(gm, engine.lattice(z, s_), engine.dists(dl), [0], sh, 0, n, want_samples=True)
    want, smp_w = oracle.sample_batch(om, z, dl, [0], sh, 0, n, want_samples=True)
    assert np.array_equal(smp, smp_w, equal_nan=True)
    assert np.any(np.isinf(smp_w))
    assert np.array_equal(np.isfinite(got[..., 2]), np.isfinite(want[..., 2])) and not np.all(np.isfinite(want[..., 2]))


def test_lognormal2d_blocked_dmma_solver_equals_register_window_solver(oracle):
    """64 x 64 cells, where the register-window solver is the default: the blocked DMMA solver (MLE_2D_DMMA=2) and the solver
    without any DMMA blocking (MLE_2D_DMMA=0) give the same bits as the oracle"""
    outs = []
    for mode in (0, 2):
        eng = _engine_with_env(MLE_2D_DMMA=mode)
        z, lat, dl, d, gm, om = lognormal2d_setup(eng, oracle, 40, [4])
        sh = splitmix_uniforms(32, 2 * 40).reshape(2, 40)
        got, smp = eng.sample_batch(gm, lat, d, [4], sh, 3, 9, want_samples=True)
        outs.append(smp)
        eng.close()
    want, smp_w = oracle.sample_batch(om, z, dl, [4], sh, 3, 9, want_samples=True)
    assert np.array_equal(outs[0], smp_w) and np.array_equal(outs[1], smp_w)
