"""Multi-GPU context behind the C ABI (mle_init_multi, the reference's `nb_of_workers`, src/parse.jl:224-234): requests are split
over the member devices inside the library -- whole shifts per member, or point slices merged in member order -- and small
requests stay on member 0.  On a box with one GPU the members are two contexts on the same device (MLE_ALLOW_DUP_DEVICES)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engines():
    import torch
    import mle_b200
    ndev = torch.cuda.device_count()
    if ndev < 2:
        os.environ["MLE_ALLOW_DUP_DEVICES"] = "1"
    os.environ["MLE_SHARD_MIN"] = "64"
    devs = [0, 1] if ndev >= 2 else [0, 0]
    multi = mle_b200.Engine(devs)
    single = mle_b200.Engine(0)
    yield single, multi
    multi.close()
    single.close()
    os.environ.pop("MLE_SHARD_MIN", None)


def _setup(eng, m=40, level=2):
    import mle_b200.kl as kl
    z = eng.genvec("K_3600_32")[:m]
    lat, d = eng.lattice(z, m), eng.dists([(1, [0.0, 1.0])] * m)
    gm = eng.model("lognormal1d", 1, [16])
    for lev in range(level + 1):
        gm.set_basis(lev, kl.kl_basis_1d(lev, m=m))
    return lat, d, gm


def test_device_count(engines):
    single, multi = engines
    assert single.device_count == 1 and multi.device_count == 2


@pytest.mark.parametrize("R,n", [(16, 500), (5, 333), (2, 64), (3, 7)])
def test_shift_split_matches_one_gpu(engines, R, n):
    """R >= members in use: every member computes whole shifts; its triplets are final and land in the caller's table"""
    single, multi = engines
    rng = np.random.default_rng(R * 1000 + n)
    sh = rng.random((R, 40))
    a = single.sample_batch(*(lambda l, d, g: (g, l, d))(*_setup(single)), [2], sh, 3, n)
    b = multi.sample_batch(*(lambda l, d, g: (g, l, d))(*_setup(multi)), [2], sh, 3, n)
    assert np.array_equal(a[..., 0], b[..., 0])
    np.testing.assert_allclose(b[..., 1], a[..., 1], rtol=1e-13)
    np.testing.assert_allclose(b[..., 2], a[..., 2], rtol=1e-12)


def test_point_split_mc_matches_one_gpu(engines):
    """R = 1 (MC): the point range is sliced over the members and the triplets merged in member order (Chan)"""
    single, multi = engines
    m = 12
    w = [2.0] + [0.3 / (j + 1) for j in range(m)]
    res = []
    for eng in (single, multi):
        d = eng.dists([(1, [0.0, 1.0])] * m)
        gm = eng.model("expsum", 1, w)
        res.append(eng.sample_batch(gm, None, d, [2], None, 11, 10001, m=m, mc_seed=77))
    a, b = res
    assert np.array_equal(a[..., 0], b[..., 0]) and a[0, 0, 0, 0] == 10001
    np.testing.assert_allclose(b[..., 1], a[..., 1], rtol=1e-13)
    np.testing.assert_allclose(b[..., 2], a[..., 2], rtol=1e-12)


def test_update_samples_and_small_requests(engines):
    """several entries in one call: big ones split, a tiny one (below MLE_SHARD_MIN per extra member) stays on member 0"""
    single, multi = engines
    rng = np.random.default_rng(5)
    sh = [rng.random((8, 40)) for _ in range(3)]
    entries = [([0], sh[0], 0, 700, 40), ([1], sh[1], 5, 300, 40), ([2], sh[2], 0, 4, 40)]
    outs = []
    for eng in (single, multi):
        lat, d, gm = _setup(eng)
        outs.append(eng.update_samples(gm, lat, d, entries))
    for a, b in zip(*outs):
        assert np.array_equal(a[..., 0], b[..., 0])
        np.testing.assert_allclose(b[..., 1], a[..., 1], rtol=1e-13)
        np.testing.assert_allclose(b[..., 2], a[..., 2], rtol=1e-12)
    # the 4-point entry ran on one member: bit-identical to the single-GPU result
    assert np.array_equal(outs[0][2], outs[1][2])


def test_run_on_a_multi_gpu_context_takes_the_same_decisions(engines):
    """run(estimator, tol) through GpuBackend on a two-member context vs one GPU: identical sample counts, estimates to 1e-12"""
    import mle_b200.host as H
    import mle_b200.kl as kl
    single, multi = engines
    m = 30
    spec = H.ModelSpec("lognormal1d", 1, [16], {lev: kl.kl_basis_1d(lev, m=m) for lev in range(5)})
    hs = []
    for eng in (single, multi):
        est = H.Estimator(H.ML(), H.QMC(), spec, [H.Normal()] * m, H.GpuBackend(eng), max_index_set_param=4, nb_of_shifts=8,
                          cost_model=lambda i: 2.0 ** i[0], nb_of_tols=3, shift_seed=3)
        hs.append(H.run(est, 2e-4))
    a, b = hs
    assert a["nb_of_samples"] == b["nb_of_samples"]
    assert abs(a["mean"] - b["mean"]) <= 1e-12 * abs(a["mean"])
    assert abs(a["varest"] - b["varest"]) <= 1e-9 * abs(a["varest"])
