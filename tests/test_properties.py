"""Property tests (hypothesis) of the host-side pieces of the path: range sharding, moment algebra, index sets.  No GPU."""
import math
import os
import sys

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def host():
    import mle_b200.host as H
    return H


@settings(max_examples=200, deadline=None, derandomize=True)
@given(k0=st.integers(0, 2**32), n=st.integers(0, 10**7), world=st.integers(1, 16))
def test_split_range_is_a_partition(k0, n, world):
    """the ranks' slices are contiguous, ordered, cover k0 .. k0+n-1 exactly once and differ by at most one point"""
    H = host()
    parts = [H.split_range(k0, n, world, r) for r in range(world)]
    assert parts[0][0] == k0 and sum(c for _, c in parts) == n
    for (s0, c0), (s1, _c1) in zip(parts, parts[1:]):
        assert s1 == s0 + c0
    counts = [c for _, c in parts]
    assert max(counts) - min(counts) <= 1 and counts == sorted(counts, reverse=True)


def triplet(x):
    x = np.asarray(x, dtype=np.float64)
    return np.array([len(x), x.mean(), ((x - x.mean()) ** 2).sum()]) if len(x) else np.zeros(3)


@settings(max_examples=100, deadline=None, derandomize=True)
@given(data=st.lists(st.floats(-1e6, 1e6, allow_nan=False, width=64), min_size=1, max_size=200),
       cuts=st.lists(st.integers(0, 200), min_size=0, max_size=5))
def test_moment_merge_equals_one_pass(data, cuts):
    """mle_moments_merge (Chan): merging the triplets of any split of a sample vector, in order, gives the triplet of the whole
    vector -- including empty pieces (a rank without points sends zeros)"""
    from mle_b200.engine import moments_merge
    cuts = sorted(min(c, len(data)) for c in cuts)
    pieces = [data[a:b] for a, b in zip([0] + cuts, cuts + [len(data)])]
    acc = triplet(pieces[0]).reshape(1, 3)
    for p in pieces[1:]:
        acc = moments_merge(acc, triplet(p).reshape(1, 3))
    want = triplet(data)
    scale = max(1.0, np.abs(np.asarray(data)).max())
    assert acc[0, 0] == want[0]
    assert abs(acc[0, 1] - want[1]) <= 1e-12 * scale
    assert abs(acc[0, 2] - want[2]) <= 1e-10 * max(want[2], scale * scale * 1e-4)


WEIGHTS = st.lists(st.sampled_from([1.0, 0.9, 0.75, 0.6, 0.5, 1 / 3]), min_size=2, max_size=3)


@settings(max_examples=60, deadline=None, derandomize=True)
@given(kind=st.sampled_from(["FT", "TD", "HC", "ZC"]), weights=WEIGHTS, sz=st.integers(0, 6))
def test_index_sets_are_nested_and_downward_closed(kind, weights, sz):
    """src/index_set.jl:77-193: every a-priori index set grows with its size parameter, contains the origin, is downward closed
    (each backward neighbour of a member is a member -- what `is_admissable` and the telescoping sum rely on) and lists its
    members in CartesianIndices (column-major) order"""
    H = host()
    t = getattr(H, kind)(*weights)
    cur, nxt = t.get_index_set(sz), t.get_index_set(sz + 1)
    s = set(cur)
    assert cur[0] == (0,) * t.ndims and s <= set(nxt) and len(s) == len(cur)
    assert cur == sorted(cur, key=lambda i: i[::-1])
    for idx in cur:
        for k in range(t.ndims):
            if idx[k] > 0:
                assert tuple(v - (1 if j == k else 0) for j, v in enumerate(idx)) in s
    for idx in set(nxt) - s:
        below = [tuple(v - (1 if j == k else 0) for j, v in enumerate(idx)) for k in range(t.ndims) if idx[k] > 0]
        if all(b in s for b in below):
            assert H.is_admissable(s, idx)
    assert not any(H.is_admissable(s, idx) for idx in cur)


@settings(max_examples=100, deadline=None, derandomize=True)
@given(index=st.lists(st.integers(0, 5), min_size=1, max_size=3))
def test_diff_signs_telescope(index):
    """src/index.jl:49-58: sum over kappa in the box max(0, index-1) .. index of sign(kappa) is 0 unless index = 0 (the mixed
    difference of a constant vanishes), every kappa is a backward neighbour pattern, signs are (-1)^|index - kappa|"""
    H = host()
    index = tuple(index)
    d = H.diff(index)
    assert len(d) == math.prod(2 if c > 0 else 1 for c in index) - 1
    assert 1 + sum(d.values()) == (1 if all(c == 0 for c in index) else 0)
    for kappa, sign in d.items():
        assert all(0 <= a - b <= 1 for a, b in zip(index, kappa)) and sign == (-1) ** sum(a - b for a, b in zip(index, kappa))


@settings(max_examples=50, deadline=None, derandomize=True)
@given(seed=st.integers(0, 2**64 - 1), a=st.lists(st.integers(0, 12), min_size=1, max_size=3),
       b=st.lists(st.integers(0, 12), min_size=1, max_size=3))
def test_index_seeds_do_not_collide(seed, a, b):
    H = host()
    sa, sb = H.index_seed(seed, tuple(a)), H.index_seed(seed, tuple(b))
    assert 0 <= sa < 2**64 and (sa == sb) == (tuple(a) == tuple(b))
