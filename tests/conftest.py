import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """a fresh checkout has no built artefacts (they are git-ignored): build libmle_cuda.so once (nvcc cross-compiles without
    a GPU, ~90 s) so that the suite does not depend on `__graft_entry__.build()` having been run first"""
    lib = os.path.join(ROOT, "multilevelestimators.jl_b200", "lib", "libmle_cuda.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "multilevelestimators.jl_b200", "csrc")], check=True)


@pytest.fixture(scope="session")
def goldens():
    with open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def engine():
    import mle_b200
    eng = mle_b200.Engine(0)
    yield eng
    eng.close()


def splitmix_uniforms(seed, count):
    """SplitMix64(seed) stream -> (u >> 11) * 2^-53, the synthetic-input generator of SURVEY.md 8(d)"""
    out = np.empty(count)
    x = np.uint64(seed)
    M = (1 << 64) - 1
    s = int(seed)
    for i in range(count):
        s = (s + 0x9E3779B97F4A7C15) & M
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        z = z ^ (z >> 31)
        out[i] = (z >> 11) * 2.0**-53
    return out


def ulp_diff(a, b):
    """distance in units in the last place between float64 arrays (same-sign finite values)"""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    ia = a.view(np.int64).astype(np.int64)
    ib = b.view(np.int64).astype(np.int64)
    ia = np.where(ia < 0, np.int64(-2**63) - ia, ia)
    ib = np.where(ib < 0, np.int64(-2**63) - ib, ib)
    return np.abs(ia - ib)
