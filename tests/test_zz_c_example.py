"""examples/abi_example.c: a strict-C99 consumer of include/mle_cuda.h (the boundary is usable from plain C, no C++ or
torch types), compiled here with gcc and linked against the in-tree library."""
import math
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "multilevelestimators.jl_b200", "lib")


def compile_example(tmp_path):
    exe = str(tmp_path / "abi_example")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "abi_example.c"), "-L", LIBDIR, "-lmle_cuda",
                           "-Wl,-rpath," + LIBDIR, "-lm", "-o", exe])
    return exe


def test_c_example_compiles_as_c99_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    exe = compile_example(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the gpu test runs the example")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 1
    assert "mle_init failed (-2)" in p.stderr and "no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_c_example_matches_the_oracle(tmp_path):
    """the example's two merged mle_sample_batch calls (points 0..n-1 and n..2n-1) against one oracle batch of 2n points"""
    import sys
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    exe = compile_example(tmp_path)
    level, n, M, R = 2, 4096, 64, 8
    p = subprocess.run([exe, str(level), str(n)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    got_f = float(re.search(r"E\[Q_l\]\s+=\s+([0-9.eE+-]+)", p.stdout).group(1))
    got_d = float(re.search(r"E\[dQ_l\]\s+=\s+([0-9.eE+-]+)", p.stdout).group(1))
    state = [2024]

    def splitmix():
        state[0] = (state[0] + 0x9E3779B97F4A7C15) & (2**64 - 1)
        z = state[0]
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
        z ^= z >> 31
        return (z >> 11) / 9007199254740992.0

    shifts = np.array([splitmix() for _ in range(R * M)]).reshape(R, M)
    a = [0.5 / (j + 1) for j in range(M)]
    om = O.Model(O.MODEL_EXPSUM, 1, [4.0] + a)
    z = np.array(O.genvec("K_3600_32")[:M], dtype=np.uint32)
    mom = O.sample_batch(om, z, [(1, [0.0, 1.0])] * M, [level], shifts, 0, 2 * n)   # [R, q, X, (n, mean, M2)]
    want_f, want_d = mom[:, 0, 1, 1].mean(), mom[:, 0, 0, 1].mean()
    assert abs(got_f - want_f) <= 2e-12 * abs(want_f) and abs(got_d - want_d) <= 1e-11 * abs(want_f)   # 12 printed digits
    assert "n per shift after merge = %d" % (2 * n) in p.stdout
    assert math.isfinite(got_f)
