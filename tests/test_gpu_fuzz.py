"""Randomised differential parity: tools/fuzz_parity.py draws (model, index, shifts, point range, dimensions, distribution
mix, lattice / MC) at random and compares the CUDA path with the oracle: per-sample values bit-identical, moments to 1e-12."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [11, 12])
def test_random_cases_match_the_oracle(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), "1500", str(seed)], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "1500 cases, 0 mismatches" in r.stdout
