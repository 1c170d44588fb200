#!/usr/bin/env python3
"""Pin the device models that are DEFINED by this repository (they are not in the reference tree): a handful of
raw samples (dQ, Qf) per model and index, as produced by the CPU oracle, stored as hex floats in model_goldens.json.

The GPU tests compare the kernels with the oracle; this file pins the oracle itself, so a change of the shared arithmetic
(KL accumulation order, the exp / log tables, the flux sums, the frontal elimination order) cannot slip through by changing
oracle and kernels together.  Regenerate ONLY when the model definition is changed on purpose:

    python tests/golden/make_model_goldens.py
"""
import json
import os
import sys

import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(here)))
sys.path.insert(0, os.path.dirname(here))

from conftest import splitmix_uniforms  # noqa: E402
from oracle import oracle as O  # noqa: E402


def basis(rows, m):
    """a synthetic KL basis from exact binary fractions only (no libm: identical on every platform), decaying like 1/(k+1)"""
    i, k = np.arange(rows).reshape(-1, 1), np.arange(m).reshape(1, -1)
    return ((i * 7 + k * 3) % 11 - 5) / 64.0 / (k + 1.0)


def cases():
    """(name, model, ndims, params, bases, dists, index, R, k0, n, lattice?)"""
    normal, unif = (1, [0.0, 1.0]), (0, [-0.5, 0.5])
    w = [0.5 / (j + 1) for j in range(24)]
    yield "expsum_l1_qmc", O.MODEL_EXPSUM, 1, [4.0] + w, {}, [normal] * 24, [1], 2, 7, 3, True
    yield "expsum_l2_mc", O.MODEL_EXPSUM, 1, [4.0] + w, {}, [normal] * 24, [2], 1, 100, 3, False
    yield "problem18_l2_qmc", O.MODEL_PROBLEM18, 1, [], {}, [unif], [2], 2, 0, 3, True
    yield "problem18_i21_qmc", O.MODEL_PROBLEM18, 2, [], {}, [unif] * 2, [2, 1], 2, 5, 2, True
    m = 40
    yield "lognormal1d_l0", O.MODEL_LOGNORMAL1D, 1, [16.0], {0: basis(17, m)}, [normal] * m, [0], 2, 3, 3, True
    yield "lognormal1d_l3", O.MODEL_LOGNORMAL1D, 1, [16.0], {3: basis(129, m)}, [normal] * m, [3], 2, 11, 3, True
    yield "lognormal2d_l2", O.MODEL_LOGNORMAL2D, 1, [4.0], {2: basis(17 * 17, m)}, [normal] * m, [2], 1, 4, 3, False
    yield ("lognormal2d_i12", O.MODEL_LOGNORMAL2D, 2, [4.0], {(1, 2): basis(9 * 17, m)}, [normal] * m, [1, 2], 2, 9, 2,
           True)


def evaluate(case):
    name, model, ndims, params, bases, dists, index, R, k0, n, lattice = case
    om = O.Model(model, ndims, params)
    for key, phi in bases.items():
        om.set_basis(key, phi)
    m = len(dists)
    z = O.genvec("K_3600_32", m) if lattice else None
    shifts = splitmix_uniforms(1000 + len(name), R * m).reshape(R, m) if lattice else None
    mom, smp = O.sample_batch(om, z, dists, index, shifts, k0, n, mc_seed=77, want_samples=True)
    return name, np.asarray(smp)   # [R, q, 2, n]


def main():
    out = {}
    for case in cases():
        name, smp = evaluate(case)
        out[name] = {"shape": list(smp.shape), "hex": [float(v).hex() for v in smp.reshape(-1)]}
    with open(os.path.join(here, "model_goldens.json"), "w") as f:
        json.dump(out, f, indent=0)
    print({k: v["shape"] for k, v in out.items()})


if __name__ == "__main__":
    main()
