"""A backend for the host control loop that evaluates `parallel_sample!` with the CPU oracle (test infrastructure):
lets the CPU suite exercise run()/History/sharding without a GPU and gives the GPU suite its end-to-end checker."""
import numpy as np

from oracle import oracle as O


class _OModel:
    def __init__(self, spec):
        kind = {"expsum": O.MODEL_EXPSUM, "problem18": O.MODEL_PROBLEM18, "lognormal1d": O.MODEL_LOGNORMAL1D, "lognormal2d": O.MODEL_LOGNORMAL2D}[spec.name]
        self.m = O.Model(kind, spec.ndims, [float(p) for p in spec.params])
        for level, phi in spec.bases.items():
            self.m.set_basis(level, phi)
        self.nqoi = self.m.nqoi


class OracleBackend:
    def lattice(self, z, s, nmax=2**32 - 1):
        z = O.genvec("K_3600_32", s) if z is None else np.asarray(z, dtype=np.uint32)[:s]
        return (z, nmax)

    def dists(self, dists):
        return list(dists)

    def model(self, spec):
        return _OModel(spec)

    def update_samples(self, model, lattice, dists, entries, mc_seed):
        """the batched call of the device backend, entry by entry (entries = [(index, shifts, k0, n, m)])"""
        self.batched_calls = getattr(self, "batched_calls", 0) + 1
        return [self.sample_batch(model, lattice, dists, index, shifts, k0, n, m, mc_seed)[0] for index, shifts, k0, n, m in entries]

    def sample_batch(self, model, lattice, dists, index, shifts, k0, n, m, mc_seed, want_samples=False):
        z, nmax = lattice if lattice is not None else (None, 2**32 - 1)
        if want_samples:
            mom, smp = O.sample_batch(model.m, z, dists, index, shifts, k0, n, m=m, nmax=nmax, mc_seed=mc_seed, want_samples=True)
            return mom, 0.0, smp
        mom = O.sample_batch(model.m, z, dists, index, shifts, k0, n, m=m, nmax=nmax, mc_seed=mc_seed)
        return mom, 0.0
