"""Host-side mirror of the reference interface for the hot path: `Estimator(index_set, sample_method, sample_function,
distributions; options...)`, `run(estimator, tol)` and the `History` it returns.

This is the Python twin of the Julia glue in julia/MLECuda.jl (Julia is not installed in the build or GPU images, so the
control loop is restated here to make `run` testable end to end).  It follows the reference line by line where the hot
path is concerned:

  Index / Level / diff          src/index.jl:19-58
  SL, ML, FT, TD, HC, ZC, AD, U src/index_set.jl:49-227 (get_index_set, is_admissable, length, show, print)
  option parsing                src/options.jl:10-43, src/parse.jl:35-307 (same keys per estimator type, defaults, rules)
  sample! / update_samples      src/sample.jl:8-34      point numbers continue across calls (Istart = N + 1; U: total + 1)
  parallel_sample!              src/sample.jl:39-56, :79-102   -> ONE `sample_batch` call on the backend (libmle_cuda)
  append_samples!               src/sample.jl:58-74, :104-124  -> Chan merge of (n, mean, M2) per (qoi, shift); U: per-sample sums
  mean / var / varest / cost    src/methods.jl:44-68, :282
  _run (SL / ML / multi-index)  src/run.jl:41-142;  unbiased estimators src/run.jl:145-200
  rates, regression, bias, MSE splitting, optimal_nb_of_samples   src/methods.jl:70-201
  adaptive index sets (AD)      src/methods.jl:206-261, src/internals.jl:150-206
  unbiased estimators (U)       src/methods.jl:287-351, src/internals.jl:208-225
  History keys, save            src/history.jl:55-107 (JSON instead of JLD2)
  backends                      GpuBackend (one GPU through the C ABI), ShardedBackend (one rank of a torch.distributed group)

  verbose report                src/print.jl (report.py: the same tables and messages, off by default)

Out of scope: the JLD2 / Reporter.jl file formats.
The sample function is a device model (name + parameters + KL bases), not an arbitrary closure.
"""
import itertools
import math
import os
import random as _random
import time as _time

import numpy as np


# ---------------------------------------------------------------------------------------------------------------
# index sets (src/index_set.jl) and sample methods (src/sample_method.jl)
# ---------------------------------------------------------------------------------------------------------------
def Index(*i):
    """Index(i...) (src/index.jl:19): a zero-based multi-index; plain tuples of ints everywhere in this module."""
    if not i or not all(isinstance(c, (int, np.integer)) and not isinstance(c, (bool, np.bool_)) for c in i):
        raise TypeError("Index: integer components required")  # MethodError in the reference, test/index.jl:13,25
    return tuple(int(c) for c in i)


def Level(l):
    """Level(l) (src/index.jl:36): a one-dimensional Index."""
    return Index(l)


def diff(index):
    """diff(index) (src/index.jl:49-58): the indices kappa in max(0, index - 1) : index other than index itself, each with the
    sign (-1)^|index - kappa| -- so that dQ(index) = Q(index) + sum_kappa sign * Q(kappa) is the mixed difference."""
    index = Index(*index)
    out = {}
    ranges = [range(max(0, c - 1), c + 1) for c in index]
    for kappa in itertools.product(*ranges):
        if kappa != index:
            out[kappa] = -1 if (sum(index) - sum(kappa)) % 2 else 1
    return out


class IndexSetNotImplemented(ValueError, NotImplementedError):
    """the reference throws ArgumentError here (src/index_set.jl:259-264): a ValueError in this mirror"""


class _IndexSetBase:
    """what every index set type shares: the short name shown by `show` (src/index_set.jl:319-327), `length(idxset, sz)`
    (:262) and the 2-d picture of `print(idxset, sz)` (:294-317)"""

    def __repr__(self):
        name = type(self).__name__
        return name if name in ("SL", "ML") else "%s{%d}" % (name, self.ndims)

    def length(self, sz):
        return len(self.get_index_set(sz))

    def print(self, sz, file=None):
        print(format_index_set(self.get_index_set(sz)), end="", file=file)


def format_index_set(idxset, idxset2=()):
    """the picture `print(idxset, idxset2)` draws for 2-d index sets (src/index_set.jl:296-317): one row per second component,
    top row = largest, a filled square per index of idxset and a framed one per index of idxset2"""
    idxset, idxset2 = list(idxset), list(idxset2)
    if any(len(i) != 2 for i in idxset + idxset2):
        raise ValueError("print only available for index sets with d = 2")
    if not idxset and not idxset2:
        return ""
    n = tuple(max(i[c] for i in idxset + idxset2) for c in range(2))   # `maximum` of CartesianIndex is component-wise
    a, b = set(idxset), set(idxset2)
    rows = []
    for j in range(n[1], -1, -1):
        row = "  "
        for i in range(n[0] + 1):
            if (i, j) in a:
                row += "\u25FC "
            if (i, j) in b:
                row += "\u25A3 "
        rows.append(row + "\n")
    return "".join(rows)


class SL(_IndexSetBase):
    ndims = 1

    def __init__(self):
        pass

    def get_index_set(self, sz):
        return [(0,)]


class ML(_IndexSetBase):
    ndims = 1

    def __init__(self):
        pass

    def get_index_set(self, sz):  # src/index_set.jl:49-55
        return [(l,) for l in range(sz + 1)]


class _Weighted(_IndexSetBase):
    """weighted multi-index sets: candidate box 0..sz+1 per dimension, filtered (src/index_set.jl:257); weights in (0, 1]"""

    def __init__(self, *weights):
        if len(weights) == 1 and isinstance(weights[0], int):
            weights = (1.0,) * weights[0]
        if len(weights) < 2:
            raise ValueError("to use %s, dimension must be greater than 1, got %d" % (type(self).__name__, len(weights)))
        if any(w <= 0 or w > 1 for w in weights):
            raise ValueError("weights must be in (0, 1]")
        self.weights = tuple(float(w) for w in weights)
        self.ndims = len(weights)

    def get_index_set(self, sz):
        import itertools
        if sz < 0:
            return []
        box = itertools.product(*[range(sz + 2)] * self.ndims)
        # CartesianIndices order: first dimension fastest
        return sorted((idx for idx in box if self.filter(idx, sz)), key=lambda t: t[::-1])


class FT(_Weighted):  # full tensor, src/index_set.jl:81
    def filter(self, idx, sz):
        return all(i / w <= sz for w, i in zip(self.weights, idx))


class TD(_Weighted):  # total degree, src/index_set.jl:114
    def filter(self, idx, sz):
        return sum(i / w for w, i in zip(self.weights, idx)) <= sz


class HC(_Weighted):  # hyperbolic cross, src/index_set.jl:147
    def filter(self, idx, sz):
        return math.prod(i / w + 1 for w, i in zip(self.weights, idx)) - 1 <= sz


class ZC(_Weighted):  # Zaremba cross, src/index_set.jl:180
    def filter(self, idx, sz):
        if sz == 0:
            return all(i == 0 for i in idx)
        return math.prod(max(1.0, i / w) for w, i in zip(self.weights, idx)) <= sz


class AD(_IndexSetBase):
    """adaptive index set AD(d), src/index_set.jl:195-210: no a-priori shape; grown one index at a time by profit
    (src/methods.jl:206-251) inside `max_search_space` at size `max_index_set_param`"""

    def __init__(self, d):
        if d <= 1:
            raise ValueError("to use AD, dimension must be greater than 1, got %d" % d)
        self.ndims = int(d)

    def get_index_set(self, sz):
        raise IndexSetNotImplemented("get_index_set not implemented for index sets of type AD")  # src/index_set.jl:259


class U(_IndexSetBase):
    """unbiased multilevel / multi-index index set U(d), src/index_set.jl:213-227: no bias, indices drawn at random from a
    probability mass function over `max_search_space` that is adapted to sqrt(V/C) (src/methods.jl:287-351)"""

    def __init__(self, d):
        if d <= 0:
            raise ValueError("to use U, dimension must be greater than 0, got %d" % d)
        self.ndims = int(d)

    def get_index_set(self, sz):
        raise IndexSetNotImplemented("get_index_set not implemented for index sets of type U")  # src/index_set.jl:261


def is_admissable(index_set, index):
    """src/index_set.jl:265-277: index not yet in the set and all its backward neighbours are"""
    s = set(index_set)
    if index in s or any(i < 0 for i in index):
        return False
    for k in range(len(index)):
        if index[k] > 0:
            nb = tuple(v - (1 if t == k else 0) for t, v in enumerate(index))
            if nb not in s:
                return False
    return True


AbstractIndexSet = (SL, ML, _Weighted, AD, U)


class MC:
    pass


class QMC:
    pass


# distributions (src/distribution.jl): constructors with the reference's argument checks
def Uniform(a=0.0, b=1.0):
    if not (math.isfinite(a) and math.isfinite(b)) or not a < b:
        raise ValueError("Uniform: a and b must be finite with a < b")
    return (0, [float(a), float(b)])


def Normal(mu=0.0, sigma=1.0):
    if not (math.isfinite(mu) and math.isfinite(sigma)) or not sigma > 0:
        raise ValueError("Normal: mu finite, sigma finite and > 0")
    return (1, [float(mu), float(sigma)])


def TruncatedNormal(mu=0.0, sigma=1.0, a=-2.0, b=2.0):
    if not all(math.isfinite(v) for v in (mu, sigma, a, b)) or not sigma > 0 or not a < b:
        raise ValueError("TruncatedNormal: finite parameters, sigma > 0, a < b")
    return (2, [float(mu), float(sigma), float(a), float(b)])


def Weibull(k=2.0, lam=math.sqrt(2.0)):
    if not (math.isfinite(k) and math.isfinite(lam)) or not (k > 0 and lam > 0):
        raise ValueError("Weibull: k and lambda finite and > 0")
    return (3, [float(k), float(lam)])


class ModelSpec:
    """Device sample function: registry name, index dimension, parameters, KL bases per level (lognormal models)."""

    def __init__(self, name, ndims=1, params=(), bases=None):
        self.name, self.ndims, self.params, self.bases = name, ndims, list(params), dict(bases or {})


# ---------------------------------------------------------------------------------------------------------------
# backend: where `parallel_sample!` runs.  GpuBackend = libmle_cuda on one GPU; ShardedBackend = one rank of a
# torch.distributed job (each rank takes a contiguous slice of the point numbers, one all-gather of the moments).
# ---------------------------------------------------------------------------------------------------------------
class GpuBackend:
    def __init__(self, engine, time_kernels=False):
        """time_kernels: also return the device time of every call (CUDA events, ~4 us per call); the estimator's default cost
        model is the wall time of the call, like the reference's (src/sample.jl:27), and does not need it"""
        self.engine, self.time_kernels = engine, time_kernels

    @classmethod
    def for_workers(cls, nb_of_workers=1, first_device=0, **kw):
        """the reference's `nb_of_workers` option (src/parse.jl:224-234: the size of the pmap worker pool of src/sample.jl:43-45,
        :89-91) as a number of GPUs: ONE multi-GPU context (mle_init_multi) on devices first_device .. first_device+nb_of_workers-1;
        the library splits every request over them (whole shifts per GPU, small requests stay on one GPU)"""
        from .engine import Engine
        devs = list(range(first_device, first_device + int(nb_of_workers)))
        return cls(Engine(devs if len(devs) > 1 else devs[0]), **kw)

    def lattice(self, z, s, nmax=2**32 - 1):
        return self.engine.lattice(z, s, nmax)

    def dists(self, dists):
        return self.engine.dists(dists)

    def model(self, spec):
        m = self.engine.model(spec.name, spec.ndims, spec.params)
        for level, phi in spec.bases.items():
            m.set_basis(level, phi)
        return m

    def update_samples(self, model, lattice, dists, entries, mc_seed):
        """entries = [(index, shifts, k0, n, m)] -> list of moment tables; one device synchronisation for all of them"""
        return self.engine.update_samples(model, lattice, dists, entries, mc_seed=mc_seed)

    def sample_unbiased(self, model, lattice, dists, index, shifts, k0, n, m, mc_seed, inv_prob):
        """unbiased estimators: every idx <= index on the same points, accumulator reduced on the device (mle_sample_unbiased)
        -> (moments[nidx, R, q, 2, 3], acc[R, q, 3], seconds)"""
        out = self.engine.sample_unbiased(model, lattice, dists, index, shifts, k0, n, inv_prob, m=m, mc_seed=mc_seed,
                                          want_time=self.time_kernels)
        return out if self.time_kernels else (out[0], out[1], 0.0)

    def sample_batch(self, model, lattice, dists, index, shifts, k0, n, m, mc_seed, want_samples=False):
        """-> (moments, seconds) or, with want_samples (the `save_samples` option), (moments, seconds, samples[R, q, 2, n])"""
        out = self.engine.sample_batch(model, lattice, dists, index, shifts, k0, n, m=m, mc_seed=mc_seed,
                                       want_time=self.time_kernels, want_samples=want_samples)
        out = out if isinstance(out, tuple) else (out,)
        t = out[-1] if self.time_kernels else 0.0
        if want_samples:
            return out[0], t, out[1]
        return out[0], t


def split_range(k0, n, world, rank):
    """contiguous slice of the point numbers k0 .. k0+n-1 owned by `rank` (first n % world ranks get one more)"""
    base, rem = divmod(n, world)
    cnt = base + (1 if rank < rem else 0)
    start = k0 + rank * base + min(rank, rem)
    return start, cnt


class ShardedBackend:
    """One rank of a multi-GPU run: shards every `sample_batch` over the ranks of a torch.distributed group and combines the
    per-rank moment triplets with ONE all-gather, merged in rank order (deterministic).  `inner` is the local backend."""

    def __init__(self, inner, group=None, device=None, merge=None):
        import torch.distributed as dist
        self.inner, self.dist, self.group, self.device = inner, dist, group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        # every rank must take the same adaptive decisions, so the time charged to an index (the reference's default cost model is
        # the wall time of the call, src/sample.jl:27) has to be the SAME number on all ranks: the calls below return the maximum
        # over the ranks of the local wall time (it travels with the moments in the one all-gather) and the estimator uses it
        self.consensus_time = True
        self.last_time = 0.0
        if merge is None:
            from .engine import moments_merge as merge
        self.merge = merge

    def lattice(self, *a, **k):
        return self.inner.lattice(*a, **k)

    def dists(self, *a, **k):
        return self.inner.dists(*a, **k)

    def model(self, spec):
        return self.inner.model(spec)

    def update_samples(self, model, lattice, dists, entries, mc_seed):
        """all due indices of one `update_samples` (src/sample.jl:8-13) at once: every rank runs its slice of every entry
        (one device synchronisation through the inner backend's batched call when it has one) and ONE all-gather carries the
        moment tables of all entries"""
        import torch
        q = getattr(model, "nqoi", 1)
        mine, local = [], []
        for index, shifts, k0, n, m in entries:
            start, cnt = split_range(k0, n, self.world, self.rank)
            if cnt > 0:
                mine.append(len(local))
                local.append((index, shifts, start, cnt, m))
            else:
                mine.append(None)
        t0 = _time.perf_counter()
        if local and hasattr(self.inner, "update_samples"):
            moms = self.inner.update_samples(model, lattice, dists, local, mc_seed)
        else:
            moms = [self.inner.sample_batch(model, lattice, dists, i, sh, k0, n, m, mc_seed)[0] for i, sh, k0, n, m in local]
        t_local = _time.perf_counter() - t0
        shapes = [((1 if e[1] is None else len(e[1])), q, 2, 3) for e in entries]
        parts = [np.asarray(moms[j], dtype=np.float64).reshape(-1) if j is not None else np.zeros(int(np.prod(sh)))
                 for j, sh in zip(mine, shapes)]
        payload = torch.from_numpy(np.concatenate(parts + [np.array([t_local])]))
        if self.device is not None:
            payload = payload.to(self.device)
        out = [torch.empty_like(payload) for _ in range(self.world)]
        self.dist.all_gather(out, payload, group=self.group)
        flat = [o.cpu().numpy() for o in out]
        self.last_time = max(float(f[-1]) for f in flat)
        result, off = [], 0
        for sh in shapes:
            size = int(np.prod(sh))
            acc = flat[0][off:off + size].reshape(sh).copy()
            for f in flat[1:]:
                acc = self.merge(acc, f[off:off + size].reshape(sh))
            result.append(acc)
            off += size
        return result

    def sample_batch(self, model, lattice, dists, index, shifts, k0, n, m, mc_seed, want_samples=False):
        """moments [R, q, 2, 3] of the whole point range (and, on request, the raw samples [R, q, 2, n] of all ranks in point
        order: `save_samples` and the unbiased estimators' per-sample sums need them on every rank)"""
        import torch
        start, cnt = split_range(k0, n, self.world, self.rank)
        R = 1 if shifts is None else len(shifts)
        q = getattr(model, "nqoi", 1)
        smp = None
        t0 = _time.perf_counter()
        if cnt > 0:
            out = self.inner.sample_batch(model, lattice, dists, index, shifts, start, cnt, m, mc_seed, want_samples=want_samples)
            mom = out[0]
            smp = out[2] if want_samples else None
        else:
            mom = np.zeros((R, q, 2, 3))
        t = _time.perf_counter() - t0   # local wall time; the maximum over the ranks is what every rank charges
        payload = torch.from_numpy(np.concatenate([np.asarray(mom, dtype=np.float64).reshape(-1), [t]]))
        if self.device is not None:
            payload = payload.to(self.device)
        out = [torch.empty_like(payload) for _ in range(self.world)]
        self.dist.all_gather(out, payload, group=self.group)
        parts = [o.cpu().numpy() for o in out]
        acc = parts[0][:-1].reshape(mom.shape).copy()
        for p in parts[1:]:
            acc = self.merge(acc, p[:-1].reshape(mom.shape))
        tmax = max(float(p[-1]) for p in parts)
        self.last_time = tmax
        if not want_samples:
            return acc, tmax
        # second all-gather: the raw samples, padded to the largest slice (slices differ by at most one point)
        counts = [split_range(k0, n, self.world, r)[1] for r in range(self.world)]
        width = max(counts)
        mine = np.zeros((R, q, 2, width))
        if cnt > 0:
            mine[..., :cnt] = smp
        send = torch.from_numpy(mine)
        if self.device is not None:
            send = send.to(self.device)
        recv = [torch.empty_like(send) for _ in range(self.world)]
        self.dist.all_gather(recv, send, group=self.group)
        full = np.concatenate([r.cpu().numpy()[..., :c] for r, c in zip(recv, counts)], axis=-1)
        return acc, tmax, full


# ---------------------------------------------------------------------------------------------------------------
# Estimator
# ---------------------------------------------------------------------------------------------------------------
_DEFAULTS = dict(nb_of_qoi=None, aggregation="max", qoi_with_max_var=1, continuate=True, nb_of_tols=10,
                 continuation_mul_factor=1.3, min_splitting=0.5, max_splitting=0.99, robustify_bias_estimate=True,
                 do_mse_splitting=True, do_regression=True, min_index_set_param=2, max_index_set_param=None,
                 sample_mul_factor=1.2, nb_of_warm_up_samples=None, nb_of_shifts=10, nb_of_uncertainties=None,
                 cost_model=None, point_generator=None, shifts=None, shift_seed=0, mc_seed=1, name=None, folder=None,
                 save=False, verbose=False, nb_of_workers=None, max_search_space=None, penalization=0.5, acceptance_rate=1.0,
                 save_samples=False)
# which keys an estimator accepts, src/options.jl:10-41
_COMMON_KEYS = ("nb_of_warm_up_samples", "nb_of_qoi", "qoi_with_max_var", "aggregation", "nb_of_tols", "continuation_mul_factor",
                "continuate", "save", "save_samples", "verbose", "folder", "name", "cost_model", "nb_of_workers",
                "nb_of_uncertainties", "max_index_set_param", "min_index_set_param")
_INDEX_SET_KEYS = ("min_splitting", "max_splitting", "robustify_bias_estimate", "do_mse_splitting", "do_regression")
_AD_ONLY = ("max_search_space", "penalization", "acceptance_rate")  # src/options.jl:36
_U_KEYS = ("max_search_space", "sample_mul_factor")                 # src/options.jl:37
_QMC_KEYS = ("nb_of_shifts", "point_generator", "sample_mul_factor")  # src/options.jl:41-43
_ENGINE_KEYS = ("shifts", "shift_seed", "mc_seed")  # not in the reference: its random inputs come from Julia's global RNG


def valid_options(index_set, sample_method):
    """get_valid_options(index_set, sample_method), src/options.jl:27-43, plus the engine's own seeding keys"""
    keys = list(_COMMON_KEYS)
    if isinstance(index_set, AD):
        keys += _INDEX_SET_KEYS + _AD_ONLY
    elif isinstance(index_set, U):
        keys += _U_KEYS
    elif not isinstance(index_set, SL):
        keys += _INDEX_SET_KEYS
    if isinstance(sample_method, QMC):
        keys += _QMC_KEYS
    return list(dict.fromkeys(keys + list(_ENGINE_KEYS)))


def _is_int(v):
    return isinstance(v, (int, np.integer)) and not isinstance(v, (bool, np.bool_))


def _is_real(v):
    return isinstance(v, (int, float, np.integer, np.floating)) and not isinstance(v, (bool, np.bool_))


def _check_type(key, val, ok, wanted):  # check_type, src/check_input.jl:25
    if not ok:
        raise ValueError("in Estimator, optional key %s must be of type %s, got %s" % (key, wanted, type(val).__name__))


def _check(key, ok, what):  # check_larger_than & co., src/check_input.jl:9-13
    if not ok:
        raise ValueError("in Estimator, optional key %s must be %s." % (key, what))


def _check_real(key, val, lo=None, lo_open=True, hi=None):
    _check_type(key, val, _is_real(val), "Real")
    _check(key, math.isfinite(val), "finite")
    if lo is not None:
        _check(key, val > lo if lo_open else val >= lo, ("larger than %s" if lo_open else "larger than or equal to %s") % lo)
    if hi is not None:
        _check(key, val <= hi, "smaller than or equal to %s" % hi)


def parse_options(index_set, sample_method, distributions, options):
    """The option checks of src/parse.jl:35-307 (same defaults, same rules, ArgumentError -> ValueError).  Returns the complete
    option dict.  Differences, all on the control-plane side: `save` defaults to False and writes JSON instead of JLD2,
    `verbose` defaults to False (nothing is printed), `nb_of_workers` is validated and ignored (the device backend decides)."""
    unknown = [k for k in options if k not in valid_options(index_set, sample_method)]
    if unknown:
        raise ValueError("in Estimator, invalid option %s found" % sorted(unknown)[0])  # src/estimator.jl:126-130
    qmc, unbiased, adaptive = isinstance(sample_method, QMC), isinstance(index_set, U), isinstance(index_set, AD)
    d = index_set.ndims
    o = dict(_DEFAULTS)
    o.update(options)
    if o["nb_of_warm_up_samples"] is None:
        o["nb_of_warm_up_samples"] = (2 if unbiased else 1) if qmc else 20          # src/parse.jl:35-36
    else:
        _check_type("nb_of_warm_up_samples", o["nb_of_warm_up_samples"], _is_int(o["nb_of_warm_up_samples"]), "Signed")
        _check("nb_of_warm_up_samples", o["nb_of_warm_up_samples"] > 1, "larger than 1")  # src/parse.jl:39
    if o["nb_of_qoi"] is not None:                                                  # src/parse.jl:44-51
        _check_type("nb_of_qoi", o["nb_of_qoi"], _is_int(o["nb_of_qoi"]), "Signed")
        _check("nb_of_qoi", o["nb_of_qoi"] > 0, "larger than 0")
    _check_type("aggregation", o["aggregation"], isinstance(o["aggregation"], str), "String")
    if o["aggregation"] not in ("max", "sum", "norm"):
        raise ValueError("in Estimator, optional key aggregation must be 'sum', 'max' or 'norm'")  # src/parse.jl:58
    for key in ("continuate", "save_samples", "save", "verbose", "robustify_bias_estimate", "do_mse_splitting", "do_regression"):
        _check_type(key, o[key], isinstance(o[key], (bool, np.bool_)), "Bool")
    _check_type("nb_of_tols", o["nb_of_tols"], _is_int(o["nb_of_tols"]), "Integer")
    _check("nb_of_tols", o["nb_of_tols"] > 0, "larger than 0")                       # src/parse.jl:70-77
    _check_real("continuation_mul_factor", o["continuation_mul_factor"], lo=1)      # src/parse.jl:79-87
    for key in ("min_splitting", "max_splitting"):                                  # src/parse.jl:89-113
        _check_real(key, o[key], lo=0, hi=1)
    if not o["min_splitting"] <= o["max_splitting"]:
        raise ValueError("in Estimator, optional key min_splitting must be smaller than or equal to optional key max_splitting")
    if o["folder"] is None:
        o["folder"] = os.getcwd()                                                   # src/parse.jl:116-117
    _check_type("folder", o["folder"], isinstance(o["folder"], str), "String")
    if o["save"] and not os.path.isdir(o["folder"]):
        raise ValueError("%s is not a directory!" % o["folder"])                    # src/parse.jl:122
    if o["name"] is None:                                                           # get_valid_filename, src/parse.jl:142-153
        stem, cntr = "UntitledEstimator", 0
        if o["save"]:
            while os.path.isfile(os.path.join(o["folder"], "%s%s.json" % (stem, cntr or ""))):
                cntr += 1
        o["name"] = "%s%s.jld2" % (stem, cntr or "")
    else:
        _check_type("name", o["name"], isinstance(o["name"], str), "String")
        if not o["name"].endswith(".jld2"):
            if "." in o["name"]:
                raise ValueError("in Estimator, optional key name must not contain a .")  # src/parse.jl:134
            o["name"] += ".jld2"
    if o["cost_model"] is not None:
        _check_type("cost_model", o["cost_model"], callable(o["cost_model"]), "Function")  # src/parse.jl:173-176
    if o["max_index_set_param"] is None:
        o["max_index_set_param"] = 10 * d                                           # src/parse.jl:193-194
    for key in ("max_index_set_param", "min_index_set_param"):                      # src/parse.jl:193-213
        _check_type(key, o[key], _is_int(o[key]), "Signed")
        _check(key, o[key] > 0, "larger than 0")
    if not o["min_index_set_param"] <= o["max_index_set_param"]:
        raise ValueError("in Estimator, optional key min_index_set_param must be smaller than or equal to "
                         "optional key max_index_set_param")
    _check_real("sample_mul_factor", o["sample_mul_factor"])                        # src/parse.jl:215-222
    if o["nb_of_workers"] is None:
        o["nb_of_workers"] = lambda index: 1
    elif not callable(o["nb_of_workers"]):                                          # src/parse.jl:224-235
        _check_type("nb_of_workers", o["nb_of_workers"], _is_int(o["nb_of_workers"]), "Union{Integer, Function}")
        _check("nb_of_workers", o["nb_of_workers"] > 0, "larger than 0")
        nw = int(o["nb_of_workers"])
        o["nb_of_workers"] = lambda index: nw
    if callable(o["nb_of_shifts"]):
        if unbiased:
            raise ValueError("in Estimator, optional key nb_of_shifts cannot be a Function when using unbiased estimation")
    else:                                                                           # src/parse.jl:237-249
        _check_type("nb_of_shifts", o["nb_of_shifts"], _is_int(o["nb_of_shifts"]), "Union{Integer, Function}")
        _check("nb_of_shifts", o["nb_of_shifts"] > 1, "larger than 1")
        R = int(o["nb_of_shifts"])
        o["nb_of_shifts"] = lambda index: R
    if o["point_generator"] is not None:                                            # src/parse.jl:251-255
        # the reference wants an AbstractRNG (a LatticeRule32); here the rule is given by its generating vector
        pg = o["point_generator"]
        ok = isinstance(pg, (list, tuple, np.ndarray)) and len(pg) > 0 and all(_is_int(v) and 0 <= v < 2**32 for v in pg)
        _check_type("point_generator", pg, ok, "a generating vector (sequence of UInt32)")
        if len(pg) < len(distributions):
            raise ValueError("in Estimator, optional key point_generator must have at least %d dimensions." % len(distributions))
    if o["nb_of_uncertainties"] is None:
        m_all = len(distributions)
        o["nb_of_uncertainties"] = lambda index: m_all                              # src/parse.jl:272-275
    else:
        _check_type("nb_of_uncertainties", o["nb_of_uncertainties"], callable(o["nb_of_uncertainties"]), "Function")
    if adaptive or unbiased:
        if o["max_search_space"] is None:
            o["max_search_space"] = TD(d)                                           # src/parse.jl:263-264 (fails for d = 1 like the reference)
        _check_type("max_search_space", o["max_search_space"], isinstance(o["max_search_space"], AbstractIndexSet),
                    "AbstractIndexSet")
        if o["max_search_space"].ndims != d:
            raise ValueError("in Estimator, dimensions of max_search_space must be equal to %d." % d)  # src/parse.jl:267
    for key in ("penalization", "acceptance_rate"):                                 # src/parse.jl:278-298
        _check_real(key, o[key], lo=0, lo_open=False, hi=1)
    _check("qoi_with_max_var", _is_int(o["qoi_with_max_var"]) and o["qoi_with_max_var"] > 0, "larger than 0")  # src/parse.jl:300-307
    if o["nb_of_qoi"] is not None:
        _check("qoi_with_max_var", o["qoi_with_max_var"] <= o["nb_of_qoi"], "smaller than or equal to %d" % o["nb_of_qoi"])
    return o


def index_seed(mc_seed, index):
    """The MC stream of one index: SplitMix64 finaliser over (mc_seed, index).  The device's uniforms are keyed by (seed, point
    number, coordinate) only, and point numbers restart at 0 on every index, so each index needs its own seed -- the reference
    draws rand(m) from one global stream (src/sample.jl:42), i.e. independent inputs on every level."""
    x = int(mc_seed) & 0xFFFFFFFFFFFFFFFF
    for c in index:
        x = (x + 0x9E3779B97F4A7C15 * (int(c) + 1)) & 0xFFFFFFFFFFFFFFFF
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        x ^= x >> 31
    return x


class Estimator:
    """Estimator(index_set, sample_method, sample_function, distributions; options...)  (src/estimator.jl:121)"""

    def __init__(self, index_set, sample_method, sample_function, distributions, backend, **options):
        if not isinstance(sample_function, ModelSpec):
            raise TypeError("sample_function must be a ModelSpec naming a device model (arbitrary closures cannot run on the GPU)")
        if isinstance(distributions, tuple):
            distributions = [distributions]
        self.adaptive = isinstance(index_set, AD)
        self.unbiased = isinstance(index_set, U)
        self.qmc = isinstance(sample_method, QMC)
        o = parse_options(index_set, sample_method, list(distributions), options)   # src/estimator.jl:123-134
        self.index_set, self.sample_method, self.spec, self.distributions, self.backend = (
            index_set, sample_method, sample_function, list(distributions), backend)
        d = index_set.ndims
        self.options = o
        self.report = None
        if o["verbose"]:   # src/print.jl: the progress report of a verbose run
            from .report import Reporter
            self.report = Reporter(self)
        self.model = backend.model(sample_function)
        self.nqoi = o["nb_of_qoi"] or getattr(self.model, "nqoi", 1)
        _check("qoi_with_max_var", o["qoi_with_max_var"] <= self.nqoi, "smaller than or equal to %d" % self.nqoi)  # src/parse.jl:305
        self.dists = backend.dists(self.distributions)
        s = len(self.distributions)
        self.lattice = None
        if self.qmc:
            z = o["point_generator"]  # a generating vector; None -> default LatticeRule32(s), src/parse.jl:251-254
            self.lattice = backend.lattice(z, s)
        # internals (src/internals.jl:35-54): per index cumulative moments [R, q, 2, 3], counters, and one shift per (index, r)
        space = o["max_search_space"] if (self.adaptive or self.unbiased) else index_set   # src/internals.jl:36,133
        box = space.get_index_set(o["max_index_set_param"])
        self.search_space = set(box)
        # ADInternals, src/internals.jl:150-171
        self.old_set, self.active_set, self.max_index_set, self.logbook = set(), set(), set(), []
        self.ad_boundary = {(0,) * d}
        self._ad_rng = _random.Random(o["shift_seed"])
        self.moments, self.nb_of_samples, self.total_work, self.total_time = {}, {}, {}, {}
        # the loop queries mean / var / varest / rates of every index many times between two sample! calls (bias, splitting, mse,
        # history); the values only change when samples arrive, so they are memoised until then (the reference recomputes them
        # from the stored sample vectors on every query)
        self._memo = {}
        self._no_memo = False
        self._bias_deps = {}     # bias memo key -> indices whose means entered it
        self._keys_cache = None
        self._boundary_cache = {}
        self.samples = {}  # save_samples: index -> list of [R, q, 2, n] blocks (X = 0: differences, 1: fine values)
        self.shifts = {}
        rng = np.random.default_rng(o["shift_seed"])
        for index in sorted(box, key=lambda t: t[::-1]):  # column-major order of the index box, src/internals.jl:138
            R = o["nb_of_shifts"](index) if self.qmc else 1
            self.nb_of_samples[index], self.total_work[index], self.total_time[index] = 0, 0.0, 0.0
            if self.qmc:
                given = o["shifts"]
                self.shifts[index] = np.asarray(given[index] if isinstance(given, dict) else given(index), dtype=np.float64) \
                    if given is not None else rng.random((R, s))
        self.current_index_set = set()
        self.sz = self.max_sz = 0
        # UInternals, src/internals.jl:208-225: pmf = normalised product of Geometric(1 - exp(-1.5)) weights; the accumulator is
        # kept as (n, mean, M2) per (shift, qoi) of the per-sample sums  sum_idx dQ_idx / Prob(idx)
        self.pmf, self.acc = {}, None
        if self.unbiased:
            pg = 1.0 - math.exp(-1.5)
            self.pmf = {i: math.prod((1.0 - pg) ** k * pg for k in i) for i in sorted(box, key=lambda t: t[::-1])}
            tot = sum(self.pmf.values())
            self.pmf = {i: v / tot for i, v in self.pmf.items()}
            self._u_rng = np.random.default_rng(o["shift_seed"] + 1)

    def __getitem__(self, key):  # estimator[:key], src/options.jl:46
        return self.options[key]

    # ---- the hot path -----------------------------------------------------------------------------------------------
    def Prob(self, index):  # src/methods.jl:333: CartesianIndex `>=` is the column-major (last index most significant) order
        return sum(v for k, v in self.pmf.items() if k[::-1] >= index[::-1])

    def _sample_unbiased(self, index, n):
        """sample! for U estimators (src/sample.jl:21-34 with the append_samples! of :65-74, :113-124): the sample function of
        the reference returns (dQ, Qf) for EVERY idx <= index from one input vector.  Here the device model is evaluated once per
        idx on the same points (same shifts, same point numbers, same leading coordinates) and the per-sample sums
        Z = sum_idx (1 / Prob(idx)) dQ_idx  are reduced to (n, mean, M2) on the device (mle_sample_unbiased); backends without
        that call (the CPU oracle of the tests, the torch.distributed sharding) and `save_samples` runs take the raw samples and
        reduce on the host."""
        import itertools
        from .engine import moments_merge
        istart = sum(self.nb_of_samples.values()) + 1   # src/sample.jl:25: U counts the samples of ALL indices
        m = self.options["nb_of_uncertainties"](index)
        seed = index_seed(self.options["mc_seed"], index)   # one input vector per sample, shared by every idx <= index
        t0 = _time.perf_counter()
        box = sorted(itertools.product(*[range(i + 1) for i in index]), key=lambda t: t[::-1])   # first entry fastest
        if hasattr(self.backend, "sample_unbiased") and not self.options["save_samples"]:
            # device reduction: ONE call, every idx on the same points, Z = sum_idx (1/Prob(idx)) dQ_idx summed per sample on the
            # GPU; only the moment tables come back (include/mle_cuda.h: mle_sample_unbiased)
            inv_prob = [1.0 / self.Prob(idx) for idx in box]
            moms, part, t_cons = self.backend.sample_unbiased(self.model, self.lattice, self.dists, list(index),
                                                              self.shifts.get(index), istart - 1, n, m, seed, inv_prob)
            for idx, mom in zip(box, moms):
                self.moments[idx] = moments_merge(self.moments[idx], mom) if idx in self.moments else np.array(mom, dtype=np.float64)
            part = np.array(part, dtype=np.float64)
        else:
            Z, t_cons = None, 0.0
            for idx in box:
                mom, _t, smp = self.backend.sample_batch(self.model, self.lattice, self.dists, list(idx), self.shifts.get(index),
                                                         istart - 1, n, m, seed, want_samples=True)
                self.moments[idx] = moments_merge(self.moments[idx], mom) if idx in self.moments else np.array(mom, dtype=np.float64)
                if self.options["save_samples"]:
                    self.samples.setdefault(idx, []).append(np.array(smp))
                term = (1.0 / self.Prob(idx)) * smp[:, :, 0, :]   # src/sample.jl:71: 1 / Prob(estimator, idx) * s_diff
                Z = term if Z is None else Z + term
                t_cons += _t
            zm = Z.mean(axis=-1)
            part = np.stack([np.full(zm.shape, float(n)), zm, ((Z - zm[..., None]) ** 2).sum(axis=-1)], axis=-1)   # [R, q, 3]
        t = t_cons if getattr(self.backend, "consensus_time", False) else _time.perf_counter() - t0
        self.acc = part if self.acc is None else moments_merge(self.acc, part)
        self._memo.clear()
        self.nb_of_samples[index] += n
        cm = self.options["cost_model"]
        self.total_work[index] += 0.0 if cm is None else n * cm(index)
        self.total_time[index] += t

    def sample(self, index, n):
        """sample!(estimator, index, n), src/sample.jl:21-34"""
        if self.report:
            self.report.sample_header(index, n, not self.has_samples_at_index(index))   # src/sample.jl:22-23
        if self.unbiased:
            self._sample_unbiased(index, n)
        else:
            self._sample_index(index, n)
        if self.report:
            self.report.sample_footer()

    def _sample_index(self, index, n):
        istart = self.nb_of_samples[index] + 1
        m = self.options["nb_of_uncertainties"](index)
        seed = index_seed(self.options["mc_seed"], index)
        t0 = _time.perf_counter()
        if self.options["save_samples"]:  # src/history.jl:92-95: keep the raw samples as well
            mom, _gpu_t, smp = self.backend.sample_batch(self.model, self.lattice, self.dists, list(index),
                                                         self.shifts.get(index), istart - 1, n, m, seed, want_samples=True)
            self.samples.setdefault(index, []).append(np.array(smp))
        else:
            mom, _gpu_t = self.backend.sample_batch(self.model, self.lattice, self.dists, list(index),
                                                    self.shifts.get(index), istart - 1, n, m, seed)
        # wall time of the call (src/sample.jl:27); a sharded backend returns the rank-agreed maximum instead (see ShardedBackend)
        t = _gpu_t if getattr(self.backend, "consensus_time", False) else _time.perf_counter() - t0
        if index in self.moments:
            from .engine import moments_merge
            self.moments[index] = moments_merge(self.moments[index], mom)
        else:
            self.moments[index] = np.array(mom, dtype=np.float64)
        self._invalidate(index)
        self.nb_of_samples[index] += n                                  # src/internals.jl:90
        cm = self.options["cost_model"]
        self.total_work[index] += 0.0 if cm is None else n * cm(index)  # src/internals.jl:98
        self.total_time[index] += t                                     # src/internals.jl:106

    def _sample_many(self, due):
        from .engine import moments_merge
        entries = [(list(index), self.shifts.get(index), self.nb_of_samples[index], n, self.options["nb_of_uncertainties"](index))
                   for index, n in due]
        t0 = _time.perf_counter()
        moms = self.backend.update_samples(self.model, self.lattice, self.dists, entries, self.options["mc_seed"])
        t = self.backend.last_time if getattr(self.backend, "consensus_time", False) else _time.perf_counter() - t0
        cm = self.options["cost_model"]
        work = [n * cm(index) for index, n in due]
        if not sum(work) > 0:   # a cost model that returns 0: share the time out by sample counts
            work = [float(n) for _, n in due]
        self._memo.clear()
        self._bias_deps.clear()
        for (index, n), mom, w in zip(due, moms, work):
            self.moments[index] = moments_merge(self.moments[index], mom) if index in self.moments else np.array(mom, dtype=np.float64)
            self.nb_of_samples[index] += n
            self.total_work[index] += w
            self.total_time[index] += t * w / sum(work)   # the call's wall time, shared out by modelled work

    def update_samples(self, n_opt):  # src/sample.jl:8-13, :15-19 (U: n_opt are numbers of ADDITIONAL samples)
        if self.unbiased:
            for index in n_opt:
                if n_opt[index] > 0:
                    self.sample(index, n_opt[index])
            return
        due = [(index, n_opt[index] - self.nb_of_samples[index]) for index in n_opt if n_opt[index] > self.nb_of_samples[index]]
        # with a deterministic cost model the per-index wall time is not needed, so all due indices go to the device in ONE
        # call (mle_update_samples: batches back to back, a single synchronisation); otherwise index by index like the reference,
        # whose default cost model is the measured time of each sample! call (src/sample.jl:27)
        # (QMC only: mle_update_samples takes ONE mc_seed, and MC needs a different stream on every index)
        if len(due) > 1 and self.qmc and self.options["cost_model"] is not None and hasattr(self.backend, "update_samples") \
                and not self.options["save_samples"]:
            return self._sample_many(due)
        for index, n_due in due:
            self.sample(index, n_due)

    def has_samples_at_index(self, index):
        return self.nb_of_samples.get(index, 0) > 0

    # ---- statistics (src/methods.jl:28-68, :282) -------------------------------------------------------------------------
    def _aggregate(self, values):
        a = self.options["aggregation"]
        if a == "sum":
            return float(np.sum(values))
        if a == "norm":
            return float(np.linalg.norm(values))
        return float(values[self.options["qoi_with_max_var"] - 1])

    def _cached(self, key, fn):
        v = None if self._no_memo else self._memo.get(key)   # (_no_memo: tests run the loop without any memoisation)
        if v is None:
            v = self._memo[key] = fn()
        return v

    def _invalidate(self, index=None):
        """new samples at `index` (None: the index sets changed): per-index statistics of the other indices stay valid"""
        if index is None:
            drop = [k for k in self._memo if k[0] in ("rates", "bias")]
            self._keys_cache = None
        else:
            # a bias estimate depends on the means of a few boundary indices only (recorded when it was computed): new samples
            # elsewhere -- the QMC loop mostly feeds the coarse levels -- leave it valid
            deps = self._bias_deps
            drop = [k for k in self._memo if k[0] == "rates" or k[1] == index or (k[0] == "bias" and index in deps.get(k, (index,)))]
        for k in drop:
            del self._memo[k]
            if k[0] == "bias":
                self._bias_deps.pop(k, None)

    def _per_qoi(self, index, X, f):
        return self._cached(("pq", index, X, f), lambda: self._per_qoi_now(index, X, f))

    def _per_qoi_now(self, index, X, f):
        if index not in self.moments:           # mean / var of an empty sample vector: NaN like the reference
            return np.full(self.nqoi, float("nan"))
        mom = self.moments[index][:, :, X, :]  # [R, q, 3]
        with np.errstate(divide="ignore", invalid="ignore"):
            if f == "mean":
                per_shift = mom[..., 1]
            else:
                per_shift = mom[..., 2] / (mom[..., 0] - 1.0)  # corrected variance; NaN with one sample like the reference
        return per_shift.mean(axis=0)  # MC: R == 1

    def mean(self, index=None):
        if index is None and self.unbiased:
            # src/methods.jl:351 reads accumulator(estimator, qoi) = accumulator[qoi, 1]: the FIRST shift only (replicated)
            return self._aggregate(self.acc[0, :, 1]) if self.acc is not None else float("nan")
        if index is None:
            return sum(self.mean(i) for i in self.keys())
        return self._aggregate(self._per_qoi(index, 0, "mean"))

    def var(self, index):
        return self._aggregate(self._per_qoi(index, 0, "var"))

    def mean0(self, index):
        return self._aggregate(self._per_qoi(index, 1, "mean"))

    def var0(self, index):
        return self._aggregate(self._per_qoi(index, 1, "var"))

    def varest(self, index=None):
        if index is None and self.unbiased:  # src/methods.jl:341-343
            if self.acc is None:
                return float("nan")
            with np.errstate(divide="ignore", invalid="ignore"):
                if not self.qmc:
                    return self._aggregate(self.acc[0, :, 2] / (self.acc[0, :, 0] - 1.0) / self.acc[0, :, 0])
                return self._aggregate(self.acc[:, :, 1].var(axis=0, ddof=1) / self.acc.shape[0])
        if index is None:
            return sum(self.varest(i) for i in self.keys())
        if not self.qmc:
            return self.var(index) / self.nb_of_samples[index]          # src/methods.jl:66

        def now():
            means = self.moments[index][:, :, 0, 1]                     # per-shift means [R, q]
            return self._aggregate(means.var(axis=0, ddof=1) / means.shape[0])   # src/methods.jl:282
        return self._cached(("varest", index), now)

    def cost(self, index):  # src/methods.jl:64
        tot = self.total_time if self.options["cost_model"] is None else self.total_work
        n = self.nb_of_samples[index]
        return tot[index] / n if n else float("nan")  # 0/0 = NaN in the reference (U: indices only seen below a larger one)

    def time(self, index):
        n = self.nb_of_samples[index]
        return self.total_time[index] / n if n else float("nan")

    def keys(self):
        # (the set only grows between two clear() calls, so its size is its version)
        c = self._keys_cache
        if c is None or c[0] != len(self.current_index_set) or self._no_memo:
            c = self._keys_cache = (len(self.current_index_set), sorted(self.current_index_set, key=lambda t: t[::-1]))
        return list(c[1])

    def all_keys(self):
        if self.unbiased:  # every idx <= a sampled index holds samples
            return sorted(self.moments, key=lambda t: t[::-1])
        return sorted((i for i, n in self.nb_of_samples.items() if n > 0), key=lambda t: t[::-1])

    # ---- rates and regression (src/methods.jl:70-163) -----------------------------------------------------------------------
    @staticmethod
    def _interp1(x, y):
        n = len(x)
        if n >= 2 and len(set(x)) >= 2 and all(math.isfinite(v) for v in y):
            # straight-line least squares in closed form (the generic solver below costs 60 us per call, and the loop fits
            # a line two or three times per sample! call)
            xm, ym = math.fsum(x) / n, math.fsum(y) / n
            sxx = math.fsum((a - xm) ** 2 for a in x)
            b = math.fsum((a - xm) * (c - ym) for a, c in zip(x, y)) / sxx
            return np.array([ym - b * xm, b])
        A = np.column_stack([np.ones(len(x)), np.asarray(x, dtype=float)])
        return np.linalg.lstsq(A, np.asarray(y, dtype=float), rcond=None)[0]

    def _rates(self, g, idx, dir):
        return self._cached(("rates", getattr(g, "__name__", None) or id(g), idx, dir), lambda: self._rates_now(g, idx, dir))

    def _rates_now(self, g, idx, dir):
        m = idx[dir] - 1
        if m < 2:
            return (float("nan"), float("nan"))
        e = tuple(1 if i == dir else 0 for i in range(len(idx)))
        xs, ys = [], []
        for k in range(1, m + 1):
            index = tuple(a - k * b for a, b in zip(idx, e))
            y = g(index)
            x = m - k + 1
            if not (math.isnan(y) or y == 0):
                xs.append(x)
                ys.append(math.log2(abs(y)))
        return tuple(self._interp1(xs, ys))

    def _rate(self, g, sgn):
        if isinstance(self.index_set, SL):
            return None
        out = []
        for dir in range(self.index_set.ndims):
            top = max(k[dir] for k in self.keys()) + 1
            idx = tuple(top if i == dir else 0 for i in range(self.index_set.ndims))
            out.append(sgn * self._rates(g, idx, dir)[1])
        return out

    def alpha(self):
        return self._rate(self.mean, -1)

    def beta(self):
        return self._rate(self.var, -1)

    def gamma(self):
        return self._rate(self.cost, 1)

    def _interp(self, f):
        # the reference filters on the LENGTH OF THE STORED SAMPLE VECTORS (src/methods.jl:121-122), which for U estimators also
        # grow on every idx below a drawn index; n of the first (shift, qoi) moment triplet is that length
        stored = {i: (int(self.moments[i][0, 0, 1, 0]) if i in self.moments else 0) for i in self.nb_of_samples}
        have = [i for i, n in stored.items() if (n > 0 if self.qmc else n > 1)]
        A = np.array([[1.0] + list(i) for i in have]).reshape(len(have), 1 + self.index_set.ndims)
        vals = [f(i) for i in have]
        # log2(0) = -Inf rows are dropped (`valid = .!isinf.(y)`), NaN rows are NOT: they poison the fit, as in the reference
        y = np.array([float("nan") if v != v else (math.log2(v) if v > 0 else float("-inf")) for v in vals])
        ok = ~np.isinf(y)
        if np.isnan(y[ok]).any():
            return np.full(A.shape[1], float("nan"))
        try:
            return np.linalg.lstsq(A[ok], y[ok], rcond=None)[0]
        except Exception:  # noqa: BLE001
            return np.full(A.shape[1], float("nan"))

    def _regress(self, g, index):
        est = []
        for dir in range(len(index)):
            p = self._rates(g, index, dir)
            est.append(2.0 ** (p[0] + index[dir] * p[1]))
        est = [v for v in est if not math.isnan(v)]
        if est:
            return float(np.mean(est))
        p = self._interp(g)
        return 2.0 ** (p[0] + float(np.dot(p[1:], index)))

    def regress_var(self, index):
        return self._regress(self.var, index)

    def regress_cost(self, index):
        cm = self.options["cost_model"]
        return self._regress(self.cost, index) if cm is None else cm(index)

    def regress_nb_of_samples(self, index_set, eps, theta, L):  # src/methods.jl:134-163
        warm = self.options["nb_of_warm_up_samples"]
        if isinstance(self.index_set, SL):
            return {(0,): warm}
        if self.qmc or not (self.options["do_regression"] and L > 2):
            return {i: warm for i in index_set}
        vars_ = {i: self.regress_var(i) for i in index_set}
        costs = {i: self.regress_cost(i) for i in index_set}
        sig = self.Sigma() + sum(math.sqrt(vars_[i] * costs[i]) for i in index_set)
        return {i: max(2, min(self._n_opt(eps, theta, vars_[i], costs[i], sig), warm)) for i in index_set}

    # ---- MSE splitting, sample allocation, bias (src/methods.jl:165-201) ------------------------------------------------------
    def Sigma(self):
        return sum(math.sqrt(self.var(i) * self.cost(i)) for i in self.keys())

    @staticmethod
    def _n_opt(eps, theta, v, c, sig):
        return int(math.ceil(1.0 / (theta * eps**2) * math.sqrt(v / c) * sig))

    def optimal_nb_of_samples(self, index, eps, theta):
        return self._n_opt(eps, theta, self.var(index), self.cost(index), self.Sigma())

    # ---- adaptivity (src/methods.jl:206-261) -------------------------------------------------------------------------------
    def profit(self, index):
        return abs(self.mean(index)) / math.sqrt(self.var(index) * self.cost(index)) ** self.options["penalization"]

    def find_index_with_max_profit(self):
        indices = sorted(self.active_set, key=lambda t: t[::-1])
        profits = [self.profit(i) for i in indices]
        best = int(np.argmax(profits))
        if self._ad_rng.random() < self.options["acceptance_rate"] or len(profits) == 1:
            max_index = indices[best]
        else:
            max_index = indices[self._ad_rng.choice([t for t in range(len(profits)) if t != best])]
        if self.report:
            self.report.largest_profit(max_index, profits[best], indices, profits)
        return max_index

    def new_index_set(self, sz):
        """the indices entering at iteration sz: the boundary of the a-priori set, or, for AD, the index with the largest
        profit moves from the active to the old set and its admissable forward neighbours become active"""
        if not self.adaptive:
            return self.boundary(sz)
        d = self.index_set.ndims
        if not self.active_set:
            max_index = (0,) * d
            self.active_set.add(max_index)
            new = [max_index]
        else:
            max_index = self.find_index_with_max_profit()
            self.old_set.add(max_index)
            self.active_set.discard(max_index)
            new = [max_index]
            for k in range(d):
                nb = tuple(v + (1 if t == k else 0) for t, v in enumerate(max_index))
                if is_admissable(self.old_set, nb):
                    if nb in self.search_space:
                        self.active_set.add(nb)
                        new.append(nb)
                    else:
                        import warnings
                        warnings.warn("Trying to add index %s outside search space to index set, ignoring..." % (nb,),
                                      stacklevel=2)  # warn_max_index, src/print.jl:247 (unconditional in the reference)
                        self.max_index_set.add(nb)
        self.logbook.append((set(self.old_set), set(self.active_set), max_index))
        self._invalidate()
        return new

    def set_sz(self, n):  # src/internals.jl:256-262
        if self.adaptive and n > self.max_sz:
            self.ad_boundary = set(self.active_set)
        self.sz, self.max_sz = n, max(n, self.max_sz)

    def max_level_exceeded(self):  # src/methods.jl:208, :203
        if self.adaptive:
            return not (self.search_space - self.current_index_set)
        return self.sz >= self.options["max_index_set_param"]

    def clear(self):  # src/internals.jl:116, :200-206: samples persist across tolerances, the index sets do not
        self.current_index_set.clear()
        self._keys_cache = None
        self.active_set.clear()
        self.old_set.clear()
        self.max_index_set.clear()
        self.logbook.clear()

    def boundary(self, cntr):
        b = self._boundary_cache.get(cntr)
        if b is None:
            b = self._boundary_cache[cntr] = self._boundary_now(cntr)
        return b

    def _boundary_now(self, cntr):
        prev = set(self.index_set.get_index_set(cntr - 1)) if cntr > 0 else set()
        return [i for i in self.index_set.get_index_set(cntr) if i not in prev]

    def bias(self, sz=None):
        if isinstance(self.index_set, SL):
            return 0.0
        if self.unbiased:
            return 0.0     # src/methods.jl:347
        sz = self.sz if sz is None else sz
        key = ("bias", sz, self.max_sz)
        v = None if self._no_memo else self._memo.get(key)
        if v is None:
            used = set()
            v = self._memo[key] = self._bias_now(sz, used)
            self._bias_deps[key] = used
        return v

    def _bias_now(self, sz, used=None):
        used = set() if used is None else used   # the indices whose means enter the estimate (see _invalidate)
        if self.adaptive:  # src/methods.jl:253-260
            used.update(self.active_set)
            used.update(self.ad_boundary)
            b = lambda indices: abs(sum(self.mean(i) for i in sorted(indices, key=lambda t: t[::-1])))
            return b(self.active_set) if sz != self.max_sz else min(b(self.active_set), b(self.ad_boundary))
        nxt = [i for i in self.boundary(sz + 1) if i in self.current_index_set]
        if nxt and not self.options["robustify_bias_estimate"]:
            used.update(self.boundary(sz + 1))
            return abs(sum(self.mean(i) for i in self.boundary(sz + 1)))
        xs, ys = [], []
        for x in range(max(1, sz - 2), sz + 1):
            used.update(self.boundary(x))
            v = abs(sum(self.mean(i) for i in self.boundary(x)))   # plain floats: no numpy warning to silence
            y = math.log2(v) if v > 0 else float("-inf")
            if math.isfinite(y):
                xs.append(x)
                ys.append(y)
        if not xs:
            return float("nan")
        p = self._interp1(xs, ys)
        return 2.0 ** (p[0] + (sz + 1) * p[1])

    def compute_splitting(self, eps):
        b = self.bias(self.max_sz)
        a, bmax = self.options["min_splitting"], self.options["max_splitting"]
        s = 1 - b**2 / eps**2
        return a if math.isnan(s) else min(bmax, max(a, s))

    def mse(self):
        return self.varest() + self.bias() ** 2

    def rmse(self):
        return math.sqrt(self.mse())

    def converged(self, eps, theta):
        if self.unbiased:
            return self.mse() <= eps**2  # src/methods.jl:349
        return self.bias() ** 2 <= (1 - theta) * eps**2 or self.mse() <= eps**2

    # ---- QMC sample growth (src/methods.jl:265-280) -----------------------------------------------------------------------------
    def next_number_of_samples(self, index):
        f, n = self.options["sample_mul_factor"], self.nb_of_samples[index]
        if f == 2:
            return {index: 1 << n.bit_length()}  # nextpow(2, n + 1)
        if f <= 1:
            return {index: n + 1}
        return {index: int(math.ceil(n * f))}

    # ---- unbiased estimation (src/methods.jl:287-331) -------------------------------------------------------------------------
    def next_number_of_samples_unbiased(self):
        n_total = sum(self.nb_of_samples.values())
        n0 = max(self.options["nb_of_warm_up_samples"], n_total)
        f = self.options["sample_mul_factor"]
        n = 1 if f <= 1 else int(math.ceil(n0 * (f - 1)))
        u = self._u_rng.random(n)
        counts, psum = {}, 0.0
        for index, p in self.pmf.items():  # column-major order (a Dict in the reference: its iteration order is unspecified)
            counts[index] = int(np.count_nonzero((psum <= u) & (u <= psum + p)))
            psum += p
        return counts

    def update_pmf(self):
        def f(index):
            with np.errstate(divide="ignore", invalid="ignore"):
                if index not in self.moments or self.nb_of_samples[index] == 0:
                    return float("nan")
                return math.sqrt(self.var(index) / self.cost(index)) if self.var(index) >= 0 else float("nan")
        p = self._interp(f)
        if not np.any(np.isnan(p)) and np.all(p[1:] < 0):
            for index in self.keys():
                fi = f(index)
                self.pmf[index] = 2.0 ** (p[0] + sum(a * b for a, b in zip(p[1:], index))) if (math.isnan(fi) or math.isinf(fi)) else fi
            tot = sum(self.pmf.values())
            self.pmf = {i: v / tot for i, v in self.pmf.items()}

    def find_index_with_max_var_over_cost(self):
        keys = self.keys()
        best, bv = 0, None
        for pos, i in enumerate(keys):  # np.argmax semantics: the first maximum, and a NaN counts as the maximum
            v = self.varest(i) / self.cost(i)
            if v != v:
                best = pos
                break
            if bv is None or v > bv:
                best, bv = pos, v
        return keys[best]


# ---------------------------------------------------------------------------------------------------------------
# run (src/run.jl:20-142) and History (src/history.jl)
# ---------------------------------------------------------------------------------------------------------------
class History:
    """History (src/history.jl:8-117): one dict per tolerance; history[key] reads the last entry, history[i] the i-th."""

    def __init__(self):
        self.data = []

    def __getitem__(self, key):
        return self.data[key] if isinstance(key, int) else self.data[-1][key]

    def __len__(self):
        return len(self.data)

    def __contains__(self, key):
        return key in self.data[-1]

    def __repr__(self):
        return "MultilevelEstimators.jl history file"  # src/history.jl:117

    @property
    def path(self):
        """Where save() writes: folder/name with the reference's .jld2 suffix replaced by .json (JLD2 itself is a Julia format
        this mirror does not produce; the Julia glue keeps the reference's own `@save`, src/history.jl:107)."""
        name = self["name"]
        return os.path.join(self["folder"], (name[:-5] if name.endswith(".jld2") else name) + ".json")

    def save(self, path=None):
        """save(history), src/history.jl:107.  Index keys become their string form, arrays become nested lists."""
        import json

        def enc(v):
            if isinstance(v, dict):
                return {(str(tuple(int(c) for c in k)) if isinstance(k, tuple) else str(k)): enc(x) for k, x in v.items()}
            if isinstance(v, np.ndarray):
                return enc(v.tolist())
            if isinstance(v, (list, tuple, set)):
                return [enc(x) for x in (sorted(v) if isinstance(v, set) else v)]
            if isinstance(v, np.generic):
                return v.item()
            return v

        path = path or self.path
        tmp = path + ".tmp"
        with open(tmp, "w") as f:
            json.dump([enc(h) for h in self.data], f)
        os.replace(tmp, path)
        return path

    @classmethod
    def load(cls, path):
        import ast
        import json

        def is_index(k):
            return k.startswith("(") and k.endswith(")")

        def dec(v, key=None):
            if isinstance(v, dict):
                return {(ast.literal_eval(k) if is_index(k) else k): dec(x, k) for k, x in v.items()}
            if isinstance(v, list):
                if key in ("current_index_set", "index_set"):
                    return [tuple(x) for x in v]
                return [dec(x) for x in v]
            return v

        h = cls()
        with open(path) as f:
            h.data = [dec(e) for e in json.load(f)]
        return h


def get_tols(est, tol):  # src/methods.jl:15
    if est["continuate"]:
        n, p = est["nb_of_tols"], est["continuation_mul_factor"]
        return [p ** (n - 1 - i) * tol for i in range(n)]
    return [tol]


def _run_unbiased(est, eps):  # src/run.jl:145-200
    rep = est.report
    if rep:
        rep.header(eps)
    index_set = sorted(est.pmf, key=lambda t: t[::-1])
    if rep:
        rep.index_set(index_set)
    est.current_index_set.update(index_set)
    est._invalidate()
    v = est.varest()
    while not v <= eps**2:  # NaN (no samples yet) keeps looping, like `is_converged = varest(estimator) <= eps^2` in the reference
        if rep:
            rep.status()
        n_opt = est.next_number_of_samples_unbiased()
        if rep:
            rep.optimal_nb_of_samples(n_opt)
        est.update_samples(n_opt)
        est.update_pmf()
        if rep:
            rep.pmf()
            rep.unbiased_convergence(eps)
        v = est.varest()
        if rep:
            rep.index_set(index_set)
    if rep:
        rep.convergence(True)


def _run(est, eps):  # src/run.jl:41-142
    if est.unbiased:
        return _run_unbiased(est, eps)
    rep = est.report
    if rep:
        rep.header(eps)
    sl = isinstance(est.index_set, SL)

    def splitting():
        return 0.5 if sl else (est.compute_splitting(eps) if est["do_mse_splitting"] else est["min_splitting"])

    L = 0
    theta = 0.5 if sl else est["min_splitting"]
    done = False
    while not done:
        if rep:
            rep.level(L)
        index_set = est.new_index_set(L)
        if rep:
            rep.index_set(index_set)
        ns = est.regress_nb_of_samples(index_set, eps, theta, L)
        for index in index_set:
            if not est.has_samples_at_index(index):
                est.sample(index, ns[index])
        est.current_index_set.update(index_set)
        est._invalidate()
        est.set_sz(L)
        if rep:
            rep.status()
        theta = splitting()
        if not est.qmc:
            n_opt = {i: est.optimal_nb_of_samples(i, eps, theta) for i in est.keys()}
            if rep:
                rep.optimal_nb_of_samples(n_opt)
            est.update_samples(n_opt)
        else:
            while est.varest() > theta * eps**2:
                n_opt = est.next_number_of_samples(est.find_index_with_max_var_over_cost())
                if rep:
                    rep.optimal_nb_of_samples(n_opt)
                est.update_samples(n_opt)
                theta = splitting()
                if rep:
                    rep.qmc_convergence(eps, theta)
        if sl:
            done = True
        else:
            if rep:
                rep.status()
                if L >= 2:
                    rep.mse_analysis(eps, theta)
            done = L > est["min_index_set_param"] and est.converged(eps, theta)
            if not done and est.max_level_exceeded():
                if est["verbose"]:  # warn_max_level, src/print.jl:241
                    import warnings
                    warnings.warn("Maximum %s L = %d reached, no convergence yet." % (
                        "level" if est.index_set.ndims == 1 else "index set parameter", est["max_index_set_param"]))
                done = True
            elif not done:
                L += 1
    if est.adaptive:
        est.ad_boundary = set(est.active_set)  # update_boundary, src/run.jl:138
    if rep:
        rep.convergence(True if sl else est.converged(eps, theta))


def update_history(history, est, tol, elapsed):  # src/history.jl:55-103
    allk = est.all_keys()
    h = {"type": "Estimator{%s, %s}" % (type(est.index_set).__name__, type(est.sample_method).__name__),
         "ndims": est.index_set.ndims, "name": est["name"], "folder": est["folder"], "elapsed": elapsed, "tol": tol,
         "current_index_set": est.keys(), "index_set": allk, "mse": est.mse(), "rmse": est.rmse(), "mean": est.mean(),
         "varest": est.varest(), "bias": est.bias(),
         "E": {i: est.mean0(i) for i in allk}, "V": {i: est.var0(i) for i in allk},
         "dE": {i: est.mean(i) for i in allk}, "dV": {i: est.var(i) for i in allk},
         "T": {i: est.time(i) for i in allk}, "alpha": est.alpha(), "beta": est.beta(), "gamma": est.gamma(),
         "nb_of_samples": dict(est.nb_of_samples)}
    if est["cost_model"] is not None:
        h["W"] = {i: est.cost(i) for i in allk}
    if est.adaptive:
        h["logbook"] = list(est.logbook)  # src/history.jl:87
    if est["save_samples"]:  # src/history.jl:92-95; arrays [R, q, n] per index
        cat = {i: np.concatenate(b, axis=-1) for i, b in est.samples.items()}
        h["samples"] = {i: a[:, :, 1, :] for i, a in cat.items()}
        h["samples_diff"] = {i: a[:, :, 0, :] for i, a in cat.items()}
    history.data.append(h)
    if est["save"]:  # src/history.jl:99-102: the file is rewritten after every tolerance
        history.save()


def run(est, tol):
    """run(estimator, tol) -> History  (src/run.jl:20-38)"""
    tols = list(tol) if isinstance(tol, (list, tuple, np.ndarray)) else get_tols(est, tol)
    if any(t <= 0 for t in tols):
        raise ValueError("in run, supplied tolerance(s) must be larger than 0")  # src/run.jl:23
    history = History()
    for t in tols:
        est.clear()
        t0 = _time.perf_counter()
        _run(est, t)
        update_history(history, est, t, _time.perf_counter() - t0)
    return history
