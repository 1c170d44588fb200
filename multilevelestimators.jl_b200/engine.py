"""Thin object layer over the C ABI (include/mle_cuda.h): one `Engine` per GPU, handles for lattice rules,
distribution vectors and device models, and the two hot-path calls `points` / `sample_batch`.

This is host plumbing only -- every number is produced by the CUDA kernels in csrc/.  The calls mirror
`parallel_sample!` (src/sample.jl:39-56, :79-102 of the reference): zero-based point numbers, one shift vector
per (index, shift), `m = nb_of_uncertainties(index)` leading coordinates.
"""
import ctypes as C

import numpy as np

from . import _lib

_dp = C.POINTER(C.c_double)


class MLEError(RuntimeError):
    """CUDA / state failure inside libmle_cuda (ErrorException in the Julia wrapper)."""


class MLEArgumentError(ValueError):
    """Invalid argument (the reference's ArgumentError, src/check_input.jl)."""


def _as_dp(a):
    return a.ctypes.data_as(_dp) if a is not None else None


class _Handle:
    def __init__(self, engine, handle, destroy):
        self.engine, self.handle, self._destroy = engine, handle, destroy

    def close(self):
        if self.handle is not None and self.engine.ctx is not None:
            self._destroy(self.engine.ctx, self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceLattice(_Handle):
    s = 0
    nmax = 0


class DeviceDists(_Handle):
    m = 0


class DeviceModel(_Handle):
    """Handle of a built-in CUDA sample function (stands where the Julia closure `sample_function` stood)."""
    name = ""
    ndims = 1
    nqoi = 1

    def set_basis(self, level, phi):
        """level: int, or the multi-index (lx, ly) of the 2-D multi-index model (packed as lx + 16*ly for the C ABI)"""
        if not isinstance(level, (int, np.integer)):
            level = int(level[0]) + 16 * int(level[1]) if len(level) == 2 else int(level[0])
        phi = np.ascontiguousarray(phi, dtype=np.float64)
        self.engine._check(self.engine.lib.mle_model_set_basis(self.engine.ctx, self.handle, int(level), phi.shape[0],
                                                               phi.shape[1], _as_dp(phi)))
        return self


class Engine:
    def __init__(self, device=0):
        """device: a CUDA ordinal, or a list of ordinals for ONE multi-GPU context (mle_init_multi: the reference's
        `nb_of_workers`, src/parse.jl:224-234 -- requests are split over the devices inside the library)"""
        self.lib = _lib.load()
        ctx = C.c_void_p()
        if isinstance(device, (list, tuple)):
            devs = (C.c_int * len(device))(*[int(x) for x in device])
            rc = self.lib.mle_init_multi(devs, len(device), C.byref(ctx))
        else:
            rc = self.lib.mle_init(int(device), C.byref(ctx))
        if rc != _lib.MLE_OK:
            msg = self.lib.mle_last_error(None).decode()
            raise MLEError("mle_init failed (%d): %s" % (rc, msg))
        self.ctx = ctx
        self.device = device
        self._idx_cache, self._shift_cache = {}, {}
        sm, hbm, cc, nd = C.c_int(), C.c_size_t(), C.c_int(), C.c_int()
        self.lib.mle_device_info(self.ctx, C.byref(sm), C.byref(hbm), C.byref(cc))
        self.lib.mle_device_count(self.ctx, C.byref(nd))
        self.sm_count, self.hbm_bytes, self.cc, self.device_count = sm.value, hbm.value, cc.value, nd.value

    # ---- plumbing ----------------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc == _lib.MLE_OK:
            return
        msg = self.lib.mle_last_error(self.ctx).decode()
        if rc in (_lib.MLE_ERR_ARG, _lib.MLE_ERR_RANGE):
            raise MLEArgumentError(msg)
        raise MLEError("libmle_cuda error %d: %s" % (rc, msg))

    def close(self):
        if getattr(self, "ctx", None) is not None:
            self.lib.mle_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        s = C.c_void_p()
        self._check(self.lib.mle_stream(self.ctx, C.byref(s)))
        return s.value or 0

    @property
    def launch_count(self):
        c = C.c_uint64()
        self._check(self.lib.mle_launch_count(self.ctx, C.byref(c)))
        return c.value

    def sync(self):
        self._check(self.lib.mle_sync(self.ctx))

    # ---- handles -----------------------------------------------------------------------------------------------
    def genvec(self, name="K_3600_32"):
        p, n = C.POINTER(C.c_uint32)(), C.c_int()
        rc = self.lib.mle_genvec(name.encode(), C.byref(p), C.byref(n))
        if rc:
            raise MLEArgumentError("unknown generating vector %r" % name)
        return np.ctypeslib.as_array(p, shape=(n.value,)).astype(np.uint32)

    def lattice(self, z=None, s=None, nmax=2**32 - 1):
        """LatticeRule32(z, s, n); z None -> the default rule LatticeRule32(s) (src/lattice.jl:88)"""
        zp = None
        if z is not None:
            z = np.ascontiguousarray(z, dtype=np.uint32)
            s = len(z) if s is None else s
            if s > len(z):
                raise MLEArgumentError("LatticeRule32: number of dimensions s must be smaller than or equal to the "
                                       "length of the generating vector z")
            zp = z.ctypes.data_as(C.POINTER(C.c_uint32))
        h = C.c_void_p()
        self._check(self.lib.mle_lattice_create(self.ctx, zp, int(s), int(nmax), C.byref(h)))
        L = DeviceLattice(self, h, self.lib.mle_lattice_destroy)
        L.s, L.nmax = int(s), int(nmax)
        return L

    def dists(self, dists):
        """dists: sequence of (kind, params) pairs"""
        arr = (_lib.mle_dist * len(dists))()
        for j, (kind, pars) in enumerate(dists):
            arr[j].kind = int(kind)
            for t, v in enumerate(pars):
                arr[j].p[t] = float(v)
        h = C.c_void_p()
        self._check(self.lib.mle_dists_create(self.ctx, arr, len(dists), C.byref(h)))
        D = DeviceDists(self, h, self.lib.mle_dists_destroy)
        D.m = len(dists)
        return D

    def model(self, name, ndims=1, params=()):
        p = np.ascontiguousarray(params, dtype=np.float64)
        h = C.c_void_p()
        self._check(self.lib.mle_model_create(self.ctx, name.encode(), int(ndims), _as_dp(p) if len(p) else None, len(p),
                                              C.byref(h)))
        M = DeviceModel(self, h, self.lib.mle_model_destroy)
        q = C.c_int()
        self.lib.mle_model_nqoi(h, C.byref(q))
        M.name, M.ndims, M.nqoi = name, int(ndims), q.value
        return M

    # ---- stage 1-2 parity hook -------------------------------------------------------------------------------------
    def points(self, lattice, dists, shifts, k0, n, m=None, mc_seed=0):
        """-> float64[R, n, m]: transform.(D[1:m], get_point(gen[r], k)[1:m]) for k = k0 .. k0+n-1 (zero-based)"""
        if shifts is not None:
            shifts = np.ascontiguousarray(shifts, dtype=np.float64)
            if lattice is None or shifts.ndim != 2 or shifts.shape[1] != lattice.s:
                raise MLEArgumentError("shifts must have shape [R, s]")
            R = shifts.shape[0]
        else:
            R = 1
        if m is None:
            m = dists.m if dists is not None else lattice.s
        out = np.empty((R, int(n), int(m)))
        self._check(self.lib.mle_points(self.ctx, lattice.handle if lattice else None, dists.handle if dists else None,
                                        _as_dp(shifts), R, int(k0), int(n), int(m), int(mc_seed), _as_dp(out)))
        return out

    def points_dev(self, lattice, dists, shifts_ptr, R, k0, n, m, out_ptr, mc_seed=0):
        """device-pointer variant (asynchronous on self.stream)"""
        self._check(self.lib.mle_points_dev(self.ctx, lattice.handle if lattice else None,
                                            dists.handle if dists else None, shifts_ptr, int(R), int(k0), int(n),
                                            int(m), int(mc_seed), out_ptr))

    # ---- the hot path ----------------------------------------------------------------------------------------------
    def sample_batch(self, model, lattice, dists, index, shifts, k0, n, m=None, mc_seed=0, want_samples=False,
                     want_time=False):
        """One `parallel_sample!` call.  -> moments float64[R, nqoi, 2, 3] = (n, mean, M2) for X in (dQ, Qf)
        [, samples float64[R, nqoi, 2, n]] [, gpu_seconds]"""
        # the adaptive loop calls this thousands of times with the same index / shift arrays and 16 x (1 .. 100) samples, where a
        # call is ~23 us in the library: addresses of the (kept alive) input arrays are remembered, the output goes through
        # plain addresses (mle_sample_batch_addr) -- ~5 instead of ~15 us of marshalling per call
        key = index if isinstance(index, tuple) else tuple(np.atleast_1d(index).tolist())
        ent = self._idx_cache.get(key)
        if ent is None:
            arr = np.ascontiguousarray(key, dtype=np.int32)
            ent = self._idx_cache[key] = (arr, arr.ctypes.data, len(key))
        if shifts is not None:
            sent = self._shift_cache.get(id(shifts))
            if sent is None or sent[0] is not shifts:
                a = np.ascontiguousarray(shifts, dtype=np.float64)
                if lattice is None or a.ndim != 2 or a.shape[1] != lattice.s:
                    raise MLEArgumentError("shifts must have shape [R, s]")
                sent = (shifts, a, a.ctypes.data, a.shape[0])
                if a is shifts:  # remembered only when no copy was made: the library then reads the caller's own memory on
                    if len(self._shift_cache) > 4096:  # every call (in-place changes are seen); the object is kept alive with
                        self._shift_cache.clear()      # the entry, so its id cannot be reused while the entry exists
                    self._shift_cache[id(shifts)] = sent
            sh_addr, R = sent[2], sent[3]
        else:
            sh_addr, R = None, 1
        if m is None:
            m = dists.m if dists is not None else lattice.s
        mom = np.empty((R, model.nqoi, 2, 3))
        smp = np.empty((R, model.nqoi, 2, int(n))) if want_samples else None
        t = C.c_double(0.0) if want_time else None
        rc = self.lib.mle_sample_batch_addr(self.ctx, model.handle, lattice.handle if lattice else None,
                                            dists.handle if dists else None, ent[1], ent[2], sh_addr, R, int(k0), int(n), int(m),
                                            int(mc_seed), mom.__array_interface__["data"][0],
                                            smp.__array_interface__["data"][0] if want_samples else None,
                                            C.addressof(t) if want_time else None)
        if rc:
            self._check(rc)
        out = (mom,)
        if want_samples:
            out += (smp,)
        if want_time:
            out += (t.value,)
        return out[0] if len(out) == 1 else out

    def update_samples(self, model, lattice, dists, entries, mc_seed=0):
        """One `update_samples` call: entries = [(index, shifts or None, k0, n, m), ...] -> list of moment tables
        float64[R, nqoi, 2, 3]; all batches are issued back to back with a single synchronisation (mle_update_samples)"""
        cnt = len(entries)
        if cnt == 0:
            return []
        d = len(np.atleast_1d(entries[0][0]))
        idx = np.ascontiguousarray([np.atleast_1d(e[0]) for e in entries], dtype=np.int32).reshape(cnt, d)
        sh = [None if e[1] is None else np.ascontiguousarray(e[1], dtype=np.float64) for e in entries]
        for a in sh:
            if a is not None and (lattice is None or a.ndim != 2 or a.shape[1] != lattice.s):
                raise MLEArgumentError("shifts must have shape [R, s]")
        R = (C.c_int * cnt)(*[1 if a is None else a.shape[0] for a in sh])
        k0 = (C.c_uint64 * cnt)(*[int(e[2]) for e in entries])
        n = (C.c_uint64 * cnt)(*[int(e[3]) for e in entries])
        m = (C.c_int * cnt)(*[int(e[4]) for e in entries])
        moms = [np.empty((R[i], model.nqoi, 2, 3)) for i in range(cnt)]
        shp = (_dp * cnt)(*[_as_dp(a) if a is not None else None for a in sh])
        mop = (_dp * cnt)(*[_as_dp(a) for a in moms])
        self._check(self.lib.mle_update_samples(self.ctx, model.handle, lattice.handle if lattice else None,
                                                dists.handle if dists else None, cnt, idx.ctypes.data_as(C.POINTER(C.c_int32)), d,
                                                shp, R, k0, n, m, int(mc_seed), mop))
        return moms

    def sample_unbiased(self, model, lattice, dists, index, shifts, k0, n, inv_prob, m=None, mc_seed=0, want_time=False):
        """One `sample!` call of an unbiased estimator (mle_sample_unbiased; src/sample.jl:65-74, :113-124): every idx in
        zero(index):index on the same points.  inv_prob[idx] = 1 / Prob(idx), column-major over the box (first entry fastest).
        -> (moments float64[nidx, R, nqoi, 2, 3], acc float64[R, nqoi, 3][, seconds])"""
        idx = np.ascontiguousarray(np.atleast_1d(index), dtype=np.int32)
        nidx = int(np.prod(idx.astype(np.int64) + 1))
        inv_prob = np.ascontiguousarray(inv_prob, dtype=np.float64).ravel()
        if inv_prob.size != nidx:
            raise MLEArgumentError("inv_prob must hold one entry per idx in zero(index):index")
        if shifts is not None:
            shifts = np.ascontiguousarray(shifts, dtype=np.float64)
            if lattice is None or shifts.ndim != 2 or shifts.shape[1] != lattice.s:
                raise MLEArgumentError("shifts must have shape [R, s]")
            R = shifts.shape[0]
        else:
            R = 1
        if m is None:
            m = dists.m if dists is not None else lattice.s
        mom = np.empty((nidx, R, model.nqoi, 2, 3))
        acc = np.empty((R, model.nqoi, 3))
        secs = C.c_double(0.0)
        self._check(self.lib.mle_sample_unbiased(self.ctx, model.handle, lattice.handle if lattice else None,
                                                 dists.handle if dists else None, idx.ctypes.data_as(C.POINTER(C.c_int32)),
                                                 len(idx), _as_dp(shifts), int(R), int(k0), int(n), int(m), int(mc_seed),
                                                 _as_dp(inv_prob), _as_dp(mom), _as_dp(acc),
                                                 C.byref(secs) if want_time else None))
        return (mom, acc, secs.value) if want_time else (mom, acc)

    def sample_batch_dev(self, model, lattice, dists, index, shifts_ptr, R, k0, n, m, moments_ptr, samples_ptr=None,
                         mc_seed=0):
        """device-pointer variant: asynchronous on self.stream, nothing crosses PCIe"""
        idx = np.ascontiguousarray(np.atleast_1d(index), dtype=np.int32)
        self._check(self.lib.mle_sample_batch_dev(self.ctx, model.handle, lattice.handle if lattice else None,
                                                  dists.handle if dists else None,
                                                  idx.ctypes.data_as(C.POINTER(C.c_int32)), len(idx), shifts_ptr, int(R),
                                                  int(k0), int(n), int(m), int(mc_seed), moments_ptr, samples_ptr))


def moments_merge(acc, part):
    """-> acc (+) part for arrays of (n, mean, M2) triplets (last axis 3); Chan merge in libmle_cuda; the inputs are not
    modified"""
    lib = _lib.load()
    acc = np.array(acc, dtype=np.float64, order="C", copy=True)
    if not (isinstance(part, np.ndarray) and part.dtype == np.float64 and part.flags.c_contiguous):
        part = np.ascontiguousarray(part, dtype=np.float64)
    assert acc.shape == part.shape and acc.shape[-1] == 3
    # plain addresses (the marshalling of two ndarray.ctypes objects costs more than the merge of a 16 x 2 table)
    lib.mle_moments_merge_addr(acc.__array_interface__["data"][0], part.__array_interface__["data"][0], acc.size // 3)
    return acc
