"""ctypes front-end for the CPU oracle (oracle/mle_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs -- never by the product package.  See the header of mle_oracle.c for the
reference file:line each function restates.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmle_oracle.so")

UNIFORM, NORMAL, TRUNCNORMAL, WEIBULL = 0, 1, 2, 3
MODEL_EXPSUM, MODEL_PROBLEM18, MODEL_LOGNORMAL1D, MODEL_LOGNORMAL2D = 0, 1, 2, 3


def build(force=False):
    src = os.path.join(_HERE, "mle_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        dp, ip, u32p = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_uint32)
        L.orc_reversebits32.restype = C.c_uint32
        L.orc_reversebits32.argtypes = [C.c_uint32]
        L.orc_phi.restype = C.c_uint32
        L.orc_phi.argtypes = [C.c_uint32]
        L.orc_get_point.restype = C.c_int
        L.orc_get_point.argtypes = [u32p, C.c_int, C.c_uint64, dp, C.c_uint64, dp]
        for name in ("orc_erfinv", "orc_norminv", "orc_normcdf"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_double]
        L.orc_log_pos.restype = C.c_double
        L.orc_log_pos.argtypes = [C.c_double, dp]
        L.orc_pow_pos.restype = C.c_double
        L.orc_pow_pos.argtypes = [C.c_double, C.c_double]
        L.orc_transform.restype = C.c_double
        L.orc_transform.argtypes = [C.c_int, dp, C.c_double]
        L.orc_splitmix64.restype = C.c_uint64
        L.orc_splitmix64.argtypes = [C.c_uint64]
        L.orc_mc_uniform.restype = C.c_double
        L.orc_mc_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.orc_points.restype = C.c_int
        L.orc_points.argtypes = [u32p, C.c_int, C.c_uint64, ip, dp, C.c_int, dp, C.c_int, C.c_uint64, C.c_uint64,
                                 C.c_uint64, dp]
        L.orc_model_create.restype = C.c_void_p
        L.orc_model_create.argtypes = [C.c_int, C.c_int, dp, C.c_int]
        L.orc_model_set_basis.restype = C.c_int
        L.orc_model_set_basis.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp]
        L.orc_model_destroy.restype = None
        L.orc_model_destroy.argtypes = [C.c_void_p]
        L.orc_model_nqoi.restype = C.c_int
        L.orc_model_nqoi.argtypes = [C.c_void_p]
        L.orc_model_eval.restype = C.c_int
        L.orc_model_eval.argtypes = [C.c_void_p, ip, dp, C.c_int, dp, dp]
        L.orc_tridiag_mid.restype = C.c_double
        L.orc_tridiag_mid.argtypes = [dp, C.c_int, C.c_int]
        L.orc_frontal_mid.restype = C.c_double
        L.orc_frontal_mid.argtypes = [dp, C.c_int, C.c_int, C.c_int]
        L.orc_frontal_mid_rect.restype = C.c_double
        L.orc_frontal_mid_rect.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_tridiag_mid_full.restype = C.c_double
        L.orc_tridiag_mid_full.argtypes = [dp, C.c_int]
        L.orc_sample_batch.restype = C.c_int
        L.orc_sample_batch.argtypes = [C.c_void_p, u32p, C.c_int, C.c_uint64, ip, dp, C.c_int, ip, dp, C.c_int,
                                       C.c_uint64, C.c_uint64, C.c_uint64, dp, dp]
        L.orc_num_threads.restype = C.c_int
        L.orc_genvec.restype = u32p
        L.orc_genvec.argtypes = [C.c_int, ip]
        L.orc_selftest_nocontract.restype = C.c_int
        assert L.orc_selftest_nocontract() == 1, "oracle was built with floating-point contraction"
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def _u32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32)) if a is not None else None


def genvec(name="K_3600_32", s=None):
    """first s entries of a published generating vector (default LatticeRule32(s), src/lattice.jl:88)"""
    n = C.c_int(0)
    p = lib().orc_genvec({"K_3600_32": 0, "CKN_250_20": 1}[name], C.byref(n))
    z = np.ctypeslib.as_array(p, shape=(n.value,)).astype(np.uint32)
    return z if s is None else z[:s].copy()


def get_point(z, k, shift=None, nmax=2**32 - 1):
    z = np.ascontiguousarray(z, dtype=np.uint32)
    sh = None if shift is None else np.ascontiguousarray(shift, dtype=np.float64)
    x = np.empty(len(z))
    rc = lib().orc_get_point(_u32p(z), len(z), nmax, _dp(sh), k, _dp(x))
    if rc:
        raise ValueError("k > n: the reference returns rand(s) here (src/lattice.jl:166,175)")
    return x


def transform(kind, pars, x):
    p = np.zeros(4)
    p[:len(pars)] = pars
    return lib().orc_transform(kind, _dp(p), float(x))


def dist_arrays(dists, m=None):
    """dists: list of (kind, [params]) -> (kinds int32[m], pars float64[m,4])"""
    if dists is None:
        return None, None
    m = len(dists) if m is None else m
    kinds = np.zeros(m, dtype=np.int32)
    pars = np.zeros((m, 4))
    for j in range(m):
        kd, pr = dists[j]
        kinds[j] = kd
        pars[j, :len(pr)] = pr
    return kinds, pars


def points(z, dists, shifts, k0, n, m=None, nmax=2**32 - 1, mc_seed=0):
    """out[R, n, m]; z None -> MC counter stream (R = 1)"""
    if z is not None:
        z = np.ascontiguousarray(z, dtype=np.uint32)
        s = len(z)
    else:
        s = 0
    if m is None:
        m = s if dists is None else len(dists)
    kinds, pars = dist_arrays(dists, m)
    if shifts is not None:
        shifts = np.ascontiguousarray(shifts, dtype=np.float64)
        R = shifts.shape[0]
        assert shifts.shape[1] == s
    else:
        R = 1
    out = np.empty((R, n, m))
    rc = lib().orc_points(_u32p(z), s, nmax, _ip(kinds), _dp(pars), m, _dp(shifts), R, k0, n, mc_seed, _dp(out))
    if rc:
        raise ValueError("orc_points failed: %d" % rc)
    return out


class Model:
    def __init__(self, kind, ndims=1, params=()):
        p = np.ascontiguousarray(params, dtype=np.float64)
        self.kind, self.ndims = kind, ndims
        self.h = lib().orc_model_create(kind, ndims, _dp(p) if len(p) else None, len(p))
        self.nqoi = lib().orc_model_nqoi(self.h)

    def set_basis(self, level, phi):
        """level: int, or (lx, ly) for the multi-index 2-D model (stored under lx + 16*ly)"""
        if not isinstance(level, (int, np.integer)):
            level = int(level[0]) + 16 * int(level[1])
        phi = np.ascontiguousarray(phi, dtype=np.float64)
        rc = lib().orc_model_set_basis(self.h, level, phi.shape[0], phi.shape[1], _dp(phi))
        assert rc == 0

    def eval(self, index, xi):
        idx = np.ascontiguousarray(np.atleast_1d(index), dtype=np.int32)
        xi = np.ascontiguousarray(xi, dtype=np.float64)
        dq, qf = np.empty(self.nqoi), np.empty(self.nqoi)
        rc = lib().orc_model_eval(self.h, _ip(idx), _dp(xi), len(xi), _dp(dq), _dp(qf))
        if rc:
            raise ValueError("orc_model_eval failed: %d" % rc)
        return dq, qf

    def __del__(self):
        try:
            lib().orc_model_destroy(self.h)
        except Exception:
            pass


def sample_batch(model, z, dists, index, shifts, k0, n, m=None, nmax=2**32 - 1, mc_seed=0, want_samples=False):
    """moments[R, nqoi, 2, 3] = (n, mean, M2) for X = dQ, Qf; optionally samples[R, nqoi, 2, n]"""
    if z is not None:
        z = np.ascontiguousarray(z, dtype=np.uint32)
        s = len(z)
    else:
        s = 0
    if m is None:
        m = len(dists) if dists is not None else s
    kinds, pars = dist_arrays(dists, m)
    if shifts is not None:
        shifts = np.ascontiguousarray(shifts, dtype=np.float64)
        R = shifts.shape[0]
    else:
        R = 1
    idx = np.ascontiguousarray(np.atleast_1d(index), dtype=np.int32)
    mom = np.empty((R, model.nqoi, 2, 3))
    smp = np.empty((R, model.nqoi, 2, n)) if want_samples else None
    rc = lib().orc_sample_batch(model.h, _u32p(z), s, nmax, _ip(kinds), _dp(pars), m, _ip(idx), _dp(shifts), R, k0, n,
                                mc_seed, _dp(mom), _dp(smp))
    if rc:
        raise ValueError("orc_sample_batch failed: %d" % rc)
    return (mom, smp) if want_samples else mom


def num_threads():
    return lib().orc_num_threads()
