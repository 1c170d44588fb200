"""ctypes binding of libmle_cuda.so -- the Python twin of the Julia `ccall` glue in julia/MLECuda.jl.

Every function declared in include/mle_cuda.h is bound here with the same name and argument order.
The library is loaded from multilevelestimators.jl_b200/lib/ (built in-tree by csrc/Makefile); importing this module
fails loudly if it is missing -- there is no CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmle_cuda.so")
# tuning builds (tools/build_variant.sh) are selected with MLE_CUDA_LIB=<path to an alternative libmle_cuda.so>
LIB_PATH = os.environ.get("MLE_CUDA_LIB", LIB_PATH)

MLE_OK, MLE_ERR_ARG, MLE_ERR_CUDA, MLE_ERR_RANGE, MLE_ERR_STATE, MLE_ERR_NOMEM = 0, -1, -2, -3, -4, -5
UNIFORM, NORMAL, TRUNCNORMAL, WEIBULL = 0, 1, 2, 3


class mle_dist(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("p", C.c_double * 4)]


# name -> (restype, argtypes); the list the "every symbol of include/mle_cuda.h is exported" test walks
_vp, _dp, _u32p, _i32p = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
SIGNATURES = {
    "mle_abi_version": (C.c_int, []),
    "mle_init": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "mle_destroy": (C.c_int, [_vp]),
    "mle_last_error": (C.c_char_p, [_vp]),
    "mle_device_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    "mle_stream": (C.c_int, [_vp, C.POINTER(_vp)]),
    "mle_init_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_vp)]),
    "mle_device_count": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "mle_launch_count": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "mle_genvec": (C.c_int, [C.c_char_p, C.POINTER(_u32p), C.POINTER(C.c_int)]),
    "mle_lattice_create": (C.c_int, [_vp, _u32p, C.c_int, C.c_uint64, C.POINTER(_vp)]),
    "mle_lattice_destroy": (C.c_int, [_vp, _vp]),
    "mle_dists_create": (C.c_int, [_vp, C.POINTER(mle_dist), C.c_int, C.POINTER(_vp)]),
    "mle_dists_destroy": (C.c_int, [_vp, _vp]),
    "mle_model_create": (C.c_int, [_vp, C.c_char_p, C.c_int, _dp, C.c_int, C.POINTER(_vp)]),
    "mle_model_set_basis": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _dp]),
    "mle_model_nqoi": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "mle_model_destroy": (C.c_int, [_vp, _vp]),
    "mle_points": (C.c_int, [_vp, _vp, _vp, _dp, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, _dp]),
    "mle_points_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, _vp]),
    "mle_sample_batch": (C.c_int, [_vp, _vp, _vp, _vp, _i32p, C.c_int, _dp, C.c_int, C.c_uint64, C.c_uint64, C.c_int,
                                   C.c_uint64, _dp, _dp, _dp]),
    "mle_sample_batch_dev": (C.c_int, [_vp, _vp, _vp, _vp, _i32p, C.c_int, _vp, C.c_int, C.c_uint64, C.c_uint64, C.c_int,
                                       C.c_uint64, _vp, _vp]),
    "mle_update_samples": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _i32p, C.c_int, C.POINTER(_dp), C.POINTER(C.c_int),
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.c_uint64, C.POINTER(_dp)]),
    "mle_sample_unbiased": (C.c_int, [_vp, _vp, _vp, _vp, _i32p, C.c_int, _dp, C.c_int, C.c_uint64, C.c_uint64, C.c_int,
                                      C.c_uint64, _dp, _dp, _dp, _dp]),
    "mle_sync": (C.c_int, [_vp]),
    "mle_moments_merge": (C.c_int, [_dp, _dp, C.c_size_t]),
}

_lib = None


def load():
    """dlopen libmle_cuda.so and attach the prototypes.  Raises if the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libmle_cuda.so not found at %s -- build it with `make -C multilevelestimators.jl_b200/csrc` "
                "(or python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.mle_abi_version() != 1:
            raise ImportError("libmle_cuda.so ABI version mismatch")
        # second prototype of the hot call with every pointer as a plain address: the per-call marshalling of ndarray.ctypes /
        # POINTER objects costs more than the launch itself for the tiny batches of the adaptive loop (engine.Engine.sample_batch)
        proto = C.CFUNCTYPE(C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64,
                            _vp, _vp, _vp)
        lib.mle_sample_batch_addr = proto(("mle_sample_batch", lib))
        lib.mle_moments_merge_addr = C.CFUNCTYPE(C.c_int, _vp, _vp, C.c_size_t)(("mle_moments_merge", lib))
        _lib = lib
    return _lib
