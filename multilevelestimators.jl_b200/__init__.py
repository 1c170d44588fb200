"""multilevelestimators.jl_b200 -- B200-native sampling engine for MultilevelEstimators.jl.

Only the hot path lives here: csrc/ (CUDA kernels + the C ABI of include/mle_cuda.h), a ctypes binding (_lib),
a thin object layer (engine) and the host-side mirror of the reference interface for that path (host).
The directory name contains a dot, so import it through the `mle_b200` alias module at the repository root.
"""
from . import _lib  # noqa: F401
from .engine import Engine, MLEArgumentError, MLEError, moments_merge  # noqa: F401
from ._lib import NORMAL, TRUNCNORMAL, UNIFORM, WEIBULL  # noqa: F401
