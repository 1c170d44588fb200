"""Synthetic Karhunen-Loeve bases for the lognormal-diffusion device models (host-side input preparation).

In the reference workflow the user builds one `GaussianRandomField` per level with GaussianRandomFields.jl
(docs/src/example.md:73-80) and the sample function evaluates `sample(grf, xi = omega)`.  The device models take the
same object in matrix form: Phi_l[node, k] = sqrt(lambda_k) * phi_k(x_node), log a = Phi_l @ xi.  Any basis can be
supplied; this helper provides the smooth synthetic spectrum used by the tests and the benchmark:
sqrt(lambda_k) phi_k(x) = sigma * k^(-decay) * sqrt(2) * cos(pi k x), k = 1..m  (SURVEY.md 8d).
"""
import numpy as np


def kl_basis_1d(level, m=100, n0=16, sigma=0.5, decay=1.5):
    """-> float64[N+1, m] for the level's grid of N = n0 * 2^level cells on [0, 1]"""
    N = n0 << level
    x = np.arange(N + 1, dtype=np.float64) / N
    k = np.arange(1, m + 1, dtype=np.float64)
    return np.ascontiguousarray(sigma * np.sqrt(2.0) * k**(-decay) * np.cos(np.pi * np.outer(x, k)))


def kl_basis_2d(level, m=400, n0=4, sigma=0.5, decay=1.5):
    """-> float64[(Nx+1)(Ny+1), m] for the grid of Nx x Ny cells on [0,1]^2, Nx = n0 * 2^lx, Ny = n0 * 2^ly with
    level = l (lx = ly = l) or level = (lx, ly); node (i, j) is row j*(Nx+1) + i.  Separable cosine modes
    cos(pi p x) cos(pi q y) ordered by p + q (then p), amplitude sigma * c_p c_q * (1 + p + q)^(-decay), c_0 = 1, c_p = sqrt(2)."""
    lx, ly = (level, level) if isinstance(level, (int, np.integer)) else level
    Nx, Ny = n0 << lx, n0 << ly
    x = np.arange(Nx + 1, dtype=np.float64) / Nx
    y = np.arange(Ny + 1, dtype=np.float64) / Ny
    modes, t = [], 0
    while len(modes) < m + 1:  # + the constant mode that is dropped below
        for p in range(t + 1):
            modes.append((p, t - p))
        t += 1
    modes = modes[1:m + 1]  # skip the constant mode (0, 0)
    cols = []
    for p, q in modes:
        cp = 1.0 if p == 0 else np.sqrt(2.0)
        cq = 1.0 if q == 0 else np.sqrt(2.0)
        amp = sigma * cp * cq * (1.0 + p + q) ** (-decay)
        cols.append(amp * np.outer(np.cos(np.pi * q * y), np.cos(np.pi * p * x)).reshape(-1))  # [j][i] -> j*(Nx+1)+i
    return np.ascontiguousarray(np.stack(cols, axis=1))
