/* mle_cuda.h -- C ABI of libmle_cuda, the B200 (sm_100a) sampling engine behind MultilevelEstimators.jl.
 *
 * The library replaces ONE path of the reference: the per-index, per-shift sample batch that
 * `parallel_sample!` drives (src/sample.jl:39-56 MC, :79-102 QMC) together with the statistics that
 * `mean/var/varest` later recompute from the stored samples (src/methods.jl:44-68, :282).  Everything else
 * (index sets, the adaptive loop, History) stays host code.  All `file:line` citations are relative to the
 * reference repository root.
 *
 * Conventions
 *   - every function returns an `int` status (MLE_OK == 0, negative = error); no exception crosses the boundary;
 *     `mle_last_error(ctx)` gives the message.  Argument errors correspond to the reference's `ArgumentError`
 *     (src/check_input.jl:9-25), CUDA / NCCL failures to an `ErrorException`.
 *   - all pointers are caller-owned HOST memory unless the name says `_dev`; the library never returns memory
 *     it allocated.  Handles are opaque and freed by the matching `*_destroy`.
 *   - arrays are dense, first index fastest (Julia / column-major order) and documented as [fastest x ... x slowest].
 *   - one context per process and GPU (or per process and set of GPUs, mle_init_multi); calls on one context are blocking and
 *     not re-entrant
 *     (the reference has a single caller too: `sample!` <- `update_samples` <- `_run`, src/sample.jl:8-34).
 *   - there is NO CPU fallback: every entry point that computes fails with MLE_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef MLE_CUDA_H
#define MLE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLE_ABI_VERSION 1

typedef struct mle_ctx_s *mle_ctx;
typedef struct mle_lattice_s *mle_lattice;
typedef struct mle_dists_s *mle_dists;
typedef struct mle_model_s *mle_model;

enum mle_status {
    MLE_OK = 0,
    MLE_ERR_ARG = -1,    /* invalid argument (ArgumentError in the reference)                              */
    MLE_ERR_CUDA = -2,   /* CUDA runtime / no usable device / kernel failure                               */
    MLE_ERR_RANGE = -3,  /* point number beyond the lattice size n: the reference returns rand(s) there
                            (src/lattice.jl:166,175), which cannot be reproduced -- refused instead          */
    MLE_ERR_STATE = -4,  /* model not fully configured (e.g. KL basis missing for the requested level)      */
    MLE_ERR_NOMEM = -5
};

/* Distribution kinds of src/distribution.jl and their parameters p[0..3]:
 *   MLE_UNIFORM      {a, b}            transform = a + (b - a) x                               (:84)
 *   MLE_NORMAL       {mu, sigma}       transform = mu + sigma * sqrt(2) erfinv(2x - 1)          (:112, :151)
 *   MLE_TRUNCNORMAL  {mu, sigma, a, b} transform = mu + sigma * Phi^-1(Phi(al) + (Phi(be)-Phi(al)) x)   (:145-149)
 *   MLE_WEIBULL      {k, lambda}       transform = lambda * (-log(1 - x))^(1/k)                 (:182)          */
enum mle_dist_kind { MLE_UNIFORM = 0, MLE_NORMAL = 1, MLE_TRUNCNORMAL = 2, MLE_WEIBULL = 3 };
typedef struct mle_dist {
    int32_t kind;
    int32_t reserved;
    double p[4];
} mle_dist;

/* ---- context -------------------------------------------------------------------------------------------------- */
int mle_abi_version(void);
/* Create a context on CUDA device `device` (ordinal as seen by this process).  Replaces nothing in the reference:
 * it is where the `pmap` worker pool of src/sample.jl:43-45 / :89-91 turns into one GPU. */
int mle_init(int device, mle_ctx *ctx);
/* Multi-GPU context on `ndev` distinct devices of this process (devices[0] is member 0).  This is the reference's
 * `nb_of_workers` option (src/parse.jl:224-234, consumed by the pmap calls of src/sample.jl:43-45, :89-91): the worker pool
 * becomes a set of GPUs behind ONE context and one calling thread.  Handles created through the context are replicated on
 * every member.  mle_sample_batch (without `samples`) and mle_update_samples split each request over the members -- whole
 * shifts per member when R >= members in use (each member's moment triplets are final: no merge), else slices of the point
 * range merged on the host in member order (Chan et al.) -- and use only as many members as the request has multiples of
 * MLE_SHARD_MIN (default 24576) samples: small requests stay on member 0.  All members are issued before the first wait.
 * The device-resident entry points (`*_dev`), mle_points and raw-sample output run on member 0. */
int mle_init_multi(const int *devices, int ndev, mle_ctx *ctx);
/* Number of member devices of the context (1 for mle_init). */
int mle_device_count(mle_ctx ctx, int *ndev);
int mle_destroy(mle_ctx ctx);
/* Message of the last failing call on `ctx` (ctx == NULL: last failing mle_init of this thread). */
const char *mle_last_error(mle_ctx ctx);
/* SM count, HBM bytes and compute capability (major*10 + minor) of the context's device. */
int mle_device_info(mle_ctx ctx, int *sm_count, size_t *hbm_bytes, int *cc);
/* The CUDA stream (cudaStream_t) all work of this context is issued on, for callers that time with events. */
int mle_stream(mle_ctx ctx, void **stream);
/* Number of kernels this context has launched so far (bench.py's `gpu_launches`). */
int mle_launch_count(mle_ctx ctx, uint64_t *count);

/* ---- stage 1: rank-1 lattice rule  (LatticeRule32{s}, src/lattice.jl:11-14, :48-50, :88-90) -------------------- */
/* The published generating vectors the reference ships in src/generating_vectors/ ("K_3600_32", "CKN_250_20"),
 * embedded in the library.  `*z` points to static storage. */
int mle_genvec(const char *name, const uint32_t **z, int *len);
/* LatticeRule32(z, s, n): first `s` entries of `z`, at most `nmax` points (reference default typemax(UInt32)).
 * z == NULL selects the default rule LatticeRule32(s) = first s entries of K_3600_32 (src/lattice.jl:88, s <= 3600). */
int mle_lattice_create(mle_ctx ctx, const uint32_t *z, int s, uint64_t nmax, mle_lattice *lat);
int mle_lattice_destroy(mle_ctx ctx, mle_lattice lat);

/* ---- stage 2: distributions  (Vector{<:AbstractDistribution}, src/estimator.jl:116, src/sample.jl:42,88) -------- */
int mle_dists_create(mle_ctx ctx, const mle_dist *d, int m, mle_dists *dists);
int mle_dists_destroy(mle_ctx ctx, mle_dists dists);

/* ---- stage 3: device model registry (replaces the Julia closure `sample_function(index, xi)`,
 *      src/estimator.jl:115, called at src/sample.jl:42,88) ------------------------------------------------------
 *   "expsum"       Q_l = exp(sum_{j<m_l} a_j xi_j), m_l = min(m, m0 2^l); params = {m0, a_1..a_m}; 1 QoI; ndims = 1
 *   "problem18"    the reference's own analytic test function, test/regression.jl:17-46 (ndims = 1) and
 *                  :69-103 (ndims = 2) with the multi-index difference of src/index.jl:49-58; 13 QoI; no params
 *   "lognormal1d"  -(a u')' = 1 on (0,1), u(0)=u(1)=0, log a = KL expansion; level l has n0 2^l cells; params = {n0};
 *                  3-point stencil with arithmetic-mean interface coefficients, u(1/2) from the closed-form flux sums
 *                  (field values must stay in the normal double range, |log a| < 700);
 *                  structure of docs/src/example.md:106-128 (coarse field = even nodes of the fine field, QoI = u(1/2));
 *                  needs mle_model_set_basis for every level used; 1 QoI; ndims = 1
 *   "lognormal2d"  -div(a grad u) = 1 on the unit square, u = 0 on the boundary, QoI = u(1/2,1/2); params = {n0} (default 4);
 *                  ndims = 1 (multilevel): level l has n0 2^l cells per direction, coarse field = even nodes;
 *                  ndims = 2 (multi-index, docs/src/example.md:221-243): index (lx, ly) has n0 2^lx x n0 2^ly cells and
 *                  dQ is the mixed difference over diff(index) (src/index.jl:49-58), every coarse solve a strided view
 *                  of the one fine field; up to 128 cells in x; basis rows = nodes (i,j) -> j*(Nx+1) + i;
 *                  frontal elimination; 1 QoI (field values in the normal double range, like lognormal1d: the kernels' division
 *                  sequence has no Inf / zero / subnormal path, an overflowing field gives NaN where IEEE division would not)
 * Every model returns (dQ, Qf) per sample exactly like the reference's sample functions. */
int mle_model_create(mle_ctx ctx, const char *name, int ndims, const double *params, int nparams, mle_model *model);
/* KL basis of one level: phi[cols x rows] = row-major [rows = nodes][cols = KL terms], log a(node i) = sum_k phi[i][k] xi_k
 * (lognormal1d: rows = n0 2^level + 1 nodes; lognormal2d: rows = (n0 2^level + 1)^2 nodes; lognormal2d with ndims = 2:
 * `level` carries the multi-index packed as lx + 16*ly and rows = (n0 2^lx + 1)(n0 2^ly + 1)) */
int mle_model_set_basis(mle_ctx ctx, mle_model model, int level, int rows, int cols, const double *phi);
int mle_model_nqoi(mle_model model, int *nqoi);
int mle_model_destroy(mle_ctx ctx, mle_model model);

/* ---- parity hook for stages 1-2:  transform.(D[1:m], get_point(gen[r], k)[1:m])  (src/sample.jl:84-88) ----------
 * out[m x n x R]: out[(r*n + i)*m + j] = coordinate j of point k0+i of shift r (k zero-based, src/sample.jl:84).
 *   lat == NULL   -> MC: counter-based uniforms keyed by (mc_seed, k, j) instead of lattice points (R must be 1)
 *                    (the index is NOT part of the key: callers give every index its own mc_seed, as the bindings do)
 *   dists == NULL -> raw points in [0,1) (no transform)
 *   shifts        -> [s x R] (shift r at shifts + r*s, the `shift` field of ShiftedLatticeRule, src/lattice.jl:132-137);
 *                    NULL -> unshifted rule, R must be 1 (src/lattice.jl:142-145)
 * Fails with MLE_ERR_RANGE if k0+n-1 > nmax, MLE_ERR_ARG if m > s (the reference pads with rand(), src/sample.jl:85). */
int mle_points(mle_ctx ctx, mle_lattice lat, mle_dists dists, const double *shifts, int R, uint64_t k0, uint64_t n,
               int m, uint64_t mc_seed, double *out);
/* Same, device-resident: `shifts_dev` and `out_dev` are device pointers; asynchronous on the context's stream. */
int mle_points_dev(mle_ctx ctx, mle_lattice lat, mle_dists dists, const double *shifts_dev, int R, uint64_t k0,
                   uint64_t n, int m, uint64_t mc_seed, double *out_dev);

/* ---- the hot path: one `parallel_sample!` call  (src/sample.jl:39-56, :79-102) ---------------------------------
 * Evaluates (dQ, Qf) = model(index, xi) for every (shift r, point k0..k0+n-1) and reduces them on the device.
 *   index[d]       the level / multi-index (zero-based like the reference's Index)
 *   moments        [3 x 2 x q x R]: moments[((r*q + c)*2 + X)*3 + {0,1,2}] = {n, mean, M2 = sum (x-mean)^2} of the n
 *                  values of X (0: dQ, 1: Qf) for QoI c and shift r -- the sufficient statistics from which
 *                  mean/var/varest of src/methods.jl:44-68,282 follow (var = M2/(n-1), Chan-mergeable across calls)
 *   samples        optional [n x 2 x q x R] (same ordering, n fastest): the raw samples that `append_samples!`
 *                  stores (src/sample.jl:104-111); NULL to skip (the default: nothing is written to HBM per sample)
 *   gpu_seconds    optional: device time of the call's kernels (CUDA events) -- the reference uses the wall time of
 *                  this call as its default cost model (src/sample.jl:27, src/internals.jl:106)
 * lat / dists / shifts / mc_seed as for mle_points. */
int mle_sample_batch(mle_ctx ctx, mle_model model, mle_lattice lat, mle_dists dists, const int32_t *index, int d,
                     const double *shifts, int R, uint64_t k0, uint64_t n, int m, uint64_t mc_seed, double *moments,
                     double *samples, double *gpu_seconds);
/* Same with device-resident shifts / moments / samples, asynchronous on the context's stream (no host copies). */
int mle_sample_batch_dev(mle_ctx ctx, mle_model model, mle_lattice lat, mle_dists dists, const int32_t *index, int d,
                         const double *shifts_dev, int R, uint64_t k0, uint64_t n, int m, uint64_t mc_seed,
                         double *moments_dev, double *samples_dev);
/* One `sample!` call of an UNBIASED estimator (Estimator{<:U}; src/sample.jl:15-19, :65-74, :113-124; accumulator and pmf in
 * src/methods.jl:287-351).  The reference's sample function returns (dQ, Qf) for EVERY idx in zero(index):index from one input
 * vector, `append_samples!` stores them per idx and adds  1/Prob(idx) * dQ_idx  to a per-sample accumulator.  Here:
 *   inv_prob   [prod(index + 1)]: 1 / Prob(idx) for idx in zero(index):index, column-major (first entry of idx fastest)
 *   moments    [prod(index + 1)][3 x 2 x q x R]: one moment table per idx, laid out like mle_sample_batch's
 *   acc        [3 x q x R]: (n, mean, M2) of the accumulator values Z_i = sum_idx inv_prob[idx] * dQ_idx,i per (QoI, shift)
 * Every idx is evaluated on the same points (same shifts, point numbers k0 .. k0+n-1, same leading m coordinates); the terms
 * are multiplied and added in idx order with separately rounded operations, like the reference.  The per-sample values stay
 * on the device (a buffer the library keeps across calls): only the tables come back.  On a multi-GPU context the call runs
 * on member 0. */
int mle_sample_unbiased(mle_ctx ctx, mle_model model, mle_lattice lat, mle_dists dists, const int32_t *index, int d,
                        const double *shifts, int R, uint64_t k0, uint64_t n, int m, uint64_t mc_seed, const double *inv_prob,
                        double *moments, double *acc, double *gpu_seconds);
/* One `update_samples` call (src/sample.jl:8-13): `count` independent `sample!` requests -- entry e has its own index
 * (indices + e*d), shifts (shifts[e], [s x R[e]], or NULL entries for MC), point range k0[e] .. k0[e]+n[e]-1, dimension count
 * m[e] and moment table moments[e] ([3 x 2 x q x R[e]]) -- issued back to back on the context's stream with ONE
 * synchronisation at the end instead of one per index.  Results are identical to `count` calls of mle_sample_batch.
 * All entries share `mc_seed`: meant for QMC entries (for MC, entries of different indices would see the same uniforms). */
int mle_update_samples(mle_ctx ctx, mle_model model, mle_lattice lat, mle_dists dists, int count, const int32_t *indices,
                       int d, const double *const *shifts, const int *R, const uint64_t *k0, const uint64_t *n, const int *m,
                       uint64_t mc_seed, double *const *moments);
/* Block until everything issued on the context's stream has finished. */
int mle_sync(mle_ctx ctx);

/* Merge (n, mean, M2) triplets: acc <- acc (+) part, `count` triplets (Chan et al.); pure host arithmetic used by the
 * host layers to accumulate calls and ranks in a fixed order. */
int mle_moments_merge(double *acc, const double *part, size_t count);

#ifdef __cplusplus
}
#endif
#endif /* MLE_CUDA_H */
