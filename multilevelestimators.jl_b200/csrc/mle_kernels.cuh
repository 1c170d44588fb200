// mle_kernels.cuh -- the CUDA kernels of libmle_cuda (sm_100a only).
//
//   points_kernel        K1/K2: lattice (or counter-RNG) point -> shift -> inverse CDF -> dense [m x n x R] in HBM,
//                        tile staged in shared memory and written with one TMA bulk store per tile
//   sample_simple_kernel K5+K6: fused generate -> analytic sample function -> per-CTA moments (no per-sample HBM traffic)
//   lognormal2d_kernel   2-D lognormal diffusion, shared-memory-window solver (fallback of mle_kernels2d.cuh)
// (the headline kernel, lognormal1d_kernel, lives in mle_kernel_ln1d.cuh; the per-CTA partial statistics of every kernel are
//  reduced by the last CTA of each shift, cta_finalize in mle_device.cuh -- there is no separate finalize launch)
#pragma once
#include "mle_device.cuh"
#include "mle_kernel_ln1d.cuh"
#include "mle_kernel_ln1d_block.cuh"
#include "mle_kernel_ln1d_small.cuh"

namespace mle {


// ---- TMA (bulk async copy) helpers --------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// ===================================================================================================================
// K1 / K2: points -> HBM
//   out[(r*n + i)*m + j].  One tile = P consecutive points of one shift = P*m consecutive doubles of `out`.
//   Persistent CTAs loop over tiles; the tile leaves through one TMA bulk store (five resident CTAs overlap each other's stores).
//   Algorithmic traffic: 8 B written per coordinate (z and the shifts stay in L1/L2).
// ===================================================================================================================
#ifndef MLE_POINTS_THREADS
#define MLE_POINTS_THREADS 256
#endif
constexpr int POINTS_THREADS = MLE_POINTS_THREADS;
#ifndef MLE_POINTS_TILE_CAP
#define MLE_POINTS_TILE_CAP 4096
#endif
constexpr int POINTS_TILE_CAP = MLE_POINTS_TILE_CAP;  // elements per tile buffer (32 KB)

template <int THREADS, int MODE>
// (plain __launch_bounds__(THREADS): with an explicit minimum of 1 block ptxas spends 56 instead of 48 registers, which costs the
// fifth resident CTA and 17 % of the throughput)
__global__ void __launch_bounds__(THREADS) points_kernel(GenArgs g, double *__restrict__ out, int P,
                                                         long long tiles_per_shift, long long ntiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *buf0 = reinterpret_cast<double *>(smem_raw);
#ifdef MLE_POINTS_DOUBLE_BUF  // A/B switch: two tile buffers (73.7 KB per CTA, 3 resident CTAs), the store of tile t overlapping tile t+1
    double *buf1 = buf0 + POINTS_TILE_CAP;
#else
    // ONE tile buffer: 41 KB per CTA -> 5 resident CTAs instead of 3.  The other CTAs of the SM cover the wait for the bulk store
    // better than a second buffer did: raw points 5.28 -> 5.80 TB/s (0.81 -> 0.89 of the measured copy peak), with the Normal
    // inverse CDF 1.81 -> 1.88 TB/s
    double *buf1 = buf0;
#endif
    GenScratch sc;
    sc.queue = reinterpret_cast<unsigned short *>(buf1 + POINTS_TILE_CAP);
    sc.wtot = reinterpret_cast<unsigned *>(sc.queue + POINTS_TILE_CAP);

    const int m = g.m;
    int it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int r = (int)(tile / tiles_per_shift);
        const long long ti = tile - (long long)r * tiles_per_shift;
        const unsigned long long i0 = (unsigned long long)ti * (unsigned long long)P;
        const unsigned long long left = g.n - i0;
        const int Pt = left < (unsigned long long)P ? (int)left : P;
        double *buf = (it & 1) ? buf1 : buf0;
        // the previous bulk store of this CTA read this buffer: thread 0 waits until it has been read, and nobody writes the
        // buffer before that wait has returned (the barrier below)
#ifdef MLE_POINTS_DOUBLE_BUF
        if (threadIdx.x == 0) bulk_wait_read<1>();
#else
        if (threadIdx.x == 0) bulk_wait_read<0>();
#endif
        __syncthreads();
        gen_tile_dense<THREADS, MODE>(buf, Pt, m, g.k0 + i0, g, r, sc);
        double *dst = out + ((size_t)r * g.n + i0) * (size_t)m;
        const uint32_t bytes = (uint32_t)Pt * (uint32_t)m * 8u;
        const bool tma_ok = ((reinterpret_cast<uintptr_t>(dst) | bytes) & 15u) == 0;
        if (tma_ok) {
            fence_proxy_async_smem();  // make this thread's generic-proxy writes visible to the async proxy
            __syncthreads();
            if (threadIdx.x == 0) { bulk_store_s2g(dst, buf, bytes); bulk_commit(); }
        } else {  // odd m with an odd starting element: plain coalesced stores
            for (int e = threadIdx.x; e < Pt * m; e += THREADS) dst[e] = buf[e];
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) bulk_wait_all<0>();
}

// ===================================================================================================================
// analytic sample functions (one sample per thread), fused with generation and moments
// ===================================================================================================================
enum : int { MODEL_EXPSUM = 0, MODEL_PROBLEM18 = 1, MODEL_LOGNORMAL1D = 2, MODEL_LOGNORMAL2D = 3 };

struct SimpleModelArgs {
    const double *weights;  // expsum: a_j
    int nweights;
    int ml;    // expsum: active dims at this level
    int mlc;   // expsum: active dims at level-1 (0: no coarse term)
    int ndims; // problem18
    int idx0, idx1;
};

constexpr int SIMPLE_THREADS = 128;       // = points per tile: every thread owns one point
#ifndef MLE_SIMPLE_MC
#define MLE_SIMPLE_MC 32
#endif
constexpr int SIMPLE_MC = MLE_SIMPLE_MC;  // dims per chunk; tile[jj][point] (point fastest: conflict-free); 32 dims =
                                          // one 32-bit class mask per thread and 40 KB per CTA (5 CTAs per SM)

// problem_18                                                      test/regression.jl:17-36 (1-d), :69-103 (2-d)
__device__ __forceinline__ double p18_coef(int l, double xj) {
    if (l == 3) return 1.0;
    if (l == 2) return 1.1;
    if (l == 1) return __dadd_rn(__ddiv_rn(xj, 60.0), 1.2);
    return 1.5;
}
__device__ __forceinline__ void problem18_eval(int ndims, int i0, int i1, double xi0, double xi1, double *f) {
    double x13 = __dmul_rn(__dmul_rn(xi0, xi0), xi0);
    double x23 = (ndims == 2) ? __dmul_rn(__dmul_rn(xi1, xi1), xi1) : 0.0;
#pragma unroll
    for (int j = 0; j < 13; ++j) {
        const double xj = 0.5 * j;  // range(0, 6, 13)[j]
        // f_j = x_j <= 3 ? (x_j - 2)^2 : 2*log(x_j - 2) + 1: thirteen constants (correctly rounded logs, generated table)
        constexpr double base[13] = MLE_P18_BASE_INIT;
        const double fj = base[j];
        double acc = __dadd_rn(fj, __dmul_rn(p18_coef(i0, xj), x13));
        if (ndims == 2) {
            acc = __dadd_rn(acc, __dmul_rn(p18_coef(i1, xj), x23));
            if (i0 > 0 && i1 > 0) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(0.03125, x13), x23));
        }
        f[j] = acc;
    }
}

// Batched corrected two-pass statistics of V values per thread: provisional tile mean, then sums of d and d^2 about it.  The
// tile's (cnt, mean0, sum d, sum d^2) IS a pivoted partial (mle_device.cuh) with pivot mean0; thread v merges it into acc[v].
template <int THREADS, int V>
__device__ __forceinline__ void block_moments_merge(const double (&x)[V], int cnt, double *red /* [3][V][W] */,
                                                    Partial *acc) {
    constexpr int W = THREADS / 32;
    const bool valid = (int)threadIdx.x < cnt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *r0 = red, *r1 = red + V * W, *r2 = red + 2 * V * W;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        double s = warp_sum(valid ? x[v] : 0.0);
        if (lane == 0) r0[v * W + warp] = s;
    }
    __syncthreads();
    double d[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < W; ++w) t += r0[v * W + w];
        double mean0 = t / (double)cnt;
        d[v] = valid ? x[v] - mean0 : 0.0;
        double sd = warp_sum(d[v]), sdd = warp_sum(d[v] * d[v]);
        if (lane == 0) { r1[v * W + warp] = sd; r2[v * W + warp] = sdd; }
    }
    __syncthreads();
    if (threadIdx.x < V) {
        const int v = threadIdx.x;
        double t = 0.0, sd = 0.0, sdd = 0.0;
        for (int w = 0; w < W; ++w) { t += r0[v * W + w]; sd += r1[v * W + w]; sdd += r2[v * W + w]; }
        Partial o;
        o.n = (double)cnt;
        o.c = t / (double)cnt;  // the same operations as mean0 above: the same value
        o.s1 = sd;
        o.s2 = sdd;
        acc[v] = partial_merge(acc[v], o);
    }
    __syncthreads();
}

// grid = (G, R).  CTA (g, r) handles point tiles g, g+G, ... of shift r, SIMPLE_THREADS points per tile, dims in chunks.
// fa.partials[(r*V + v)*G + g] = pivoted partial of this CTA; the last CTA of a shift writes its moments (cta_finalize).
// samples (optional): [(r*V + v)*n + i].
template <int MODEL, int V, int MODE>
__global__ void __launch_bounds__(SIMPLE_THREADS) sample_simple_kernel(GenArgs g, SimpleModelArgs ma, FinArgs fa,
                                                                       double *__restrict__ samples,
                                                                       long long tiles_per_shift) {
    constexpr int THREADS = SIMPLE_THREADS;
    constexpr int W = THREADS / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *tile = reinterpret_cast<double *>(smem_raw);                    // [SIMPLE_MC][THREADS]
    double *red = tile + THREADS * SIMPLE_MC;                               // [3][V][W]
    Partial *acc = reinterpret_cast<Partial *>(red + 3 * V * W);            // [V]
    GenScratch sc;
    sc.queue = reinterpret_cast<unsigned short *>(acc + V);                 // [THREADS*SIMPLE_MC]
    sc.wtot = reinterpret_cast<unsigned *>(sc.queue + THREADS * SIMPLE_MC);
    // {z_j, shift_j} of the current and the next chunk of dims (lattice modes): staged while the previous chunk is consumed
    double2 *zsbuf = reinterpret_cast<double2 *>(sc.wtot + ((W + 3) & ~3));  // [2][SIMPLE_MC]
    constexpr bool lattice = (MODE & GEN_MC) == 0;

    const int r = blockIdx.y;
    const int tid = threadIdx.x;
    const double *shift_r = lattice ? g.shifts + (size_t)r * g.s : nullptr;
    auto stage = [&](int buf, int j0) {
        if (lattice && tid < SIMPLE_MC && j0 + tid < g.m)
            zsbuf[buf * SIMPLE_MC + tid] = make_double2(__ldg(g.zd + j0 + tid), __ldg(shift_r + j0 + tid));
    };
    int cb = 0;
    if (tid < V) acc[tid] = Partial{0.0, 0.0, 0.0, 0.0};
    stage(0, 0);
    __syncthreads();

    for (long long ti = blockIdx.x; ti < tiles_per_shift; ti += gridDim.x) {
        const unsigned long long i0 = (unsigned long long)ti * THREADS;
        const unsigned long long left = g.n - i0;
        const int Pt = left < (unsigned long long)THREADS ? (int)left : THREADS;
        double val[V];
        double s_acc = 0.0, s_coarse = 0.0;
        double xi0 = 0.0, xi1 = 0.0;
        for (int j0 = 0; j0 < g.m; j0 += SIMPLE_MC) {
            const int mc = min(SIMPLE_MC, g.m - j0);
            // always a full tile of points: surplus points (tid >= Pt) are masked out of the statistics
            gen_tile_owned<THREADS, THREADS, MODE, THREADS, unsigned, 4, lattice>(tile, j0, mc, g.k0 + i0, g, r, sc,
                                                                                  zsbuf + cb * SIMPLE_MC - j0);
            stage(cb ^ 1, j0 + SIMPLE_MC < g.m ? j0 + SIMPLE_MC : 0);  // next chunk (or chunk 0 of the next tile)
            cb ^= 1;
            if (MODEL == MODEL_EXPSUM) {
                for (int jj = 0; jj < mc; ++jj) {
                    const int j = j0 + jj;
                    if (j < ma.ml) s_acc = __fma_rn(__ldg(ma.weights + j), tile[jj * THREADS + tid], s_acc);
                    if (j == ma.mlc - 1) s_coarse = s_acc;
                }
            } else {
                if (j0 == 0) { xi0 = tile[tid]; if (mc > 1) xi1 = tile[THREADS + tid]; }
            }
            __syncthreads();  // the tile is overwritten by the next chunk
        }
        if (MODEL == MODEL_EXPSUM) {
            double qf = exp_det(s_acc);
            val[1] = qf;
            val[0] = ma.mlc > 0 ? __dsub_rn(qf, exp_det(s_coarse)) : qf;
        } else {
            // (dQ, Qf): dQ = Qf + sum over diff(index) of +-Q_kappa         src/index.jl:49-58, test/regression.jl:39-46
            double f[13], dq[13];
            problem18_eval(ma.ndims, ma.idx0, ma.idx1, xi0, xi1, f);
#pragma unroll
            for (int j = 0; j < 13; ++j) dq[j] = f[j];
            const int lo0 = ma.idx0 > 0 ? ma.idx0 - 1 : 0;
            const int hi1 = ma.ndims == 2 ? ma.idx1 : 0;
            const int lo1 = (ma.ndims == 2 && ma.idx1 > 0) ? ma.idx1 - 1 : 0;
            for (int b = lo1; b <= hi1; ++b)
                for (int a = lo0; a <= ma.idx0; ++a) {
                    if (a == ma.idx0 && b == hi1) continue;
                    const double sgn = (((ma.idx0 - a) + (hi1 - b)) & 1) ? -1.0 : 1.0;
                    double fc[13];
                    problem18_eval(ma.ndims, a, b, xi0, xi1, fc);
#pragma unroll
                    for (int j = 0; j < 13; ++j) dq[j] = __dadd_rn(dq[j], __dmul_rn(sgn, fc[j]));
                }
#pragma unroll
            for (int j = 0; j < 13; ++j) {
                if (2 * j + 1 < V) { val[2 * j] = dq[j]; val[2 * j + 1] = f[j]; }
            }
        }
        if (samples && tid < Pt) {
#pragma unroll
            for (int v = 0; v < V; ++v) samples[((size_t)r * V + v) * g.n + i0 + tid] = val[v];
        }
        block_moments_merge<THREADS, V>(val, Pt, red, acc);
    }
    if (tid < V) fa.partials[((size_t)r * V + tid) * gridDim.x + blockIdx.x] = acc[tid];
    cta_finalize(fa, V);
}

// ===================================================================================================================
// 2-D lognormal diffusion (config C3/C4 shape; model definition in oracle/mle_oracle.c and DESIGN.md):
//   -div(a grad u) = 1 on the unit square, u = 0 on the boundary, Nx x Ny = n0 2^lx x n0 2^ly cells (multilevel: lx = ly = l;
//   multi-index: anisotropic, mixed differences over diff(index) with strided coarse fields, docs/src/example.md:221-243),
//   5-point stencil with edge coefficients = mean of the two node values scaled by 1/h^2, QoI = u(1/2, 1/2).
// One GROUP of TG threads owns one sample (TG = 32: a warp, 4 samples per CTA, N <= 32;  TG = 256: the whole CTA, N >= 64):
//   1. xi (m inverse normals) -> shared memory;
//   2. field: thread-per-node dot products over the KL terms in ascending k (fused multiply-adds, the oracle's order) from the
//      TRANSPOSED basis [k][node] (coalesced), then exp;
//   3. frontal Gaussian elimination towards the centre node: two half sweeps (the second on the y-mirrored field) carrying a
//      sliding (n+1)x(n+1) window in shared memory -- per pivot one phase that loads the entering unknown and forms
//      v_q = M[q][p], l_q = v_q / d, and one phase of rank-1 updates M[q][r] = fma(-l_q, v_r, M[q][r]) -- followed by a dense
//      two-sided elimination along the centre line.  There is no reduction anywhere, so the thread mapping cannot change a
//      single bit of the result.
// ===================================================================================================================
struct Lognormal2dArgs {
    const double *basisT;  // [m][nodes_pad]: transposed KL basis of this index
    int Nx, Ny;            // cells per direction (Nx = n0 2^lx, Ny = n0 2^ly; multilevel model: lx = ly)
    int nodes_pad;
    int ndims;             // 1: multilevel (one coarse solve), 2: multi-index (mixed differences, src/index.jl:49-58)
    int lx, ly;
    int use_scratch;       // field / centre blocks in global scratch (Nx >= 64) instead of shared memory
    double *scratch;       // [groups in grid][scratch_per_group]
    long long scratch_per_group;
};

template <int TG>
__device__ __forceinline__ void group_sync() {
    if (TG == 32) __syncwarp();
    else __syncthreads();
}

// node values a[(j*sty)*lda + i*stx] of an Nx x Ny-cell grid (optionally y-mirrored) and its scaled edge coefficients
struct Field2d {
    const double *a;
    int lda, stx, sty, Nx, Ny, mirror;
    double wx, wy;  // 1/hx^2 = Nx^2, 1/hy^2 = Ny^2
    __device__ __forceinline__ double at(int i, int j) const {
        const int jj = mirror ? Ny - j : j;
        return a[(size_t)(jj * sty) * lda + (size_t)i * stx];
    }
    __device__ __forceinline__ double kx(int i, int j) const { return __dmul_rn(wx, __dmul_rn(0.5, __dadd_rn(at(i, j), at(i + 1, j)))); }
    __device__ __forceinline__ double ky(int i, int j) const { return __dmul_rn(wy, __dmul_rn(0.5, __dadd_rn(at(i, j), at(i, j + 1)))); }
    __device__ __forceinline__ double diag(int i, int j) const {
        return __dadd_rn(__dadd_rn(__dadd_rn(kx(i - 1, j), kx(i, j)), ky(i, j - 1)), ky(i, j));
    }
};

// rank-1 update of the cnt x cnt block of M whose rows/cols are the window slots of unknowns first .. first+cnt-1
// (ld = row stride of M, mod = slot modulus: W for the sliding window, anything > first+cnt for a plain matrix)
template <int TG>
__device__ __forceinline__ void rank1_update(double *M, double *gv, const double *l, const double *v, double gp, int ld,
                                             int mod, int first, int cnt, int t) {
    if (TG == 32) {
        if (t < cnt) {
            const int sr = (first + t) % mod;
            const double vr = v[sr];
            for (int qo = 0; qo < cnt; ++qo) {
                const int sq = (first + qo) % mod;
                M[sq * ld + sr] = __fma_rn(-l[sq], vr, M[sq * ld + sr]);
            }
            gv[sr] = __fma_rn(-l[sr], gp, gv[sr]);
        }
    } else {
        const int tx = t & 31, ty = t >> 5;
        for (int qo = ty; qo < cnt; qo += TG / 32) {
            const int sq = (first + qo) % mod;
            const double lq = -l[sq];
            for (int ro = tx; ro < cnt; ro += 32) {
                const int sr = (first + ro) % mod;
                M[(size_t)sq * ld + sr] = __fma_rn(lq, v[sr], M[(size_t)sq * ld + sr]);
            }
        }
        for (int qo = t; qo < cnt; qo += TG) { const int sq = (first + qo) % mod; gv[sq] = __fma_rn(-l[sq], gp, gv[sq]); }
    }
}

template <int TG>
__device__ void frontal_half_dev(const Field2d &f, double *M, double *gv, double *l, double *v, double *S, double *gs, int t) {
    const int n = f.Nx - 1, c = f.Ny / 2, W = n + 1;
    const int P = (c - 1) * n, last = P + n - 1;
    // entering unknown e: row/column slot(e) of the window <- its original matrix entries (one phase, no read of M)
    auto load = [&](int e) {
        const int ie = e % n + 1, je = e / n + 1, se = e % W;
        const bool own = !(je == c && f.mirror);
        for (int u = t; u < W; u += TG) {
            double val = 0.0;
            if (u == se) val = own ? f.diag(ie, je) : 0.0;
            else if (ie > 1 && own && u == (e - 1) % W) val = -f.kx(ie - 1, je);
            else if (je > 1 && u == (e - n) % W) val = -f.ky(ie, je - 1);
            M[se * W + u] = val;
            M[u * W + se] = val;
        }
        if (t == 0) gv[se] = own ? 1.0 : 0.0;
    };
    for (int e = 0; e < n && e <= last; ++e) load(e);
    group_sync<TG>();
    for (int p = 0; p < P; ++p) {
        const int sp = p % W;
        const bool enter = p + n <= last;
        const int hi = enter ? p + n : last;
        const double rd = __ddiv_rn(1.0, M[sp * W + sp]), gp = gv[sp];  // multipliers l = v * (1/d)
        // phase A: v_q = M[q][p], l_q = v_q / d for the window p+1..hi (the entering unknown's coupling to p is known
        // directly), and the entering unknown's row/column (the slot of pivot p-1; disjoint from everything read here)
        for (int qo = t; qo < hi - p; qo += TG) {
            const int q = p + 1 + qo, sq = q % W;
            double vq;
            if (enter && q == p + n) { const int ie = q % n + 1, je = q / n + 1; vq = -f.ky(ie, je - 1); }
            else vq = M[sq * W + sp];
            v[sq] = vq;
            l[sq] = __dmul_rn(vq, rd);
        }
        if (enter) load(p + n);
        group_sync<TG>();
        // phase B: rank-1 update of the window
        rank1_update<TG>(M, gv, l, v, gp, W, W, p + 1, hi - p, t);
        group_sync<TG>();
    }
    for (int e = t; e < n * n; e += TG) { const int q = e / n, r = e - q * n; S[e] = M[((P + q) % W) * W + (P + r) % W]; }
    for (int q = t; q < n; q += TG) gs[q] = gv[(P + q) % W];
    group_sync<TG>();
}

// u at the centre node (Nx/2, Ny/2)
template <int TG>
__device__ double frontal_mid_dev(const double *a, int lda, int stx, int sty, int Nx, int Ny, double *M, double *gv,
                                  double *l, double *v, double *S, double *gs, double *S2, double *gs2, int t) {
    const int n = Nx - 1, tc = Nx / 2 - 1;
    Field2d f{a, lda, stx, sty, Nx, Ny, 1, (double)Nx * (double)Nx, (double)Ny * (double)Ny};
    frontal_half_dev<TG>(f, M, gv, l, v, S2, gs2, t);  // upper half first (y-mirrored field)
    f.mirror = 0;
    frontal_half_dev<TG>(f, M, gv, l, v, S, gs, t);    // lower half
    for (int e = t; e < n * n; e += TG) S[e] = __dadd_rn(S[e], S2[e]);
    for (int q = t; q < n; q += TG) gs[q] = __dadd_rn(gs[q], gs2[q]);
    group_sync<TG>();
    // dense two-sided elimination along the centre line (row-major S)
    for (int tt = 0; tt < tc; ++tt) {
        const double rd = __ddiv_rn(1.0, S[tt * n + tt]), gp = gs[tt];
        for (int q = tt + 1 + t; q < n; q += TG) { v[q] = S[q * n + tt]; l[q] = __dmul_rn(v[q], rd); }
        group_sync<TG>();
        rank1_update<TG>(S, gs, l, v, gp, n, 1 << 30, tt + 1, n - tt - 1, t);
        group_sync<TG>();
    }
    for (int tt = n - 1; tt > tc; --tt) {
        const double rd = __ddiv_rn(1.0, S[tt * n + tt]), gp = gs[tt];
        for (int q = tc + t; q < tt; q += TG) { v[q] = S[q * n + tt]; l[q] = __dmul_rn(v[q], rd); }
        group_sync<TG>();
        rank1_update<TG>(S, gs, l, v, gp, n, 1 << 30, tc, tt - tc, t);
        group_sync<TG>();
    }
    const double u = __ddiv_rn(gs[tc], S[tc * n + tc]);
    group_sync<TG>();  // the buffers are reused by the next solve
    return u;
}

template <int MODE, int TG>
__global__ void __launch_bounds__(TG == 32 ? 128 : 256) lognormal2d_kernel(GenArgs g, Lognormal2dArgs la, FinArgs fa,
                                                                           double *__restrict__ samples) {
    constexpr int THREADS = TG == 32 ? 128 : 256, GROUPS = THREADS / TG;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int Nx = la.Nx, Ny = la.Ny, n = Nx - 1, W = n + 1, nodes = (Nx + 1) * (Ny + 1);
    const int t = threadIdx.x % TG, grp = threadIdx.x / TG, r = blockIdx.y;
    // per-group shared memory: xi[mpad] | M[W*W] | gv[W] | l[W] | v[W] | gs[W] gs2[W] and, unless they live in the global
    // scratch (Nx >= 64): S[n*n] | S2[n*n] | field[(Nx+1)(Ny+1)]
    const int mpad = (g.m + 3) & ~3;
    const size_t per_group = (size_t)mpad + (size_t)W * W + 5 * (size_t)W + (la.use_scratch ? 0 : 2 * (size_t)n * n + (size_t)nodes);
    double *base = reinterpret_cast<double *>(smem_raw) + (size_t)grp * per_group;
    double *xi = base, *M = xi + mpad, *gv = M + (size_t)W * W, *l = gv + W, *v = l + W, *gs = v + W, *gs2 = gs + W;
    double *S, *S2, *A;
    if (la.use_scratch) {
        S = la.scratch + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * GROUPS + grp) * la.scratch_per_group;
    } else {
        S = gs2 + W;
    }
    S2 = S + (size_t)n * n;
    A = S2 + (size_t)n * n;
    Partial *mred = reinterpret_cast<Partial *>(reinterpret_cast<double *>(smem_raw) + (size_t)GROUPS * per_group);  // [2][GROUPS]
    constexpr bool lattice = (MODE & GEN_MC) == 0;
    const double *shift_r = lattice ? g.shifts + (size_t)r * g.s : nullptr;
    Partial pdq{0.0, 0.0, 0.0, 0.0}, pqf{0.0, 0.0, 0.0, 0.0};

    // GROUPS samples per CTA iteration; every group runs the same number of iterations (barriers are CTA-wide for TG = 256)
    for (unsigned long long i = (unsigned long long)blockIdx.x * GROUPS + grp;; i += (unsigned long long)gridDim.x * GROUPS) {
        const bool active = i < g.n;
        if (TG != 32) { if (!active) break; }         // one group per CTA: uniform exit
        else if (!__any_sync(0xffffffffu, active)) break;
        if (active) {
            const unsigned long long k = g.k0 + i;
            // 1. inputs
            const double tphi = phi_to_unit(radical_phi((uint32_t)k));
            for (int j = t; j < g.m; j += TG) {
                const double x = lattice ? lattice_coord(tphi, __ldg(g.zd + j), __ldg(shift_r + j))
                                         : mc_uniform(g.seed, k, (uint32_t)g.mkey, (uint32_t)j);
                int cls;
                double val = elem_first<MODE>(x, j, g, cls);
                if (cls == 2) val = elem_finish<MODE, 2>(val, j, g);
                else if (cls == 3) val = elem_finish<MODE, 3>(val, j, g);
                xi[j] = val;
            }
            group_sync<TG>();
            // 2. field
            for (int node = t; node < nodes; node += TG) {
                const double *col = la.basisT + node;
                double acc = 0.0;
                for (int kk = 0; kk < g.m; ++kk) acc = __fma_rn(__ldg(col + (size_t)kk * la.nodes_pad), xi[kk], acc);
                A[node] = exp_det(acc);
            }
            group_sync<TG>();
            // 3. solves: Qf and the coarse terms of diff(index), added in the order (lx-1,ly-1), (lx,ly-1), (lx-1,ly)
            const double qf = frontal_mid_dev<TG>(A, Nx + 1, 1, 1, Nx, Ny, M, gv, l, v, S, gs, S2, gs2, t);
            double dq = qf;
            if (la.ndims == 1) {
                if (la.lx > 0) dq = __dsub_rn(qf, frontal_mid_dev<TG>(A, Nx + 1, 2, 2, Nx / 2, Ny / 2, M, gv, l, v, S, gs, S2, gs2, t));
            } else {
                if (la.lx > 0 && la.ly > 0) dq = __dadd_rn(dq, frontal_mid_dev<TG>(A, Nx + 1, 2, 2, Nx / 2, Ny / 2, M, gv, l, v, S, gs, S2, gs2, t));
                if (la.ly > 0) dq = __dsub_rn(dq, frontal_mid_dev<TG>(A, Nx + 1, 1, 2, Nx, Ny / 2, M, gv, l, v, S, gs, S2, gs2, t));
                if (la.lx > 0) dq = __dsub_rn(dq, frontal_mid_dev<TG>(A, Nx + 1, 2, 1, Nx / 2, Ny, M, gv, l, v, S, gs, S2, gs2, t));
            }
            if (t == 0) {
                if (samples) {
                    samples[((size_t)r * 2 + 0) * g.n + i] = dq;
                    samples[((size_t)r * 2 + 1) * g.n + i] = qf;
                }
                partial_add(pdq, dq);
                partial_add(pqf, qf);
            }
            group_sync<TG>();
        }
    }
    __syncthreads();
    if (t == 0) {
        mred[grp] = pdq;
        mred[GROUPS + grp] = pqf;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        Partial a{0.0, 0.0, 0.0, 0.0};
        for (int q = 0; q < GROUPS; ++q) a = partial_merge(a, mred[threadIdx.x * GROUPS + q]);
        fa.partials[((size_t)r * 2 + threadIdx.x) * gridDim.x + blockIdx.x] = a;
    }
    cta_finalize(fa, 2);
}


// ---------------------------------------------------------------------------------------------------------------
// Unbiased estimators (src/sample.jl:65-74, :113-124): per-sample accumulator  Z_i = sum_{idx <= index} dQ_idx,i / Prob(idx).
// zacc_kernel adds one idx's term to Z from the raw samples that idx's launch left on the device ([R][q][2][n], X = 0 is dQ);
// the product and the sum are rounded separately, term by term in idx order, exactly as `add_to_accumulator!` does with
// `1 / Prob(idx) * s_diff`.  zacc_moments_kernel reduces Z to (n, mean, M2) per (shift, QoI): two passes over the values (they
// are in memory here), fixed tree order.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) zacc_kernel(double *__restrict__ Z, const double *__restrict__ samples, double w, int first,
                                                   unsigned long long n, unsigned long long total) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * 256) {
        const unsigned long long rq = i / n, k = i - rq * n;
        const double term = __dmul_rn(w, samples[(rq * 2) * n + k]);
        Z[i] = first ? term : __dadd_rn(Z[i], term);
    }
}

__global__ void __launch_bounds__(256) zacc_moments_kernel(const double *__restrict__ Z, unsigned long long n, double *__restrict__ out3) {
    __shared__ double red[8];
    const double *z = Z + (size_t)blockIdx.x * n;
    double s = 0.0;
    for (unsigned long long i = threadIdx.x; i < n; i += 256) s += z[i];
    const double mean = block_sum<256>(s, red) / (double)n;
    double q = 0.0;
    for (unsigned long long i = threadIdx.x; i < n; i += 256) { const double d = z[i] - mean; q = __fma_rn(d, d, q); }
    const double m2 = block_sum<256>(q, red);
    if (threadIdx.x == 0) { out3[blockIdx.x * 3 + 0] = (double)n; out3[blockIdx.x * 3 + 1] = mean; out3[blockIdx.x * 3 + 2] = m2; }
}

}  // namespace mle
