// mle_kernels.cuh -- the CUDA kernels of libmle_cuda (sm_100a only).
//
//   points_kernel        K1/K2: lattice (or counter-RNG) point -> shift -> inverse CDF -> dense [m x n x R] in HBM,
//                        tile staged in shared memory and written with one TMA bulk store per tile
//   sample_simple_kernel K5+K6: fused generate -> analytic sample function -> per-CTA moments (no per-sample HBM traffic)
//   lognormal1d_kernel   K1+K3+K4+K6 fused: generate xi tile -> KL contraction (FP64 register-tiled GEMM fed by TMA bulk
//                        loads of the basis) -> exp -> two-sided flux sums for fine and coarse grids -> moments
//   finalize_kernel      fixed-order Chan merge of the per-CTA partial moments
#pragma once
#include "mle_device.cuh"

namespace mle {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

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

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load_g2s(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ===================================================================================================================
// K1 / K2: points -> HBM
//   out[(r*n + i)*m + j].  One tile = P consecutive points of one shift = P*m consecutive doubles of `out`.
//   Persistent CTAs loop over tiles; two shared-memory tile buffers so the TMA store of tile t overlaps tile t+1.
//   Algorithmic traffic: 8 B written per coordinate (z and the shifts stay in L1/L2).
// ===================================================================================================================
constexpr int POINTS_THREADS = 256;
constexpr int POINTS_TILE_CAP = 4096;  // elements per tile buffer (32 KB)

template <int THREADS, int MODE>
__global__ void __launch_bounds__(THREADS) points_kernel(GenArgs g, double *__restrict__ out, int P,
                                                         long long tiles_per_shift, long long ntiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *buf0 = reinterpret_cast<double *>(smem_raw);
    double *buf1 = buf0 + POINTS_TILE_CAP;
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
        // the bulk store issued two iterations ago read this buffer: wait until at most one store is still reading
        if (threadIdx.x == 0) bulk_wait_read<1>();
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
constexpr int SIMPLE_MC = 32;             // dims per chunk; tile[jj][point] (point fastest: conflict-free); 32 dims =
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

// Batched corrected two-pass moments of V values per thread; thread 0 merges them into acc[V] (shared memory).
template <int THREADS, int V>
__device__ __forceinline__ void block_moments_merge(const double (&x)[V], int cnt, double *red /* [3][V][W] */,
                                                    Moment *acc) {
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
        Moment o;
        o.n = (double)cnt;
        o.mean = t / (double)cnt + sd / (double)cnt;
        o.m2 = fmax(sdd - sd * sd / (double)cnt, 0.0);
        acc[v] = moment_merge(acc[v], o);
    }
    __syncthreads();
}

// grid = (G, R).  CTA (g, r) handles point tiles g, g+G, ... of shift r, SIMPLE_THREADS points per tile, dims in chunks.
// partials[(r*V + v)*G + g] = merged (n, mean, M2) of this CTA.  samples (optional): [(r*V + v)*n + i].
template <int MODEL, int V, int MODE>
__global__ void __launch_bounds__(SIMPLE_THREADS) sample_simple_kernel(GenArgs g, SimpleModelArgs ma,
                                                                       Moment *__restrict__ partials,
                                                                       double *__restrict__ samples,
                                                                       long long tiles_per_shift) {
    constexpr int THREADS = SIMPLE_THREADS;
    constexpr int W = THREADS / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *tile = reinterpret_cast<double *>(smem_raw);                    // [SIMPLE_MC][THREADS]
    double *red = tile + THREADS * SIMPLE_MC;                               // [3][V][W]
    Moment *acc = reinterpret_cast<Moment *>(red + 3 * V * W);              // [V]
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
    if (tid < V) acc[tid] = Moment{0.0, 0.0, 0.0};
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
    if (tid < V) partials[((size_t)r * V + tid) * gridDim.x + blockIdx.x] = acc[tid];
}

// ===================================================================================================================
// finalize: merge the G per-CTA partials of every (shift, value) in a fixed order -> moments[(r*V + v)*3 + {n,mean,M2}]
// one warp per (r, v): lane l merges partials [l*chunk, (l+1)*chunk) in order, then a fixed shuffle tree.
// ===================================================================================================================
__global__ void finalize_kernel(const Moment *__restrict__ partials, int G, int RV, double *__restrict__ moments) {
    const int rv = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (rv >= RV) return;
    const int lane = threadIdx.x & 31;
    const int chunk = (G + 31) / 32;
    Moment a{0.0, 0.0, 0.0};
    const Moment *p = partials + (size_t)rv * G;
    for (int i = lane * chunk; i < min(G, (lane + 1) * chunk); ++i) a = moment_merge(a, p[i]);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Moment b;
        b.n = __shfl_down_sync(0xffffffffu, a.n, o);
        b.mean = __shfl_down_sync(0xffffffffu, a.mean, o);
        b.m2 = __shfl_down_sync(0xffffffffu, a.m2, o);
        if ((lane & (2 * o - 1)) == 0) a = moment_merge(a, b);
    }
    if (lane == 0) {
        moments[(size_t)rv * 3 + 0] = a.n;
        moments[(size_t)rv * 3 + 1] = a.mean;
        moments[(size_t)rv * 3 + 2] = a.m2;
    }
}

// ===================================================================================================================
// 1-D lognormal diffusion, fully fused (model definition: see the header of include/mle_cuda.h and DESIGN.md)
//
//  per CTA tile: BT = 64 samples of one shift, 256 threads (8 warps), 2 CTAs per SM.
//   A. xi tile [mpad][BT+4] (k-major, sample fastest) generated in shared memory: lattice -> shift -> Phi^-1.
//   B. KL contraction on the FP64 tensor instruction: G[row][sample] = sum_k A[row][k] xi[k][sample] with
//      mma.sync.m8n8k4.f64 (DMMA), accumulators chained over k.  On B200 one DMMA is bit-identical to four fused
//      multiply-adds in ascending k (tools/dmma_order_test.cu: 262144/262144), so the field equals the oracle's
//      sequential-FMA field bit for bit.  A "row" is (half-node i', side): side 0 = node i', side 1 = the mirrored node
//      N-i', so the sweeps of BOTH halves consume rows in ascending order.  Rows are
//      processed in chunks of RC = 32 (4 row tiles x 8 sample tiles, each warp a 2x2 block of 8x8 tiles).  The basis is
//      stored in HBM in fragment order, so an A fragment is one fully coalesced 256-byte load served by L1/L2; the B
//      fragment comes from the xi tile (padded stride: conflict-free).  Epilogue: a = exp(G) -> F[row][sample] (smem).
//   C. sweeps: thread (side, sample) adds one edge per half-node to the fine flux sums and one to the coarse ones on every
//      even half-node (coarse field = even nodes of the fine field, docs/src/example.md:120); only running scalars are
//      kept -- the field never leaves the SM and nothing per-sample is written to HBM.
//   D. the two sides meet at the mid node: Qf, Qc, dQ = Qf - Qc; corrected two-pass block moments; Chan merge.
// ===================================================================================================================
struct Lognormal1dArgs {
    const double *basis;  // fragment-ordered basis of this level: [row tile][k step][lane] (see mle_model_set_basis)
    int N;                // fine cells
    int mpad;             // KL terms padded to a multiple of 4
    int nrows;            // N + 2 rows (both sides, mid node twice)
    int nchunks;          // row chunks of LN_RC rows
    int level;
    int skip;             // debug (MLE_SKIP env): 1 = generator, 2 = contraction, 4 = exp, 8 = sweeps; results invalid when set
    double h2, h2c;       // 1/N^2, 1/(N/2)^2
};

constexpr int LN_RC = 32;  // rows per chunk (16 half-nodes x 2 sides)
#ifndef LN_GEN_U
#define LN_GEN_U 4  // coordinates a generator thread keeps in flight
#endif
// samples per tile BT (32 or 64) is a template parameter; THREADS = 4*BT; xi row stride BT+4 (the 4 k-rows of a B fragment
// fall into different banks); F row stride BT+8 (conflict-free 16-byte epilogue stores)
__host__ __device__ constexpr int ln_threads(int bt) { return 4 * bt; }
__host__ __device__ constexpr int ln_ldx(int bt) { return bt + 4; }
__host__ __device__ constexpr int ln_ldf(int bt) { return bt + 8; }

struct SweepState {
    double a_prev, A, D;  // last node value, sum of r = 1/kappa, sum of w*r over the edges seen so far
};

// one edge on arrival of node value `a` (node number ip on this side, counted from the boundary; M = N/2): the closed form of
// the mid value (oracle/mle_oracle.c:tridiag_mid) needs only two running sums per side; the division is off the chain
__device__ __forceinline__ void sweep_step(SweepState &s, int ip, double a, int M) {
    if (ip > 0) {
        const double kap = __dmul_rn(0.5, __dadd_rn(s.a_prev, a));
        const double r = ddiv_normal(1.0, kap);  // == 1.0 / kap for every kappa in the normal range (a = exp(g), |g| < 700); branch-free,
                                                 // so the divisions of successive edges pipeline
        s.A = __dadd_rn(s.A, r);
        s.D = __fma_rn((double)(M - ip) + 0.5, r, s.D);
    }
    s.a_prev = a;
}

// u(1/2) = h^2 (D_R A_L + D_L A_R) / (A_L + A_R) from the left state L and the right (mirrored) state Rr
__device__ __forceinline__ double sweep_combine(const SweepState &L, const SweepState &Rr, double h2) {
    const double num = __fma_rn(Rr.D, L.A, __dmul_rn(L.D, Rr.A));
    return __dmul_rn(h2, __ddiv_rn(num, __dadd_rn(L.A, Rr.A)));
}

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// per-thread running statistics (Welford): n is the same for every thread of a tile slot, so 1/n is computed once
struct Welford {
    double mean, m2;
};
__device__ __forceinline__ void welford_add(Welford &w, double x, double inv_n) {
    const double d = x - w.mean;
    w.mean += d * inv_n;
    w.m2 += d * (x - w.mean);
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

#ifdef MLE_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[8];
__device__ unsigned long long g_cta_times[3 * 4096];  // per CTA of the last lognormal1d launch: start ns, end ns, SM id
#define MLE_PH_DECL long long ph_t = clock64(); long long ph_acc[6] = {0, 0, 0, 0, 0, 0}; if (threadIdx.x == 0) { unsigned smid_; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_)); const unsigned l_ = blockIdx.y * gridDim.x + blockIdx.x; if (l_ < 4096) { g_cta_times[3 * l_] = global_timer_ns(); g_cta_times[3 * l_ + 2] = smid_; } }
#define MLE_PH(i) do { long long ph_n = clock64(); ph_acc[i] += ph_n - ph_t; ph_t = ph_n; } while (0)
#define MLE_PH_FLUSH do { if (threadIdx.x == 0) { for (int i_ = 0; i_ < 6; ++i_) atomicAdd(&g_phase_cycles[i_], (unsigned long long)ph_acc[i_]); const unsigned l_ = blockIdx.y * gridDim.x + blockIdx.x; if (l_ < 4096) g_cta_times[3 * l_ + 1] = global_timer_ns(); } } while (0)
#else
#define MLE_PH_DECL
#define MLE_PH(i)
#define MLE_PH_FLUSH
#endif

template <int MODE, int BT>
__global__ void __launch_bounds__(4 * BT, (BT == 64) ? 2 : 5) lognormal1d_kernel(GenArgs g, Lognormal1dArgs la,
                                                                                 Moment *__restrict__ partials,
                                                                                 double *__restrict__ samples,
                                                                                 long long tiles_per_shift) {
    constexpr int THREADS = ln_threads(BT), RC = LN_RC, LDX = ln_ldx(BT), LDF = ln_ldf(BT), W = THREADS / 32;
    constexpr int WC = BT / 16;  // warp grid: 2 (row-tile pairs) x WC (sample-tile pairs)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int mpad = la.mpad;
    double *xi = reinterpret_cast<double *>(smem_raw);           // [mpad][LDX]
    double *F = xi + (size_t)mpad * LDX;                         // [RC][LDF]
    SweepState *xch = reinterpret_cast<SweepState *>(F + RC * LDF);  // [3][BT]: fine right, coarse left, coarse right states
    Moment *mred = reinterpret_cast<Moment *>(xch + 3 * BT);     // [2][BT] final per-thread moments
    GenScratch sc;
    sc.wtot = reinterpret_cast<unsigned *>(mred + 2 * BT);       // [W]
    // generator queue [BT * min(m, 256)]: lives inside F when it fits (F is idle while the inputs are generated)
    const int qbytes = BT * (g.m < 256 ? ((g.m + 3) & ~3) : 256) * 2;
    // {z_j, shift_j} of this CTA's shift (lattice modes), 16-byte aligned behind wtot; then the queue when it does not fit in F
    double2 *zs = reinterpret_cast<double2 *>(sc.wtot + ((W + 3) & ~3));
    constexpr bool lattice = (MODE & GEN_MC) == 0;
    unsigned short *qtail = reinterpret_cast<unsigned short *>(zs + (lattice ? g.m : 0));
    sc.queue = qbytes <= RC * LDF * 8 ? reinterpret_cast<unsigned short *>(F) : qtail;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, r = blockIdx.y;
    if (lattice) stage_lattice_row<THREADS>(zs, g, r);
    const int M = la.N / 2;  // mid node
    const int nks = mpad >> 2;
    // warp -> 2x2 block of 8x8 tiles inside the RC x BT chunk: row tiles {2*wr, 2*wr+1}, sample tiles {2*wc, 2*wc+1}
    const int wr = warp / WC, wc = warp % WC;
    const int lr = lane >> 2, lk = lane & 3;
    for (int e = tid; e < (mpad - g.m) * LDX; e += THREADS) xi[(size_t)g.m * LDX + e] = 0.0;  // k padding rows
    __syncthreads();
    Welford wdq{0.0, 0.0}, wqf{0.0, 0.0};  // threads tid < BT: running statistics of sample slot tid over this CTA's tiles
    double ncount = 0.0;
    MLE_PH_DECL

    for (long long ti = blockIdx.x; ti < tiles_per_shift; ti += gridDim.x) {
        const unsigned long long i0 = (unsigned long long)ti * BT;
        const unsigned long long left = g.n - i0;
        const int Pt = left < (unsigned long long)BT ? (int)left : BT;
        MLE_PH(5);
        // A. inputs (always a full tile: surplus points are masked out of the statistics)
        if (la.skip & 1) {
        } else if (g.m <= 32 * (THREADS / BT)) {  // at most 32 dims per thread: 32-bit class masks
            gen_tile_owned<THREADS, BT, MODE, LDX, unsigned, LN_GEN_U, lattice>(xi, 0, g.m, g.k0 + i0, g, r, sc, zs);
        } else {
            constexpr int GEN_DIMS = 64 * (THREADS / BT);  // 64 iterations per thread fit the 64-bit class masks
            for (int j0 = 0; j0 < g.m; j0 += GEN_DIMS)
                gen_tile_owned<THREADS, BT, MODE, LDX, unsigned long long, LN_GEN_U, lattice>(xi + (size_t)j0 * LDX, j0, min(GEN_DIMS, g.m - j0), g.k0 + i0, g, r, sc, zs);
        }

        MLE_PH(0);
        // sweep roles: thread (grid, side, sample) = (tid / (2 BT), (tid / BT) & 1, tid % BT); grid 0 = fine, 1 = coarse (even nodes)
        SweepState st{0, 0, 0};
        const int side = (tid / BT) & 1, coarse = tid / (2 * BT), b = tid & (BT - 1);

        for (int c = 0; c < la.nchunks; ++c) {
            // B. contraction for rows [c*RC, c*RC + RC).  A full chunk (4 live row tiles) is split 2x2 per warp: row tiles
            // {2wr, 2wr+1} x sample tiles {2wc, 2wc+1}.  A partial chunk (the only chunk of the coarse levels, the last chunk
            // of the others) would leave half the warps idle that way, so there warp w owns sample tile w and ALL live row tiles.
            const int live = min(RC / 8, (la.nrows - c * RC + 7) / 8);
            const bool full = live == RC / 8;
            const int rt0 = c * (RC / 8) + 2 * wr;
            double c00a = 0, c00b = 0, c01a = 0, c01b = 0, c10a = 0, c10b = 0, c11a = 0, c11b = 0;
            if (!(la.skip & 2)) {
                if (full) {
                    const double *A0 = la.basis + (size_t)rt0 * nks * 32 + lane;
                    const double *A1 = A0 + (size_t)nks * 32;
                    const double *B0 = xi + lk * LDX + 16 * wc + lr;
                    double a0 = __ldg(A0), a1 = __ldg(A1);
                    double b0 = B0[0], b1 = B0[8];
#pragma unroll 5
                    for (int ks = 0; ks < nks; ++ks) {
                        const int kn = (ks + 1 < nks) ? ks + 1 : ks;  // prefetch the next k step's fragments
                        const double na0 = __ldg(A0 + kn * 32), na1 = __ldg(A1 + kn * 32);
                        const double nb0 = B0[kn * 4 * LDX], nb1 = B0[kn * 4 * LDX + 8];
                        dmma_m8n8k4(c00a, c00b, a0, b0);
                        dmma_m8n8k4(c01a, c01b, a0, b1);
                        dmma_m8n8k4(c10a, c10b, a1, b0);
                        dmma_m8n8k4(c11a, c11b, a1, b1);
                        a0 = na0; a1 = na1; b0 = nb0; b1 = nb1;
                    }
                } else {
                    // accumulators: (c00), (c01), (c10) = row tiles 0, 1, 2 of the chunk for sample tile `warp`
                    const double *Ab = la.basis + (size_t)(c * (RC / 8)) * nks * 32 + lane;
                    const double *Bp = xi + lk * LDX + 8 * warp + lr;
                    const size_t ts = (size_t)nks * 32;
                    double a0 = __ldg(Ab), a1 = live > 1 ? __ldg(Ab + ts) : 0.0, a2 = live > 2 ? __ldg(Ab + 2 * ts) : 0.0;
                    double b0 = Bp[0];
#pragma unroll 5
                    for (int ks = 0; ks < nks; ++ks) {
                        const int kn = (ks + 1 < nks) ? ks + 1 : ks;
                        const double na0 = __ldg(Ab + kn * 32);
                        const double na1 = live > 1 ? __ldg(Ab + ts + kn * 32) : 0.0;
                        const double na2 = live > 2 ? __ldg(Ab + 2 * ts + kn * 32) : 0.0;
                        const double nb0 = Bp[kn * 4 * LDX];
                        dmma_m8n8k4(c00a, c00b, a0, b0);
                        if (live > 1) dmma_m8n8k4(c01a, c01b, a1, b0);
                        if (live > 2) dmma_m8n8k4(c10a, c10b, a2, b0);
                        a0 = na0; a1 = na1; a2 = na2; b0 = nb0;
                    }
                }
            }
            MLE_PH(1);
            __syncthreads();  // the previous chunk's sweeps have finished reading F
            MLE_PH(2);
            // epilogue: a = exp(G) -> F[row in chunk][sample]; lane holds row lr, samples 2*lk, 2*lk+1 of each 8x8 tile
            if (la.skip & 4) { c00a = c00b = c01a = c01b = c10a = c10b = c11a = c11b = -1e300; }
            if (full) {
                double *f0 = F + (16 * wr + lr) * LDF + 16 * wc + 2 * lk;
                double2 v;
                v.x = exp_det(c00a); v.y = exp_det(c00b);
                *reinterpret_cast<double2 *>(f0) = v;
                v.x = exp_det(c01a); v.y = exp_det(c01b);
                *reinterpret_cast<double2 *>(f0 + 8) = v;
                v.x = exp_det(c10a); v.y = exp_det(c10b);
                *reinterpret_cast<double2 *>(f0 + 8 * LDF) = v;
                v.x = exp_det(c11a); v.y = exp_det(c11b);
                *reinterpret_cast<double2 *>(f0 + 8 * LDF + 8) = v;
            } else {
                double *f0 = F + lr * LDF + 8 * warp + 2 * lk;
                double2 v;
                v.x = exp_det(c00a); v.y = exp_det(c00b);
                *reinterpret_cast<double2 *>(f0) = v;
                if (live > 1) { v.x = exp_det(c01a); v.y = exp_det(c01b); *reinterpret_cast<double2 *>(f0 + 8 * LDF) = v; }
                if (live > 2) { v.x = exp_det(c10a); v.y = exp_det(c10b); *reinterpret_cast<double2 *>(f0 + 16 * LDF) = v; }
            }
            __syncthreads();
            MLE_PH(3);
            // C. sweeps over the RC/2 half-nodes of this chunk: warps 0-1 the fine grid (every half-node), warps 2-3 the coarse
            // grid (every second one; idle at level 0)
            if (!(la.skip & 8)) {
                const int ip0 = c * (RC / 2);
                if (!coarse) {
#pragma unroll 8
                    for (int q = 0; q < RC / 2; ++q) {
                        const int ip = ip0 + q;
                        if (ip <= M) sweep_step(st, ip, F[(2 * q + side) * LDF + b], M);
                    }
                } else if (la.level > 0) {
#pragma unroll 4
                    for (int q = 0; q < RC / 2; q += 2) {  // ip0 is even: the even half-nodes of the chunk
                        const int ip = ip0 + q;
                        if (ip <= M) sweep_step(st, ip >> 1, F[(2 * q + side) * LDF + b], M / 2);
                    }
                }
            }
        }
        MLE_PH(4);
        // D. meet in the middle: thread b collects the fine right state and both coarse states
        if (tid >= BT) xch[(tid / BT - 1) * BT + b] = st;
        __syncthreads();
        if (tid < BT) {
            const double qf = sweep_combine(st, xch[b], la.h2);
            double dq = qf;
            if (la.level > 0) dq = __dsub_rn(qf, sweep_combine(xch[BT + b], xch[2 * BT + b], la.h2c));
            if (b < Pt) {
                if (samples) {
                    samples[((size_t)r * 2 + 0) * g.n + i0 + b] = dq;
                    samples[((size_t)r * 2 + 1) * g.n + i0 + b] = qf;
                }
                ncount += 1.0;
                const double inv_n = 1.0 / ncount;
                welford_add(wdq, dq, inv_n);
                welford_add(wqf, qf, inv_n);
            }
        }
        // (the barriers of the next tile's generator order the xch reads above before xch is rewritten)
    }
    MLE_PH(5);
    MLE_PH_FLUSH;
    // merge the BT per-thread statistics in a fixed order: slot b holds every BT-th sample of this CTA's tiles
    if (tid < BT) {
        mred[tid] = Moment{ncount, wdq.mean, wdq.m2};
        mred[BT + tid] = Moment{ncount, wqf.mean, wqf.m2};
    }
    __syncthreads();
    if (tid < 2) {
        Moment a{0.0, 0.0, 0.0};
        for (int i = 0; i < BT; ++i) a = moment_merge(a, mred[tid * BT + i]);
        partials[((size_t)r * 2 + tid) * gridDim.x + blockIdx.x] = a;
    }
}


#ifdef MLE_WITH_WS  // experiment, off by default: make EXTRA=-DMLE_WITH_WS, then MLE_LN_WS=1 at run time
// ===================================================================================================================
// 1-D lognormal diffusion, warp-specialised variant (MLE_LN_WS=1; an experiment kept for A/B runs -- measured 8 % slower
// than the phase-synchronous kernel on the C2 step): same arithmetic as lognormal1d_kernel, different schedule.  ncu showed the fused kernel above spending 35-50 % of its non-generator time at the barrier that separates the
// KL contraction from the elimination sweeps: the sweeps are a latency-bound dependent chain (one IEEE division per row) that
// only 2 of 4 warps execute while the other two idle.  Here a CTA is 5 warps for a 32-sample tile:
//   warps 0-3 (producers): KL contraction with DMMA + exp epilogue into a DOUBLE-BUFFERED field chunk F[2][32 rows][32];
//   warp 4 (consumer): all four elimination chains of sample `lane` (fine/coarse x left/right -> 4-way ILP in one
//                      thread), meeting in the middle without any shared-memory exchange, then Welford statistics.
// Producers and consumer hand chunks over with named barriers (bar.arrive / bar.sync, ids 1-4), so the contraction of chunk
// c+1 overlaps the sweeps of chunk c and nobody waits at a CTA-wide barrier inside the chunk loop.
// ===================================================================================================================
constexpr int LNW_BT = 32;
constexpr int LNW_THREADS = 160;
constexpr int LNW_LDX = LNW_BT + 4;
constexpr int LNW_LDF = LNW_BT + 8;

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <int MODE>
__global__ void __launch_bounds__(LNW_THREADS, 4) lognormal1d_ws_kernel(GenArgs g, Lognormal1dArgs la,
                                                                         Moment *__restrict__ partials,
                                                                         double *__restrict__ samples,
                                                                         long long tiles_per_shift) {
    constexpr int THREADS = LNW_THREADS, BT = LNW_BT, RC = LN_RC, LDX = LNW_LDX, LDF = LNW_LDF;
    constexpr int G = THREADS / BT;  // 5 generator threads per point
    constexpr int FULL0 = 1, EMPTY0 = 3;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int mpad = la.mpad;
    double *xi = reinterpret_cast<double *>(smem_raw);            // [mpad][LDX]
    double *F = xi + (size_t)mpad * LDX;                          // [2][RC][LDF]
    Moment *mred = reinterpret_cast<Moment *>(F + 2 * RC * LDF);  // [2][BT]
    GenScratch sc;
    sc.wtot = reinterpret_cast<unsigned *>(mred + 2 * BT);        // [W]
    sc.queue = reinterpret_cast<unsigned short *>(sc.wtot + 8);   // [BT * dims per call]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, r = blockIdx.y;
    const int M = la.N / 2;
    const int nks = mpad >> 2;
    const bool consumer = warp == 4;
    const int wr = warp >> 1, wc = warp & 1;  // producers: row tiles {2wr, 2wr+1}, sample tiles {2wc, 2wc+1}
    const int lr = lane >> 2, lk = lane & 3;
    for (int e = tid; e < (mpad - g.m) * LDX; e += THREADS) xi[(size_t)g.m * LDX + e] = 0.0;  // k padding rows
    Welford wdq{0.0, 0.0}, wqf{0.0, 0.0};  // consumer lanes: running statistics of sample slot `lane`
    double ncount = 0.0;

    for (long long ti = blockIdx.x; ti < tiles_per_shift; ti += gridDim.x) {
        const unsigned long long i0 = (unsigned long long)ti * BT;
        const unsigned long long left = g.n - i0;
        const int Pt = left < (unsigned long long)BT ? (int)left : BT;
        __syncthreads();  // previous tile fully consumed (xi, F), padding rows written
        if (g.m <= 32 * G) {
            gen_tile_owned<THREADS, BT, MODE, LDX, unsigned, LN_GEN_U>(xi, 0, g.m, g.k0 + i0, g, r, sc);
        } else {
            for (int j0 = 0; j0 < g.m; j0 += 64 * G)
                gen_tile_owned<THREADS, BT, MODE, LDX, unsigned long long, LN_GEN_U>(xi + (size_t)j0 * LDX, j0, min(64 * G, g.m - j0), g.k0 + i0, g, r, sc);
        }
        // (gen_tile_owned ends with a CTA barrier: xi is complete)

        if (!consumer) {
            for (int c = 0; c < la.nchunks; ++c) {
                const int rt0 = c * (RC / 8) + 2 * wr;
                const int rows_left = la.nrows - rt0 * 8;
                double c00a = 0, c00b = 0, c01a = 0, c01b = 0, c10a = 0, c10b = 0, c11a = 0, c11b = 0;
                if (rows_left > 0) {
                    const bool two = rows_left > 8;
                    const double *A0 = la.basis + (size_t)rt0 * nks * 32 + lane;
                    const double *A1 = two ? A0 + (size_t)nks * 32 : A0;
                    const double *B0 = xi + lk * LDX + 16 * wc + lr;
                    double a0 = __ldg(A0), a1 = __ldg(A1);
                    double b0 = B0[0], b1 = B0[8];
#pragma unroll 5
                    for (int ks = 0; ks < nks; ++ks) {
                        const int kn = (ks + 1 < nks) ? ks + 1 : ks;  // prefetch the next k step's fragments
                        const double na0 = __ldg(A0 + kn * 32), na1 = __ldg(A1 + kn * 32);
                        const double nb0 = B0[kn * 4 * LDX], nb1 = B0[kn * 4 * LDX + 8];
                        dmma_m8n8k4(c00a, c00b, a0, b0);
                        dmma_m8n8k4(c01a, c01b, a0, b1);
                        if (two) {
                            dmma_m8n8k4(c10a, c10b, a1, b0);
                            dmma_m8n8k4(c11a, c11b, a1, b1);
                        }
                        a0 = na0; a1 = na1; b0 = nb0; b1 = nb1;
                    }
                }
                if (c >= 2) named_bar_sync(EMPTY0 + (c & 1), THREADS);  // the consumer has released this F buffer
                if (rows_left > 0) {
                    double *f0 = F + (c & 1) * RC * LDF + (16 * wr + lr) * LDF + 16 * wc + 2 * lk;
                    double2 v;
                    v.x = exp_det(c00a); v.y = exp_det(c00b);
                    *reinterpret_cast<double2 *>(f0) = v;
                    v.x = exp_det(c01a); v.y = exp_det(c01b);
                    *reinterpret_cast<double2 *>(f0 + 8) = v;
                    if (rows_left > 8) {
                        v.x = exp_det(c10a); v.y = exp_det(c10b);
                        *reinterpret_cast<double2 *>(f0 + 8 * LDF) = v;
                        v.x = exp_det(c11a); v.y = exp_det(c11b);
                        *reinterpret_cast<double2 *>(f0 + 8 * LDF + 8) = v;
                    }
                }
                named_bar_arrive(FULL0 + (c & 1), THREADS);
            }
        } else {
            SweepState fl{0, 0, 0}, fr{0, 0, 0}, cl{0, 0, 0}, cr{0, 0, 0};
            for (int c = 0; c < la.nchunks; ++c) {
                named_bar_sync(FULL0 + (c & 1), THREADS);
                const double *Fc = F + (c & 1) * RC * LDF + lane;
                const int ip0 = c * (RC / 2);
#pragma unroll 2
                for (int q = 0; q < RC / 2; ++q) {
                    const int ip = ip0 + q;
                    if (ip <= M) {
                        const double al = Fc[(2 * q) * LDF], ar = Fc[(2 * q + 1) * LDF];
                        sweep_step(fl, ip, al, M);
                        sweep_step(fr, ip, ar, M);
                        if (la.level > 0 && (ip & 1) == 0) {
                            sweep_step(cl, ip >> 1, al, M / 2);
                            sweep_step(cr, ip >> 1, ar, M / 2);
                        }
                    }
                }
                if (c + 2 < la.nchunks) named_bar_arrive(EMPTY0 + (c & 1), THREADS);
            }
            const double qf = sweep_combine(fl, fr, la.h2);
            double dq = qf;
            if (la.level > 0) dq = __dsub_rn(qf, sweep_combine(cl, cr, la.h2c));
            if (lane < Pt) {
                if (samples) {
                    samples[((size_t)r * 2 + 0) * g.n + i0 + lane] = dq;
                    samples[((size_t)r * 2 + 1) * g.n + i0 + lane] = qf;
                }
                ncount += 1.0;
                const double inv_n = 1.0 / ncount;
                welford_add(wdq, dq, inv_n);
                welford_add(wqf, qf, inv_n);
            }
        }
    }
    if (consumer) {
        mred[lane] = Moment{ncount, wdq.mean, wdq.m2};
        mred[BT + lane] = Moment{ncount, wqf.mean, wqf.m2};
    }
    __syncthreads();
    if (tid < 2) {
        Moment a{0.0, 0.0, 0.0};
        for (int i = 0; i < BT; ++i) a = moment_merge(a, mred[tid * BT + i]);
        partials[((size_t)r * 2 + tid) * gridDim.x + blockIdx.x] = a;
    }
}


#endif  // MLE_WITH_WS

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
__global__ void __launch_bounds__(TG == 32 ? 128 : 256) lognormal2d_kernel(GenArgs g, Lognormal2dArgs la,
                                                                           Moment *__restrict__ partials,
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
    Moment *mred = reinterpret_cast<Moment *>(reinterpret_cast<double *>(smem_raw) + (size_t)GROUPS * per_group);  // [2][GROUPS]
    constexpr bool lattice = (MODE & GEN_MC) == 0;
    const double *shift_r = lattice ? g.shifts + (size_t)r * g.s : nullptr;
    Welford wdq{0.0, 0.0}, wqf{0.0, 0.0};
    double ncount = 0.0;

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
                                         : mc_uniform(g.seed, k, (uint32_t)g.m, (uint32_t)j);
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
                ncount += 1.0;
                const double inv_n = 1.0 / ncount;
                welford_add(wdq, dq, inv_n);
                welford_add(wqf, qf, inv_n);
            }
            group_sync<TG>();
        }
    }
    __syncthreads();
    if (t == 0) {
        mred[grp] = Moment{ncount, wdq.mean, wdq.m2};
        mred[GROUPS + grp] = Moment{ncount, wqf.mean, wqf.m2};
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        Moment a{0.0, 0.0, 0.0};
        for (int q = 0; q < GROUPS; ++q) a = moment_merge(a, mred[threadIdx.x * GROUPS + q]);
        partials[((size_t)r * 2 + threadIdx.x) * gridDim.x + blockIdx.x] = a;
    }
}

}  // namespace mle
