// mle_kernel_ln1d_small.cuh -- 1-D lognormal diffusion, LATENCY variant for the tiny batches of the adaptive loop (sm_100a).
//
// The QMC loop of the reference grows one index by a factor 1.2 per iteration (src/run.jl:96-113, src/methods.jl:265-273): most
// `sample!` calls carry 1 .. 100 points per shift, and on the fine levels a call is a handful of samples on a 512- or 1024-cell
// grid.  The throughput kernels (mle_kernel_ln1d.cuh, mle_kernel_ln1d_block.cuh) walk such a grid in 32-row chunks -- contraction,
// exp, sweeps, two barriers per chunk, 33 chunks in sequence at level 6 -- which is the right schedule when the SMs are full and
// the wrong one when a call is sixteen CTAs: 88 us for ONE sample per shift at level 6 (profiles/r02_call_latency.txt).
//
// Same model, same arithmetic per sample and the same basis layout; the schedule is organised around the critical path instead:
//   tile = 8 samples of one shift (ONE 8-wide DMMA sample tile), 256 threads, the whole field of the tile in shared memory.
//   A. inputs: gen_tile_owned with 32 threads per point (3-4 coordinates each instead of 25-100 in sequence).
//   B. KL contraction: the basis streams through two shared-memory panels of 8 row tiles (TMA bulk loads + mbarriers, two panels
//      in flight); warp w takes row tile w of the panel: one chain of `nks` DMMAs (ascending k, as everywhere); the raw sums go
//      to G[row][sample].  The mid-node chain rides along in warp 0's first panel.
//   C. exp over the whole field, all threads, in place (the same exp_det on the same sums).
//   D. all reciprocals r = 1/kappa of the fine AND the coarse grid, all threads (rcp_kappa: the Newton sequence of the sweeps).
//   E. the only sequential part: per (grid, side, sample) one thread adds its edges from the boundary towards the mid node,
//      A += r, D = fma(w, r, D) -- the oracle's order (oracle/mle_oracle.c:tridiag_mid), two dependent operations per edge.
//   F. the sides meet, (dQ, Qf), per-thread pivoted statistics, cta_finalize.
// Per-sample values are bit-identical to the other two schedules and to the oracle (tests/test_gpu_batch.py).
#pragma once
#include "mle_kernel_ln1d.cuh"

namespace mle {

constexpr int LNS_BT = 8;        // samples per tile
constexpr int LNS_THREADS = 256;
constexpr int LNS_PT = 8;        // row tiles per panel (one per warp)

// bytes of the region shared by the two basis panels (phase B) and the reciprocals (phases D-E)
__host__ __device__ constexpr size_t lns_region_bytes(int mpad, int ntiles) {
    const size_t panels = (size_t)2 * (ntiles < LNS_PT ? ntiles : LNS_PT) * (size_t)(mpad / 4) * 32 * 8;
    const size_t recips = (size_t)ntiles * 8 * LNS_BT * 8 * 3 / 2 + 2 * 16 * 8;
    return panels > recips ? panels : recips;
}
__host__ __device__ constexpr size_t lns_smem_bytes(int mpad, int m_lattice, int ntiles, int m) {
    return (size_t)mpad * LNS_BT * 8            // xi
           + (size_t)ntiles * 8 * LNS_BT * 8    // G / F
           + (size_t)LNS_BT * 8 + (size_t)mpad * 8  // amid, smid
           + 3 * LNS_BT * 16                    // exchanged (A, D) states
           + 2 * LNS_BT * sizeof(Partial)       // final per-thread statistics
           + 32                                 // wtot
           + (size_t)m_lattice * 16             // {z, shift}
           + (((size_t)LNS_BT * ((m + 3) & ~3) * 2 + 15) & ~(size_t)15)  // generator queue
           + 128 + lns_region_bytes(mpad, ntiles) + 16;                  // alignment slack, panels / reciprocals, mbarriers
}

template <int MODE>
__global__ void __launch_bounds__(LNS_THREADS, 1) lognormal1d_small_kernel(GenArgs g, Lognormal1dArgs la, FinArgs fa,
                                                                           double *__restrict__ samples, long long tiles_per_shift) {
    constexpr int BT = LNS_BT, THREADS = LNS_THREADS, W = THREADS / 32, PT = LNS_PT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int mpad = la.mpad, nks = mpad >> 2, ntiles = la.ntiles;
    const int M = la.N / 2, Mc = M / 2;
    double *xi = reinterpret_cast<double *>(smem_raw);                  // [mpad][BT]
    double *F = xi + (size_t)mpad * BT;                                 // [ntiles * 8 rows r' = 2 i' + side][BT]
    double *amid = F + (size_t)ntiles * 8 * BT;                         // [BT]
    double *smid = amid + BT;                                           // [mpad]
    double2 *xch = reinterpret_cast<double2 *>(smid + mpad);            // [3][BT] (A, D): fine right, coarse left, coarse right
    Partial *mred = reinterpret_cast<Partial *>(xch + 3 * BT);          // [2][BT]
    GenScratch sc;
    sc.wtot = reinterpret_cast<unsigned *>(mred + 2 * BT);              // [W]
    double2 *zs = reinterpret_cast<double2 *>(sc.wtot + W);
    constexpr bool lattice = (MODE & GEN_MC) == 0;
    sc.queue = reinterpret_cast<unsigned short *>(zs + (lattice ? g.m : 0));
    const size_t qbytes = (((size_t)BT * ((g.m + 3) & ~3) * 2) + 15) & ~(size_t)15;
    const size_t used = (reinterpret_cast<unsigned char *>(sc.queue) - smem_raw) + qbytes;
    double *region = reinterpret_cast<double *>(smem_raw + ((used + 127) & ~(size_t)127));
    const int ptiles = ntiles < PT ? ntiles : PT;
    const int pdoubles = ptiles * nks * 32;                             // one panel
    double *Rf = region;                                                // [M edges][2 sides][BT]: edge ip-1 -> node ip
    double *Rc = region + (size_t)M * 2 * BT;                           // [Mc edges][2 sides][BT]
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(region) + lns_region_bytes(mpad, ntiles));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, r = blockIdx.y;
    const int lr = lane >> 2, lk = lane & 3;
    if (tid == 0) { mbar_init(mbar, 1); mbar_init(mbar + 1, 1); mbar_fence_init(); }
    if (lattice) stage_lattice_row<THREADS>(zs, g, r);
    for (int e = tid; e < mpad; e += THREADS) smid[e] = __ldg(la.mid + e);
    for (int e = tid; e < (mpad - g.m) * BT; e += THREADS) xi[(size_t)g.m * BT + e] = 0.0;  // k padding rows
    __syncthreads();
    const int npanels = (ntiles + PT - 1) / PT;
    unsigned cc = 0;  // panels consumed so far by this CTA (buffer cc & 1, mbarrier phase (cc >> 1) & 1)
    Partial pdq{0.0, 0.0, 0.0, 0.0}, pqf{0.0, 0.0, 0.0, 0.0};  // threads tid < BT

    for (long long ti = blockIdx.x; ti < tiles_per_shift; ti += gridDim.x) {
        const unsigned long long i0 = (unsigned long long)ti * BT;
        const unsigned long long left = g.n - i0;
        const int Pt = left < (unsigned long long)BT ? (int)left : BT;
        const unsigned cc0 = cc;
        auto issue_panel = [&](int p) {  // thread 0: panel p of this tile -> buffer (cc0 + p) & 1
            const int live = min(PT, ntiles - p * PT);
            const uint32_t bytes = (uint32_t)live * (uint32_t)nks * 32u * 8u;
            const unsigned b = (cc0 + (unsigned)p) & 1u;
            mbar_expect_tx(mbar + b, bytes);
            bulk_load_g2s(region + (size_t)b * pdoubles, la.basis + (size_t)p * PT * nks * 32, bytes, mbar + b);
        };
        // (the barrier that ended the previous tile's phase E ordered every read of the region before these loads)
        if (tid == 0) { issue_panel(0); if (npanels > 1) issue_panel(1); }
        // A. inputs
        gen_tile_owned<THREADS, BT, MODE, BT, unsigned, 4, lattice>(xi, 0, g.m, g.k0 + i0, g, r, sc, zs);
        // B. contraction
        double gmid = 0.0;
        const double *Bp = xi + lk * BT + lr;
        const double *xm = xi + (lane & (BT - 1));
        for (int p = 0; p < npanels; ++p) {
            const unsigned b = cc & 1u;
            mbar_wait(mbar + b, (cc >> 1) & 1u);
            ++cc;
            const int rt = p * PT + warp;
            if (rt < ntiles) {
                const double *Ap = region + (size_t)b * pdoubles + (size_t)warp * nks * 32 + lane;
                double c0 = 0.0, c1 = 0.0;
                if (p == 0 && warp == 0) {
#pragma unroll 5
                    for (int ks = 0; ks < nks; ++ks) {
                        dmma_m8n8k4(c0, c1, Ap[ks * 32], Bp[ks * 4 * BT]);
                        const double2 p01 = *reinterpret_cast<const double2 *>(smid + 4 * ks), p23 = *reinterpret_cast<const double2 *>(smid + 4 * ks + 2);
                        gmid = __fma_rn(p01.x, xm[(4 * ks) * BT], gmid);
                        gmid = __fma_rn(p01.y, xm[(4 * ks + 1) * BT], gmid);
                        gmid = __fma_rn(p23.x, xm[(4 * ks + 2) * BT], gmid);
                        gmid = __fma_rn(p23.y, xm[(4 * ks + 3) * BT], gmid);
                    }
                } else {
#pragma unroll 5
                    for (int ks = 0; ks < nks; ++ks) dmma_m8n8k4(c0, c1, Ap[ks * 32], Bp[ks * 4 * BT]);
                }
                // lane holds row lr, samples 2 lk and 2 lk + 1 of the 8x8 tile
                *reinterpret_cast<double2 *>(F + ((size_t)rt * 8 + lr) * BT + 2 * lk) = make_double2(c0, c1);
            }
            __syncthreads();  // every warp has left buffer b: the panel after next may land there
            if (tid == 0 && p + 2 < npanels) issue_panel(p + 2);
        }
        // C. a = exp(G), in place; four independent evaluations per thread and iteration
        {
            const int total = ntiles * 8 * BT;
            for (int e = tid; e < total; e += 4 * THREADS) {
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = e + u * THREADS < total ? F[e + u * THREADS] : 0.0;
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = exp_det(v[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (e + u * THREADS < total) F[e + u * THREADS] = v[u];
            }
            if (warp == 0 && lane < BT) amid[lane] = exp_det(gmid);
        }
        __syncthreads();
        // D. reciprocals of the interface coefficients.  Fine grid: element e = (edge ip-1, side, sample) reads nodes ip-1 and ip
        // of its side: F[e] and F[e + 16] (rows 2 i' + side), the mid node closing the last edge.  Coarse grid (even nodes):
        // element (edge q-1, side, sample) reads F[32 (q-1) + low] and F[32 q + low].
        {
            const int nf = M * 2 * BT;
            for (int e = tid; e < nf; e += 4 * THREADS) {
                double a0[4], a1[4], rr[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ee = min(e + u * THREADS, nf - 1);
                    a0[u] = F[ee];
                    a1[u] = (ee >> 4) == M - 1 ? amid[ee & (BT - 1)] : F[ee + 16];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) rr[u] = rcp_kappa(a0[u], a1[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (e + u * THREADS < nf) Rf[e + u * THREADS] = rr[u];
            }
            if (la.level > 0) {
                const int nc = Mc * 2 * BT;
                for (int e = tid; e < nc; e += 4 * THREADS) {
                    double a0[4], a1[4], rr[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int ee = min(e + u * THREADS, nc - 1);
                        const int q1 = ee >> 4, low = ee & 15;
                        a0[u] = F[32 * q1 + low];
                        a1[u] = q1 == Mc - 1 ? amid[ee & (BT - 1)] : F[32 * (q1 + 1) + low];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) rr[u] = rcp_kappa(a0[u], a1[u]);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (e + u * THREADS < nc) Rc[e + u * THREADS] = rr[u];
                }
            }
        }
        __syncthreads();
        // E. the flux sums, from the boundary towards the mid node: thread (grid, side, sample) = (tid / 16, (tid / 8) & 1, tid % 8)
        if (tid < 4 * BT) {
            const int coarse = tid >> 4;
            const int ne = coarse ? (la.level > 0 ? Mc : 0) : M;  // edges of this side; edge i (0-based) arrives at node i + 1
            const double *Rp = (coarse ? Rc : Rf) + (tid & 15);
            double A = 0.0, D = 0.0;
            const double w0 = (double)ne - 0.5;                   // weight of edge 0: (M - 1) + 1/2
            int i = 0;
            for (; i + 8 <= ne; i += 8) {
                double rr[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) rr[u] = Rp[(i + u) * 16];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    A = __dadd_rn(A, rr[u]);
                    D = __fma_rn(w0 - (double)(i + u), rr[u], D);
                }
            }
            for (; i < ne; ++i) {
                const double rv = Rp[i * 16];
                A = __dadd_rn(A, rv);
                D = __fma_rn(w0 - (double)i, rv, D);
            }
            if (tid >= BT) xch[tid - BT] = make_double2(A, D);
            __syncwarp();
            // F. the sides meet
            if (tid < BT) {
                const double qf = sweep_combine(A, D, xch[tid].x, xch[tid].y, la.h2);
                double dq = qf;
                if (la.level > 0) dq = __dsub_rn(qf, sweep_combine(xch[BT + tid].x, xch[BT + tid].y, xch[2 * BT + tid].x, xch[2 * BT + tid].y, la.h2c));
                if (tid < Pt) {
                    if (samples) {
                        samples[((size_t)r * 2 + 0) * g.n + i0 + tid] = dq;
                        samples[((size_t)r * 2 + 1) * g.n + i0 + tid] = qf;
                    }
                    partial_add(pdq, dq);
                    partial_add(pqf, qf);
                }
            }
        }
        __syncthreads();  // the region (reciprocals) and F are free for the next tile
    }
    if (tid < BT) {
        mred[tid] = pdq;
        mred[BT + tid] = pqf;
    }
    __syncthreads();
    if (tid < 2) {
        Partial a{0.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < BT; ++i) a = partial_merge(a, mred[tid * BT + i]);
        fa.partials[((size_t)r * 2 + tid) * gridDim.x + blockIdx.x] = a;
    }
    cta_finalize(fa, 2);
}

}  // namespace mle
