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
//   A. inputs: gen_tile_wide (below): 32 threads per point, 3-4 coordinates each instead of 25-100 in sequence.
//   B. KL contraction: the basis streams through two shared-memory panels of 8 row tiles (TMA bulk loads + mbarriers, two panels
//      in flight); warp w takes row tile w of the panel: one chain of `nks` DMMAs (ascending k, as everywhere); the raw sums go
//      to G[row][sample].  The mid-node chain rides along in warp 0's first panel.
//   C. exp over the whole field, all threads, in place (the same exp_det on the same sums).
//   D. all reciprocals r = 1/kappa of the fine AND the coarse grid, all threads (rcp_kappa: the Newton sequence of the sweeps).
//   E. the only sequential part: per (grid, side, sample) one thread adds its edges from the boundary towards the mid node,
//      A += r, D = fma(w, r, D) -- the oracle's order (oracle/mle_oracle.c:tridiag_mid), two dependent operations per edge.
//   F. the sides meet, (dQ, Qf), per-thread pivoted statistics, cta_finalize (moments in the CTA when it saw the whole shift).
// Fine grids (>= 16 row tiles): a thread-block CLUSTER shares one sample tile -- every rank contracts, exponentiates and inverts
// 1/cs of the rows and stores its reciprocals into the leader's shared memory (DSMEM), the leader runs E-F.  One SM takes the basis
// in at ~28 bytes per clock, which at level 6 (819 KB) is 14 us: the split, not a faster inner loop, is what removes it.
// Per-sample values are bit-identical to the other two schedules and to the oracle (tests/test_gpu_batch.py).
#pragma once
#include "mle_kernel_ln1d.cuh"

namespace mle {

#ifdef MLE_LNS_TIMING  // debug builds only (tools/build_variant.sh lnst -DMLE_LNS_TIMING): SM clock at the phase boundaries of CTA (0, 0)
__device__ long long g_lns_t[16];
#define LNS_T(i) do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) g_lns_t[i] = clock64(); } while (0)
#else
#define LNS_T(i) do { } while (0)
#endif

constexpr int LNS_BT = 8;        // samples per tile
constexpr int LNS_THREADS = 256;
constexpr int LNS_PT = 8;        // row tiles per basis panel (one per warp)

// A cluster of `cs` CTAs shares one sample tile: rank k contracts, exponentiates and inverts the row tiles [k tpr, (k+1) tpr)
// (plus one halo tile: the node behind its last edge) and stores its reciprocals into the leader's shared memory; cs = 1: no
// cluster.  Row tiles a CTA keeps in its own F:
__host__ __device__ constexpr int lns_local_tiles(int ntiles, int cs) {
    return cs <= 1 ? ntiles : ((ntiles + cs - 1) / cs + 1 < ntiles ? (ntiles + cs - 1) / cs + 1 : ntiles);
}
// bytes of the region shared by the two basis panels (phase B) and the reciprocals of the WHOLE grid (phases D-E, leader)
__host__ __device__ constexpr size_t lns_region_bytes(int mpad, int ntiles, int cs) {
    const int lt = lns_local_tiles(ntiles, cs);
    const size_t panels = (size_t)2 * (lt < LNS_PT ? lt : LNS_PT) * (size_t)(mpad / 4) * 32 * 8;
    const size_t recips = (size_t)ntiles * 8 * LNS_BT * 8 * 3 / 2 + 2 * 16 * 8;
    return panels > recips ? panels : recips;
}
__host__ __device__ constexpr size_t lns_smem_bytes(int mpad, int ntiles, int m, int cs) {
    return (size_t)mpad * LNS_BT * 8            // xi
           + (size_t)lns_local_tiles(ntiles, cs) * 8 * LNS_BT * 8    // G / F (the CTA's own row tiles)
           + (size_t)LNS_BT * 8 + (size_t)mpad * 8  // amid, smid
           + 3 * LNS_BT * 16                    // exchanged (A, D) states
           + 2 * LNS_BT * sizeof(Partial)       // final per-thread statistics
           + 32                                 // wtot
           + (((size_t)LNS_BT * ((m + 3) & ~3) * 2 + 15) & ~(size_t)15)  // generator queue
           + 128 + lns_region_bytes(mpad, ntiles, cs) + 16;              // alignment slack, panels / reciprocals, mbarriers
}

// the same into the shared memory of CTA `rank` of the cluster (distributed shared memory): `p` is the address the datum has in
// THIS CTA's window -- every CTA of the cluster lays its shared memory out identically
__device__ __forceinline__ void st_cluster_if(double *p, unsigned rank, double v, bool ok) {
    asm volatile("{\n.reg .pred q;\n.reg .b32 ra;\nsetp.ne.u32 q, %3, 0;\nmapa.shared::cluster.u32 ra, %0, %1;\n@q st.shared::cluster.f64 [ra], %2;\n}" ::"r"(smem_u32(p)), "r"(rank), "d"(v),
                 "r"((unsigned)ok)
                 : "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
// all threads of all CTAs of the cluster; release / acquire: the shared-memory stores before it (local and remote) are visible after it
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Point-owned generator for a tile of P = 8 points and 32 threads per point (the layout of gen_tile_owned, mle_device.cuh, for a
// CTA whose threads hold 3-4 coordinates each): the coordinates of a thread -- jj = gg, gg + G, ... -- are evaluated four at a
// time in ONE straight-line block, surplus slots predicated instead of walked by a scalar remainder loop, z_j / shift_j straight
// from global memory (all loads of a thread issue together: no staging pass, no barrier in front), and the parked middle- and
// tail-branch coordinates are finished in ONE pass over both queues (different warps take different branches: the pass costs
// the longer chain, not the sum).  Same functions on the same inputs as every other generator: identical values.
// Requires mc <= 32 * THREADS / P.
template <int THREADS, int P, int MODE>
__device__ __forceinline__ void gen_tile_wide(double *__restrict__ tile, int mc, unsigned long long kfirst, const GenArgs &g, int r,
                                              GenScratch sc) {
    constexpr int G = THREADS / P, U = 4;
    constexpr bool lattice = (MODE & GEN_MC) == 0;
    const int tid = threadIdx.x, p = tid & (P - 1), gg = tid / P;
    const double *shift_r = lattice ? g.shifts + (size_t)r * g.s : nullptr;
    const unsigned long long k = kfirst + (unsigned long long)p;
    const double t = phi_to_unit(radical_phi((uint32_t)k));
    unsigned m2 = 0, m3 = 0;
    int it = 0;
    for (int jj = gg; jj < mc; jj += U * G, it += U) {
        double v[U];
        int cls[U];
#pragma unroll
        for (int e = 0; e < U; ++e) {
            const int j = jj + e * G < mc ? jj + e * G : jj;  // surplus slots repeat the first coordinate (never stored)
            double x;
            if (!lattice) x = mc_uniform(g.seed, k, (uint32_t)g.mkey, (uint32_t)j);
            else x = lattice_coord(t, __ldg(g.zd + j), __ldg(shift_r + j));
            v[e] = elem_first<MODE>(x, j, g, cls[e]);
        }
#pragma unroll
        for (int e = 0; e < U; ++e) {
            const bool ok = jj + e * G < mc;
            st_shared_if(tile + (ok ? jj + e * G : jj) * P + p, v[e], ok);
            m2 |= (unsigned)(ok && cls[e] == 2) << (it + e);
            m3 |= (unsigned)(ok && cls[e] == 3) << (it + e);
        }
    }
    if ((MODE & 3) == GEN_RAW) { __syncthreads(); return; }
    int n2, n3;
    const int total = mc * P;
    compact_classes<THREADS, unsigned>(m2, m3, total, sc, n2, n3);  // slot = it * THREADS + tid = (it * G + gg) * P + p
    for (int u = tid; u < n2 + n3; u += THREADS) {
        if (u < n2) {
            const int le = sc.queue[u];
            tile[le] = elem_finish<MODE, 2>(tile[le], le / P, g);
        } else {
            const int le = sc.queue[total - 1 - (u - n2)];
            tile[le] = elem_finish<MODE, 3>(tile[le], le / P, g);
        }
    }
    __syncthreads();
}

// moments of ONE partial (a shift whose samples all went through one CTA): the operations of finalize_value for G = 1
__device__ __forceinline__ void finalize_single(const Partial &a, double *out3) {
    const double mu = __dadd_rn(a.c, __ddiv_rn(a.s1, a.n));
    const double dc = __dsub_rn(a.c, mu);
    const double t = __dmul_rn(a.n, dc);
    const double s1 = __dadd_rn(0.0, __dadd_rn(a.s1, t));
    const double s2 = __dadd_rn(0.0, __fma_rn(dc, __fma_rn(2.0, a.s1, t), a.s2));
    out3[0] = a.n;
    out3[1] = __dadd_rn(mu, __ddiv_rn(s1, a.n));
    const double m2 = __dsub_rn(s2, __ddiv_rn(__dmul_rn(s1, s1), a.n));
    out3[2] = m2 < 0.0 ? 0.0 : m2;  // (not fmax: NaN stays NaN)
}

template <int MODE, bool CL>
__global__ void __launch_bounds__(LNS_THREADS, 1) lognormal1d_small_kernel(GenArgs g, Lognormal1dArgs la, FinArgs fa,
                                                                           double *__restrict__ samples, long long tiles_per_shift) {
    constexpr int BT = LNS_BT, THREADS = LNS_THREADS, W = THREADS / 32, PT = LNS_PT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int mpad = la.mpad, nks = mpad >> 2, ntiles = la.ntiles;
    const int M = la.N / 2, Mc = M / 2;
    // cluster geometry: CS CTAs per sample tile along x; rank 0 leads (flux sums, statistics)
    const unsigned CS = CL ? cluster_nctarank() : 1u, rank = CL ? cluster_ctarank() : 0u;
    const int tpr = (ntiles + (int)CS - 1) / (int)CS;                   // row tiles per rank
    const int t0 = (int)rank * tpr, t1 = min(ntiles, t0 + tpr);         // this CTA's edges start in the row tiles [t0, t1)
    const int ltiles = lns_local_tiles(ntiles, (int)CS);                // layout: identical in every CTA of the cluster
    const int nloc = max(0, min(ntiles, t1 + 1) - t0);                  // tiles it contracts: its own + the halo tile
    double *xi = reinterpret_cast<double *>(smem_raw);                  // [mpad][BT]
    double *F = xi + (size_t)mpad * BT;                                 // [ltiles * 8 rows r' = 2 i' + side, from row 8 t0][BT]
    double *amid = F + (size_t)ltiles * 8 * BT;                         // [BT]
    double *smid = amid + BT;                                           // [mpad]
    double2 *xch = reinterpret_cast<double2 *>(smid + mpad);            // [3][BT] (A, D): fine right, coarse left, coarse right
    Partial *mred = reinterpret_cast<Partial *>(xch + 3 * BT);          // [2][BT]
    GenScratch sc;
    sc.wtot = reinterpret_cast<unsigned *>(mred + 2 * BT);              // [W]
    sc.queue = reinterpret_cast<unsigned short *>(sc.wtot + W);
    const size_t qbytes = (((size_t)BT * ((g.m + 3) & ~3) * 2) + 15) & ~(size_t)15;
    const size_t used = (reinterpret_cast<unsigned char *>(sc.queue) - smem_raw) + qbytes;
    double *region = reinterpret_cast<double *>(smem_raw + ((used + 127) & ~(size_t)127));
    const int ptiles = ltiles < PT ? ltiles : PT;
    const int pdoubles = ptiles * nks * 32;                             // one panel
    double *Rf = region;                                                // [M edges][2 sides][BT]: edge ip-1 -> node ip (leader's copy is used)
    double *Rc = region + (size_t)M * 2 * BT;                           // [Mc edges][2 sides][BT]
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(region) + lns_region_bytes(mpad, ntiles, (int)CS));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, r = blockIdx.y;
    const int lr = lane >> 2, lk = lane & 3;
    const int cid = CL ? (int)(blockIdx.x / CS) : (int)blockIdx.x, ncl = CL ? (int)(gridDim.x / CS) : (int)gridDim.x;  // sample-tile slot
    const bool has_mid = t1 == ntiles && nloc > 0;                      // the rank whose last edge ends at the mid node
    LNS_T(0);
    if (tid == 0) { mbar_init(mbar, 1); mbar_init(mbar + 1, 1); mbar_fence_init(); }
    for (int e = tid; e < mpad; e += THREADS) smid[e] = __ldg(la.mid + e);
    for (int e = tid; e < (mpad - g.m) * BT; e += THREADS) xi[(size_t)g.m * BT + e] = 0.0;  // k padding rows
    // (no barrier here: the generator's own barriers order these writes before the contraction reads them, and thread 0 itself
    // issues the first bulk loads on the mbarriers it has just initialised)
    LNS_T(1);
    const int npanels = (nloc + PT - 1) / PT;
    unsigned cc = 0;  // panels consumed so far by this CTA (buffer cc & 1, mbarrier phase (cc >> 1) & 1)
    Partial pdq{0.0, 0.0, 0.0, 0.0}, pqf{0.0, 0.0, 0.0, 0.0};  // leader, threads tid < BT

    for (long long ti = cid; ti < tiles_per_shift; ti += ncl) {
        const unsigned long long i0 = (unsigned long long)ti * BT;
        const unsigned long long left = g.n - i0;
        const int Pt = left < (unsigned long long)BT ? (int)left : BT;
        const unsigned cc0 = cc;
        auto issue_panel = [&](int p) {  // thread 0: panel p of this CTA's row tiles -> buffer (cc0 + p) & 1
            const int live = min(PT, nloc - p * PT);
            const uint32_t bytes = (uint32_t)live * (uint32_t)nks * 32u * 8u;
            const unsigned b = (cc0 + (unsigned)p) & 1u;
            mbar_expect_tx(mbar + b, bytes);
            bulk_load_g2s(region + (size_t)b * pdoubles, la.basis + (size_t)(t0 + p * PT) * nks * 32, bytes, mbar + b);
        };
        // (the barrier that ended the previous tile's phase E ordered every read of the region before these loads)
        if (tid == 0) { if (npanels > 0) issue_panel(0); if (npanels > 1) issue_panel(1); }
        // A. inputs (every CTA of a cluster generates the same tile: a microsecond of redundant work instead of an exchange)
        gen_tile_wide<THREADS, BT, MODE>(xi, g.m, g.k0 + i0, g, r, sc);
        LNS_T(2);
        // B. contraction: warp w takes local row tile 8 p + w of panel p
        double gmid = 0.0;
        const double *Bp = xi + lk * BT + lr;
        const double *xm = xi + (lane & (BT - 1));
        for (int p = 0; p < npanels; ++p) {
            const unsigned b = cc & 1u;
            mbar_wait(mbar + b, (cc >> 1) & 1u);
            ++cc;
            const int lt = p * PT + warp;
            if (lt < nloc) {
                const double *Ap = region + (size_t)b * pdoubles + (size_t)warp * nks * 32 + lane;
                double c0 = 0.0, c1 = 0.0;
                if (p == 0 && warp == 0 && has_mid) {  // the mid-node chain rides along (lane % 8 = sample)
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
                *reinterpret_cast<double2 *>(F + ((size_t)lt * 8 + lr) * BT + 2 * lk) = make_double2(c0, c1);
            }
            __syncthreads();  // every warp has left buffer b: the panel after next may land there
            if (tid == 0 && p + 2 < npanels) issue_panel(p + 2);
        }
        LNS_T(3);
        // C. a = exp(G), in place; four independent evaluations per thread and iteration
        {
            const int total = nloc * 8 * BT;
            for (int e = tid; e < total; e += 4 * THREADS) {
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = F[min(e + u * THREADS, total - 1)];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = exp_det(v[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) st_shared_if(F + min(e + u * THREADS, total - 1), v[u], e + u * THREADS < total);
            }
            if (warp == 0 && lane < BT && has_mid) amid[lane] = exp_det(gmid);
        }
        // cluster: the leader's panels share the region the reciprocals are about to be stored into -- everybody has left phase B
        if (CL) cluster_sync_all(); else __syncthreads();
        LNS_T(4);
        // D. reciprocals of the interface coefficients, into the LEADER's region.  Fine grid: element e = (edge ip-1, side,
        // sample) reads nodes ip-1 and ip of its side: F[e] and F[e + 16] (rows 2 i' + side; F is local: minus 64 t0), the mid
        // node closing the last edge; this CTA owns the elements whose first node lies in its row tiles: [64 t0, 64 t1).
        // Coarse grid (even nodes): element (edge q-1, side, sample) reads F[32 (q-1) + low] and F[32 q + low]: [32 t0, 32 t1).
        {
            const int nf = M * 2 * BT;
            const int e0 = 64 * t0, e1 = min(nf, 64 * t1), fo = 64 * t0;
            for (int e = e0 + tid; e < e1; e += 4 * THREADS) {
                double a0[4], a1[4], rr[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ee = min(e + u * THREADS, e1 - 1);
                    a0[u] = F[ee - fo];
                    a1[u] = (ee >> 4) == M - 1 ? amid[ee & (BT - 1)] : F[ee + 16 - fo];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) rr[u] = rcp_kappa(a0[u], a1[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (CL) st_cluster_if(Rf + min(e + u * THREADS, e1 - 1), 0u, rr[u], e + u * THREADS < e1);
                    else st_shared_if(Rf + min(e + u * THREADS, e1 - 1), rr[u], e + u * THREADS < e1);
                }
            }
            if (la.level > 0) {
                const int nc = Mc * 2 * BT;
                const int c0 = 32 * t0, c1 = min(nc, 32 * t1);
                for (int e = c0 + tid; e < c1; e += 4 * THREADS) {
                    double a0[4], a1[4], rr[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int ee = min(e + u * THREADS, c1 - 1);
                        const int q1 = ee >> 4, low = ee & 15;
                        a0[u] = F[32 * q1 + low - fo];
                        a1[u] = q1 == Mc - 1 ? amid[ee & (BT - 1)] : F[32 * (q1 + 1) + low - fo];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) rr[u] = rcp_kappa(a0[u], a1[u]);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (CL) st_cluster_if(Rc + min(e + u * THREADS, c1 - 1), 0u, rr[u], e + u * THREADS < c1);
                        else st_shared_if(Rc + min(e + u * THREADS, c1 - 1), rr[u], e + u * THREADS < c1);
                    }
                }
            }
        }
        if (CL) cluster_sync_all(); else __syncthreads();
        LNS_T(5);
        // E. the flux sums, from the boundary towards the mid node: thread (grid, side, sample) = (tid / 16, (tid / 8) & 1, tid % 8)
        // of the leader.  Two dependent operations per edge; the next eight reciprocals are loaded while the current eight are added.
        if (rank == 0 && tid < 4 * BT) {
            const int coarse = tid >> 4;
            const int ne = coarse ? (la.level > 0 ? Mc : 0) : M;  // edges of this side; edge i (0-based) arrives at node i + 1
            const double *Rp = (coarse ? Rc : Rf) + (tid & 15);
            double A = 0.0, D = 0.0;
            const double w0 = (double)ne - 0.5;                   // weight of edge 0: (M - 1) + 1/2; edge i: w0 - i (exact)
            int i = 0;
            if (ne >= 8) {
                double cur[8], nxt[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) cur[u] = Rp[u * 16];
                double wg = w0;
                for (; i + 8 <= ne; i += 8) {
                    const int inext = i + 16 <= ne ? i + 8 : i;   // the last group reloads itself (unused)
#pragma unroll
                    for (int u = 0; u < 8; ++u) nxt[u] = Rp[(inext + u) * 16];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        A = __dadd_rn(A, cur[u]);
                        D = __fma_rn(wg - (double)u, cur[u], D);
                    }
                    wg -= 8.0;
#pragma unroll
                    for (int u = 0; u < 8; ++u) cur[u] = nxt[u];
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
        LNS_T(6);
    }
    if (CL && rank != 0) return;  // (the last cluster barrier is behind every store into the leader's shared memory)
    if (tid < BT) {
        mred[tid] = pdq;
        mred[BT + tid] = pqf;
    }
    __syncthreads();
    if (tid < 2) {
        Partial a{0.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < BT; ++i) a = partial_merge(a, mred[tid * BT + i]);
        if (ncl == 1) finalize_single(a, fa.moments + ((size_t)r * 2 + tid) * 3);  // this CTA saw every sample of the shift
        else fa.partials[((size_t)r * 2 + tid) * ncl + cid] = a;
    }
    LNS_T(7);
    cta_finalize(fa, 2, ncl == 1, ncl);
    LNS_T(8);
}

}  // namespace mle
