// mle_kernels2d.cuh -- register-resident frontal solver for the 2-D lognormal model (configs C3 / C4).
//
// Same arithmetic as frontal_half_dev / frontal_mid_dev in mle_kernels.cuh (and as oracle/mle_oracle.c:frontal_mid): every
// window entry receives fma(-l_q, v_r, M[q][r]) in pivot order with l_q = v_q * (1/d), so the result is bit-identical for
// any thread mapping.  What changes is WHERE the sliding window lives:
//
//   * row q of the window belongs to LPR adjacent lanes (row group q % NW); each lane keeps C = NW / LPR consecutive window
//     columns in REGISTERS, indexed by position = column - pivot.  One elimination step is
//         reg[pos - 1] = fma(-l_q, v[pos - 1], reg[pos])          (static register indices: the shift by one position IS the
//                                                                   advance of the window)
//     with l_q from the lane's own position-0 entry and the column vector v (every row's position-0 entry) broadcast through
//     shared memory: C FMAs against C/2 16-byte LDS per lane and step, ONE barrier per pivot (double-buffered v).
//   * the reciprocal of the next pivot is computed by its row right after that row's update, off the critical path.
//   * stencil coefficients are computed a grid line at a time by all lanes (lane i: unknown i+1 of the line) and handed to
//     the entering row through the same shared-memory exchange.
//   * both half sweeps run in the same registers (the upper half's centre block is parked in a second register set); the
//     two-sided centre-line elimination continues in place: top-down with the same step, bottom-up with the mirrored shift.
//
// Kernel: one CTA iteration = S samples.  A: inputs (gen_tile_owned) -> xi[m][S] in shared memory.  B: field = exp(Phi xi)
// with mma.sync.m8n8k4.f64 over node tiles (basis in HBM/L2 in A-fragment order) -> F[node][S] (shared memory, or an
// L2-resident global scratch for the large grids).  C: CS = THREADS / (NW*LPR) samples are solved concurrently, S / CS
// rounds; fine solve + the coarse solves of diff(index) on strided views of F.  D: per-slot pivoted sums, fixed-order merge.
#pragma once
#include <cuda/std/type_traits>

#include "mle_kernels.cuh"

// layout knobs (compile with -DMLE_ROWS_FAT=0 / -DMLE_ROWS_RPL=1 for the A/B measurements quoted in DESIGN.md 3.4)
#ifndef MLE_ROWS_FAT
#define MLE_ROWS_FAT 1  // 1: 64 window columns per lane from 64 cells up (255 registers, one resident CTA); 0: 32 (two CTAs)
#endif
#ifndef MLE_ROWS_RPL
#define MLE_ROWS_RPL 2  // window rows per lane for the 16-cell window (1: one row per lane)
#endif

namespace mle {

struct Lognormal2dRowsArgs {
    const double *basis;  // fragment order [node tile][k step][lane]
    int Nx, Ny, ndims, lx, ly;
    int mpad;             // KL terms padded to a multiple of 4
    int ntiles;           // node tiles of 8 rows
    int use_scratch;      // F in global scratch instead of shared memory
    double *scratch;      // [CTAs in grid][ntiles*8*LDF]
    double *park;         // [CTAs in grid][ROWS_PARK][THREADS]: the upper half's centre block of every thread
    int spt;              // samples per tile, a multiple of the CS samples solved concurrently, <= S: small batches are spread
                          // over more CTAs instead of filling the S slots of a few (a tile's surplus slots are never solved)
};

// node values a[((j*sty)*lda + i*stx) * ns] of an Nx x Ny-cell grid (optionally y-mirrored); ns = distance between nodes
struct FieldView {
    const double *a;
    int ns, lda, stx, sty, Nx, Ny, mirror;
    double wx, wy;
    __device__ __forceinline__ double at(int i, int j) const {
        const int jj = mirror ? Ny - j : j;
        return a[(size_t)((jj * sty) * lda + i * stx) * ns];
    }
    __device__ __forceinline__ double kx(int i, int j) const { return __dmul_rn(wx, __dmul_rn(0.5, __dadd_rn(at(i, j), at(i + 1, j)))); }
    __device__ __forceinline__ double ky(int i, int j) const { return __dmul_rn(wy, __dmul_rn(0.5, __dadd_rn(at(i, j), at(i, j + 1)))); }
    __device__ __forceinline__ double diag(int i, int j) const {
        return __dadd_rn(__dadd_rn(__dadd_rn(kx(i - 1, j), kx(i, j)), ky(i, j - 1)), ky(i, j));
    }
};

constexpr int ROWS_PARK = 64;  // window columns a lane can hold

template <int GT>
__device__ __forceinline__ void rows_sync(int barid) {
    if (GT <= 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(barid), "n"(GT) : "memory");
}

// 2 v buffers + 8 for the pivot exchange + (rows spanning LPR = w/32 lanes, from 64 cells up) 2 x (LPR-1) x w for the
// window entries that cross a chunk boundary
__host__ __device__ constexpr int rows_sh_doubles(int w) { return 2 * w + 8 + (w > 32 ? 2 * (w / 32 - 1) * w : 0); }

// u at the centre node (Nx/2, Ny/2) of the NW x Ny-cell grid.  Called by all GT threads of the sample's group
// (tg = 0 .. GT-1, GT = 32 stands for "some lanes of one warp"); lanes with tg >= NW*LPR idle (coarse solves run on the fine
// solve's group).  sh: rows_sh_doubles(NW) doubles of group-private shared memory, 16-byte aligned.  park: this thread's
// column of a global scratch [C][pstride] that holds the upper half's centre block while the lower half is eliminated.
template <int NW, int LPR, int GT>
__device__ __noinline__ double frontal_mid_rows(FieldView f, double *sh, double *park, int pstride, int tg, int barid) {
    constexpr int C = NW / LPR, n = NW - 1, tc = NW / 2 - 1, L = NW / 2 - 1;  // L = n - 1 - tc: position of the last pivot column
    static_assert(C >= 2 && C <= ROWS_PARK && C * LPR == NW && (C % 2) == 0, "NW / LPR even, >= 2, <= ROWS_PARK");
    // row group q, chunk s: with several lanes per row, chunk s of all rows sits in consecutive threads, so the lanes of a
    // warp read the same v entries (broadcast) and the entry that crosses a chunk boundary travels through shared memory
    const int q = LPR == 1 ? tg : tg % NW, s = LPR == 1 ? 0 : tg / NW;
    const bool lane_on = tg < NW * LPR;
    constexpr int VS = NW;
    auto vi = [](int slot) { return slot; };
    double *vs = sh;            // [2][VS]: column vector v of the current pivot (slot = row - p - 1)
    double *piv = sh + 2 * VS;  // [2][4]: 1/d, g_p, -kx and diagonal of the entering unknown
    double *xch = piv + 8;      // [2][LPR-1][NW]: position s*C of row q before the step, for the lane that holds chunk s-1
    const int c = f.Ny / 2, P = (c - 1) * n, last = P + n - 1;
    double reg[C], g = 0.0, g2 = 0.0, rcp = 0.0;
    int row = -1, step = 0;  // step: running parity of the exchange buffers
#pragma unroll
    for (int j = 0; j < C; ++j) reg[j] = 0.0;

    // One shift-left elimination step on pivot row p, written branch-free: rows (p, ..] get
    // reg[pos-1] = fma(-l, v[pos-1], reg[pos]); every other lane runs the same FMAs with l = 0 (its registers are dead).
    // ENTER: unknown e = p + n joins the window in row group `se` (the one freed by pivot p - 1) with coefficients published
    // by the group that holds them (ie - 1); its registers are cleared by predicated moves inside the same basic block.
    auto eliminate = [&](auto enter_tag, int p, int e, int se, int ie, double cdg, double ckx, double cky, double gv) {
        constexpr bool ENTER = decltype(enter_tag)::value;
        const int buf = step & 1;
        ++step;
        double *vb = vs + buf * VS, *pb = piv + buf * 4;
        if (lane_on && s == 0) {
            if (row > p) vb[vi(row - p - 1)] = reg[0];
            if (row == p) *reinterpret_cast<double2 *>(pb) = make_double2(rcp, g);
            if (ENTER && q == ie - 1) { vb[vi(n - 1)] = cky; *reinterpret_cast<double2 *>(pb + 2) = make_double2(ckx, cdg); }
        }
        if (LPR > 1 && lane_on && s > 0) xch[(buf * (LPR - 1) + s - 1) * NW + q] = reg[0];
        rows_sync<GT>(barid);
        const double2 pg = *reinterpret_cast<const double2 *>(pb);  // 1/d, g_p
        double top = 0.0;                                            // what position n holds for this row
        bool me = false;
        if (ENTER) {
            const double2 ent = *reinterpret_cast<const double2 *>(pb + 2);  // -kx, diagonal of the entering unknown
            me = lane_on && q == se;
            if (me) {
                row = e;
                g = gv;
#pragma unroll
                for (int j = 0; j < C; ++j) reg[j] = 0.0;
                if (s == LPR - 1) reg[C - 2] = ent.x;
                if (s == 0) reg[0] = vb[vi(n - 1)];
            }
            top = me ? ent.y : (row == e - 1 ? ent.x : 0.0);
        }
        if (s == LPR - 1) reg[C - 1] = top;
        double nb = 0.0;  // position (s+1)*C lives in the lane that holds the next chunk of the row
        if (LPR > 1 && lane_on && s < LPR - 1 && !me) nb = xch[(buf * (LPR - 1) + s) * NW + q];
        const bool active = row > p;
        const double v = LPR == 1 ? reg[0] : vb[active ? vi(row - p - 1) : 0];
        const double l = active ? __dmul_rn(v, pg.x) : 0.0;
        g = __fma_rn(-l, pg.y, g);
        // v is read in groups of PF 16-byte loads issued ahead of the FMAs that consume them (volatile: the order is kept,
        // so PF loads are in flight instead of one)
        constexpr int PF = C >= 16 ? 4 : C / 2, NG = (C / 2) / PF;
        const uint32_t va = smem_u32(vb + vi(s * C));
        double w[2][2 * PF];
        auto load_group = [&](int gi, double *dst) {
#pragma unroll
            for (int t = 0; t < PF; ++t)
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(dst[2 * t]), "=d"(dst[2 * t + 1]) : "r"(va + 16 * (gi * PF + t)));
        };
        load_group(0, w[0]);
#pragma unroll
        for (int gi = 0; gi < NG; ++gi) {
            if (gi + 1 < NG) load_group(gi + 1, w[(gi + 1) & 1]);
            const double *wg = w[gi & 1];
#pragma unroll
            for (int t = 0; t < 2 * PF; ++t) {
                const int i = gi * 2 * PF + t;  // reg[i] = fma(-l, v[i], reg[i + 1])
                if (i + 1 < C) reg[i] = __fma_rn(-l, wg[t], reg[i + 1]);
                else if (LPR > 1 && s < LPR - 1) reg[i] = __fma_rn(-l, wg[t], nb);
                if (i == 0) {
                    const double r1 = __ddiv_rn(1.0, reg[0]);  // the next pivot's reciprocal: overlaps the FMAs below
                    rcp = (s == 0 && row == p + 1) ? r1 : rcp;
                }
            }
        }
        row = row == p ? -1 : row;
    };
    using with_enter = cuda::std::true_type;
    using no_enter = cuda::std::false_type;

    for (int half = 0; half < 2; ++half) {
        f.mirror = half == 0;  // upper half first (y-mirrored field)
        // line 1: unknowns 0 .. n-1 are rows 0 .. n-1
        const bool own1 = !(c == 1 && f.mirror);
        row = (lane_on && q < n) ? q : -1;
        {
            double dg = 0.0, kl = 0.0, kr = 0.0;
            if (row >= 0 && own1) {
                dg = f.diag(q + 1, 1);
                if (q > 0) kl = -f.kx(q, 1);
                if (q + 1 < n) kr = -f.kx(q + 1, 1);
            }
#pragma unroll
            for (int j = 0; j < C; ++j) {
                const int pos = s * C + j;
                reg[j] = pos == q ? dg : pos == q - 1 ? kl : pos == q + 1 ? kr : 0.0;
            }
            g = (row >= 0 && own1) ? 1.0 : 0.0;
            if (s == 0 && row == 0) rcp = __ddiv_rn(1.0, reg[0]);
        }
        // lines 2 .. c enter one unknown per pivot: line je during pivots (je-2)*n .. (je-1)*n - 1
        int p = 0, se = n;
        for (int je = 2; je <= c; ++je) {
            const bool own = !(je == c && f.mirror);  // the centre line's own entries are loaded once (lower half)
            double cdg = 0.0, ckx = 0.0, cky = 0.0;   // coefficients of unknown (q+1, je), held by row group q
            if (lane_on && q < n) {
                if (own) { cdg = f.diag(q + 1, je); if (q > 0) ckx = -f.kx(q, je); }
                cky = -f.ky(q + 1, je - 1);
            }
            const double gv = own ? 1.0 : 0.0;
            for (int ie = 1; ie <= n; ++ie, ++p) {
                eliminate(with_enter{}, p, p + n, se, ie, cdg, ckx, cky, gv);
                if (++se == NW) se = 0;
            }
        }
        if (half == 0) {
#pragma unroll
            for (int j = 0; j < C; ++j) park[(size_t)j * pstride] = reg[j];
            g2 = g;
            rows_sync<GT>(barid);  // the exchange buffers are re-initialised by the second half
        }
    }
    // centre line: rows P .. last hold S (lower) in reg and S2 (upper) in the park, columns P .. last at positions 0 .. n-1
#pragma unroll
    for (int j = 0; j < C; ++j) reg[j] = __dadd_rn(reg[j], park[(size_t)j * pstride]);
    g = __dadd_rn(g, g2);
    if (s == 0 && row == P) rcp = __ddiv_rn(1.0, reg[0]);
    // top-down: pivots P .. P+tc-1 with the same shift-left step
    for (int t = 0; t < tc; ++t) eliminate(no_enter{}, P + t, 0, 0, 0, 0.0, 0.0, 0.0, 0.0);
    // bottom-up: pivots last .. P+tc+1; the pivot column stays at position L, everything shifts right by one per step
    constexpr int sL = L / C, jL = L % C;
    const int rtc = P + tc;
    if (s == sL && row == last && last > rtc) rcp = __ddiv_rn(1.0, reg[jL]);
    for (int k = 0; k < n - 1 - tc; ++k) {
        const int rowt = last - k;
        const int buf = step & 1;
        ++step;
        double *vb = vs + buf * VS, *pb = piv + buf * 4;
        if (lane_on && s == sL) {
            if (row >= rtc && row < rowt) vb[vi(row - rtc + k)] = reg[jL];
            else if (row == rowt) { pb[0] = rcp; pb[1] = g; }
        }
        if (LPR > 1 && lane_on && s < LPR - 1) xch[(buf * (LPR - 1) + s) * NW + q] = reg[C - 1];
        rows_sync<GT>(barid);
        double pv = 0.0;  // position s*C - 1 lives in the lane that holds the previous chunk of the row
        if (LPR > 1 && lane_on && s > 0) pv = xch[(buf * (LPR - 1) + s - 1) * NW + q];
        if (row >= rtc && row < rowt) {
            const double l = __dmul_rn(vb[vi(row - rtc + k)], pb[0]);
            g = __fma_rn(-l, pb[1], g);
            // new[pos + 1] = fma(-l, v[pos], old[pos]) for pos = 0 .. L-1, descending so nothing is overwritten early
#pragma unroll
            for (int j = C - 1; j >= 1; --j) {
                const int pos = s * C + j - 1;  // source position
                if (pos < L) reg[j] = __fma_rn(-l, vb[vi(pos)], reg[j - 1]);
            }
            if (LPR > 1 && s > 0 && s * C - 1 < L) reg[0] = __fma_rn(-l, vb[vi(s * C - 1)], pv);
            if (s == sL && row == rowt - 1 && rowt - 1 > rtc) rcp = __ddiv_rn(1.0, reg[jL]);
        } else if (row == rowt) {
            row = -1;
        }
    }
    // u = gs[tc] / S[tc][tc]
    {
        const int buf = step & 1;
        ++step;
        if (lane_on && s == sL && row == rtc) piv[buf * 4] = __ddiv_rn(g, reg[jL]);
        rows_sync<GT>(barid);
        const double u = piv[buf * 4];
        rows_sync<GT>(barid);  // the buffers are reused by the next solve
        return u;
    }
}

// The same solver with RPL window rows per lane (NW <= 32): a sample occupies NW / RPL lanes of one warp, so the v loads and
// all of the step's control flow are shared by RPL rows (2x the FMAs per issued instruction of overhead).  Row r of the
// window lives in lane (r % NW) % TL, register set (r % NW) / TL, TL = NW / RPL.  Same arithmetic, same order.
template <int NW, int RPL>
__device__ __noinline__ double frontal_mid_rpl(FieldView f, double *sh, double *park, int pstride, int tg) {
    constexpr int C = NW, n = NW - 1, tc = NW / 2 - 1, L = NW / 2 - 1, TL = NW / RPL;
    static_assert(NW >= 4 && NW <= 32 && TL * RPL == NW && RPL * C <= ROWS_PARK, "window of at most 32 rows, RPL rows per lane");
    const bool lane_on = tg < TL;
    double *vs = sh;            // [2][NW]
    double *piv = sh + 2 * NW;  // [2][4]
    const int c = f.Ny / 2, P = (c - 1) * n, last = P + n - 1;
    double reg[RPL][C], g[RPL], g2[RPL], rcp = 0.0;
    int row[RPL], step = 0;
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
        g[k] = 0.0; g2[k] = 0.0; row[k] = -1;
#pragma unroll
        for (int j = 0; j < C; ++j) reg[k][j] = 0.0;
    }

    auto eliminate = [&](auto enter_tag, int p, int e, int se, int ie, const double (&cdg)[RPL], const double (&ckx)[RPL],
                         const double (&cky)[RPL], double gv) {
        constexpr bool ENTER = decltype(enter_tag)::value;
        const int buf = step & 1;
        ++step;
        double *vb = vs + buf * NW, *pb = piv + buf * 4;
        if (lane_on) {
#pragma unroll
            for (int k = 0; k < RPL; ++k) {
                if (row[k] > p) vb[row[k] - p - 1] = reg[k][0];
                if (row[k] == p) *reinterpret_cast<double2 *>(pb) = make_double2(rcp, g[k]);
                if (ENTER && k * TL + tg == ie - 1) { vb[n - 1] = cky[k]; *reinterpret_cast<double2 *>(pb + 2) = make_double2(ckx[k], cdg[k]); }
            }
        }
        __syncwarp();
        const double2 pg = *reinterpret_cast<const double2 *>(pb);  // 1/d, g_p
        double l[RPL];
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
            double top = 0.0;
            if (ENTER) {
                const double2 ent = *reinterpret_cast<const double2 *>(pb + 2);  // -kx, diagonal of the entering unknown
                const bool me = lane_on && k * TL + tg == se;
                if (me) {
                    row[k] = e;
                    g[k] = gv;
#pragma unroll
                    for (int j = 0; j < C; ++j) reg[k][j] = 0.0;
                    reg[k][C - 2] = ent.x;
                    reg[k][0] = vb[n - 1];
                }
                top = me ? ent.y : (row[k] == e - 1 ? ent.x : 0.0);
            }
            reg[k][C - 1] = top;
            l[k] = row[k] > p ? __dmul_rn(reg[k][0], pg.x) : 0.0;
            g[k] = __fma_rn(-l[k], pg.y, g[k]);
        }
        constexpr int PF = C >= 16 ? 4 : C / 2, NG = (C / 2) / PF;
        const uint32_t va = smem_u32(vb);
        double w[2][2 * PF];
        auto load_group = [&](int gi, double *dst) {
#pragma unroll
            for (int t = 0; t < PF; ++t)
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(dst[2 * t]), "=d"(dst[2 * t + 1]) : "r"(va + 16 * (gi * PF + t)));
        };
        load_group(0, w[0]);
#pragma unroll
        for (int gi = 0; gi < NG; ++gi) {
            if (gi + 1 < NG) load_group(gi + 1, w[(gi + 1) & 1]);
            const double *wg = w[gi & 1];
#pragma unroll
            for (int t = 0; t < 2 * PF; ++t) {
                const int i = gi * 2 * PF + t;  // reg[i] = fma(-l, v[i], reg[i + 1])
                if (i + 1 < C) {
#pragma unroll
                    for (int k = 0; k < RPL; ++k) reg[k][i] = __fma_rn(-l[k], wg[t], reg[k][i + 1]);
                }
                if (i == 0) {  // the next pivot's reciprocal: overlaps the FMAs below
                    double x = reg[0][0];
                    bool mine = row[0] == p + 1;
#pragma unroll
                    for (int k = 1; k < RPL; ++k) { if (row[k] == p + 1) { x = reg[k][0]; mine = true; } }
                    const double r1 = __ddiv_rn(1.0, x);
                    rcp = mine ? r1 : rcp;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < RPL; ++k) row[k] = row[k] == p ? -1 : row[k];
    };
    using with_enter = cuda::std::true_type;
    using no_enter = cuda::std::false_type;
    const double zero3[RPL] = {};

    for (int half = 0; half < 2; ++half) {
        f.mirror = half == 0;  // upper half first (y-mirrored field)
        const bool own1 = !(c == 1 && f.mirror);
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
            const int r = k * TL + tg;
            row[k] = (lane_on && r < n) ? r : -1;
            double dg = 0.0, kl = 0.0, kr = 0.0;
            if (row[k] >= 0 && own1) {
                dg = f.diag(r + 1, 1);
                if (r > 0) kl = -f.kx(r, 1);
                if (r + 1 < n) kr = -f.kx(r + 1, 1);
            }
#pragma unroll
            for (int j = 0; j < C; ++j) reg[k][j] = j == r ? dg : j == r - 1 ? kl : j == r + 1 ? kr : 0.0;
            g[k] = (row[k] >= 0 && own1) ? 1.0 : 0.0;
            if (row[k] == 0) rcp = __ddiv_rn(1.0, reg[k][0]);
        }
        int p = 0, se = n;
        for (int je = 2; je <= c; ++je) {
            const bool own = !(je == c && f.mirror);
            double cdg[RPL], ckx[RPL], cky[RPL];  // coefficients of unknown (q+1, je), q = k*TL + tg
#pragma unroll
            for (int k = 0; k < RPL; ++k) {
                const int q = k * TL + tg;
                cdg[k] = 0.0; ckx[k] = 0.0; cky[k] = 0.0;
                if (lane_on && q < n) {
                    if (own) { cdg[k] = f.diag(q + 1, je); if (q > 0) ckx[k] = -f.kx(q, je); }
                    cky[k] = -f.ky(q + 1, je - 1);
                }
            }
            const double gv = own ? 1.0 : 0.0;
            for (int ie = 1; ie <= n; ++ie, ++p) {
                eliminate(with_enter{}, p, p + n, se, ie, cdg, ckx, cky, gv);
                if (++se == NW) se = 0;
            }
        }
        if (half == 0) {
#pragma unroll
            for (int k = 0; k < RPL; ++k) {
#pragma unroll
                for (int j = 0; j < C; ++j) park[(size_t)(k * C + j) * pstride] = reg[k][j];
                g2[k] = g[k];
            }
            __syncwarp();
        }
    }
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
#pragma unroll
        for (int j = 0; j < C; ++j) reg[k][j] = __dadd_rn(reg[k][j], park[(size_t)(k * C + j) * pstride]);
        g[k] = __dadd_rn(g[k], g2[k]);
        if (row[k] == P) rcp = __ddiv_rn(1.0, reg[k][0]);
    }
    for (int t = 0; t < tc; ++t) eliminate(no_enter{}, P + t, 0, 0, 0, zero3, zero3, zero3, 0.0);
    const int rtc = P + tc;
#pragma unroll
    for (int k = 0; k < RPL; ++k)
        if (row[k] == last && last > rtc) rcp = __ddiv_rn(1.0, reg[k][L]);
    for (int kk = 0; kk < n - 1 - tc; ++kk) {
        const int rowt = last - kk;
        const int buf = step & 1;
        ++step;
        double *vb = vs + buf * NW, *pb = piv + buf * 4;
        if (lane_on) {
#pragma unroll
            for (int k = 0; k < RPL; ++k) {
                if (row[k] >= rtc && row[k] < rowt) vb[row[k] - rtc + kk] = reg[k][L];
                else if (row[k] == rowt) { pb[0] = rcp; pb[1] = g[k]; }
            }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
            if (row[k] >= rtc && row[k] < rowt) {
                const double lk = __dmul_rn(vb[row[k] - rtc + kk], pb[0]);
                g[k] = __fma_rn(-lk, pb[1], g[k]);
#pragma unroll
                for (int j = C - 1; j >= 1; --j)
                    if (j - 1 < L) reg[k][j] = __fma_rn(-lk, vb[j - 1], reg[k][j - 1]);
                if (row[k] == rowt - 1 && rowt - 1 > rtc) rcp = __ddiv_rn(1.0, reg[k][L]);
            } else if (row[k] == rowt) {
                row[k] = -1;
            }
        }
    }
    {
        const int buf = step & 1;
        ++step;
#pragma unroll
        for (int k = 0; k < RPL; ++k)
            if (lane_on && row[k] == rtc) piv[buf * 4] = __ddiv_rn(g[k], reg[k][L]);
        __syncwarp();
        const double u = piv[buf * 4];
        __syncwarp();
        return u;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Blocked frontal solver on the FP64 tensor instruction (windows of 64 and 128 cells).
//
// Same arithmetic as the oracle's frontal_mid (oracle/mle_oracle.c): every window entry receives fma(-l_q, v_r, M[q][r]) once per
// pivot, in pivot order.  On B200 one mma.sync.m8n8k4.f64 IS four such fused multiply-adds in ascending k (tools/dmma_order_test.cu),
// so FOUR CONSECUTIVE PIVOTS can be applied to an 8x8 tile of the window by one DMMA -- provided the multipliers l^(j)_q and the
// column values v^(j)_r of the four pivots are known, and those depend on each other.  Hence two phases per block of four pivots
// p .. p+3 (window kept in the oracle's circular slot coordinates, slot = unknown % W, as DMMA accumulator tiles in registers):
//   1. PANEL.  The four pivot columns are copied from the register tiles to shared memory; thread q (one per window row) takes its
//      four entries X[q][0..3] and the four pivots are factorised one after the other exactly as the oracle does it for these
//      entries: step j publishes the entries of the pivot rows, every active row forms l = v * (1/d) and updates its entries of
//      the later pivot columns and its right-hand side g.  Cost: four small barriers among W threads, O(W) work.
//   2. TRAILING UPDATE.  A[q][j] = -l^(j)_q and B[j][r] = v^(j)_r go to shared memory and every 8x8 tile takes ONE DMMA: 4 W^2
//      fused multiply-adds in W^2/64 tensor instructions (the register-window solver above spends ~6 issue slots per DFMA here).
// Unknown e = p+j+n enters the window at step j in the slot pivot p+j-1 has just left (the oracle's LOAD_UNKNOWN); a row or column
// that changes owner inside a block simply carries zeros in the operands of the steps before the newcomer arrived (fma(-0, v, x) =
// x exactly), and its register row / column is re-initialised to the newcomer's stencil entries before the DMMA.  The newcomer's
// coupling to its pivot (-ky) only ever lives in the panel.  The dense two-sided elimination along the centre line uses the same
// step with one pivot per block (126 of ~16000 pivots at 128 cells).  Bit-identical to the oracle (tests/test_gpu_batch.py).
// ---------------------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int dmma_sh_doubles(int w) { return 12 * w + 6 * w + 32 + 8 + 4 * 64; }

// warp-uniform dispatch of a runtime tile index to a compile-time one (the accumulators must keep static register indices)
#define MLE_DISPATCH4(idx, N, CALL)                                  \
    switch (idx) {                                                   \
        case 0: { constexpr int I_ = 0; CALL; } break;               \
        case 1: if constexpr ((N) > 1) { constexpr int I_ = 1; CALL; } break; \
        case 2: if constexpr ((N) > 2) { constexpr int I_ = 2; CALL; } break; \
        case 3: if constexpr ((N) > 3) { constexpr int I_ = 3; CALL; } break; \
        default: break;                                              \
    }

// correctly rounded 1/d for d well inside the normal range: ddiv_normal(1.0, d) without its exact product 1.0 * r
__device__ __forceinline__ double drcp_normal(double den) {
    double r = rcp_seed(den);
    double e = __fma_rn(-den, r, 1.0);
    e = __fma_rn(e, e, e);
    r = __fma_rn(r, e, r);
    e = __fma_rn(-den, r, 1.0);
    r = __fma_rn(r, e, r);
    const double rem = __fma_rn(-den, r, 1.0);
    return __fma_rn(r, rem, r);
}

template <int NW, int TW>
__device__ __noinline__ double frontal_mid_dmma(FieldView f, double *sh, double *park, int pstride, int tg, int barid) {
    constexpr int W = NW, n = NW - 1, NT = NW / 8, tc = NW / 2 - 1;
    constexpr int WR = TW >= 16 ? 4 : 2, WC = TW / WR, TR = NT / WR, TC = NT / WC;
    static_assert(WR * WC == TW && TR * WR == NT && TC * WC == NT && TR >= 1 && TR <= 4 && TC >= 1 && TC <= 4, "warp grid");
    static_assert(W % 32 == 0, "the panel barrier counts whole warps");
    const int lane = tg & 31, warp = tg >> 5, wr = warp / WC, wc = warp % WC, lr = lane >> 2, lk = lane & 3;
    double *X = sh;               // [W][4] panel as extracted from the register tiles
    double *LA = sh + 4 * W;      // [W][4] -l^(j)_q   (A operand)
    double *VB = sh + 8 * W;      // [W][4] v^(j)_r    (B operand, [slot][k])
    double *coef = sh + 12 * W;   // [2][3][W] diagonal, -kx, -ky of the unknowns of two grid lines (line L in buffer L & 1)
    double *bc = coef + 6 * W;    // [4][8] step j: entries of the pivot rows in pivot column j, [4] = g of the pivot
    double *misc = bc + 32;
    double *patch = misc + 8;     // [2][2][8][8] the tiles that hold the stencil entries of a block's newcomers, on their way through shared memory
    const int c = f.Ny / 2, P = (c - 1) * n, last = P + n - 1;
    double acc[TR][TC][2];
    const bool pt = tg < W;  // panel thread of slot q = tg
    const bool pt_warp = warp < W / 32;
    int uk = -1;             // unknown held by this slot (-1: none)
    double g = 0.0, g2 = 0.0;
    auto gsync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(barid), "n"(TW * 32) : "memory"); };

    // stencil coefficients of grid line L (unknown (i, L), i = tg + 1) -> coef buffer L & 1
    auto compute_line = [&](int L) {
        if (tg < n) {
            const bool own = !(L == c && f.mirror);  // the centre line's own entries are loaded once (lower half)
            const int i = tg + 1;
            double *cb = coef + (L & 1) * 3 * W;
            cb[tg] = own ? f.diag(i, L) : 0.0;
            cb[W + tg] = (own && i > 1) ? -f.kx(i - 1, L) : 0.0;
            cb[2 * W + tg] = L > 1 ? -f.ky(i, L - 1) : 0.0;
        }
    };
    // nb (<= 4) pivots at consecutive slots sp .. sp+nb-1 (inside one 8-slot tile), unknowns pv0, pv0 +- 1, ...
#ifdef MLE_DMMA_TIMING
    long long tacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#define MLE_TICK(i) { const long long now_ = clock64(); tacc[i] += now_ - tlast; tlast = now_; }
#else
#define MLE_TICK(i)
#endif
    // (je0, i00): grid line and 0-based position in the line of the first newcomer pv0 + n (half sweeps only)
    auto block_step = [&](int sp, int pv0, int nb, bool enter, bool desc, int lo, int je0, int i00) {
        MLE_TICK(5)
        // newcomers of this block (warp-uniform, computed once: unknown, right-hand side, coupling -ky to its pivot);
        // a thread whose slot one of them takes adopts it with selects -- a per-thread branch here would serialise the warp
        int ne[4];
        double ngv[4], nky[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = pv0 + j + n;
            const bool ok = enter && j < nb && e <= last;
            const bool wrap = i00 + j >= n;  // the newcomer starts the next grid line
            const int je = je0 + (wrap ? 1 : 0), i0 = i00 + j - (wrap ? n : 0);
            ne[j] = ok ? e : -1;
            ngv[j] = (je == c && f.mirror) ? 0.0 : 1.0;
            nky[j] = ok ? coef[(je & 1) * 3 * W + 2 * W + i0] : 0.0;
        }
        // (1) the pivot columns -> X
        {
            const int cs = sp & 7, b0 = (sp >> 3) - wc * TC;
            if (b0 >= 0 && b0 < TC) {
                MLE_DISPATCH4(b0, TC, {
_Pragma("unroll") for (int a = 0; a < TR; ++a) {
_Pragma("unroll") for (int e = 0; e < 2; ++e) {
                            const int col = 2 * lk + e - cs;
                            if (col >= 0 && col < nb) X[((wr * TR + a) * 8 + lr) * 4 + col] = acc[a][I_][e];
                        }
                    }
                })
            }
        }
        const int rel = (tg - sp + W) & (W - 1);  // 0 .. nb-1: the slot of pivot `rel`;  W-1: the slot the first newcomer takes
        if (pt && rel < nb) bc[rel] = g;          // right-hand sides of the pivot rows
        MLE_TICK(0)
        gsync();
        MLE_TICK(1)
        // (2) panel factorisation, without a barrier: EVERY panel thread factorises the 4x4 block of the pivot rows for itself
        // (the same few operations in the same order, so all threads hold identical values) and then takes its own row
        // through the four steps
        // (3) register rows / columns of the newcomers (slots s0, s0+1, ... mod W): zero, except the stencil entries -- the
        // diagonal, and -kx towards the left neighbour (the slot before).  Branch-free selects (per-lane branches cost a
        // reconvergence point per element); warps whose tile band holds no newcomer skip it, and the warp that holds the
        // diagonal entries does this work in the shadow of the other warps' DMMAs.
        auto fresh_init = [&]() {
        if (enter) {
            const int nv = max(0, min(nb, last - (pv0 + n) + 1));  // newcomers of this block (the unknowns enter in order)
            if (nv > 0) {
                const int s0 = (sp - 1 + W) & (W - 1);
                auto in_band = [&](int base, int len) {
                    const int dd = (s0 - base + W) & (W - 1);
                    return dd < len || dd + nv > W;
                };
                const bool rband = in_band(wr * TR * 8, TR * 8), cband = in_band(wc * TC * 8, TC * 8);  // warp-uniform
                if (rband || cband) {
                    int jr[TR], jc[TC][2];
#pragma unroll
                    for (int a = 0; a < TR; ++a) {
                        const int dd = ((wr * TR + a) * 8 + lr - s0 + W) & (W - 1);
                        jr[a] = dd < nv ? dd : -1;
                    }
#pragma unroll
                    for (int b2 = 0; b2 < TC; ++b2)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int dd = ((wc * TC + b2) * 8 + 2 * lk + e - s0 + W) & (W - 1);
                            jc[b2][e] = dd < nv ? dd : -1;
                        }
                    if (rband) {
#pragma unroll
                        for (int a = 0; a < TR; ++a)
#pragma unroll
                            for (int b2 = 0; b2 < TC; ++b2) {
                                acc[a][b2][0] = jr[a] >= 0 ? 0.0 : acc[a][b2][0];
                                acc[a][b2][1] = jr[a] >= 0 ? 0.0 : acc[a][b2][1];
                            }
                    }
                    if (cband) {
#pragma unroll
                        for (int a = 0; a < TR; ++a)
#pragma unroll
                            for (int b2 = 0; b2 < TC; ++b2) {
                                acc[a][b2][0] = jc[b2][0] >= 0 ? 0.0 : acc[a][b2][0];
                                acc[a][b2][1] = jc[b2][1] >= 0 ? 0.0 : acc[a][b2][1];
                            }
                    }
                    // stencil entries (at most three per newcomer, in at most 2 x 2 tiles around the diagonal): the warp that
                    // owns such a tile passes it through shared memory, where a few lanes write the entries at computed addresses
                    // (the register tiles cannot be indexed dynamically)
                    if (rband && cband) {
                        const int tlo = ((s0 - 1 + W) & (W - 1)) >> 3, thi = ((s0 + nv - 1) & (W - 1)) >> 3;
                        // patch[2][2][8][8]: the tile (ti, tj) at [ti != tlo][tj != tlo]; every tile has exactly one owner warp
#pragma unroll
                        for (int a = 0; a < TR; ++a)
#pragma unroll
                            for (int b2 = 0; b2 < TC; ++b2) {
                                const int ti = wr * TR + a, tj = wc * TC + b2;
                                if ((ti == tlo || ti == thi) && (tj == tlo || tj == thi))  // warp-uniform
                                    *reinterpret_cast<double2 *>(patch + ((ti != tlo) * 2 + (tj != tlo)) * 64 + lr * 8 + 2 * lk) =
                                        make_double2(acc[a][b2][0], acc[a][b2][1]);
                            }
                        __syncwarp();
                        if (lane < 3 * nv) {
                            const int j = lane / 3, kind = lane - 3 * j;  // 0: (s, s)  1: (s, s-1)  2: (s-1, s)
                            const int sj = (s0 + j) & (W - 1), sm = (sj - 1 + W) & (W - 1);
                            const int R = kind == 2 ? sm : sj, Cc = kind == 1 ? sm : sj;
                            const int ti = R >> 3, tj = Cc >> 3;
                            // only the warp that owns the tile writes the entry
                            if (ti / TR == wr && tj / TC == wc)
                                {
                                const bool wrap = i00 + j >= n;
                                const int je = je0 + (wrap ? 1 : 0), i0 = i00 + j - (wrap ? n : 0);
                                const double *cb = coef + (je & 1) * 3 * W;
                                patch[((ti != tlo) * 2 + (tj != tlo)) * 64 + (R & 7) * 8 + (Cc & 7)] = kind == 0 ? cb[i0] : (i0 > 0 ? cb[W + i0] : 0.0);
                            }
                        }
                        __syncwarp();
#pragma unroll
                        for (int a = 0; a < TR; ++a)
#pragma unroll
                            for (int b2 = 0; b2 < TC; ++b2) {
                                const int ti = wr * TR + a, tj = wc * TC + b2;
                                if ((ti == tlo || ti == thi) && (tj == tlo || tj == thi)) {
                                    const double2 v = *reinterpret_cast<const double2 *>(patch + ((ti != tlo) * 2 + (tj != tlo)) * 64 + lr * 8 + 2 * lk);
                                    acc[a][b2][0] = v.x; acc[a][b2][1] = v.y;
                                }
                            }
                    }
                }
            }
        }
        };
        // (the number of pivots is a compile-time constant inside: the common case of four is one straight-line block)
        auto panel = [&](auto nbc) {
            constexpr int NB = decltype(nbc)::value;
            double xr[4];
            {
                const double2 x01 = *reinterpret_cast<const double2 *>(X + tg * 4), x23 = *reinterpret_cast<const double2 *>(X + tg * 4 + 2);
                xr[0] = x01.x; xr[1] = x01.y; xr[2] = x23.x; xr[3] = x23.y;
            }
            MLE_TICK(6)
            double D[4][4], gp[4], rcp[4];  // D[k][j], k >= j: entry of pivot row k in pivot column j
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int row = k < NB ? sp + k : sp;  // rows beyond nb are not used
                const double2 d01 = *reinterpret_cast<const double2 *>(X + row * 4), d23 = *reinterpret_cast<const double2 *>(X + row * 4 + 2);
                D[k][0] = d01.x; D[k][1] = d01.y; D[k][2] = d23.x; D[k][3] = d23.y;
                gp[k] = bc[k < NB ? k : 0];
                rcp[k] = 0.0;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < NB) {
                    rcp[j] = drcp_normal(D[j][j]);  // pivots of an SPD system: positive, normal range
#pragma unroll
                    for (int k = j + 1; k < 4; ++k) {
                        if (k < NB) {
                            const double lk2 = __dmul_rn(D[k][j], rcp[j]);
                            gp[k] = __fma_rn(-lk2, gp[j], gp[k]);
#pragma unroll
                            for (int j2 = j + 1; j2 <= k; ++j2) D[k][j2] = __fma_rn(-lk2, D[j2][j], D[k][j2]);
                        }
                    }
                }
            }
            MLE_TICK(7)
            auto arrive = [&](bool me, int j) {  // newcomer j takes this thread's slot; its coupling sits in panel column j
                if (ne[j] >= 0) {                // warp-uniform
                    uk = me ? ne[j] : uk;
                    g = me ? ngv[j] : g;
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (t >= j) xr[t] = me ? (t == j ? nky[j] : 0.0) : xr[t];
                }
            };
            arrive(rel == W - 1, 0);
            MLE_TICK(8)
            double lneg[4] = {0.0, 0.0, 0.0, 0.0}, vv[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < NB) {  // warp-uniform
                    const int pv = desc ? pv0 - j : pv0 + j;
                    const bool active = desc ? (uk >= lo && uk < pv) : (uk > pv);
                    const double v = active ? xr[j] : 0.0;
                    const double l = active ? __dmul_rn(v, rcp[j]) : 0.0;
                    if (active) {
                        g = __fma_rn(-l, gp[j], g);
#pragma unroll
                        for (int k = j + 1; k < 4; ++k)
                            if (k < NB) xr[k] = __fma_rn(-l, D[k][j], xr[k]);
                    }
                    // the register row / column of a pivot slot is dead until its newcomer arrives: no terms before that
                    const bool dead = rel < NB && j <= rel;
                    lneg[j] = dead ? 0.0 : -l;
                    vv[j] = dead ? 0.0 : v;
                    // this slot's pivot is eliminated; the next unknown arrives here at step j + 1
                    uk = rel == j ? -1 : uk;
                    if (j + 1 < 4) arrive(rel == j, j + 1);
                }
            }
            MLE_TICK(9)
            *reinterpret_cast<double2 *>(LA + tg * 4) = make_double2(lneg[0], lneg[1]);
            *reinterpret_cast<double2 *>(LA + tg * 4 + 2) = make_double2(lneg[2], lneg[3]);
            *reinterpret_cast<double2 *>(VB + tg * 4) = make_double2(vv[0], vv[1]);
            *reinterpret_cast<double2 *>(VB + tg * 4 + 2) = make_double2(vv[2], vv[3]);
        };
        if (!pt_warp) fresh_init();  // warps without panel rows: in the shadow of the panel factorisation
        if (pt) {
            if (nb == 4) panel(cuda::std::integral_constant<int, 4>{});
            else if (nb == 1) panel(cuda::std::integral_constant<int, 1>{});
            else if (nb == 2) panel(cuda::std::integral_constant<int, 2>{});
            else panel(cuda::std::integral_constant<int, 3>{});
        }
        MLE_TICK(2)
        gsync();
        MLE_TICK(3)
        if (pt_warp) fresh_init();
        MLE_TICK(4)
        // (4) one DMMA per tile: the four pivots' rank-1 updates in pivot order
        {
            double a[TR], b[TC];
#pragma unroll
            for (int t = 0; t < TR; ++t) a[t] = LA[((wr * TR + t) * 8 + lr) * 4 + lk];
#pragma unroll
            for (int t = 0; t < TC; ++t) b[t] = VB[((wc * TC + t) * 8 + lr) * 4 + lk];
#pragma unroll
            for (int ta = 0; ta < TR; ++ta)
#pragma unroll
                for (int tb = 0; tb < TC; ++tb) dmma_m8n8k4(acc[ta][tb][0], acc[ta][tb][1], a[ta], b[tb]);
        }
    };

    for (int half = 0; half < 2; ++half) {
        f.mirror = half == 0;  // upper half first (y-mirrored field)
        gsync();               // coef / X of the previous phase are no longer read
        compute_line(1);
        gsync();
        {   // line 1: unknowns 0 .. n-1 in slots 0 .. n-1, tridiagonal
            const double *cb = coef + 3 * W;
#pragma unroll
            for (int a = 0; a < TR; ++a)
#pragma unroll
                for (int b = 0; b < TC; ++b)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int R = (wr * TR + a) * 8 + lr, Cc = (wc * TC + b) * 8 + 2 * lk + e;
                        double v = 0.0;
                        if (R < n && Cc < n) {
                            if (R == Cc) v = cb[R];
                            else if (Cc == R - 1) v = cb[W + R];
                            else if (Cc == R + 1) v = cb[W + Cc];
                        }
                        acc[a][b][e] = v;
                    }
            uk = (pt && tg < n) ? tg : -1;
            g = (uk >= 0 && !(c == 1 && f.mirror)) ? 1.0 : 0.0;
        }
        int lines_done = 1;
        int je0 = 2, i00 = 0;  // newcomer p + n of pivot p: grid line p / n + 2, position p % n
        for (int p = 0; p < P; p += 4) {
            const int nb = min(4, P - p);
            const int lneed = min(je0 + (i00 + nb - 1 >= n ? 1 : 0), c);  // grid line of the last unknown that enters in this block
            if (lneed > lines_done) {
                for (int L = lines_done + 1; L <= lneed; ++L) compute_line(L);
                lines_done = lneed;
                gsync();
            }
            block_step(p & (W - 1), p, nb, true, false, 0, je0, i00);
            i00 += 4;
            if (i00 >= n) { i00 -= n; ++je0; }
        }
        if (half == 0) {
#pragma unroll
            for (int a = 0; a < TR; ++a)
#pragma unroll
                for (int b = 0; b < TC; ++b) {
                    park[(size_t)((a * TC + b) * 2) * pstride] = acc[a][b][0];
                    park[(size_t)((a * TC + b) * 2 + 1) * pstride] = acc[a][b][1];
                }
            g2 = g;
        }
    }
    // centre line: S = S(lower) + S(upper), then the dense two-sided elimination, one pivot per step
#pragma unroll
    for (int a = 0; a < TR; ++a)
#pragma unroll
        for (int b = 0; b < TC; ++b) {
            acc[a][b][0] = __dadd_rn(acc[a][b][0], park[(size_t)((a * TC + b) * 2) * pstride]);
            acc[a][b][1] = __dadd_rn(acc[a][b][1], park[(size_t)((a * TC + b) * 2 + 1) * pstride]);
        }
    g = __dadd_rn(g, g2);
    for (int t = 0; t < tc; ++t) block_step((P + t) & (W - 1), P + t, 1, false, false, 0, 0, 0);
    for (int t = n - 1; t > tc; --t) block_step((P + t) & (W - 1), P + t, 1, false, true, P + tc, 0, 0);
    {   // u = gs[tc] / S[tc][tc]
        const int sp = (P + tc) % W, cs = sp & 7, b0 = (sp >> 3) - wc * TC;
        if (b0 >= 0 && b0 < TC) {
            MLE_DISPATCH4(b0, TC, {
_Pragma("unroll") for (int a = 0; a < TR; ++a)
_Pragma("unroll") for (int e = 0; e < 2; ++e)
                        if (2 * lk + e == cs) X[((wr * TR + a) * 8 + lr) * 4] = acc[a][I_][e];
            })
        }
        gsync();
        if (pt && tg == sp) misc[0] = __ddiv_rn(g, X[tg * 4]);
        gsync();
        const double u = misc[0];
        gsync();  // the buffers are reused by the next solve
#ifdef MLE_DMMA_TIMING
        if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && NW == 128)
            printf("warp %2d  extract %lld  wait1 %lld  panel %lld  wait2 %lld  fresh %lld  dmma+loop %lld | xload %lld factor %lld newcomers %lld ownrow %lld\n", warp, tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5], tacc[6], tacc[7], tacc[8], tacc[9]);
#endif
        return u;
    }
}
#undef MLE_TICK

// ---- kernel ---------------------------------------------------------------------------------------------------------
template <int W> struct Rows2dCfg {
    static constexpr int LPR = MLE_ROWS_FAT ? (W > 64 ? W / 64 : 1) : (W > 32 ? W / 32 : 1);
    // measured: two rows per lane is +37 % at 16 cells; at 32 cells its 255 registers cost the second resident CTA and the gain
    static constexpr int RPL = W == 16 ? MLE_ROWS_RPL : 1;
    static constexpr int TGS = W * LPR / RPL;                 // threads per sample
    static constexpr int THREADS = (W >= 128 && !MLE_ROWS_FAT) ? 512 : W >= 8 ? 256 : 32 * W;
    static constexpr int MINB = (W >= 64 && MLE_ROWS_FAT) ? 1 : W >= 128 ? 1 : (W >= 32 && RPL == 1) ? 2 : 1;  // resident CTAs the register budget is sized for
    static constexpr int CS = THREADS / TGS;                  // samples solved concurrently
    static constexpr int S = CS < 8 ? 8 : CS;                 // samples per CTA iteration (a multiple of the DMMA tile)
    static constexpr int LDX = S + 4, LDF = S == 8 ? 8 : S + 8;
    static constexpr int GT = TGS <= 32 ? 32 : TGS;
    static constexpr int GEN_DIMS = 32 * (THREADS / S);       // dims per generator call (32-bit class masks)
    static constexpr bool DMMA = false;
    static constexpr int SH = rows_sh_doubles(W);             // shared-memory doubles of one solver group
};

// the blocked DMMA solver: 512-thread CTAs, one sample per 16 warps at 128 cells, four samples side by side at 64 cells
template <int W> struct Dmma2dCfg {
    static_assert(W == 64 || W == 128, "blocked DMMA solver: windows of 64 and 128 cells");
    static constexpr bool DMMA = true;
    static constexpr int LPR = 1, RPL = 1;
    static constexpr int TGS = W == 128 ? 512 : 128;
    static constexpr int THREADS = 512, MINB = 1;
    static constexpr int CS = THREADS / TGS;
    static constexpr int S = 8;
    static constexpr int LDX = S + 4, LDF = 8;
    static constexpr int GT = TGS;
    static constexpr int GEN_DIMS = 32 * (THREADS / S);
    static constexpr int SH = dmma_sh_doubles(W);
};

template <int MODE, int W, typename Cfg = Rows2dCfg<W>>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB) lognormal2d_rows_kernel(GenArgs g, Lognormal2dRowsArgs la, FinArgs fa,
                                                                                double *__restrict__ samples,
                                                                                long long tiles_per_shift) {
    static_assert(W >= 4, "the coarse solve needs W / 2 >= 2");
    constexpr int THREADS = Cfg::THREADS, S = Cfg::S, LDX = Cfg::LDX, LDF = Cfg::LDF, TGS = Cfg::TGS, CS = Cfg::CS;
    constexpr int LPR = Cfg::LPR, RPL = Cfg::RPL, GT = Cfg::GT, NWARP = THREADS / 32, ST = S / 8;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int mpad = la.mpad, nks = mpad >> 2;
    double *xi = reinterpret_cast<double *>(smem_raw);                       // [mpad][LDX]
    double *gsh = xi + (size_t)mpad * LDX;                                   // [CS][Cfg::SH]
    double *res = gsh + CS * Cfg::SH;                                        // [S][2]
    Partial *mred = reinterpret_cast<Partial *>(res + 2 * S);                // [2][S]
    GenScratch sc;
    sc.wtot = reinterpret_cast<unsigned *>(mred + 2 * S);                    // [NWARP]
    double *tail = reinterpret_cast<double *>(sc.wtot + ((NWARP + 3) & ~3));
    const size_t fdoubles = (size_t)la.ntiles * 8 * LDF;
    double *F = la.use_scratch ? la.scratch + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * fdoubles : tail;  // [ntiles*8][LDF]
    // generator queue: inside F while it is idle, else behind it
    const size_t qbytes = (size_t)32 * THREADS * 2;
    sc.queue = (!la.use_scratch && qbytes <= fdoubles * 8) ? reinterpret_cast<unsigned short *>(F)
               : reinterpret_cast<unsigned short *>(la.use_scratch ? tail : tail + fdoubles);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, r = blockIdx.y;
    const int lr = lane >> 2, lk = lane & 3;
    for (int e = tid; e < (mpad - g.m) * LDX; e += THREADS) xi[(size_t)g.m * LDX + e] = 0.0;  // k padding rows
    __syncthreads();
    Partial pdq{0.0, 0.0, 0.0, 0.0}, pqf{0.0, 0.0, 0.0, 0.0};
    const int Nx = la.Nx, Ny = la.Ny;

    for (long long ti = blockIdx.x; ti < tiles_per_shift; ti += gridDim.x) {
        const unsigned long long i0 = (unsigned long long)ti * (unsigned long long)la.spt;
        const unsigned long long left = g.n - i0;
        const int Pt = left < (unsigned long long)la.spt ? (int)left : la.spt;
        // A. inputs (always S points: the field synthesis works on whole 8-sample DMMA tiles; surplus points are never solved)
        for (int j0 = 0; j0 < g.m; j0 += Cfg::GEN_DIMS)
            gen_tile_owned<THREADS, S, MODE, LDX, unsigned, 4>(xi + (size_t)j0 * LDX, j0, min(Cfg::GEN_DIMS, g.m - j0),
                                                              g.k0 + i0, g, r, sc);
        // B. field: F[node][sample] = exp(sum_k Phi[node][k] xi[k][sample]), ascending k (DMMA = 4 sequential FMAs)
        for (int rt = warp; rt < la.ntiles; rt += NWARP) {
            double acc[ST][2];
#pragma unroll
            for (int st = 0; st < ST; ++st) { acc[st][0] = 0.0; acc[st][1] = 0.0; }
            const double *A = la.basis + (size_t)rt * nks * 32 + lane;
            const double *B = xi + lk * LDX + lr;
            double a = __ldg(A);
#pragma unroll 4
            for (int ks = 0; ks < nks; ++ks) {
                const double na = __ldg(A + (ks + 1 < nks ? ks + 1 : ks) * 32);
#pragma unroll
                for (int st = 0; st < ST; ++st) dmma_m8n8k4(acc[st][0], acc[st][1], a, B[ks * 4 * LDX + 8 * st]);
                a = na;
            }
            double *f0 = F + (size_t)(8 * rt + lr) * LDF + 2 * lk;
#pragma unroll
            for (int st = 0; st < ST; ++st) {
                double2 v;
                v.x = exp_det(acc[st][0]);
                v.y = exp_det(acc[st][1]);
                *reinterpret_cast<double2 *>(f0 + 8 * st) = v;
            }
        }
        __syncthreads();
        // C. solves: CS samples at a time; rounds without a live sample are skipped (a round is a sequential elimination of
        // thousands of pivots: a 128 x 128 call of one sample used to pay for eight)
        // (how the loop is bounded matters to ptxas at the register caps of these kernels: a trip count read back from shared
        // memory keeps the blocked DMMA solver at its old spill level -- a register-held count cost it 9 % --, the constant bound
        // with an early exit is the better form for the register-window solvers)
        if constexpr (Cfg::DMMA) {
            if (tid == 0) reinterpret_cast<int *>(mred)[0] = (Pt + CS - 1) / CS;
            __syncthreads();
        }
        for (int rd = 0; Cfg::DMMA ? rd < reinterpret_cast<volatile int *>(mred)[0] : rd < S / CS; ++rd) {
            if (!Cfg::DMMA && rd * CS >= Pt) break;
            const int grp = tid / TGS, tg = tid % TGS, sigma = rd * CS + grp;
            double *sh = gsh + grp * Cfg::SH;
            FieldView fv{F + sigma, LDF, Nx + 1, 1, 1, Nx, Ny, 0, (double)Nx * (double)Nx, (double)Ny * (double)Ny};
            double *park = la.park + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * ROWS_PARK * THREADS + tid;
            // the solver of the view's window width: one row per lane (or a row over several lanes), or RPL rows per lane
            auto solve = [&](const FieldView &v, bool half_width) {
                if constexpr (Cfg::DMMA) {
                    // multilevel differences on square grids only (the host dispatches nothing else here)
                    if (!half_width) return frontal_mid_dmma<W, TGS / 32>(v, sh, park, THREADS, tg, 1 + grp);
                    if constexpr (W == 128) {
                        return frontal_mid_dmma<64, TGS / 32>(v, sh, park, THREADS, tg, 1 + grp);
                    } else {
                        // 32 cells: the register-window solver on the group's first warp; the others wait for it
                        double u = 0.0;
                        if (tg < 32) u = frontal_mid_rows<32, 1, 32>(v, sh, park, THREADS, tg, 1 + grp);
                        rows_sync<GT>(1 + grp);
                        return u;
                    }
                } else if constexpr (RPL > 1) {
                    if (half_width) return frontal_mid_rpl<W / 2, RPL>(v, sh, park, THREADS, tg);
                    return frontal_mid_rpl<W, RPL>(v, sh, park, THREADS, tg);
                } else {
                    if (half_width) return frontal_mid_rows<W / 2, (W / 2 / LPR >= 2 ? LPR : 1), GT>(v, sh, park, THREADS, tg, 1 + grp);
                    return frontal_mid_rows<W, LPR, GT>(v, sh, park, THREADS, tg, 1 + grp);
                }
            };
            const double qf = solve(fv, false);
            double dq = qf;
            // coarse terms of diff(index), added in the order (lx-1,ly-1), (lx,ly-1), (lx-1,ly); multilevel: one coarse solve
            auto coarse = [&](int stx, int sty) {
                FieldView cv = fv;
                cv.stx = stx; cv.sty = sty; cv.Nx = Nx / stx; cv.Ny = Ny / sty;
                cv.wx = (double)cv.Nx * (double)cv.Nx; cv.wy = (double)cv.Ny * (double)cv.Ny;
                return solve(cv, stx == 2);
            };
            if (la.ndims == 1) {
                if (la.lx > 0) dq = __dsub_rn(qf, coarse(2, 2));
            } else {
                if (la.lx > 0 && la.ly > 0) dq = __dadd_rn(dq, coarse(2, 2));
                if (la.ly > 0) dq = __dsub_rn(dq, coarse(1, 2));
                if (la.lx > 0) dq = __dsub_rn(dq, coarse(2, 1));
            }
            if (tg == 0) { res[2 * sigma] = dq; res[2 * sigma + 1] = qf; }
        }
        __syncthreads();
        // D. statistics: slot tid < S
        if (tid < Pt) {
            const double dq = res[2 * tid], qf = res[2 * tid + 1];
            if (samples) {
                samples[((size_t)r * 2 + 0) * g.n + i0 + tid] = dq;
                samples[((size_t)r * 2 + 1) * g.n + i0 + tid] = qf;
            }
            partial_add(pdq, dq);
            partial_add(pqf, qf);
        }
        // (the barriers of the next tile's generator order the res reads above before res is rewritten)
    }
    if (tid < S) {
        mred[tid] = pdq;
        mred[S + tid] = pqf;
    }
    __syncthreads();
    if (tid < 2) {
        Partial a{0.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < S; ++i) a = partial_merge(a, mred[tid * S + i]);
        fa.partials[((size_t)r * 2 + tid) * gridDim.x + blockIdx.x] = a;
    }
    cta_finalize(fa, 2);
}

}  // namespace mle
