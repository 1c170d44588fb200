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
};

template <int MODE, int W>
__global__ void __launch_bounds__(Rows2dCfg<W>::THREADS, Rows2dCfg<W>::MINB) lognormal2d_rows_kernel(GenArgs g, Lognormal2dRowsArgs la, FinArgs fa,
                                                                                double *__restrict__ samples,
                                                                                long long tiles_per_shift) {
    static_assert(W >= 4, "the coarse solve needs W / 2 >= 2");
    using Cfg = Rows2dCfg<W>;
    constexpr int THREADS = Cfg::THREADS, S = Cfg::S, LDX = Cfg::LDX, LDF = Cfg::LDF, TGS = Cfg::TGS, CS = Cfg::CS;
    constexpr int LPR = Cfg::LPR, RPL = Cfg::RPL, GT = Cfg::GT, NWARP = THREADS / 32, ST = S / 8;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int mpad = la.mpad, nks = mpad >> 2;
    double *xi = reinterpret_cast<double *>(smem_raw);                       // [mpad][LDX]
    double *gsh = xi + (size_t)mpad * LDX;                                   // [CS][rows_sh_doubles(W)]
    double *res = gsh + CS * rows_sh_doubles(W);                             // [S][2]
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
        const unsigned long long i0 = (unsigned long long)ti * S;
        const unsigned long long left = g.n - i0;
        const int Pt = left < (unsigned long long)S ? (int)left : S;
        // A. inputs (always a full tile: surplus points are masked out of the statistics)
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
        // C. solves: CS samples at a time
        for (int rd = 0; rd < S / CS; ++rd) {
            const int grp = tid / TGS, tg = tid % TGS, sigma = rd * CS + grp;
            double *sh = gsh + grp * rows_sh_doubles(W);
            FieldView fv{F + sigma, LDF, Nx + 1, 1, 1, Nx, Ny, 0, (double)Nx * (double)Nx, (double)Ny * (double)Ny};
            double *park = la.park + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * ROWS_PARK * THREADS + tid;
            // the solver of the view's window width: one row per lane (or a row over several lanes), or RPL rows per lane
            auto solve = [&](const FieldView &v, bool half_width) {
                if constexpr (RPL > 1) {
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
