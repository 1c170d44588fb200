// mle_kernel_ln1d.cuh -- the headline kernel: 1-D lognormal diffusion, fully fused (sm_100a).
//
// Model (defined in oracle/mle_oracle.c and DESIGN.md; structure of docs/src/example.md:106-128): -(a u')' = 1 on (0,1),
// u(0) = u(1) = 0, N = n0 2^l cells, log a(x_i) = sum_k Phi_l[i][k] xi_k, interface coefficient = mean of the two node values,
// QoI u(1/2) from the closed-form flux sums, coarse problem = even nodes of the fine field, dQ = Qf - Qc.
//
// Schedule: WARP-AUTONOMOUS.  A CTA is two warps that share nothing but the read-only {z_j, shift_j} table; each warp takes
// 16 consecutive samples through the whole pipeline on its own and only ever executes __syncwarp():
//   A. inputs: gen_warp16 (mle_device.cuh): lattice point -> shift -> Phi^-1 with all three erfinv branches compacted,
//      xi[k][16] in shared memory (XOR-swizzled, no padding).
//   B. KL contraction on the FP64 tensor instruction, rows in chunks of 32 (16 half-nodes x 2 sides): the warp owns a
//      4 x 2 block of 8x8 tiles (four row tiles x two sample tiles: 6 fragment loads per 8 DMMA), A fragments from the
//      fragment-ordered basis in L1/L2, B fragments from xi.  On B200 one mma.sync.m8n8k4.f64 is bit-identical to four
//      fused multiply-adds in ascending k (tools/dmma_order_test.cu), so the field equals the oracle's sequential-FMA field.
//      Row r' = 2 i' + side (i' < N/2): side 0 = node i', side 1 = the mirrored node N - i', so the sweeps of both halves
//      consume rows in ascending order.  The MID node is not a DMMA row (N is a multiple of 8, N + 1 is not): its value is
//      one chain of m fused multiply-adds per sample on the FP64 pipe -- the same ascending-k chain.
//   C. epilogue a = exp(G) -> F[32][16] (warp-private, swizzled), then the sweeps: lane (side, sample) adds one edge per
//      half-node to the fine flux sums and one per even half-node to the coarse ones.
//   D. the sides meet through a shuffle; per-lane pivoted statistics (mle_device.cuh), merged per CTA at the very end; the
//      CTA that finishes last for its shift reduces the per-CTA partials (cta_finalize): one launch per sample! call.
// Nothing per-sample is written to HBM unless the caller asks for the raw samples.
#pragma once
#include "mle_device.cuh"

namespace mle {

struct Lognormal1dArgs {
    const double *basis;  // [row tile][k step][lane]: rows r' = 2 i' + side, i' < N/2, zero padded to whole 8-row tiles
    const double *mid;    // [mpad] basis row of the mid node N/2 (zero padded in k)
    int N;                // fine cells
    int mpad;             // KL terms padded to a multiple of 4
    int ntiles;           // 8-row tiles = ceil(N / 8)
    int nchunks;          // chunks of 4 row tiles
    int level;
    double h2, h2c;       // 1/N^2, 1/(N/2)^2
};

constexpr int LN_BT = 32;                    // samples per CTA tile; a warp owns WS = 8 or 16 of them (template parameter)
constexpr int LN_GEN_PIECE = 128;            // dims per generator call (private lists + queue inside F)
__host__ __device__ constexpr int ln_threads(int ws) { return 32 * (LN_BT / ws); }
__host__ __device__ constexpr size_t ln1d_smem_bytes(int mpad, int m_lattice, int ws) {
    return (size_t)m_lattice * 16 + (size_t)mpad * 8 + (size_t)(LN_BT / ws) * ((size_t)mpad * ws + 32 * ws) * 8 + 16;
}

struct SweepState {
    double a_prev, A, D;  // last node value, sum of r = 1/kappa, sum of w*r over the edges seen so far
};

// one edge on arrival of node value `a` (node number ip on this side, counted from the boundary; M = N/2): the closed form of
// the mid value (oracle/mle_oracle.c:tridiag_mid) needs only two running sums per side; the division is off the chain
__device__ __forceinline__ void sweep_step(SweepState &s, int ip, double a, int M) {
    if (ip > 0) {
        const double kap = __dmul_rn(0.5, __dadd_rn(s.a_prev, a));
        const double r = ddiv_normal(1.0, kap);  // == 1.0 / kap for every kappa in the normal range (a = exp(g), |g| < 700)
        s.A = __dadd_rn(s.A, r);
        s.D = __fma_rn((double)(M - ip) + 0.5, r, s.D);
    }
    s.a_prev = a;
}

// reciprocal of kappa = (a0 + a1)/2: the Newton / Markstein sequence of ddiv_normal(1.0, kappa) without the product 1.0 * r
// (exact, so the result is the same correctly rounded 1/kappa)
__device__ __forceinline__ double rcp_kappa(double a0, double a1) {
    const double den = __dmul_rn(0.5, __dadd_rn(a0, a1));
    double r = rcp_seed(den);
    double e = __fma_rn(-den, r, 1.0);
    e = __fma_rn(e, e, e);
    r = __fma_rn(r, e, r);
    e = __fma_rn(-den, r, 1.0);
    r = __fma_rn(r, e, r);
    const double rem = __fma_rn(-den, r, 1.0);
    return __fma_rn(r, rem, r);
}

// NN consecutive nodes of one side at once (node numbers ip0 .. ip0+NN-1 counted from the boundary, values a[]): the same
// arithmetic as NN calls of sweep_step, arranged as straight-line code -- all reciprocals are independent of each other and of the
// running sums, so their Newton chains interleave; only the two accumulations per edge are sequential.
// `first`: node ip0 is the boundary node (no edge in front of it).
template <int NN>
__device__ __forceinline__ void sweep_nodes(SweepState &s, const double (&a)[NN], int ip0, int M, bool first) {
    double r[NN];
#pragma unroll
    for (int q = 0; q < NN; ++q) r[q] = rcp_kappa(q ? a[q - 1] : s.a_prev, a[q]);
    const double w0 = (double)(M - ip0) + 0.5;
#pragma unroll
    for (int q = 0; q < NN; ++q) {
        if (q == 0 && first) continue;  // warp-uniform
        s.A = __dadd_rn(s.A, r[q]);
        s.D = __fma_rn(w0 - (double)q, r[q], s.D);  // exact: small half-integers
    }
    s.a_prev = a[NN - 1];
}

// NN half-nodes of one side (rows 2 q + side of F, q = g0 .. g0+NN-1 of a chunk that starts at half-node ip0, an even number),
// optionally followed by the mid node: the fine state takes every node, the coarse state the even ones.
template <int NN, bool LAST>
__device__ __forceinline__ void sweep_group(SweepState &sf, SweepState &sc, const double *__restrict__ Fq, int stride, int ip,
                                            int M, bool coarse, double amid) {
    double a[NN + (LAST ? 1 : 0)];
#pragma unroll
    for (int q = 0; q < NN; ++q) a[q] = Fq[q * stride];
    if (LAST) a[NN] = amid;
    sweep_nodes<NN + (LAST ? 1 : 0)>(sf, a, ip, M, ip == 0);
    if (coarse) {
        double ac[NN / 2 + (LAST ? 1 : 0)];
#pragma unroll
        for (int q = 0; q < NN / 2; ++q) ac[q] = a[2 * q];
        if (LAST) ac[NN / 2] = amid;
        sweep_nodes<NN / 2 + (LAST ? 1 : 0)>(sc, ac, ip >> 1, M >> 1, ip == 0);
    }
}

// nn consecutive nodes of one grid and side from shared memory (node ip+q at Fq[q * stride]), in straight-line groups of 8 / 4
// nodes and single steps for the rest; with `last` the mid node (value amid, node number M) closes the run.
template <int NN, bool LAST>
__device__ __forceinline__ void sweep_run_group(SweepState &s, const double *__restrict__ Fq, int stride, int ip, int M, double amid) {
    double a[NN + (LAST ? 1 : 0)];
#pragma unroll
    for (int q = 0; q < NN; ++q) a[q] = Fq[q * stride];
    if (LAST) a[NN] = amid;
    sweep_nodes<NN + (LAST ? 1 : 0)>(s, a, ip, M, ip == 0);
}
__device__ __forceinline__ void sweep_run(SweepState &s, const double *__restrict__ Fq, int stride, int nn, int ip, int M, bool last,
                                          double amid) {
    int g0 = 0;
    for (; g0 + 8 <= nn; g0 += 8) {
        if (last && g0 + 8 == nn) sweep_run_group<8, true>(s, Fq + g0 * stride, stride, ip + g0, M, amid);
        else sweep_run_group<8, false>(s, Fq + g0 * stride, stride, ip + g0, M, amid);
    }
    for (; g0 + 4 <= nn; g0 += 4) {
        if (last && g0 + 4 == nn) sweep_run_group<4, true>(s, Fq + g0 * stride, stride, ip + g0, M, amid);
        else sweep_run_group<4, false>(s, Fq + g0 * stride, stride, ip + g0, M, amid);
    }
    if (g0 < nn || (last && nn == 0)) {
        for (; g0 < nn; ++g0) sweep_step(s, ip + g0, Fq[g0 * stride], M);
        if (last) sweep_step(s, M, amid, M);
    }
}

// u(1/2) = h^2 (D_R A_L + D_L A_R) / (A_L + A_R)
__device__ __forceinline__ double sweep_combine(double LA, double LD, double RA, double RD, double h2) {
    const double num = __fma_rn(RD, LA, __dmul_rn(LD, RA));
    return __dmul_rn(h2, ddiv_normal(num, __dadd_rn(LA, RA)));  // sums of reciprocals of kappa = mean of exp(.): normal range
}

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// contraction + epilogue of one chunk: RT live row tiles (1..4) x ST = WS/8 sample tiles (always both: the inputs of surplus
// points of a last, partial tile exist and are masked out of the statistics only).
// With MID the chain of fused multiply-adds of the mid-node row (smid[k] * xi[k][b], ascending k: the lane's own sample b)
// rides along in the k loop, in the shadow of the DMMAs, and gm returns the FIELD VALUE exp(.) at the mid node.
template <int RT, bool MID, int WS>
__device__ __forceinline__ void ln1d_chunk(const double *__restrict__ A, const double *__restrict__ B0,
                                           const double *__restrict__ B1, int nks, bool two, double *__restrict__ F, int lane,
                                           const double *__restrict__ smid, const double *__restrict__ xi, double &gm) {
    constexpr int ST = WS / 8, KS = 4 * WS;  // doubles of xi per k step
    const int lr = lane >> 2, lk = lane & 3;
    double acc[RT][ST][2];
#pragma unroll
    for (int t = 0; t < RT; ++t)
#pragma unroll
        for (int st = 0; st < ST; ++st) acc[t][st][0] = acc[t][st][1] = 0.0;
    const size_t ts = (size_t)nks * 32;
    double a[RT], b0, b1 = 0.0;
#pragma unroll
    for (int t = 0; t < RT; ++t) a[t] = __ldg(A + t * ts);
    b0 = B0[0];
    if (ST == 2) b1 = B1[0];
    const int b = lane & (WS - 1);
    const double *x0 = xi + swz<WS>(0, b), *x1 = xi + swz<WS>(1, b), *x2 = xi + swz<WS>(2, b), *x3 = xi + swz<WS>(3, b);
    double g0 = 0.0;
#pragma unroll 2
    for (int ks = 0; ks < nks; ++ks) {
        const int kn = (ks + 1 < nks) ? ks + 1 : ks;  // next k step's fragments
        double na[RT];
#pragma unroll
        for (int t = 0; t < RT; ++t) na[t] = __ldg(A + t * ts + kn * 32);
        const double nb0 = B0[kn * KS];
        double nb1 = 0.0;
        if (ST == 2) nb1 = B1[kn * KS];
        double2 p01, p23;
        double y0, y1, y2, y3;
        if (MID) {
            p01 = *reinterpret_cast<const double2 *>(smid + 4 * ks);
            p23 = *reinterpret_cast<const double2 *>(smid + 4 * ks + 2);
            y0 = x0[ks * KS]; y1 = x1[ks * KS]; y2 = x2[ks * KS]; y3 = x3[ks * KS];
        }
#pragma unroll
        for (int t = 0; t < RT; ++t) dmma_m8n8k4(acc[t][0][0], acc[t][0][1], a[t], b0);
        if (ST == 2) {
#pragma unroll
            for (int t = 0; t < RT; ++t) dmma_m8n8k4(acc[t][ST - 1][0], acc[t][ST - 1][1], a[t], b1);
        }
        if (MID) {
            g0 = __fma_rn(p01.x, y0, g0);
            g0 = __fma_rn(p01.y, y1, g0);
            g0 = __fma_rn(p23.x, y2, g0);
            g0 = __fma_rn(p23.y, y3, g0);
        }
#pragma unroll
        for (int t = 0; t < RT; ++t) a[t] = na[t];
        b0 = nb0; b1 = nb1;
    }
    if (MID) gm = exp_det(g0);  // field value at the mid node (straight-line with the exps below: the chains interleave)
    // epilogue: a = exp(G) -> F[row][sample]; the lane holds row lr, samples 2 lk, 2 lk + 1 of each 8x8 tile
    // (WS = 16: columns XOR-swizzled by 8 (row & 1) so that the 16-byte stores of a quarter warp hit 8 distinct bank groups)
    const int sw = WS == 16 ? (lr & 1) << 3 : 0;
    double *f0 = F + lr * WS + ((2 * lk) ^ sw), *f1 = F + lr * WS + ((2 * lk) ^ sw ^ 8);
#pragma unroll
    for (int t = 0; t < RT; ++t) {
        double2 v;
        v.x = exp_det(acc[t][0][0]); v.y = exp_det(acc[t][0][1]);
        *reinterpret_cast<double2 *>(f0 + t * 8 * WS) = v;
        if (ST == 2) {
            v.x = exp_det(acc[t][ST - 1][0]); v.y = exp_det(acc[t][ST - 1][1]);
            *reinterpret_cast<double2 *>(f1 + t * 8 * WS) = v;
        }
    }
}

template <bool MID, int WS>
__device__ __forceinline__ void ln1d_chunk_any(int live, const double *__restrict__ A, const double *__restrict__ B0,
                                               const double *__restrict__ B1, int nks, bool two, double *__restrict__ F, int lane,
                                               const double *__restrict__ smid, const double *__restrict__ xi, double &gm) {
    if (live == 4) ln1d_chunk<4, MID, WS>(A, B0, B1, nks, two, F, lane, smid, xi, gm);
    else if (live == 2) ln1d_chunk<2, MID, WS>(A, B0, B1, nks, two, F, lane, smid, xi, gm);
    else if (live == 3) ln1d_chunk<3, MID, WS>(A, B0, B1, nks, two, F, lane, smid, xi, gm);
    else ln1d_chunk<1, MID, WS>(A, B0, B1, nks, two, F, lane, smid, xi, gm);
}

template <int MODE, int WS>
__global__ void __launch_bounds__(32 * (LN_BT / WS), WS == 16 ? 6 : 5) lognormal1d_kernel(GenArgs g, Lognormal1dArgs la, FinArgs fa,
                                                                                          double *__restrict__ samples,
                                                                                          long long tiles_per_shift) {
    constexpr bool lattice = (MODE & GEN_MC) == 0;
    constexpr int WARPS = LN_BT / WS, THREADS = 32 * WARPS, FD = 32 * WS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int mpad = la.mpad, nks = mpad >> 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, r = blockIdx.y;
    double2 *zs = reinterpret_cast<double2 *>(smem_raw);                                  // [m] (lattice modes)
    double *smid = reinterpret_cast<double *>(zs + (lattice ? g.m : 0));                  // [mpad] basis row of the mid node
    double *xi = smid + mpad + (size_t)warp * ((size_t)mpad * WS + FD);                   // [mpad][WS] swizzled
    double *F = xi + (size_t)mpad * WS;                                                   // [32][WS]
    unsigned char *gscratch = reinterpret_cast<unsigned char *>(F);                       // generator lists + queue (F is idle then)
    if (lattice) stage_lattice_row<THREADS>(zs, g, r);
    for (int e = tid; e < mpad; e += THREADS) smid[e] = __ldg(la.mid + e);
    for (int e = lane; e < (mpad - g.m) * WS; e += 32) xi[(size_t)g.m * WS + e] = 0.0;    // k padding rows
    __syncthreads();

    const int M = la.N >> 1, lvl = la.level;
    const int lr = lane >> 2, lk = lane & 3;
    // B fragments: row 4 ks + lk (row & 3 = lk), sample 8 st + lr at the swizzled column
    const double *B0 = xi + swz<WS>(lk, lr), *B1 = xi + swz<WS>(lk, (lr + 8) & (WS - 1));
    // sweep roles.  WS = 16: lane (side, sample) carries the fine and the coarse state of its side;
    //               WS = 8:  lane (grid, side, sample): lanes 0..15 the fine grid, lanes 16..31 the coarse one (even half-nodes)
    const int side = WS == 16 ? lane >> 4 : (lane >> 3) & 1, b = lane & (WS - 1), grid = WS == 16 ? 0 : lane >> 4;
    const double *Fr = F + side * WS + (WS == 16 ? (b ^ (side << 3)) : b);  // row 2 q + side, sample b
    Partial pdq{0.0, 0.0, 0.0, 0.0}, pqf{0.0, 0.0, 0.0, 0.0};

    for (long long ti = blockIdx.x; ti < tiles_per_shift; ti += gridDim.x) {
        const unsigned long long i0 = (unsigned long long)ti * LN_BT + (unsigned long long)(warp * WS);
        if (i0 >= g.n) continue;  // this warp's part of the last tile is empty (warp-uniform)
        const unsigned long long left = g.n - i0;
        const int Pt = left < (unsigned long long)WS ? (int)left : WS;
        const bool two = Pt > 8;
        // A. inputs (always all WS points: surplus ones are masked out of the statistics)
        for (int j0 = 0; j0 < g.m; j0 += LN_GEN_PIECE)
            gen_warp<MODE, lattice, WS>(xi + (size_t)j0 * WS, gscratch, FD * 8, j0, min(LN_GEN_PIECE, g.m - j0), g.k0 + i0, g, r, zs);

        SweepState sf{0.0, 0.0, 0.0}, sc{0.0, 0.0, 0.0};
        double gm = 0.0;
        for (int c = 0; c < la.nchunks; ++c) {
            // B. contraction + exp for row tiles 4c .. 4c + live - 1
            const int live = min(4, la.ntiles - 4 * c);
            const double *A = la.basis + (size_t)(4 * c) * nks * 32 + lane;
            if (c == la.nchunks - 1) ln1d_chunk_any<true, WS>(live, A, B0, B1, nks, two, F, lane, smid, xi, gm);
            else ln1d_chunk_any<false, WS>(live, A, B0, B1, nks, two, F, lane, smid, xi, gm);
            __syncwarp();
            // C. sweeps over the half-nodes of this chunk (ip0 is a multiple of 16: even half-nodes are the coarse nodes), in
            // straight-line groups of 8 or 4 nodes; the mid node closes the last group
            const int ip0 = 16 * c;
            const int nh = min(16, M - ip0);
            const bool last = c == la.nchunks - 1;
            if (WS == 16) {
                const bool co = lvl > 0;
                const int st = 2 * WS;
#ifndef LN_SWEEP_NN
#define LN_SWEEP_NN 8
#endif
                constexpr int NN = LN_SWEEP_NN;  // nodes per straight-line group
                int g0 = 0;
                if (NN == 8) {
                    for (; g0 + 8 <= nh; g0 += 8) {
                        if (last && g0 + 8 == nh) sweep_group<8, true>(sf, sc, Fr + g0 * st, st, ip0 + g0, M, co, gm);
                        else sweep_group<8, false>(sf, sc, Fr + g0 * st, st, ip0 + g0, M, co, 0.0);
                    }
                }
                for (; g0 + 4 <= nh; g0 += 4) {
                    if (last && g0 + 4 == nh) sweep_group<4, true>(sf, sc, Fr + g0 * st, st, ip0 + g0, M, co, gm);
                    else sweep_group<4, false>(sf, sc, Fr + g0 * st, st, ip0 + g0, M, co, 0.0);
                }
                if (g0 < nh) {  // grids whose half is not a multiple of four nodes: the rest one node at a time
                    for (; g0 < nh; ++g0) {
                        const double a = Fr[g0 * st];
                        sweep_step(sf, ip0 + g0, a, M);
                        if (co && !(g0 & 1)) sweep_step(sc, (ip0 + g0) >> 1, a, M >> 1);
                    }
                    if (last) {
                        sweep_step(sf, M, gm, M);
                        if (co) sweep_step(sc, M >> 1, gm, M >> 1);
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    if (q < nh) {
                        const double a = Fr[q * 2 * WS];
                        if (!grid) sweep_step(sf, ip0 + q, a, M);
                        else if (!(q & 1) && lvl > 0) sweep_step(sf, (ip0 + q) >> 1, a, M >> 1);
                    }
                }
            }
            __syncwarp();  // F is rewritten by the next chunk's epilogue
        }
        // the mid node (its chain of fused multiply-adds rode along in the last chunk; WS = 16: already consumed above)
        const double amid = gm;
        double qf, dq;
        if (WS == 16) {
            // D. the two sides meet: lanes 0..15 hold the left states, their partners lane + 16 the right ones
            const double fA = __shfl_xor_sync(0xffffffffu, sf.A, 16), fD = __shfl_xor_sync(0xffffffffu, sf.D, 16);
            const double cA = __shfl_xor_sync(0xffffffffu, sc.A, 16), cD = __shfl_xor_sync(0xffffffffu, sc.D, 16);
            qf = sweep_combine(sf.A, sf.D, fA, fD, la.h2);
            dq = qf;
            if (lvl > 0) dq = __dsub_rn(qf, sweep_combine(sc.A, sc.D, cA, cD, la.h2c));
        } else {
            if (!grid) sweep_step(sf, M, amid, M);
            else if (lvl > 0) sweep_step(sf, M >> 1, amid, M >> 1);
            // D. left states in lanes with side 0, right states in lane + 8; the coarse value of sample b sits in lane b + 16
            const double xA = __shfl_xor_sync(0xffffffffu, sf.A, 8), xD = __shfl_xor_sync(0xffffffffu, sf.D, 8);
            const double q = sweep_combine(sf.A, sf.D, xA, xD, grid ? la.h2c : la.h2);
            const double qc = __shfl_down_sync(0xffffffffu, q, 16);
            qf = q;
            dq = lvl > 0 ? __dsub_rn(q, qc) : q;
        }
        if (lane < Pt) {
            if (samples) {
                samples[((size_t)r * 2 + 0) * g.n + i0 + lane] = dq;
                samples[((size_t)r * 2 + 1) * g.n + i0 + lane] = qf;
            }
            partial_add(pdq, dq);
            partial_add(pqf, qf);
        }
        __syncwarp();
    }
    // merge the per-lane statistics of the CTA in slot order (slot = warp * WS + lane)
    __syncthreads();
    Partial *slots = reinterpret_cast<Partial *>(smid + mpad);  // [2][LN_BT] over warp 0's xi / F
    if (lane < WS) {
        slots[warp * WS + lane] = pdq;
        slots[LN_BT + warp * WS + lane] = pqf;
    }
    __syncthreads();
    if (tid < 2) {
        Partial a{0.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < LN_BT; ++i) a = partial_merge(a, slots[tid * LN_BT + i]);
        fa.partials[((size_t)r * 2 + tid) * gridDim.x + blockIdx.x] = a;
    }
    cta_finalize(fa, 2);
}

}  // namespace mle
