/* mle_oracle.c -- CPU restatement of the MultilevelEstimators.jl sampling hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (multilevelestimators.jl_b200/,
 * libmle_cuda) may link, import or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do, and only as the checker / CPU baseline.
 *
 * What it restates (file:line relative to /root/reference):
 *   stage 1  rank-1 lattice points            src/lattice.jl:142-145 (unshifted), :168-171 (shifted),
 *                                             :166,:175 (k > n), :178 (reversebits)
 *   stage 2  inverse-CDF transforms           src/distribution.jl:84 (Uniform), :112 (Normal),
 *                                             :145-149 (TruncatedNormal), :151 (Phi^-1), :153 (Phi), :182 (Weibull)
 *            erfinv                           SpecialFunctions.jl (compat "2", Project.toml:25; NOT in the tree):
 *                                             Blair-Edwards-Johnson rational approximants, @horner = fused muladd chain
 *   stage 3  sample functions                 test/regression.jl:17-46, :69-103 (problem_18, the only executable sample
 *                                             function in the reference); docs/src/example.md:106-128 (structure of the
 *                                             lognormal-diffusion function: fine field, coarse = stride-2 view, QoI at the
 *                                             mid index, return (dQ, Qf)); src/index.jl:49-58 (diff signs)
 *   stage 4  per-(qoi, shift) statistics      src/sample.jl:39-56, :79-111 (loop order, zero-based k, (shift, k) product);
 *                                             src/methods.jl:44-57, :282 (mean / corrected var per shift)
 *
 * Parity pinning: stages 1-2 are pinned by the reference's doctest golden vectors (tests/golden/, SURVEY.md 8c:
 * G1, G2, G4 bit-exact; G5-G8 to the reference's own tolerance).  problem_18 restates reference test code.
 * The lognormal-diffusion models are NOT in the reference tree (they live in GaussianRandomFields.jl /
 * SimpleMultigrid.jl): for those this oracle is "parity unpinned" -- it defines the model that the CUDA path must match.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off -mfma; every fused multiply-add is an explicit fma()).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* the published generating vectors the reference reads at src/lattice.jl:88-90 (embedded, see tools/embed_genvecs.py) */
#include "../include/mle_genvec_data.inc"
ORC_API const uint32_t *orc_genvec(int which, int *len) {
    if (which == 0) { *len = MLE_GENVEC_K3600_LEN; return MLE_GENVEC_K3600; }
    if (which == 1) { *len = MLE_GENVEC_CKN250_LEN; return MLE_GENVEC_CKN250; }
    *len = 0;
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* Elementary functions the reference takes from Julia's runtime (Base.log1p, Base.exp -- not in the tree). */
/* Any two 1-ulp implementations of them differ in the last bit now and then, and the erfinv tail approximant  */
/* amplifies a 1-ulp difference of the logarithm to ~10 ulp.  To make "the same arithmetic" checkable bit for */
/* bit, the oracle and the CUDA kernels both use the algorithms defined here (tables: tools/gen_math_tables.py, */
/* include/mle_math_tables.inc); tests/ checks them against mpmath / libm.                                      */
/* ------------------------------------------------------------------------------------------------ */
#include "../include/mle_math_tables.inc"
static const double LOGT[3 * MLE_LOGT_ROWS] = MLE_LOGT_TABLE_INIT;

static inline double u64_as_double(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
static inline uint64_t double_as_u64(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }

/* log(w) for a normal double w in (0, 1/16]; error ~2^-66 absolute on |log w| >= 2.77 (correctly rounded except
 * with probability ~1e-4).  w = 2^e m, m in [1,2); c = table centre of m's 1/128 bin; m/c - 1 = r exactly as
 * r_hi + p_lo (two-product); log w = e ln2 + log c + log1p(r). */
ORC_API double orc_log_tail(double w) {
    uint64_t bits = double_as_u64(w);
    int e = (int)(bits >> 52) - 1023;
    double m = u64_as_double((bits & 0x000fffffffffffffull) | 0x3ff0000000000000ull);
    int i = (int)((bits >> 45) & 127);
    double invc = LOGT[3 * i], logc_hi = LOGT[3 * i + 1], logc_lo = LOGT[3 * i + 2];
    double p_hi = m * invc;
    double p_lo = fma(m, invc, -p_hi); /* exact */
    double r_hi = p_hi - 1.0;          /* exact */
    double r = r_hi + p_lo;
    double q = -0.125;                 /* log1p(r) - r = r^2 (-1/2 + r/3 - r^2/4 + ... - r^6/8) */
    q = fma(q, r, 0x1.2492492492492p-3);  /* 1/7 */
    q = fma(q, r, -0x1.5555555555555p-3); /* 1/6 */
    q = fma(q, r, 0.2);
    q = fma(q, r, -0.25);
    q = fma(q, r, 0x1.5555555555555p-2);  /* 1/3 */
    q = fma(q, r, -0.5);
    q = (r * r) * q;
    double ed = (double)e;
    double A = ed * MLE_LOGT_LN2_HI + logc_hi; /* both exact: 32-bit ln2_hi, logc_hi multiple of 2^-40 */
    double s = A + r_hi;
    double err = r_hi - (s - A);               /* Fast2Sum, |A| > |r_hi| */
    double lo = fma(ed, MLE_LOGT_LN2_LO, logc_lo) + p_lo;
    lo = lo + q;
    lo = lo + err;
    return s + lo;
}

/* log(w) for ANY positive normal double, as an unevaluated sum hi + lo (|lo| <= ulp(hi)/2) -- Weibull's log(1 - x) and the
 * logarithm inside pow (src/distribution.jl:182; Julia's Base.log / Base.^ are not in the reference tree).  Away from 1 this
 * is the table algorithm of orc_log_tail (absolute error ~2^-68, so relative error < 2^-59 because |log w| >= 2^-7.02
 * there); for w in [1 - 1/128, 1 + 1/128) the centre is c = 1: r = w - 1 exactly and log w = r + r^2 P(r) with ten series
 * terms (relative error < 2^-60).  Rounded to one double the result is correctly rounded except with probability ~1e-3. */
static const double LOG1P_C[10] = {-0.5, 0x1.5555555555555p-2, -0.25, 0x1.999999999999ap-3, -0x1.5555555555555p-3,
                                   0x1.2492492492492p-3, -0.125, 0x1.c71c71c71c71cp-4, -0x1.999999999999ap-4,
                                   0x1.745d1745d1746p-4};
ORC_API double orc_log_pos(double w, double *lo_out) {
    uint64_t bits = double_as_u64(w);
    int e = (int)(bits >> 52) - 1023;
    int i = (int)((bits >> 45) & 127);
    double hi, lo;
    if ((e == 0 && i == 0) || (e == -1 && i >= 126)) {
        double r = w - 1.0; /* exact */
        double q = LOG1P_C[9];
        for (int k = 8; k >= 0; --k) q = fma(q, r, LOG1P_C[k]);
        q = (r * r) * q;
        hi = r + q;
        lo = (r - hi) + q; /* Fast2Sum, |r| >= |q| */
    } else {
        double m = u64_as_double((bits & 0x000fffffffffffffull) | 0x3ff0000000000000ull);
        double invc = LOGT[3 * i], logc_hi = LOGT[3 * i + 1], logc_lo = LOGT[3 * i + 2];
        double p_hi = m * invc;
        double p_lo = fma(m, invc, -p_hi);
        double r_hi = p_hi - 1.0;
        double r = r_hi + p_lo;
        double q = -0.125;
        q = fma(q, r, 0x1.2492492492492p-3);
        q = fma(q, r, -0x1.5555555555555p-3);
        q = fma(q, r, 0.2);
        q = fma(q, r, -0.25);
        q = fma(q, r, 0x1.5555555555555p-2);
        q = fma(q, r, -0.5);
        q = (r * r) * q;
        double ed = (double)e;
        double A = ed * MLE_LOGT_LN2_HI + logc_hi; /* exact; |A| >= 2^-7.02 > |r_hi| outside the branch above */
        double s = A + r_hi;
        double err = r_hi - (s - A);
        double l = fma(ed, MLE_LOGT_LN2_LO, logc_lo) + p_lo;
        l = l + q;
        l = l + err;
        hi = s + l;
        lo = (s - hi) + l;
    }
    if (lo_out) *lo_out = lo;
    return hi;
}

/* L^p for a positive normal double L: exp(p log L) with the product carried as th + tl (two-product of p and the hi part, p
 * times the lo part), the exponential of th by orc_exp and the factor 1 + tl applied with one fused multiply-add.
 * Error < 1 ulp (tests/test_oracle_golden.py checks it against mpmath). */
ORC_API double orc_exp(double x);
ORC_API double orc_pow_pos(double L, double p) {
    double ll, lh = orc_log_pos(L, &ll);
    double th = p * lh;
    double tl = fma(p, lh, -th) + p * ll;
    double E = orc_exp(th);
    return fma(E, tl, E);
}

/* exp(x): k = rint(x log2 e), r = x - k ln2 (two-step), degree-13 Taylor polynomial, scaling by 2^k in two exact-then-
 * rounded steps.  Error < 1 ulp; the lognormal / expsum models are DEFINED with this exp on both sides. */
ORC_API double orc_exp(double x) {
    if (x != x) return x;
    if (x > 710.0) x = 710.0;
    if (x < -746.0) x = -746.0;
    double kd = rint(x * 0x1.71547652b82fep+0);
    double r = fma(kd, -MLE_LOGT_LN2_HI, x);
    r = fma(kd, -MLE_LOGT_LN2_LO, r);
    double p = 0x1.6124613a86d09p-33;      /* 1/13! */
    p = fma(p, r, 0x1.1eed8eff8d898p-29);  /* 1/12! */
    p = fma(p, r, 0x1.ae64567f544e4p-26);  /* 1/11! */
    p = fma(p, r, 0x1.27e4fb7789f5cp-22);  /* 1/10! */
    p = fma(p, r, 0x1.71de3a556c734p-19);  /* 1/9! */
    p = fma(p, r, 0x1.a01a01a01a01ap-16);  /* 1/8! */
    p = fma(p, r, 0x1.a01a01a01a01ap-13);  /* 1/7! */
    p = fma(p, r, 0x1.6c16c16c16c17p-10);  /* 1/6! */
    p = fma(p, r, 0x1.1111111111111p-7);   /* 1/5! */
    p = fma(p, r, 0x1.5555555555555p-5);   /* 1/4! */
    p = fma(p, r, 0x1.5555555555555p-3);   /* 1/3! */
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    int k = (int)kd, k1 = k / 2, k2 = k - k1;
    double s1 = u64_as_double((uint64_t)(k1 + 1023) << 52), s2 = u64_as_double((uint64_t)(k2 + 1023) << 52);
    return (p * s1) * s2;
}

/* ------------------------------------------------------------------------------------------------ */
/* stage 1: lattice points                                                                          */
/* ------------------------------------------------------------------------------------------------ */

/* src/lattice.jl:178  reversebits(n) = parse(U, reverse(bitstring(n)), base=2) */
ORC_API uint32_t orc_reversebits32(uint32_t v) {
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}

/* src/lattice.jl:143 / :169   phi = reversebits(xor(k, k >> 1)) */
ORC_API uint32_t orc_phi(uint32_t k) { return orc_reversebits32(k ^ (k >> 1)); }

/* One coordinate.  src/lattice.jl:144 (shift == NULL):  phi * 2.0^(-32) * z .% 1
 *                  src/lattice.jl:170:                  (phi * 2.0^(-32) * z .+ shift) .% 1
 * phi*2^-32 is exact; the product with z is rounded once; the shift is added to the UNREDUCED product
 * and rounded again; `% 1` on Float64 is fmod (exact). */
static inline double orc_coord(uint32_t phi, uint32_t zj, const double *shiftj) {
    double t = (double)phi * 0x1p-32;
    double p = t * (double)zj; /* separately rounded (built with -ffp-contract=off) */
    if (shiftj) {
        double s = p + *shiftj;
        return fmod(s, 1.0);
    }
    return fmod(p, 1.0);
}

/* get_point(rule, k) for one (shift, k): x[0..s-1].  Returns -1 when k > nmax (the reference then returns
 * rand(s), src/lattice.jl:166,175 -- not reproducible, treated as an argument error by the engine). */
ORC_API int orc_get_point(const uint32_t *z, int s, uint64_t nmax, const double *shift, uint64_t k, double *x) {
    if (k > nmax || k > 0xffffffffull) return -1;
    uint32_t phi = orc_phi((uint32_t)k);
    for (int j = 0; j < s; ++j) x[j] = orc_coord(phi, z[j], shift ? shift + j : NULL);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* stage 2: transforms                                                                              */
/* ------------------------------------------------------------------------------------------------ */

/* @horner(t, c0, c1, ..., cn): acc = cn; acc = muladd(t, acc, c_{n-1}) ... (fused on every FMA machine) */
static inline double horner(double t, const double *c, int n) {
    double acc = c[n - 1];
    for (int i = n - 2; i >= 0; --i) acc = fma(t, acc, c[i]);
    return acc;
}

static const double P1[7] = {0.16030495584406622931e2, -0.9078495926296032665e2, 0.18644914861620987391e3,
                             -0.16900142734642382420e3, 0.654546628479448704e2, -0.86421301158724779e1,
                             0.17605878213905900};
static const double Q1[7] = {0.14780647071513831611e2, -0.9137416702426031394e2, 0.21015790486205317714e3,
                             -0.22210254121855132366e3, 0.10760453916055123830e3, -0.20601073032826544e2, 1.0};
static const double P2[8] = {-0.15238926344072612e-1, 0.3444556924136125216, -0.29344398672542478687e1,
                             0.11763505705217827302e2, -0.22655292823101104193e2, 0.19121334396580330163e2,
                             -0.5478927619598318769e1, 0.237516689024448};
static const double Q2[8] = {-0.10846516960205995e-1, 0.2610628885843078511, -0.24068318104393757995e1,
                             0.10695129973387014469e2, -0.23716715521596581025e2, 0.24640158943917284883e2,
                             -0.10014376349783070835e2, 1.0};
static const double P3[9] = {0.10501311523733438116e-3, 0.1053261131423333816425e-1, 0.26987802736243283544516,
                             0.23268695788919690806414e1, 0.71678547949107996810001e1, 0.85475611822167827825185e1,
                             0.68738088073543839802913e1, 0.3627002483095870893002e1, 0.886062739296515468149};
static const double Q3[10] = {0.10501266687030337690e-3, 0.1053286230093332753111e-1, 0.27019862373751554845553,
                              0.23501436397970253259123e1, 0.76078028785801277064351e1, 0.111815861040569078273451e2,
                              0.119487879184353966678438e2, 0.81922409747269907893913e1, 0.4099387907636801536145e1,
                              1.0};

/* SpecialFunctions.erfinv(::Float64) -- see the header comment; |x| > 1 returns NaN here (DomainError there). */
ORC_API double orc_erfinv(double x) {
    double a = fabs(x);
    if (a >= 1.0) {
        if (x == 1.0) return INFINITY;
        if (x == -1.0) return -INFINITY;
        return NAN;
    }
    if (a <= 0.75) {
        double xx = x * x;
        double t = xx - 0.5625;
        double num = x * horner(t, P1, 7);
        return num / horner(t, Q1, 7);
    }
    if (a <= 0.9375) {
        double xx = x * x;
        double t = xx - 0.87890625;
        double num = x * horner(t, P2, 8);
        return num / horner(t, Q2, 8);
    }
    double t = 1.0 / sqrt(-orc_log_tail(1.0 - a)); /* log1p(-a): 1 - a is exact for a >= 1/2 */
    double den = copysign(t, x) * horner(t, Q3, 10);
    return horner(t, P3, 9) / den;
}

/* src/distribution.jl:151   Phi^-1(x) = sqrt(2) * erfinv(2*x - 1) */
ORC_API double orc_norminv(double x) {
    double u = 2.0 * x - 1.0; /* 2*x is exact, so one rounding either way */
    return 1.4142135623730951 * orc_erfinv(u);
}

/* src/distribution.jl:153   Phi(x) = 1/2*(1 + erf(x/sqrt(2))) */
ORC_API double orc_normcdf(double x) {
    double q = x / 1.4142135623730951;
    double e = 1.0 + erf(q);
    return 0.5 * e;
}

enum { ORC_UNIFORM = 0, ORC_NORMAL = 1, ORC_TRUNCNORMAL = 2, ORC_WEIBULL = 3 };

/* transform(D, x).  p = {a,b} | {mu,sigma} | {mu,sigma,a,b} | {k,lambda}.  Every operation individually rounded
 * (Julia does not contract a + b*c). */
ORC_API double orc_transform(int kind, const double *p, double x) {
    switch (kind) {
    case ORC_UNIFORM: { /* :84  a + (b - a)*x */
        double w = p[1] - p[0];
        double wx = w * x;
        return p[0] + wx;
    }
    case ORC_NORMAL: { /* :112  mu + sigma*Phi^-1(x) */
        double sy = p[1] * orc_norminv(x);
        return p[0] + sy;
    }
    case ORC_TRUNCNORMAL: { /* :145-149 */
        double al = (p[2] - p[0]) / p[1];
        double be = (p[3] - p[0]) / p[1];
        double Fa = orc_normcdf(al);
        double Fb = orc_normcdf(be);
        double d = Fb - Fa;
        double dx = d * x;
        double arg = Fa + dx;
        double sy = p[1] * orc_norminv(arg);
        return p[0] + sy;
    }
    case ORC_WEIBULL: { /* :182  lambda * (-log(1 - x))^(1/k) */
        double om = 1.0 - x;
        double l = -orc_log_pos(om, 0);
        double ik = 1.0 / p[0];
        double pw = (l > 0.0) ? orc_pow_pos(l, ik) : 0.0; /* x = 0: log 1 = 0, 0^(1/k) = 0 */
        return p[1] * pw;
    }
    default:
        return x; /* identity: raw points */
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* MC uniforms: counter-based SplitMix64.  The reference draws rand(m) from Julia's task-local Xoshiro     */
/* (src/sample.jl:42) which is not reproducible outside Julia; the engine defines its own counter stream.  */
/* u(seed, k, j) = ((mix(seed + GOLDEN*(k*m + j + 1)) >> 12) + 0.5) * 2^-52  in (0, 1).                     */
/* ------------------------------------------------------------------------------------------------ */
ORC_API uint64_t orc_splitmix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

ORC_API double orc_mc_uniform(uint64_t seed, uint64_t k, uint32_t m, uint32_t j) {
    uint64_t ctr = k * (uint64_t)m + j + 1ull;
    uint64_t v = orc_splitmix64(seed + 0x9e3779b97f4a7c15ull * ctr);
    return ((double)(v >> 12) + 0.5) * 0x1p-52;
}

/* ------------------------------------------------------------------------------------------------ */
/* points for a whole (shift, k) rectangle: the parity hook for stages 1-2                             */
/* out[(r*n + i)*m + j],  k = k0 + i.  kinds == NULL -> raw points.  z == NULL -> MC uniforms (R = 1). */
/* ------------------------------------------------------------------------------------------------ */
ORC_API int orc_points(const uint32_t *z, int s, uint64_t nmax, const int *kinds, const double *pars /*4 per dim*/,
                       int m, const double *shifts /* [R][s] or NULL */, int R, uint64_t k0, uint64_t n,
                       uint64_t mc_seed, double *out) {
    if (z && m > s) return -2; /* reference pads with rand(): not reproducible (src/sample.jl:85) */
    if (z && n && (k0 + n - 1 > nmax || k0 + n - 1 > 0xffffffffull)) return -1;
    for (int r = 0; r < R; ++r) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            uint64_t k = k0 + (uint64_t)i;
            double *o = out + ((size_t)r * n + (size_t)i) * m;
            uint32_t phi = z ? orc_phi((uint32_t)k) : 0;
            for (int j = 0; j < m; ++j) {
                double x = z ? orc_coord(phi, z[j], shifts ? shifts + (size_t)r * s + j : NULL)
                             : orc_mc_uniform(mc_seed, k, (uint32_t)m, (uint32_t)j);
                o[j] = kinds ? orc_transform(kinds[j], pars + 4 * j, x) : x;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* stage 3: sample functions (the device-model registry, restated on the CPU)                        */
/* ------------------------------------------------------------------------------------------------ */
enum { ORC_MODEL_EXPSUM = 0, ORC_MODEL_PROBLEM18 = 1, ORC_MODEL_LOGNORMAL1D = 2, ORC_MODEL_LOGNORMAL2D = 3 };

#define ORC_MAX_LEVELS 256 /* levels, or packed multi-indices lx + 16*ly */
typedef struct {
    int kind;
    int nqoi;
    int ndims;     /* dimension of the index (1 = multilevel, 2 = multi-index) */
    int n0;        /* coarsest number of cells per direction (lognormal) / coarsest active dims (expsum) */
    int nweights;  /* expsum */
    double *weights;
    /* lognormal: KL basis per level, row-major [nodes][m] */
    int basis_rows[ORC_MAX_LEVELS];
    int basis_cols[ORC_MAX_LEVELS];
    double *basis[ORC_MAX_LEVELS];
} orc_model;

ORC_API orc_model *orc_model_create(int kind, int ndims, const double *params, int nparams) {
    orc_model *M = (orc_model *)calloc(1, sizeof(orc_model));
    M->kind = kind;
    M->ndims = ndims;
    M->nqoi = 1;
    if (kind == ORC_MODEL_EXPSUM) { /* params = {m0, a_1 ... a_m} */
        M->n0 = (int)params[0];
        M->nweights = nparams - 1;
        M->weights = (double *)malloc(sizeof(double) * (size_t)(nparams - 1));
        memcpy(M->weights, params + 1, sizeof(double) * (size_t)(nparams - 1));
    } else if (kind == ORC_MODEL_PROBLEM18) {
        M->nqoi = 13;
    } else { /* lognormal: params = {n0} */
        M->n0 = nparams > 0 ? (int)params[0] : (kind == ORC_MODEL_LOGNORMAL2D ? 4 : 16);
    }
    return M;
}

ORC_API int orc_model_set_basis(orc_model *M, int level, int rows, int cols, const double *phi) {
    if (level < 0 || level >= ORC_MAX_LEVELS) return -1;
    free(M->basis[level]);
    M->basis[level] = (double *)malloc(sizeof(double) * (size_t)rows * cols);
    memcpy(M->basis[level], phi, sizeof(double) * (size_t)rows * cols);
    M->basis_rows[level] = rows;
    M->basis_cols[level] = cols;
    return 0;
}

ORC_API void orc_model_destroy(orc_model *M) {
    if (!M) return;
    free(M->weights);
    for (int l = 0; l < ORC_MAX_LEVELS; ++l) free(M->basis[l]);
    free(M);
}

ORC_API int orc_model_nqoi(const orc_model *M) { return M->nqoi; }

/* --- expsum: Q_l(xi) = exp(sum_{j < m_l} a_j xi_j), m_l = min(m, m0 * 2^l); dQ = Q_l - Q_{l-1} ---------- */
static double expsum_q(const orc_model *M, int level, const double *xi, int m) {
    long ml = (long)M->n0 << level;
    if (ml > m) ml = m;
    if (ml > M->nweights) ml = M->nweights;
    double acc = 0.0;
    for (long j = 0; j < ml; ++j) acc = fma(M->weights[j], xi[j], acc);
    return orc_exp(acc);
}

/* --- problem_18: test/regression.jl:17-36 (1-d), :69-103 (2-d) --------------------------------------- */
static double p18_coef(int l, double xj) { /* the level-dependent factor of xi^3 */
    if (l == 3) return 1.0;
    if (l == 2) return 1.1;
    if (l == 1) {
        double q = xj / 60.0;
        return q + 1.2;
    }
    return 1.5; /* 3/2 */
}

static void problem18_eval(int ndims, const int *idx, const double *xi, double *f) {
    const int n = 13;
    double x1sq = xi[0] * xi[0];
    double x13 = x1sq * xi[0]; /* xi * xi * xi, left to right */
    double x23 = 0.0;
    if (ndims == 2) {
        double x2sq = xi[1] * xi[1];
        double t = x2sq * xi[1];
        x23 = t;
    }
    for (int j = 0; j < n; ++j) {
        /* range(0, 6, 13)[j] = j * 0.5 exactly */
        double xj = 0.5 * j;
        /* f_j = x_j <= 3 ? (x_j - 2)^2 : 2*log(x_j - 2) + 1: thirteen constants, tabulated with correctly rounded logs */
        static const double P18_BASE[13] = MLE_P18_BASE_INIT;
        double fj = P18_BASE[j];
        double t1 = p18_coef(idx[0], xj) * x13;
        double acc = fj + t1;
        if (ndims == 2) {
            double t2 = p18_coef(idx[1], xj) * x23;
            acc = acc + t2;
            if (idx[0] > 0 && idx[1] > 0) {
                double c = 0.03125 * x13; /* 1/32 * xi13 * xi23, left to right */
                double t3 = c * x23;
                acc = acc + t3;
            }
        }
        f[j] = acc;
    }
}

/* --- lognormal diffusion, 1-D:  -(a u')' = 1 on (0,1), u(0) = u(1) = 0 ------------------------------------
 * level l: N = n0 * 2^l cells, nodes 0..N; log a at node i = sum_k Phi_l[i][k] xi_k (k ascending, fused);
 * interface coefficient kappa_{i+1/2} = 0.5*(a_i + a_{i+1});
 * (kappa_{i-1/2} + kappa_{i+1/2}) u_i - kappa_{i-1/2} u_{i-1} - kappa_{i+1/2} u_{i+1} = h^2, i = 1..N-1, h = 1/N;
 * QoI = u at node N/2, in closed form: the discrete flux F_{i+1/2} = kappa_{i+1/2} (u_{i+1} - u_i) / h drops by h per cell,
 * so with r_i = 1/kappa_{i+1/2} and the distance w_i = |i + 1/2 - N/2| of edge i from the mid node (in cells)
 *     u_{N/2} = h^2 (D_R A_L + D_L A_R) / (A_L + A_R),   A = sum r_i,  D = sum w_i r_i  over the left / right half
 * -- the exact solution of the tridiagonal system above (all terms positive: no cancellation).  Both halves are summed from
 * the boundary towards the mid node, one edge at a time (A += r; D = fma(w, r, D)): the only sequential dependence per edge is
 * one addition, the divisions are independent -- on the GPU the two running sums replace the division-latency-bound
 * elimination recurrence.  The coarse problem uses the even nodes of the fine field (docs/src/example.md:120) with N/2 cells. */
static double tridiag_mid(const double *a /* node values, stride st */, int st, int N) {
    int M = N / 2;
    double h2 = 1.0 / ((double)N * (double)N);
    double AL = 0.0, DL = 0.0, AR = 0.0, DR = 0.0;
    /* left half: nodes 0 .. M arrive in ascending order; edge (ip-1, ip) has weight M - ip + 1/2 */
    for (int ip = 1; ip <= M; ++ip) {
        double kap = 0.5 * (a[(size_t)(ip - 1) * st] + a[(size_t)ip * st]);
        double r = 1.0 / kap;
        AL = AL + r;
        DL = fma((double)(M - ip) + 0.5, r, DL);
    }
    /* right half: nodes N .. M arrive in descending order (mirrored node number ip = N - i); same weights */
    for (int ip = 1; ip <= M; ++ip) {
        double kap = 0.5 * (a[(size_t)(N - ip + 1) * st] + a[(size_t)(N - ip) * st]);
        double r = 1.0 / kap;
        AR = AR + r;
        DR = fma((double)(M - ip) + 0.5, r, DR);
    }
    double num = fma(DR, AL, DL * AR);
    return h2 * (num / (AL + AR));
}

/* full Thomas solve (all unknowns) -- used by tests to cross-check the twisted mid value */
ORC_API double orc_tridiag_mid_full(const double *a, int N) {
    int n = N - 1;
    if (n < 1) return 0.0;
    double *c = (double *)calloc((size_t)n, sizeof(double)), *d = (double *)calloc((size_t)n, sizeof(double));
    double h2 = 1.0 / ((double)N * (double)N);
    for (int i = 1; i <= n; ++i) {
        double km = 0.5 * (a[i - 1] + a[i]), kp = 0.5 * (a[i] + a[i + 1]);
        double b = km + kp, lo = -km, up = -kp;
        if (i == 1) { c[0] = up / b; d[0] = h2 / b; }
        else {
            double den = b - lo * c[i - 2];
            c[i - 1] = up / den;
            d[i - 1] = (h2 - lo * d[i - 2]) / den;
        }
    }
    double *u = (double *)malloc(sizeof(double) * (size_t)n);
    u[n - 1] = d[n - 1];
    for (int i = n - 2; i >= 0; --i) u[i] = d[i] - c[i] * u[i + 1];
    double r = u[N / 2 - 1];
    free(c); free(d); free(u);
    return r;
}

ORC_API double orc_tridiag_mid(const double *a, int st, int N) { return tridiag_mid(a, st, N); }

static int lognormal1d_eval(const orc_model *M, int level, const double *xi, int m, double *dQ, double *Qf,
                            double *work /* >= nodes */) {
    const double *Phi = M->basis[level];
    if (!Phi) return -3;
    int N = M->n0 << level;
    if (M->basis_rows[level] != N + 1 || M->basis_cols[level] < m) return -3;
    int ld = M->basis_cols[level];
    for (int i = 0; i <= N; ++i) {
        double g = 0.0;
        const double *row = Phi + (size_t)i * ld;
        for (int k = 0; k < m; ++k) g = fma(row[k], xi[k], g);
        work[i] = orc_exp(g);
    }
    double qf = tridiag_mid(work, 1, N);
    *Qf = qf;
    *dQ = qf;
    if (level > 0) {
        double qc = tridiag_mid(work, 2, N / 2);
        *dQ = qf - qc;
    }
    return 0;
}

/* --- lognormal diffusion, 2-D:  -div(a grad u) = 1 on (0,1)^2, u = 0 on the boundary ---------------------------------
 * level l: N = n0 * 2^l cells per direction, nodes (i,j), i,j = 0..N (node number j*(N+1) + i); log a at a node =
 * sum_k Phi_l[node][k] xi_k (k ascending, fused); edge coefficients = arithmetic mean of the two node values; 5-point
 * stencil; QoI = u at the centre node (N/2, N/2); the coarse problem uses the even nodes (docs/src/example.md:120:
 * `view(af, 2:2:end, 2:2:end)`).  The centre value comes from a FRONTAL (banded, fill-in carrying) Gaussian elimination in
 * two independent half sweeps -- lines 1..N/2-1 upwards and, on the y-mirrored field, lines N-1..N/2+1 downwards --
 * followed by a dense two-sided elimination along the centre line.  Only a sliding (n+1)x(n+1) window is ever stored
 * (n = N-1); every entry is updated by fma(-l_q, v_r, M[q][r]) in pivot order, so there is no reduction whose order a
 * parallel implementation could change: the CUDA kernel reproduces the result bit for bit. */
typedef struct {
    const double *a; /* node values, row j (y) has stride lda */
    int lda, stx, sty, Nx, Ny, mirror;
    double wx, wy; /* 1/hx^2 = Nx^2, 1/hy^2 = Ny^2 */
} field2d;
static inline double F2(const field2d *f, int i, int j) {
    int jj = f->mirror ? f->Ny - j : j;
    return f->a[(size_t)(jj * f->sty) * f->lda + (size_t)i * f->stx];
}
/* scaled edge coefficients: x-edge (i,j)-(i+1,j) and y-edge (i,j)-(i,j+1); every product and sum rounded on its own */
static inline double KX(const field2d *f, int i, int j) { double k = 0.5 * (F2(f, i, j) + F2(f, i + 1, j)); return f->wx * k; }
static inline double KY(const field2d *f, int i, int j) { double k = 0.5 * (F2(f, i, j) + F2(f, i, j + 1)); return f->wy * k; }
static inline double DIAG2(const field2d *f, int i, int j) {
    return ((KX(f, i - 1, j) + KX(f, i, j)) + KY(f, i, j - 1)) + KY(f, i, j);
}

/* one half sweep over the lines 1..c-1 (c = Ny/2); on return S[n*n], gs[n] hold this half's centre-line block
 * (row-major, unknowns i = 1..n, n = Nx-1) */
static void frontal_half(const field2d *f, double *M, double *g, double *l, double *v, double *S, double *gs) {
    const int n = f->Nx - 1, c = f->Ny / 2, W = n + 1;
    const int P = (c - 1) * n;     /* pivots of this half */
    const int last = P + n - 1;    /* last unknown (end of the centre line) */
#define SLOT(e) ((e) % W)
#define LOAD_UNKNOWN(e)                                                                          \
    do {                                                                                         \
        int ie_ = (e) % n + 1, je_ = (e) / n + 1, se_ = SLOT(e);                                 \
        int centre_ = (je_ == c);                                                                \
        int own_ = !(centre_ && f->mirror); /* the centre line's own entries are loaded once */  \
        for (int t_ = 0; t_ < W; ++t_) { M[se_ * W + t_] = 0.0; M[t_ * W + se_] = 0.0; }         \
        M[se_ * W + se_] = own_ ? DIAG2(f, ie_, je_) : 0.0;                                      \
        g[se_] = own_ ? 1.0 : 0.0;                                                               \
        if (ie_ > 1 && own_) { double k_ = -KX(f, ie_ - 1, je_); M[se_ * W + SLOT((e)-1)] = k_; M[SLOT((e)-1) * W + se_] = k_; } \
        if (je_ > 1) { double k_ = -KY(f, ie_, je_ - 1); M[se_ * W + SLOT((e)-n)] = k_; M[SLOT((e)-n) * W + se_] = k_; }         \
    } while (0)
    for (int e = 0; e < n && e <= last; ++e) LOAD_UNKNOWN(e);
    for (int p = 0; p < P; ++p) {
        if (p + n <= last) LOAD_UNKNOWN(p + n);
        const int sp = SLOT(p);
        const int hi = (p + n <= last) ? p + n : last;
        const double rd = 1.0 / M[sp * W + sp], gp = g[sp]; /* multipliers l = v * (1/d): one division per pivot */
        for (int q = p + 1; q <= hi; ++q) { v[SLOT(q)] = M[SLOT(q) * W + sp]; l[SLOT(q)] = v[SLOT(q)] * rd; }
        for (int q = p + 1; q <= hi; ++q) {
            const int sq = SLOT(q);
            const double lq = l[sq];
            g[sq] = fma(-lq, gp, g[sq]);
            for (int r = p + 1; r <= hi; ++r) M[sq * W + SLOT(r)] = fma(-lq, v[SLOT(r)], M[sq * W + SLOT(r)]);
        }
    }
    for (int q = 0; q < n; ++q) {
        gs[q] = g[SLOT(P + q)];
        for (int r = 0; r < n; ++r) S[q * n + r] = M[SLOT(P + q) * W + SLOT(P + r)];
    }
#undef LOAD_UNKNOWN
#undef SLOT
}

/* u at the centre node (Nx/2, Ny/2) of the Nx x Ny-cell grid whose node values are a[(j*sty)*lda + i*stx] */
static double frontal_mid(const double *a, int lda, int stx, int sty, int Nx, int Ny) {
    const int n = Nx - 1, W = n + 1, tc = Nx / 2 - 1;
    double *M = (double *)malloc(sizeof(double) * (size_t)W * W), *g = (double *)malloc(sizeof(double) * (size_t)W);
    double *l = (double *)malloc(sizeof(double) * (size_t)W), *v = (double *)malloc(sizeof(double) * (size_t)W);
    double *S = (double *)malloc(sizeof(double) * (size_t)n * n), *gs = (double *)malloc(sizeof(double) * (size_t)n);
    double *S2 = (double *)malloc(sizeof(double) * (size_t)n * n), *gs2 = (double *)malloc(sizeof(double) * (size_t)n);
    field2d f = {a, lda, stx, sty, Nx, Ny, 1, (double)Nx * (double)Nx, (double)Ny * (double)Ny};
    frontal_half(&f, M, g, l, v, S2, gs2); /* upper half first (y-mirrored field) */
    f.mirror = 0;
    frontal_half(&f, M, g, l, v, S, gs);   /* lower half */
    for (int q = 0; q < n; ++q) {
        gs[q] = gs[q] + gs2[q];
        for (int r = 0; r < n; ++r) S[q * n + r] = S[q * n + r] + S2[q * n + r];
    }
    /* dense two-sided elimination along the centre line: t = 0..tc-1 ascending, then t = n-1..tc+1 descending */
    for (int t = 0; t < tc; ++t) {
        const double rd = 1.0 / S[t * n + t], gp = gs[t];
        for (int q = t + 1; q < n; ++q) { v[q] = S[q * n + t]; l[q] = v[q] * rd; }
        for (int q = t + 1; q < n; ++q) {
            gs[q] = fma(-l[q], gp, gs[q]);
            for (int r = t + 1; r < n; ++r) S[q * n + r] = fma(-l[q], v[r], S[q * n + r]);
        }
    }
    for (int t = n - 1; t > tc; --t) {
        const double rd = 1.0 / S[t * n + t], gp = gs[t];
        for (int q = tc; q < t; ++q) { v[q] = S[q * n + t]; l[q] = v[q] * rd; }
        for (int q = tc; q < t; ++q) {
            gs[q] = fma(-l[q], gp, gs[q]);
            for (int r = tc; r < t; ++r) S[q * n + r] = fma(-l[q], v[r], S[q * n + r]);
        }
    }
    const double u = gs[tc] / S[tc * n + tc];
    free(M); free(g); free(l); free(v); free(S); free(gs); free(S2); free(gs2);
    return u;
}

ORC_API double orc_frontal_mid(const double *a, int lda, int st, int N) { return frontal_mid(a, lda, st, st, N, N); }
ORC_API double orc_frontal_mid_rect(const double *a, int lda, int stx, int sty, int Nx, int Ny) {
    return frontal_mid(a, lda, stx, sty, Nx, Ny);
}

/* per-thread scratch that grows on demand: no allocation per sample in the timed CPU arm */
static double *orc_scratch(size_t n) {
    static __thread double *buf = NULL;
    static __thread size_t cap = 0;
    if (cap < n) {
        free(buf);
        buf = (double *)malloc(sizeof(double) * n);
        cap = buf ? n : 0;
    }
    return buf;
}

/* index = (l) for the multilevel model (Nx = Ny = n0 2^l) or (lx, ly) for the multi-index model (Nx = n0 2^lx,
 * Ny = n0 2^ly, anisotropic coarsening as in docs/src/example.md:221-243); the KL basis of an index is stored under the key
 * lx + 16*ly.  dQ = Qf + sum over diff(index) of +-Q_kappa, coarse fields = strided views of the fine field
 * (`step = (index - key).I .+ 1`), added in the order (lx-1, ly-1), (lx, ly-1), (lx-1, ly). */
static int lognormal2d_eval(const orc_model *M, const int *index, const double *xi, int m, double *dQ, double *Qf) {
    const int lx = index[0], ly = M->ndims == 2 ? index[1] : index[0];
    const int key = M->ndims == 2 ? lx + 16 * ly : lx;
    if (key < 0 || key >= ORC_MAX_LEVELS) return -3;
    const double *Phi = M->basis[key];
    if (!Phi) return -3;
    const int Nx = M->n0 << lx, Ny = M->n0 << ly, nodes = (Nx + 1) * (Ny + 1);
    if (M->basis_rows[key] != nodes || M->basis_cols[key] < m) return -3;
    const int ld = M->basis_cols[key];
    double *a = orc_scratch((size_t)nodes);
    if (!a) return -5;
    for (int i = 0; i < nodes; ++i) {
        double gsum = 0.0;
        const double *row = Phi + (size_t)i * ld;
        for (int k = 0; k < m; ++k) gsum = fma(row[k], xi[k], gsum);
        a[i] = orc_exp(gsum);
    }
    const double qf = frontal_mid(a, Nx + 1, 1, 1, Nx, Ny);
    *Qf = qf;
    double dq = qf;
    if (M->ndims == 1) {
        if (lx > 0) dq = qf - frontal_mid(a, Nx + 1, 2, 2, Nx / 2, Ny / 2);
    } else {
        if (lx > 0 && ly > 0) dq = dq + frontal_mid(a, Nx + 1, 2, 2, Nx / 2, Ny / 2);
        if (ly > 0) dq = dq - frontal_mid(a, Nx + 1, 1, 2, Nx, Ny / 2);
        if (lx > 0) dq = dq - frontal_mid(a, Nx + 1, 2, 1, Nx / 2, Ny);
    }
    *dQ = dq;
    return 0;
}

/* f(index, xi) -> (dQ[nqoi], Qf[nqoi]).   dQ = Qf + sum_{(kappa, +-1) in diff(index)} +- Q_kappa  (src/index.jl:49-58) */
ORC_API int orc_model_eval(const orc_model *M, const int *index, const double *xi, int m, double *dQ, double *Qf) {
    switch (M->kind) {
    case ORC_MODEL_EXPSUM: {
        double qf = expsum_q(M, index[0], xi, m);
        Qf[0] = qf;
        dQ[0] = index[0] > 0 ? qf - expsum_q(M, index[0] - 1, xi, m) : qf;
        return 0;
    }
    case ORC_MODEL_PROBLEM18: {
        double f[13];
        problem18_eval(M->ndims, index, xi, Qf);
        for (int j = 0; j < 13; ++j) dQ[j] = Qf[j];
        /* diff(index): all I in [max(0, index-1), index] except index itself, sign -1 if |index - I|_1 odd.
         * Dict iteration order in Julia is unspecified; for d = 2 the three terms are added here in the order
         * (i-1, j-1), (i, j-1), (i-1, j)  [column-major CartesianIndices order]. */
        int lo0 = index[0] > 0 ? index[0] - 1 : 0;
        int lo1 = (M->ndims == 2) ? (index[1] > 0 ? index[1] - 1 : 0) : 0;
        int hi1 = (M->ndims == 2) ? index[1] : 0;
        for (int b = lo1; b <= hi1; ++b)
            for (int a = lo0; a <= index[0]; ++a) {
                if (a == index[0] && b == hi1) continue;
                int I[2] = {a, b};
                int dist = (index[0] - a) + (hi1 - b);
                double sgn = (dist & 1) ? -1.0 : 1.0;
                problem18_eval(M->ndims, I, xi, f);
                for (int j = 0; j < 13; ++j) {
                    double t = sgn * f[j];
                    dQ[j] = dQ[j] + t;
                }
            }
        return 0;
    }
    case ORC_MODEL_LOGNORMAL1D: {
        int N = M->n0 << index[0];
        return lognormal1d_eval(M, index[0], xi, m, dQ, Qf, orc_scratch((size_t)(N + 1)));
    }
    case ORC_MODEL_LOGNORMAL2D:
        return lognormal2d_eval(M, index, xi, m, dQ, Qf);
    default:
        return -4;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* stage 4: the whole batch, the way the reference runs it                                           */
/*   for (r, i) in product(1:R, Istart:Iend):  xi = transform.(D[1:m], get_point(gen[r], i-1)[1:m])    */
/*                                             (dQ, Qf) = f(index, xi)           src/sample.jl:79-102 */
/*   then per (shift, qoi): mean and corrected variance of the n values        src/methods.jl:44-57  */
/* moments[((r*nqoi + q)*2 + X)*3 + {0,1,2}] = {n, mean, M2 = sum (x - mean)^2},  X = 0: dQ, 1: Qf      */
/* samples (optional) [((r*nqoi + q)*2 + X)*n + i]                                                     */
/* Sums are carried in long double so the oracle is (much) closer to the exact statistics than 1e-12.  */
/* ------------------------------------------------------------------------------------------------ */
ORC_API int orc_sample_batch(const orc_model *M, const uint32_t *z, int s, uint64_t nmax, const int *kinds,
                             const double *pars, int m, const int *index, const double *shifts, int R, uint64_t k0,
                             uint64_t n, uint64_t mc_seed, double *moments, double *samples_or_null) {
    int q = M->nqoi;
    if (z && m > s) return -2;
    if (z && n && (k0 + n - 1 > nmax || k0 + n - 1 > 0xffffffffull)) return -1;
    size_t per = (size_t)2 * q * n;
    double *buf = samples_or_null ? samples_or_null : (double *)malloc(sizeof(double) * per * (size_t)R);
    if (!buf) return -5;
    int rc_all = 0;
#pragma omp parallel
    {
        double *xi = (double *)malloc(sizeof(double) * (size_t)m);
        double *dq = (double *)malloc(sizeof(double) * (size_t)q * 2);
        double *qf = dq + q;
        /* the reference farms out one (shift, k) item at a time (pmap, batch_size = 1, src/sample.jl:93-95) */
#pragma omp for schedule(dynamic, 1) collapse(2)
        for (int r = 0; r < R; ++r)
            for (int64_t i = 0; i < (int64_t)n; ++i) {
                uint64_t k = k0 + (uint64_t)i;
                uint32_t phi = z ? orc_phi((uint32_t)k) : 0;
                for (int j = 0; j < m; ++j) {
                    double x = z ? orc_coord(phi, z[j], shifts ? shifts + (size_t)r * s + j : NULL)
                                 : orc_mc_uniform(mc_seed, k, (uint32_t)m, (uint32_t)j);
                    xi[j] = kinds ? orc_transform(kinds[j], pars + 4 * j, x) : x;
                }
                int rc = orc_model_eval(M, index, xi, m, dq, qf);
                if (rc) {
#pragma omp atomic write
                    rc_all = rc;
                }
                for (int c = 0; c < q; ++c) {
                    buf[(((size_t)r * q + c) * 2 + 0) * n + (size_t)i] = dq[c];
                    buf[(((size_t)r * q + c) * 2 + 1) * n + (size_t)i] = qf[c];
                }
            }
        free(xi);
        free(dq);
    }
    if (rc_all) { if (!samples_or_null) free(buf); return rc_all; }
    for (size_t v = 0; v < (size_t)R * q * 2; ++v) {
        const double *x = buf + v * n;
        long double sum = 0.0L;
        for (uint64_t i = 0; i < n; ++i) sum += x[i];
        long double mean = n ? sum / (long double)n : 0.0L;
        /* corrected two-pass: the rounding of `mean` (2^-64 relative) would otherwise add n e^2, which matters when the
         * samples agree to 15+ digits */
        long double m2 = 0.0L, sd = 0.0L;
        for (uint64_t i = 0; i < n; ++i) { long double d = x[i] - mean; m2 += d * d; sd += d; }
        if (n) m2 -= sd * sd / (long double)n;
        if (m2 < 0.0L) m2 = 0.0L;
        moments[v * 3 + 0] = (double)n;
        moments[v * 3 + 1] = (double)mean;
        moments[v * 3 + 2] = (double)m2;
    }
    if (!samples_or_null) free(buf);
    return 0;
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Returns 1 when this object was compiled WITHOUT floating-point contraction (required: every fused operation in this
 * file is an explicit fma()).  a*b + c below differs between the fused and the separately rounded evaluation. */
ORC_API int orc_selftest_nocontract(void) {
    volatile double a = 1.0 + 0x1p-30, b = 1.0 - 0x1p-30, c = -1.0;
    double r = a * b + c; /* exact product 1 - 2^-60 rounds to 1.0 -> r == 0 when not contracted */
    return r == 0.0;
}
