/* A plain C99 consumer of libmle_cuda: the calls a MultilevelEstimators.jl binding makes for one `sample!` of a
 * multilevel QMC estimator (src/sample.jl:21-34, :79-102), written against include/mle_cuda.h only.
 *
 *   gcc -std=c99 -Wall -Wextra -pedantic -Iinclude examples/abi_example.c \
 *       -Lmultilevelestimators.jl_b200/lib -lmle_cuda -Wl,-rpath,$PWD/multilevelestimators.jl_b200/lib -lm -o abi_example
 *
 * Model: "expsum", Q_l = exp(sum_{j < m_l} a_j xi_j), a_j = 0.5/j, m_l = min(m, 4 * 2^l), xi ~ N(0,1): the level-l estimate
 * of E[Q_l] = exp(0.5 * sum a_j^2) is printed next to the exact value.  Exits 1 with the library's message when no
 * sm_100 device is present (there is no CPU fallback). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "mle_cuda.h"

#define M 64  /* random dimensions */
#define R 8   /* random shifts     */

static uint64_t sm64_state = 2024;
static double sm64_uniform(void) { /* SplitMix64 -> [0,1) */
    uint64_t z = (sm64_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

static int fail(mle_ctx ctx, const char *what, int rc) {
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, mle_last_error(ctx));
    if (ctx) mle_destroy(ctx);
    return 1;
}

int main(int argc, char **argv) {
    int level = argc > 1 ? atoi(argv[1]) : 2;
    uint64_t n = argc > 2 ? (uint64_t)strtoull(argv[2], NULL, 10) : 4096; /* points per shift */
    mle_ctx ctx = NULL;
    mle_lattice lat = NULL;
    mle_dists dists = NULL;
    mle_model model = NULL;
    const uint32_t *z = NULL;
    int zlen = 0, rc, j, r, q = 0;
    mle_dist d[M];
    double params[1 + M], shifts[R * M], moments[3 * 2 * 1 * R], acc[3 * 2 * 1 * R], seconds = 0.0;
    double mean_f = 0.0, mean_d = 0.0, varest = 0.0, exact = 0.0;
    int32_t index[1];
    int m_l;

    if (mle_abi_version() != MLE_ABI_VERSION) {
        fprintf(stderr, "header / library ABI mismatch\n");
        return 1;
    }
    if ((rc = mle_init(0, &ctx)) != MLE_OK) return fail(ctx, "mle_init", rc);

    if ((rc = mle_genvec("K_3600_32", &z, &zlen)) != MLE_OK || zlen < M) return fail(ctx, "mle_genvec", rc);
    if ((rc = mle_lattice_create(ctx, z, M, 0xFFFFFFFFull, &lat)) != MLE_OK) return fail(ctx, "mle_lattice_create", rc);
    for (j = 0; j < M; j++) {
        d[j].kind = MLE_NORMAL;
        d[j].reserved = 0;
        d[j].p[0] = 0.0;
        d[j].p[1] = 1.0;
        d[j].p[2] = d[j].p[3] = 0.0;
    }
    if ((rc = mle_dists_create(ctx, d, M, &dists)) != MLE_OK) return fail(ctx, "mle_dists_create", rc);
    params[0] = 4.0;
    for (j = 0; j < M; j++) params[1 + j] = 0.5 / (double)(j + 1);
    if ((rc = mle_model_create(ctx, "expsum", 1, params, 1 + M, &model)) != MLE_OK) return fail(ctx, "mle_model_create", rc);
    if ((rc = mle_model_nqoi(model, &q)) != MLE_OK || q != 1) return fail(ctx, "mle_model_nqoi", rc);
    for (j = 0; j < R * M; j++) shifts[j] = sm64_uniform(); /* shift r at shifts + r*M */

    /* two consecutive sample! calls on the same index: points 0..n-1, then n..2n-1, merged like the reference's append! */
    index[0] = level;
    if ((rc = mle_sample_batch(ctx, model, lat, dists, index, 1, shifts, R, 0, n, M, 0, acc, NULL, &seconds)) != MLE_OK)
        return fail(ctx, "mle_sample_batch", rc);
    if ((rc = mle_sample_batch(ctx, model, lat, dists, index, 1, shifts, R, n, n, M, 0, moments, NULL, NULL)) != MLE_OK)
        return fail(ctx, "mle_sample_batch", rc);
    if ((rc = mle_moments_merge(acc, moments, (size_t)(2 * 1 * R))) != MLE_OK) return fail(ctx, "mle_moments_merge", rc);

    /* mean(estimator, index) and the QMC varest of src/methods.jl:48-57, :282 from the (n, mean, M2) triplets */
    for (r = 0; r < R; r++) {
        mean_d += acc[((r * 1 + 0) * 2 + 0) * 3 + 1] / R;
        mean_f += acc[((r * 1 + 0) * 2 + 1) * 3 + 1] / R;
    }
    for (r = 0; r < R; r++) {
        double dev = acc[((r * 1 + 0) * 2 + 1) * 3 + 1] - mean_f;
        varest += dev * dev / (R - 1) / R;
    }
    m_l = 4 << level;
    if (m_l > M) m_l = M;
    for (j = 0; j < m_l; j++) exact += 0.5 * params[1 + j] * params[1 + j];
    exact = exp(exact);
    printf("level %d: %d shifts x %llu points, n per shift after merge = %.0f\n", level, R, (unsigned long long)(2 * n), acc[0]);
    printf("E[Q_l]  = %.12f  (exact %.12f, QMC standard error %.3e)\n", mean_f, exact, sqrt(varest));
    printf("E[dQ_l] = %.12f\n", mean_d);
    printf("device time of the first call: %.3f ms\n", 1e3 * seconds);

    mle_model_destroy(ctx, model);
    mle_dists_destroy(ctx, dists);
    mle_lattice_destroy(ctx, lat);
    mle_destroy(ctx);
    return fabs(mean_f - exact) < 6.0 * sqrt(varest) + 1e-9 ? 0 : 2;
}
