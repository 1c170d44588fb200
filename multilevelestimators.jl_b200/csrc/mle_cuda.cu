// mle_cuda.cu -- C ABI of libmle_cuda (see include/mle_cuda.h).  Host side: handles, argument checks mirroring the
// reference's ArgumentErrors, device buffers, kernel launches.  There is no CPU fallback anywhere in this file.
#include "../../include/mle_cuda.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "mle_kernels.cuh"
#include "mle_kernels2d.cuh"
#include "../../include/mle_genvec_data.inc"

using namespace mle;

// kernel instantiations per input mode (raw points / all Normal(0,1) / general distribution mix)
template <int MODE> static constexpr auto points_k = points_kernel<POINTS_THREADS, MODE>;
template <int MODE> static constexpr auto expsum_k = sample_simple_kernel<MODEL_EXPSUM, 2, MODE>;
template <int MODE> static constexpr auto problem18_k = sample_simple_kernel<MODEL_PROBLEM18, 26, MODE>;
template <int MODE> static constexpr auto lognormal1d_k16 = lognormal1d_kernel<MODE, 16>;
template <int MODE> static constexpr auto lognormal1d_kb = lognormal1d_block_kernel<MODE, 32>;
template <int MODE> static constexpr auto lognormal1d_ks = lognormal1d_block_kernel<MODE, 32, true>;
template <int MODE> static constexpr auto lognormal1d_kl = lognormal1d_small_kernel<MODE, false>;
template <int MODE> static constexpr auto lognormal1d_kc = lognormal1d_small_kernel<MODE, true>;
template <int MODE> static constexpr auto lognormal2d_w = lognormal2d_kernel<MODE, 32>;
template <int MODE> static constexpr auto lognormal2d_c = lognormal2d_kernel<MODE, 256>;
template <int MODE> static constexpr auto lognormal2d_r4 = lognormal2d_rows_kernel<MODE, 4>;
template <int MODE> static constexpr auto lognormal2d_r8 = lognormal2d_rows_kernel<MODE, 8>;
template <int MODE> static constexpr auto lognormal2d_r16 = lognormal2d_rows_kernel<MODE, 16>;
template <int MODE> static constexpr auto lognormal2d_r32 = lognormal2d_rows_kernel<MODE, 32>;
template <int MODE> static constexpr auto lognormal2d_r64 = lognormal2d_rows_kernel<MODE, 64>;
template <int MODE> static constexpr auto lognormal2d_r128 = lognormal2d_rows_kernel<MODE, 128>;
template <int MODE> static constexpr auto lognormal2d_d64 = lognormal2d_rows_kernel<MODE, 64, Dmma2dCfg<64>>;
template <int MODE> static constexpr auto lognormal2d_d128 = lognormal2d_rows_kernel<MODE, 128, Dmma2dCfg<128>>;
// mode = GEN_RAW | GEN_STDNORMAL | GEN_GENERAL, optionally | GEN_MC: six instantiations per kernel
#define MLE_LAUNCH_BY_MODE(mode, K, grid, block, smem, stream, ...)                          \
    do {                                                                                       \
        switch (mode) {                                                                        \
        case GEN_RAW: K<GEN_RAW><<<grid, block, smem, stream>>>(__VA_ARGS__); break;           \
        case GEN_STDNORMAL: K<GEN_STDNORMAL><<<grid, block, smem, stream>>>(__VA_ARGS__); break; \
        case GEN_GENERAL: K<GEN_GENERAL><<<grid, block, smem, stream>>>(__VA_ARGS__); break;   \
        case GEN_RAW | GEN_MC: K<GEN_RAW | GEN_MC><<<grid, block, smem, stream>>>(__VA_ARGS__); break; \
        case GEN_STDNORMAL | GEN_MC: K<GEN_STDNORMAL | GEN_MC><<<grid, block, smem, stream>>>(__VA_ARGS__); break; \
        default: K<GEN_GENERAL | GEN_MC><<<grid, block, smem, stream>>>(__VA_ARGS__); break;   \
        }                                                                                      \
    } while (0)

// launch with a thread-block cluster of `cs` CTAs along x
template <typename... KArgs, typename... Args>
static cudaError_t launch_cluster(void (*kernel)(KArgs...), dim3 grid, int threads, size_t smem, cudaStream_t stream, int cs, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}
#define MLE_LAUNCH_CLUSTER_BY_MODE(mode, K, grid, block, smem, stream, cs, ...)                                  \
    do {                                                                                                           \
        switch (mode) {                                                                                            \
        case GEN_RAW: launch_cluster(K<GEN_RAW>, grid, block, smem, stream, cs, __VA_ARGS__); break;               \
        case GEN_STDNORMAL: launch_cluster(K<GEN_STDNORMAL>, grid, block, smem, stream, cs, __VA_ARGS__); break;   \
        case GEN_GENERAL: launch_cluster(K<GEN_GENERAL>, grid, block, smem, stream, cs, __VA_ARGS__); break;       \
        case GEN_RAW | GEN_MC: launch_cluster(K<GEN_RAW | GEN_MC>, grid, block, smem, stream, cs, __VA_ARGS__); break; \
        case GEN_STDNORMAL | GEN_MC: launch_cluster(K<GEN_STDNORMAL | GEN_MC>, grid, block, smem, stream, cs, __VA_ARGS__); break; \
        default: launch_cluster(K<GEN_GENERAL | GEN_MC>, grid, block, smem, stream, cs, __VA_ARGS__); break;       \
        }                                                                                                          \
    } while (0)

// clusters of `cs` CTAs of (kernel, block size, dynamic shared memory) the device can keep resident at once (GPC granularity:
// a cluster lives inside one GPC); remembered like resident_ctas_per_sm
template <typename K>
static int resident_clusters(K kernel, int threads, size_t smem, int cs) {
    struct Entry { const void *fn; int threads; size_t smem; int cs; int n; };
    static std::vector<Entry> cache;
    const void *fn = reinterpret_cast<const void *>(kernel);
    for (const Entry &e : cache)
        if (e.fn == fn && e.threads == threads && e.smem == smem && e.cs == cs) return e.n;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)cs);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
    if (getenv("MLE_DEBUG")) fprintf(stderr, "[mle] resident clusters of %d CTAs (%d threads, %zu bytes): %d\n", cs, threads, smem, n);
    if (cache.size() < 256) cache.push_back(Entry{fn, threads, smem, cs, n});
    return n;
}
#define MLE_CLUSTERS_BY_MODE(mode, K, threads, smem, cs)                                            \
    ((mode) == GEN_RAW ? resident_clusters(K<GEN_RAW>, threads, smem, cs)                           \
     : (mode) == GEN_STDNORMAL ? resident_clusters(K<GEN_STDNORMAL>, threads, smem, cs)             \
     : (mode) == GEN_GENERAL ? resident_clusters(K<GEN_GENERAL>, threads, smem, cs)                 \
     : (mode) == (GEN_RAW | GEN_MC) ? resident_clusters(K<GEN_RAW | GEN_MC>, threads, smem, cs)     \
     : (mode) == (GEN_STDNORMAL | GEN_MC) ? resident_clusters(K<GEN_STDNORMAL | GEN_MC>, threads, smem, cs) \
                                          : resident_clusters(K<GEN_GENERAL | GEN_MC>, threads, smem, cs))

// resident CTAs per SM of (kernel, block size, dynamic shared memory); the query costs a few microseconds, so the answers are
// remembered (one calling thread per context, see include/mle_cuda.h)
template <typename K>
static int resident_ctas_per_sm(K kernel, int threads, size_t smem) {
    struct Entry { const void *fn; int threads; size_t smem; int n; };
    static std::vector<Entry> cache;
    const void *fn = reinterpret_cast<const void *>(kernel);
    for (const Entry &e : cache)
        if (e.fn == fn && e.threads == threads && e.smem == smem) return e.n;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) { (void)cudaGetLastError(); n = 1; }
    if (cache.size() < 256) cache.push_back(Entry{fn, threads, smem, n});
    return n;
}
#define MLE_OCC_BY_MODE(mode, K, threads, smem)                                             \
    ((mode) == GEN_RAW ? resident_ctas_per_sm(K<GEN_RAW>, threads, smem)                    \
     : (mode) == GEN_STDNORMAL ? resident_ctas_per_sm(K<GEN_STDNORMAL>, threads, smem)      \
     : (mode) == GEN_GENERAL ? resident_ctas_per_sm(K<GEN_GENERAL>, threads, smem)          \
     : (mode) == (GEN_RAW | GEN_MC) ? resident_ctas_per_sm(K<GEN_RAW | GEN_MC>, threads, smem) \
     : (mode) == (GEN_STDNORMAL | GEN_MC) ? resident_ctas_per_sm(K<GEN_STDNORMAL | GEN_MC>, threads, smem) \
                                          : resident_ctas_per_sm(K<GEN_GENERAL | GEN_MC>, threads, smem))

// A grid of exactly the resident CTAs, each striding over its share of the tiles, is NOT the fastest way to run these kernels:
// the resident CTAs of an SM advance at very different rates (profiles/r01_cta_timeline.txt: lifetimes of 7.4 .. 11.8 ms for
// identical work), the favoured ones finish early and the SM spends the last third of the launch below full occupancy.  With
// several times more CTAs than slots, each doing K tiles, the hardware block scheduler refills every slot as soon as it frees up:
// all slots stay busy to the end and the launch is 3-8 % shorter (measured: K = 16 best for long launches, about 8 "waves"
// for short ones; K = 2 costs more in per-CTA set-up than it gains).  Partials stay indexed by CTA, so the result is still
// independent of the execution order.
static long long churn_grid_x(long long gx_resident, long long tiles_per_shift) {
    if (getenv("MLE_PERSISTENT_GRID")) return gx_resident;  // A/B switch
    const long long per_cta = tiles_per_shift / std::max<long long>(1, gx_resident);  // tiles of a resident-grid CTA
    const long long K = std::min<long long>(16, std::max<long long>(1, per_cta / 8));
    long long gx = (tiles_per_shift + K - 1) / K;
    gx = std::min(gx, 64 * gx_resident);  // at most 64 waves: the per-CTA partials stay small for huge batches
    return std::max(gx, gx_resident);
}

// ---------------------------------------------------------------------------------------------------------------
// handles
// ---------------------------------------------------------------------------------------------------------------
struct mle_ctx_s {
    // multi-GPU context (mle_init_multi): this object is member 0, `peers` are the contexts of the other devices.  Handles
    // created through it carry one replica per peer; mle_sample_batch / mle_update_samples split each request over the members.
    std::vector<mle_ctx> peers;
    uint64_t shard_min = 24576;  // a member is only used if it gets at least this many (shift, point) samples (MLE_SHARD_MIN)
    int device = -1;
    int sm_count = 0;
    int cc = 0;
    size_t hbm = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    uint64_t launches = 0;
    int ln_staged = -1;      // MLE_LN_STAGED=0|1: force / forbid the TMA-staged small-batch variant of the phase-synchronous kernel
    int ln_cluster = 1;      // MLE_LN_CLUSTER=0: the latency variant never spreads a sample tile over a thread-block cluster; 2|4|8: largest cluster tried
    int ln_small = -1;       // MLE_LN_SMALL=0|1: forbid / force (when it fits) the latency variant for tiny batches (mle_kernel_ln1d_small.cuh)
    int ln_block = -1;       // MLE_LN_BLOCK=0|1: force the warp-autonomous / phase-synchronous lognormal1d kernel (default: by grid size)
    int ln2d_old = 0;        // MLE_2D_WINDOW_SMEM=1: force the shared-memory-window lognormal2d kernel (A/B timing)
    int ln2d_dmma = 1;       // MLE_2D_DMMA=0: register-window solver everywhere; 1: blocked DMMA solver at 128 cells; 2: also at 64
    // reusable device / pinned scratch
    Partial *d_partials = nullptr;
    size_t partials_cap = 0;  // in Partials
    unsigned *d_tickets = nullptr;  // [R + 1] per-shift CTA tickets + finished-shift count of cta_finalize; zero between launches
    size_t tickets_cap = 0;
    // results that the host reads without synchronising the stream: a mapped pinned moment table and a mapped flag the last
    // CTA of a launch sets to the call's sequence number (mle_sample_batch / mle_update_samples poll it)
    double *h_res = nullptr, *d_res = nullptr;
    size_t res_cap = 0;  // doubles
    unsigned *h_flag = nullptr, *d_flag = nullptr;
    unsigned seq = 0;
    int zero_copy = 1;  // MLE_NO_ZEROCOPY=1: classic device moments + cudaMemcpyAsync + stream synchronisation
    double *d_shifts = nullptr;
    size_t shifts_cap = 0;
    size_t shifts_staged = 0;  // doubles of h_pinned[0..] that mirror d_shifts (0: nothing valid)
    double *d_scratch = nullptr;  // lognormal2d: per-group field / centre blocks for the large grids
    size_t scratch_cap = 0;
    double *d_zero = nullptr;  // a row of zeros standing in for the shift of an unshifted rule
    size_t zero_cap = 0;
    double *d_moments = nullptr;
    size_t moments_cap = 0;
    double *d_batch = nullptr;  // mle_update_samples: shifts and moment tables of all entries
    size_t batch_cap = 0;
    size_t batch_staged = 0;    // doubles of h_pinned[0..] that mirror the shifts in d_batch (0: nothing valid)
    double *d_samples = nullptr;  // raw samples of a mle_sample_batch call that asked for them
    size_t samples_cap = 0;
    double *d_zacc = nullptr;     // mle_sample_unbiased: per-sample accumulator Z [R][q][n]
    size_t zacc_cap = 0;
    double *h_pinned = nullptr;
    size_t pinned_cap = 0;  // doubles
};

struct mle_lattice_s {
    std::vector<mle_lattice> rep;  // replicas on the peers of a multi-GPU context
    int s = 0;
    uint64_t nmax = 0;
    double *d_z = nullptr;  // z as Float64 (exact)
};

struct mle_dists_s {
    std::vector<mle_dists> rep;
    int m = 0;
    int mode = GEN_GENERAL;
    int *d_kind = nullptr;
    double *d_par = nullptr;  // [m][6]
};

static const int MLE_MAX_LEVELS = 256;  // levels, or packed multi-indices lx + 16*ly (lognormal2d with ndims = 2)
struct mle_model_s {
    std::vector<mle_model> rep;
    int kind = 0, ndims = 1, nqoi = 1, n0 = 0;
    int nweights = 0;
    double *d_weights = nullptr;
    // lognormal: per-level permuted / chunked basis
    double *d_basis[MLE_MAX_LEVELS] = {};
    int basis_rows[MLE_MAX_LEVELS] = {};
    int basis_cols[MLE_MAX_LEVELS] = {};
    int basis_mpad[MLE_MAX_LEVELS] = {};
    int basis_nchunks[MLE_MAX_LEVELS] = {};
    // lognormal2d, power-of-two Nx <= 64: the basis in DMMA A-fragment order [node tile][k step][lane]
    double *d_frag[MLE_MAX_LEVELS] = {};
    int frag_tiles[MLE_MAX_LEVELS] = {};
    int frag_mpad[MLE_MAX_LEVELS] = {};
};

static thread_local std::string g_init_err;

// MLE_DEBUG=1: report on stderr which entry point leaves a CUDA error pending (an unchecked runtime call), and log every
// device allocation and release
static bool mle_debug() { static const bool on = getenv("MLE_DEBUG") != nullptr; return on; }
struct ApiScope {
    const char *fn;
    explicit ApiScope(const char *f) : fn(f) {
        if (mle_debug()) { cudaError_t e = cudaPeekAtLastError(); if (e != cudaSuccess) fprintf(stderr, "[mle] %s entered with pending CUDA error: %s\n", fn, cudaGetErrorString(e)); }
    }
    ~ApiScope() {
        if (mle_debug()) { cudaError_t e = cudaPeekAtLastError(); if (e != cudaSuccess) fprintf(stderr, "[mle] %s left CUDA error: %s\n", fn, cudaGetErrorString(e)); }
    }
};
#define MLE_API_SCOPE ApiScope api_scope_(__func__)
static cudaError_t mle_free_(void *p, int line) {
    cudaError_t e = cudaFree(p);
    if (p && mle_debug()) fprintf(stderr, "[mle] free %p @%d -> %s\n", p, line, cudaGetErrorString(e));
    return e;
}
static cudaError_t mle_malloc_(void **p, size_t bytes, int line) {
    cudaError_t e = cudaMalloc(p, bytes);
    if (mle_debug()) fprintf(stderr, "[mle] malloc %zu @%d -> %p %s\n", bytes, line, *p, cudaGetErrorString(e));
    return e;
}
#define cudaFree(p) mle_free_((void *)(p), __LINE__)
#define MLE_MALLOC(pp, bytes) mle_malloc_(reinterpret_cast<void **>(pp), (bytes), __LINE__)

static int fail(mle_ctx ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg; else g_init_err = msg;
    return code;
}

#define CU_TRY(ctx, call)                                                                                  \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            (void)cudaGetLastError();                                                                      \
            return fail(ctx, MLE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
        }                                                                                                  \
    } while (0)

template <typename T>
static int ensure_dev(mle_ctx ctx, T **p, size_t *cap, size_t need) {
    if (*cap >= need) return MLE_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    size_t want = need + need / 2 + 64;
    cudaError_t e = MLE_MALLOC(p, want * sizeof(T));
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(ctx, MLE_ERR_NOMEM, "cudaMalloc failed (scratch)"); }
    *cap = want;
    return MLE_OK;
}

static int ensure_pinned(mle_ctx ctx, size_t need) {
    if (ctx->pinned_cap >= need) return MLE_OK;
    if (ctx->h_pinned) { cudaStreamSynchronize(ctx->stream); cudaFreeHost(ctx->h_pinned); }  // copies from it may be in flight
    ctx->h_pinned = nullptr; ctx->pinned_cap = 0;
    ctx->shifts_staged = 0;
    ctx->batch_staged = 0;
    size_t want = need + need / 2 + 64;
    cudaError_t e = cudaMallocHost(reinterpret_cast<void **>(&ctx->h_pinned), want * sizeof(double));
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(ctx, MLE_ERR_NOMEM, "cudaMallocHost failed"); }
    ctx->pinned_cap = want;
    return MLE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------------
extern "C" int mle_abi_version(void) {
    MLE_API_SCOPE; return MLE_ABI_VERSION; }

extern "C" int mle_init(int device, mle_ctx *out) {
    MLE_API_SCOPE;
    if (!out) return fail(nullptr, MLE_ERR_ARG, "mle_init: ctx is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        return fail(nullptr, MLE_ERR_CUDA, std::string("mle_init: no CUDA device (") +
                                               (e == cudaSuccess ? "device count 0" : cudaGetErrorString(e)) +
                                               "); libmle_cuda has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(nullptr, MLE_ERR_ARG, "mle_init: device ordinal out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, MLE_ERR_CUDA, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return fail(nullptr, MLE_ERR_CUDA, "mle_init: libmle_cuda is built for sm_100a only (device is sm_" +
                                               std::to_string(prop.major * 10 + prop.minor) + ")");
    mle_ctx ctx = new (std::nothrow) mle_ctx_s();
    if (!ctx) return fail(nullptr, MLE_ERR_NOMEM, "mle_init: out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc = prop.major * 10 + prop.minor;
    ctx->hbm = prop.totalGlobalMem;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        (void)cudaGetLastError();
        delete ctx;
        return fail(nullptr, MLE_ERR_CUDA, "mle_init: stream / event creation failed");
    }
    // opt in to large dynamic shared memory once
#define MLE_SET_SMEM(K, BYTES)                                                                      \
    cudaFuncSetAttribute(K<GEN_RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES);          \
    cudaFuncSetAttribute(K<GEN_STDNORMAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES);    \
    cudaFuncSetAttribute(K<GEN_GENERAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES);      \
    cudaFuncSetAttribute(K<GEN_RAW | GEN_MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES); \
    cudaFuncSetAttribute(K<GEN_STDNORMAL | GEN_MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES); \
    cudaFuncSetAttribute(K<GEN_GENERAL | GEN_MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES)
    MLE_SET_SMEM(points_k, 200 * 1024);
    MLE_SET_SMEM(expsum_k, 200 * 1024);
    MLE_SET_SMEM(problem18_k, 200 * 1024);
    MLE_SET_SMEM(lognormal1d_k16, 220 * 1024);
    MLE_SET_SMEM(lognormal1d_kb, 220 * 1024);
    MLE_SET_SMEM(lognormal1d_ks, 220 * 1024);
    MLE_SET_SMEM(lognormal1d_kl, 220 * 1024);
    MLE_SET_SMEM(lognormal1d_kc, 220 * 1024);
    MLE_SET_SMEM(lognormal2d_w, 220 * 1024);
    MLE_SET_SMEM(lognormal2d_c, 220 * 1024);
    MLE_SET_SMEM(lognormal2d_r4, 220 * 1024);
    MLE_SET_SMEM(lognormal2d_r8, 220 * 1024);
    MLE_SET_SMEM(lognormal2d_r16, 220 * 1024);
    MLE_SET_SMEM(lognormal2d_r32, 220 * 1024);
    MLE_SET_SMEM(lognormal2d_r64, 220 * 1024);
    MLE_SET_SMEM(lognormal2d_r128, 220 * 1024);
    MLE_SET_SMEM(lognormal2d_d64, 220 * 1024);
    MLE_SET_SMEM(lognormal2d_d128, 220 * 1024);
#undef MLE_SET_SMEM
    if (cudaGetLastError() != cudaSuccess) { delete ctx; return fail(nullptr, MLE_ERR_CUDA, "mle_init: cudaFuncSetAttribute failed"); }
    if (const char *e = getenv("MLE_2D_WINDOW_SMEM")) ctx->ln2d_old = atoi(e) != 0;
    if (const char *e = getenv("MLE_2D_DMMA")) ctx->ln2d_dmma = atoi(e);
    if (const char *e = getenv("MLE_LN_BLOCK")) ctx->ln_block = atoi(e) != 0;
    if (const char *e = getenv("MLE_LN_STAGED")) ctx->ln_staged = atoi(e) != 0;
    if (const char *e = getenv("MLE_LN_SMALL")) ctx->ln_small = atoi(e) != 0;
    if (const char *e = getenv("MLE_LN_CLUSTER")) ctx->ln_cluster = atoi(e);
    if (const char *e = getenv("MLE_NO_ZEROCOPY")) ctx->zero_copy = atoi(e) == 0;
    // the mapped completion flag (zero-copy result path); without it the library falls back to copy + synchronise
    if (ctx->zero_copy) {
        if (cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_flag), 64, cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer(reinterpret_cast<void **>(&ctx->d_flag), ctx->h_flag, 0) != cudaSuccess) {
            (void)cudaGetLastError();
            if (ctx->h_flag) cudaFreeHost(ctx->h_flag);
            ctx->h_flag = nullptr; ctx->d_flag = nullptr; ctx->zero_copy = 0;
        } else {
            *ctx->h_flag = 0;
        }
    }
    *out = ctx;
    return MLE_OK;
}

extern "C" int mle_init_multi(const int *devices, int ndev, mle_ctx *out) {
    MLE_API_SCOPE;
    if (!out) return fail(nullptr, MLE_ERR_ARG, "mle_init_multi: ctx is NULL");
    *out = nullptr;
    if (!devices || ndev < 1) return fail(nullptr, MLE_ERR_ARG, "mle_init_multi: need at least one device");
    // (MLE_ALLOW_DUP_DEVICES=1 lets a test box with one GPU exercise the multi-member logic: two members on the same device)
    for (int i = 0; i < ndev && !getenv("MLE_ALLOW_DUP_DEVICES"); ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return fail(nullptr, MLE_ERR_ARG, "mle_init_multi: device listed twice");
    mle_ctx ctx = nullptr;
    int rc = mle_init(devices[0], &ctx);
    for (int i = 1; rc == MLE_OK && i < ndev; ++i) {
        mle_ctx p = nullptr;
        rc = mle_init(devices[i], &p);
        if (rc == MLE_OK) ctx->peers.push_back(p);
    }
    if (rc != MLE_OK) { if (ctx) mle_destroy(ctx); return rc; }
    if (const char *e = getenv("MLE_SHARD_MIN")) { long long v = atoll(e); if (v >= 1) ctx->shard_min = (uint64_t)v; }
    *out = ctx;
    return MLE_OK;
}

extern "C" int mle_device_count(mle_ctx ctx, int *ndev) {
    MLE_API_SCOPE;
    if (!ctx || !ndev) return MLE_ERR_ARG;
    *ndev = 1 + (int)ctx->peers.size();
    return MLE_OK;
}

extern "C" int mle_destroy(mle_ctx ctx) {
    MLE_API_SCOPE;
    if (!ctx) return MLE_OK;
    for (mle_ctx p : ctx->peers) mle_destroy(p);
    ctx->peers.clear();
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->d_partials) cudaFree(ctx->d_partials);
    if (ctx->d_tickets) cudaFree(ctx->d_tickets);
    if (ctx->h_res) cudaFreeHost(ctx->h_res);
    if (ctx->h_flag) cudaFreeHost(ctx->h_flag);
    if (ctx->d_shifts) cudaFree(ctx->d_shifts);
    if (ctx->d_zero) cudaFree(ctx->d_zero);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->d_batch) cudaFree(ctx->d_batch);
    if (ctx->d_samples) cudaFree(ctx->d_samples);
    if (ctx->d_zacc) cudaFree(ctx->d_zacc);
    if (ctx->d_moments) cudaFree(ctx->d_moments);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return MLE_OK;
}

extern "C" const char *mle_last_error(mle_ctx ctx) {
    MLE_API_SCOPE; return ctx ? ctx->err.c_str() : g_init_err.c_str(); }

extern "C" int mle_device_info(mle_ctx ctx, int *sm_count, size_t *hbm_bytes, int *cc) {
    MLE_API_SCOPE;
    if (!ctx) return MLE_ERR_ARG;
    if (sm_count) *sm_count = ctx->sm_count;
    if (hbm_bytes) *hbm_bytes = ctx->hbm;
    if (cc) *cc = ctx->cc;
    return MLE_OK;
}

extern "C" int mle_stream(mle_ctx ctx, void **stream) {
    MLE_API_SCOPE;
    if (!ctx || !stream) return MLE_ERR_ARG;
    *stream = reinterpret_cast<void *>(ctx->stream);
    return MLE_OK;
}

extern "C" int mle_launch_count(mle_ctx ctx, uint64_t *count) {
    MLE_API_SCOPE;
    if (!ctx || !count) return MLE_ERR_ARG;
    *count = ctx->launches;
    for (mle_ctx p : ctx->peers) *count += p->launches;
    return MLE_OK;
}

extern "C" int mle_sync(mle_ctx ctx) {
    MLE_API_SCOPE;
    if (!ctx) return MLE_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (mle_ctx p : ctx->peers) {
        CU_TRY(ctx, cudaSetDevice(p->device));
        CU_TRY(ctx, cudaStreamSynchronize(p->stream));
    }
    return MLE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// lattice rule
// ---------------------------------------------------------------------------------------------------------------
extern "C" int mle_genvec(const char *name, const uint32_t **z, int *len) {
    MLE_API_SCOPE;
    if (!name || !z || !len) return MLE_ERR_ARG;
    if (!strcmp(name, "K_3600_32")) { *z = MLE_GENVEC_K3600; *len = MLE_GENVEC_K3600_LEN; return MLE_OK; }
    if (!strcmp(name, "CKN_250_20")) { *z = MLE_GENVEC_CKN250; *len = MLE_GENVEC_CKN250_LEN; return MLE_OK; }
    return MLE_ERR_ARG;
}

static int lattice_create_one(mle_ctx ctx, const uint32_t *z, int s, uint64_t nmax, mle_lattice *out) {
    if (!ctx || !out) return MLE_ERR_ARG;
    *out = nullptr;
    // check_args, src/lattice.jl:181-185: s > 0, s <= length(z) (caller's contract), n <= 2^32
    if (s <= 0) return fail(ctx, MLE_ERR_ARG, "LatticeRule32: number of dimensions s must be larger than 0");
    if (nmax > 0xffffffffull) return fail(ctx, MLE_ERR_ARG, "LatticeRule32: maximum number of points n must be smaller than or equal to 2^32");
    if (!z) {  // LatticeRule32(s): src/lattice.jl:88, check_arg :187
        if (s > MLE_GENVEC_K3600_LEN)
            return fail(ctx, MLE_ERR_ARG, "LatticeRule32: number of dimensions s must be smaller than or equal to 3600, please supply your own generating vector z to proceed");
        z = MLE_GENVEC_K3600;
    }
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    mle_lattice L = new (std::nothrow) mle_lattice_s();
    if (!L) return fail(ctx, MLE_ERR_NOMEM, "out of host memory");
    L->s = s;
    L->nmax = nmax;
    std::vector<double> zd((size_t)s);
    for (int j = 0; j < s; ++j) zd[(size_t)j] = (double)z[j];
    if (MLE_MALLOC(&L->d_z, sizeof(double) * (size_t)s) != cudaSuccess) { delete L; (void)cudaGetLastError(); return fail(ctx, MLE_ERR_NOMEM, "cudaMalloc(z)"); }
    cudaError_t e = cudaMemcpyAsync(L->d_z, zd.data(), sizeof(double) * (size_t)s, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cudaFree(L->d_z); delete L; return fail(ctx, MLE_ERR_CUDA, cudaGetErrorString(e)); }
    *out = L;
    return MLE_OK;
}

static int lattice_destroy_one(mle_ctx ctx, mle_lattice lat) {
    if (!lat) return MLE_OK;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    cudaFree(lat->d_z);
    delete lat;
    return MLE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// distributions
// ---------------------------------------------------------------------------------------------------------------
static double host_normcdf(double x) {  // Phi(x) = 1/2*(1 + erf(x/sqrt(2)))              src/distribution.jl:153
    volatile double q = x / 1.4142135623730951;
    volatile double e = 1.0 + std::erf(q);
    return 0.5 * e;
}

static int dists_create_one(mle_ctx ctx, const mle_dist *d, int m, mle_dists *out) {
    if (!ctx || !out) return MLE_ERR_ARG;
    *out = nullptr;
    if (!d || m <= 0) return fail(ctx, MLE_ERR_ARG, "distributions: need at least one distribution");
    std::vector<int> kind((size_t)m);
    std::vector<double> par((size_t)m * 6, 0.0);
    bool all_std = true;
    for (int j = 0; j < m; ++j) {
        const mle_dist &D = d[j];
        const double *p = D.p;
        double *o = &par[(size_t)j * 6];
        auto finite = [](double v) { return std::isfinite(v); };
        switch (D.kind) {  // constructor checks of src/distribution.jl:36-40, :104-108, :134-141, :173-178
        case MLE_UNIFORM:
            if (!finite(p[0]) || !finite(p[1])) return fail(ctx, MLE_ERR_ARG, "Uniform: a and b must be finite");
            if (!(p[0] < p[1])) return fail(ctx, MLE_ERR_ARG, "Uniform: a must be smaller than b");
            break;
        case MLE_NORMAL:
            if (!finite(p[0]) || !finite(p[1])) return fail(ctx, MLE_ERR_ARG, "Normal: mu and sigma must be finite");
            if (!(p[1] > 0)) return fail(ctx, MLE_ERR_ARG, "Normal: sigma must be larger than 0");
            break;
        case MLE_TRUNCNORMAL: {
            if (!finite(p[0]) || !finite(p[1]) || !finite(p[2]) || !finite(p[3])) return fail(ctx, MLE_ERR_ARG, "TruncatedNormal: parameters must be finite");
            if (!(p[1] > 0)) return fail(ctx, MLE_ERR_ARG, "TruncatedNormal: sigma must be larger than 0");
            if (!(p[2] < p[3])) return fail(ctx, MLE_ERR_ARG, "TruncatedNormal: a must be smaller than b");
            // per-distribution constants of src/distribution.jl:146-148 (depend on the parameters only)
            volatile double al = (p[2] - p[0]) / p[1], be = (p[3] - p[0]) / p[1];
            double Fa = host_normcdf(al), Fb = host_normcdf(be);
            volatile double dF = Fb - Fa;
            o[4] = Fa; o[5] = dF;
            break;
        }
        case MLE_WEIBULL:
            if (!finite(p[0]) || !finite(p[1])) return fail(ctx, MLE_ERR_ARG, "Weibull: k and lambda must be finite");
            if (!(p[0] > 0) || !(p[1] > 0)) return fail(ctx, MLE_ERR_ARG, "Weibull: k and lambda must be larger than 0");
            break;
        default:
            return fail(ctx, MLE_ERR_ARG, "distributions: unknown kind");
        }
        kind[(size_t)j] = D.kind;
        for (int t = 0; t < 4; ++t) o[t] = p[t];
        if (!(D.kind == MLE_NORMAL && p[0] == 0.0 && p[1] == 1.0)) all_std = false;
    }
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    mle_dists S = new (std::nothrow) mle_dists_s();
    if (!S) return fail(ctx, MLE_ERR_NOMEM, "out of host memory");
    S->m = m;
    S->mode = all_std ? GEN_STDNORMAL : GEN_GENERAL;
    cudaError_t e = MLE_MALLOC(&S->d_kind, sizeof(int) * (size_t)m);
    if (e == cudaSuccess) e = MLE_MALLOC(&S->d_par, sizeof(double) * (size_t)m * 6);
    if (e == cudaSuccess) e = cudaMemcpyAsync(S->d_kind, kind.data(), sizeof(int) * (size_t)m, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(S->d_par, par.data(), sizeof(double) * (size_t)m * 6, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        cudaFree(S->d_kind); cudaFree(S->d_par); delete S;
        return fail(ctx, MLE_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = S;
    return MLE_OK;
}

static int dists_destroy_one(mle_ctx ctx, mle_dists dists) {
    if (!dists) return MLE_OK;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    cudaFree(dists->d_kind);
    cudaFree(dists->d_par);
    delete dists;
    return MLE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// model registry
// ---------------------------------------------------------------------------------------------------------------
static int model_create_one(mle_ctx ctx, const char *name, int ndims, const double *params, int nparams, mle_model *out) {
    if (!ctx || !out || !name) return MLE_ERR_ARG;
    *out = nullptr;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    mle_model M = new (std::nothrow) mle_model_s();
    if (!M) return fail(ctx, MLE_ERR_NOMEM, "out of host memory");
    M->ndims = ndims;
    if (!strcmp(name, "expsum")) {
        if (ndims != 1 || nparams < 2 || !params || params[0] < 1) { delete M; return fail(ctx, MLE_ERR_ARG, "expsum: ndims = 1, params = {m0 >= 1, a_1..a_m}"); }
        M->kind = MODEL_EXPSUM;
        M->n0 = (int)params[0];
        M->nweights = nparams - 1;
        cudaError_t e = MLE_MALLOC(&M->d_weights, sizeof(double) * (size_t)M->nweights);
        if (e == cudaSuccess) e = cudaMemcpyAsync(M->d_weights, params + 1, sizeof(double) * (size_t)M->nweights, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { (void)cudaGetLastError(); cudaFree(M->d_weights); delete M; return fail(ctx, MLE_ERR_CUDA, cudaGetErrorString(e)); }
    } else if (!strcmp(name, "problem18")) {
        if (ndims != 1 && ndims != 2) { delete M; return fail(ctx, MLE_ERR_ARG, "problem18: ndims must be 1 or 2"); }
        M->kind = MODEL_PROBLEM18;
        M->nqoi = 13;
    } else if (!strcmp(name, "lognormal1d")) {
        int n0 = (nparams > 0 && params) ? (int)params[0] : 16;
        if (ndims != 1 || n0 < 4 || (n0 & 1)) { delete M; return fail(ctx, MLE_ERR_ARG, "lognormal1d: ndims = 1, params = {n0}, n0 even and >= 4"); }
        M->kind = MODEL_LOGNORMAL1D;
        M->n0 = n0;
    } else if (!strcmp(name, "lognormal2d")) {
        int n0 = (nparams > 0 && params) ? (int)params[0] : 4;
        if ((ndims != 1 && ndims != 2) || n0 < 2 || (n0 & 1)) { delete M; return fail(ctx, MLE_ERR_ARG, "lognormal2d: ndims = 1 (multilevel) or 2 (multi-index), params = {n0}, n0 even and >= 2"); }
        M->kind = MODEL_LOGNORMAL2D;
        M->n0 = n0;
    } else {
        delete M;
        return fail(ctx, MLE_ERR_ARG, std::string("unknown model '") + name + "'");
    }
    *out = M;
    return MLE_OK;
}

static int model_set_basis_one(mle_ctx ctx, mle_model M, int level, int rows, int cols, const double *phi) {
    if (!ctx || !M || !phi) return MLE_ERR_ARG;
    if (M->kind != MODEL_LOGNORMAL1D && M->kind != MODEL_LOGNORMAL2D) return fail(ctx, MLE_ERR_ARG, "set_basis: model takes no KL basis");
    if (level < 0 || level >= MLE_MAX_LEVELS) return fail(ctx, MLE_ERR_ARG, "set_basis: level out of range");
    if (M->kind == MODEL_LOGNORMAL2D) {
        // multilevel: level l (Nx = Ny = n0 2^l); multi-index: level = lx + 16*ly
        const int lx = M->ndims == 2 ? (level & 15) : level, ly = M->ndims == 2 ? (level >> 4) : level;
        if (lx > 12 || ly > 12) return fail(ctx, MLE_ERR_ARG, "set_basis: level out of range");
        const long long Nx = (long long)M->n0 << lx, Ny = (long long)M->n0 << ly;
        if (Nx > 128 || Ny > 4096) return fail(ctx, MLE_ERR_ARG, "set_basis: lognormal2d supports grids up to 128 cells in x");
        const long long nodes = (Nx + 1) * (Ny + 1);
        if (rows != nodes) return fail(ctx, MLE_ERR_ARG, "set_basis: rows must be (n0*2^lx + 1)*(n0*2^ly + 1) (nodes of the index's grid)");
        if (cols < 1 || cols > 1024) return fail(ctx, MLE_ERR_ARG, "set_basis: 1 <= KL terms <= 1024");
        CU_TRY(ctx, cudaSetDevice(ctx->device));
        const long long npad = (nodes + 31) & ~31ll;
        std::vector<double> tr((size_t)cols * npad, 0.0);  // transposed [k][node]: thread-per-node reads are coalesced
        for (long long nd = 0; nd < nodes; ++nd)
            for (int k = 0; k < cols; ++k) tr[(size_t)k * npad + nd] = phi[(size_t)nd * cols + k];
        double *d2 = nullptr;
        if (MLE_MALLOC(&d2, tr.size() * sizeof(double)) != cudaSuccess) { (void)cudaGetLastError(); return fail(ctx, MLE_ERR_NOMEM, "cudaMalloc(basis)"); }
        cudaError_t e2 = cudaMemcpyAsync(d2, tr.data(), tr.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
        if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(ctx->stream);
        if (e2 != cudaSuccess) { cudaFree(d2); return fail(ctx, MLE_ERR_CUDA, cudaGetErrorString(e2)); }
        if (M->d_basis[level]) cudaFree(M->d_basis[level]);
        M->d_basis[level] = d2;
        M->basis_rows[level] = rows;
        M->basis_cols[level] = cols;
        M->basis_mpad[level] = (int)npad;
        M->basis_nchunks[level] = 0;
        if (M->d_frag[level]) { cudaFree(M->d_frag[level]); M->d_frag[level] = nullptr; }
        if (Nx >= 4 && Nx <= 128 && (Nx & (Nx - 1)) == 0) {  // register-window kernel: fragment-ordered copy
            const int mpad = (cols + 3) & ~3, nks = mpad / 4, ntiles = (int)((nodes + 7) / 8);
            std::vector<double> packed((size_t)ntiles * nks * 32, 0.0);
            for (long long nd = 0; nd < nodes; ++nd)
                for (int k = 0; k < cols; ++k)
                    packed[((size_t)(nd / 8) * nks + k / 4) * 32 + (nd % 8) * 4 + (k % 4)] = phi[(size_t)nd * cols + k];
            double *d3 = nullptr;
            if (MLE_MALLOC(&d3, packed.size() * sizeof(double)) != cudaSuccess) { (void)cudaGetLastError(); return fail(ctx, MLE_ERR_NOMEM, "cudaMalloc(basis)"); }
            cudaError_t e3 = cudaMemcpyAsync(d3, packed.data(), packed.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
            if (e3 == cudaSuccess) e3 = cudaStreamSynchronize(ctx->stream);
            if (e3 != cudaSuccess) { cudaFree(d3); return fail(ctx, MLE_ERR_CUDA, cudaGetErrorString(e3)); }
            M->d_frag[level] = d3;
            M->frag_tiles[level] = ntiles;
            M->frag_mpad[level] = mpad;
        }
        return MLE_OK;
    }
    if (level > 20) return fail(ctx, MLE_ERR_ARG, "set_basis: level out of range");
    const long long N = (long long)M->n0 << level;
    if (rows != N + 1) return fail(ctx, MLE_ERR_ARG, "set_basis: rows must be n0*2^level + 1 (nodes of the level's grid)");
    if (cols < 1 || cols > 352) return fail(ctx, MLE_ERR_ARG, "set_basis: 1 <= KL terms <= 352 (the xi tile must fit in shared memory)");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    // Row r' = 2*i' + side, i' < N/2: side 0 -> node i', side 1 -> mirrored node N - i', so both sweeps read rows in ascending
    // order.  Stored in DMMA A-fragment order: [row tile rt][k step ks][lane] holds Phi'[8 rt + lane/4][4 ks + lane%4], zero
    // padded in rows (to whole 8-row tiles) and in k (to a multiple of 4); the row of the MID node N/2 follows as a plain
    // vector of mpad doubles (the kernel evaluates it as one fused-multiply-add chain per sample).
    const int Mh = (int)(N / 2);
    const int nrows = 2 * Mh;
    const int ntiles = (nrows + 7) / 8;
    const int nchunks = (ntiles + 3) / 4;
    const int mpad = (cols + 3) & ~3;
    const int nks = mpad / 4;
    std::vector<double> packed((size_t)ntiles * nks * 32 + (size_t)mpad, 0.0);
    for (int rp = 0; rp < nrows; ++rp) {
        const int ip = rp >> 1, side = rp & 1;
        const long long node = side ? N - ip : ip;
        const int rt = rp / 8, lr = rp % 8;
        for (int k = 0; k < cols; ++k)
            packed[((size_t)rt * nks + k / 4) * 32 + lr * 4 + (k % 4)] = phi[(size_t)node * cols + k];
    }
    for (int k = 0; k < cols; ++k) packed[(size_t)ntiles * nks * 32 + k] = phi[(size_t)Mh * cols + k];
    double *d = nullptr;
    if (MLE_MALLOC(&d, packed.size() * sizeof(double)) != cudaSuccess) { (void)cudaGetLastError(); return fail(ctx, MLE_ERR_NOMEM, "cudaMalloc(basis)"); }
    cudaError_t e = cudaMemcpyAsync(d, packed.data(), packed.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cudaFree(d); return fail(ctx, MLE_ERR_CUDA, cudaGetErrorString(e)); }
    if (M->d_basis[level]) cudaFree(M->d_basis[level]);
    M->d_basis[level] = d;
    M->basis_rows[level] = rows;
    M->basis_cols[level] = cols;
    M->basis_mpad[level] = mpad;
    M->basis_nchunks[level] = nchunks;
    M->frag_tiles[level] = ntiles;
    return MLE_OK;
}

extern "C" int mle_model_nqoi(mle_model M, int *nqoi) {
    MLE_API_SCOPE;
    if (!M || !nqoi) return MLE_ERR_ARG;
    *nqoi = M->nqoi;
    return MLE_OK;
}

static int model_destroy_one(mle_ctx ctx, mle_model M) {
    if (!M) return MLE_OK;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    cudaFree(M->d_weights);
    for (int l = 0; l < MLE_MAX_LEVELS; ++l) { cudaFree(M->d_basis[l]); cudaFree(M->d_frag[l]); }
    delete M;
    return MLE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// public handle constructors: one object on the context's device plus one replica per peer of a multi-GPU context
// ---------------------------------------------------------------------------------------------------------------
static inline int group_size(mle_ctx ctx) { return 1 + (int)ctx->peers.size(); }
static inline mle_ctx member_ctx(mle_ctx ctx, int i) { return i == 0 ? ctx : ctx->peers[(size_t)i - 1]; }
template <typename H> static inline H member_handle(H h, int i) { return (!h || i == 0) ? h : h->rep[(size_t)i - 1]; }
static int peer_fail(mle_ctx ctx, mle_ctx peer, int rc) { ctx->err = "device " + std::to_string(peer->device) + ": " + peer->err; return rc; }

extern "C" int mle_lattice_create(mle_ctx ctx, const uint32_t *z, int s, uint64_t nmax, mle_lattice *out) {
    MLE_API_SCOPE;
    int rc = lattice_create_one(ctx, z, s, nmax, out);
    for (size_t i = 0; rc == MLE_OK && i < ctx->peers.size(); ++i) {
        mle_lattice r = nullptr;
        rc = lattice_create_one(ctx->peers[i], z, s, nmax, &r);
        if (rc == MLE_OK) (*out)->rep.push_back(r); else peer_fail(ctx, ctx->peers[i], rc);
    }
    if (rc != MLE_OK && out && *out) { mle_lattice_destroy(ctx, *out); *out = nullptr; }
    return rc;
}
extern "C" int mle_lattice_destroy(mle_ctx ctx, mle_lattice lat) {
    MLE_API_SCOPE;
    if (!lat) return MLE_OK;
    for (size_t i = 0; i < lat->rep.size(); ++i) lattice_destroy_one(ctx && i < ctx->peers.size() ? ctx->peers[i] : nullptr, lat->rep[i]);
    return lattice_destroy_one(ctx, lat);
}
extern "C" int mle_dists_create(mle_ctx ctx, const mle_dist *d, int m, mle_dists *out) {
    MLE_API_SCOPE;
    int rc = dists_create_one(ctx, d, m, out);
    for (size_t i = 0; rc == MLE_OK && i < ctx->peers.size(); ++i) {
        mle_dists r = nullptr;
        rc = dists_create_one(ctx->peers[i], d, m, &r);
        if (rc == MLE_OK) (*out)->rep.push_back(r); else peer_fail(ctx, ctx->peers[i], rc);
    }
    if (rc != MLE_OK && out && *out) { mle_dists_destroy(ctx, *out); *out = nullptr; }
    return rc;
}
extern "C" int mle_dists_destroy(mle_ctx ctx, mle_dists dists) {
    MLE_API_SCOPE;
    if (!dists) return MLE_OK;
    for (size_t i = 0; i < dists->rep.size(); ++i) dists_destroy_one(ctx && i < ctx->peers.size() ? ctx->peers[i] : nullptr, dists->rep[i]);
    return dists_destroy_one(ctx, dists);
}
extern "C" int mle_model_create(mle_ctx ctx, const char *name, int ndims, const double *params, int nparams, mle_model *out) {
    MLE_API_SCOPE;
    int rc = model_create_one(ctx, name, ndims, params, nparams, out);
    for (size_t i = 0; rc == MLE_OK && i < ctx->peers.size(); ++i) {
        mle_model r = nullptr;
        rc = model_create_one(ctx->peers[i], name, ndims, params, nparams, &r);
        if (rc == MLE_OK) (*out)->rep.push_back(r); else peer_fail(ctx, ctx->peers[i], rc);
    }
    if (rc != MLE_OK && out && *out) { mle_model_destroy(ctx, *out); *out = nullptr; }
    return rc;
}
extern "C" int mle_model_set_basis(mle_ctx ctx, mle_model M, int level, int rows, int cols, const double *phi) {
    MLE_API_SCOPE;
    int rc = model_set_basis_one(ctx, M, level, rows, cols, phi);
    for (size_t i = 0; rc == MLE_OK && M && i < M->rep.size() && i < ctx->peers.size(); ++i) {
        rc = model_set_basis_one(ctx->peers[i], M->rep[i], level, rows, cols, phi);
        if (rc != MLE_OK) peer_fail(ctx, ctx->peers[i], rc);
    }
    return rc;
}
extern "C" int mle_model_destroy(mle_ctx ctx, mle_model M) {
    MLE_API_SCOPE;
    if (!M) return MLE_OK;
    for (size_t i = 0; i < M->rep.size(); ++i) model_destroy_one(ctx && i < ctx->peers.size() ? ctx->peers[i] : nullptr, M->rep[i]);
    return model_destroy_one(ctx, M);
}

// ---------------------------------------------------------------------------------------------------------------
// common argument handling for the point / batch entry points
// ---------------------------------------------------------------------------------------------------------------
static int make_gen_args(mle_ctx ctx, mle_lattice lat, mle_dists dists, const double *shifts_dev, int R, uint64_t k0,
                         uint64_t n, int m, uint64_t seed, GenArgs *g) {
    if (R < 1) return fail(ctx, MLE_ERR_ARG, "number of shifts must be at least 1");
    if (m < 1) return fail(ctx, MLE_ERR_ARG, "nb_of_uncertainties must be at least 1");
    if (n > 0xffffffffull) return fail(ctx, MLE_ERR_ARG, "at most 2^32 points per call");
    if (lat) {
        if (m > lat->s) return fail(ctx, MLE_ERR_ARG, "nb_of_uncertainties exceeds the lattice dimension s (the reference pads with rand(), src/sample.jl:85, which is not reproducible)");
        if (n && (k0 + n - 1 > lat->nmax || k0 + n - 1 > 0xffffffffull))
            return fail(ctx, MLE_ERR_RANGE, "point number beyond the lattice size n (the reference returns rand(s) there, src/lattice.jl:166,175)");
        if (!shifts_dev && R != 1) return fail(ctx, MLE_ERR_ARG, "an unshifted lattice rule has a single 'shift'");
    } else {
        if (R != 1) return fail(ctx, MLE_ERR_ARG, "MC sampling has a single 'shift' slot (src/internals.jl:41)");
        if (shifts_dev) return fail(ctx, MLE_ERR_ARG, "MC sampling takes no shifts");
    }
    if (dists && m > dists->m) return fail(ctx, MLE_ERR_ARG, "nb_of_uncertainties exceeds the number of distributions");
    if (lat && !shifts_dev) {  // unshifted rule: a zero shift (x + 0.0 is exact) keeps the generator branch-free
        if (ctx->zero_cap < (size_t)lat->s) {
            if (ctx->d_zero) cudaFree(ctx->d_zero);
            ctx->d_zero = nullptr; ctx->zero_cap = 0;
            if (MLE_MALLOC(&ctx->d_zero, sizeof(double) * (size_t)lat->s) != cudaSuccess) { (void)cudaGetLastError(); return fail(ctx, MLE_ERR_NOMEM, "cudaMalloc(zero shift)"); }
            if (cudaMemsetAsync(ctx->d_zero, 0, sizeof(double) * (size_t)lat->s, ctx->stream) != cudaSuccess) return fail(ctx, MLE_ERR_CUDA, "cudaMemsetAsync");
            ctx->zero_cap = (size_t)lat->s;
        }
        shifts_dev = ctx->d_zero;
    }
    g->zd = lat ? lat->d_z : nullptr;
    g->shifts = shifts_dev;
    g->kind = dists ? dists->d_kind : nullptr;
    g->par = dists ? dists->d_par : nullptr;
    g->seed = seed;
    g->k0 = k0;
    g->n = n;
    g->s = lat ? lat->s : 0;
    g->m = m;
    g->mkey = m;
    g->R = R;
    g->mode = (dists ? dists->mode : GEN_RAW) | (lat ? 0 : GEN_MC);
    return MLE_OK;
}

static int upload_shifts(mle_ctx ctx, mle_lattice lat, const double *shifts, int R, const double **dev) {
    *dev = nullptr;
    if (!shifts) return MLE_OK;
    if (!lat) return fail(ctx, MLE_ERR_ARG, "MC sampling takes no shifts");
    const size_t cnt = (size_t)R * (size_t)lat->s;
    const double *before = ctx->d_shifts;
    int rc = ensure_dev(ctx, &ctx->d_shifts, &ctx->shifts_cap, cnt);
    if (rc) return rc;
    if (ctx->d_shifts != before) ctx->shifts_staged = 0;
    rc = ensure_pinned(ctx, cnt + 4096);
    if (rc) return rc;
    // the adaptive loop calls again and again with the same shifts of an index (src/internals.jl:138: one shift vector per
    // (index, shift) for the whole run): skip the copy when the staged image is still what the device holds
    if (ctx->shifts_staged != cnt || memcmp(ctx->h_pinned, shifts, cnt * sizeof(double)) != 0) {
        memcpy(ctx->h_pinned, shifts, cnt * sizeof(double));
        ctx->shifts_staged = 0;
        ctx->batch_staged = 0;
        CU_TRY(ctx, cudaMemcpyAsync(ctx->d_shifts, ctx->h_pinned, cnt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        ctx->shifts_staged = cnt;
    }
    *dev = ctx->d_shifts;
    return MLE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// stage 1-2 parity hook
// ---------------------------------------------------------------------------------------------------------------
extern "C" int mle_points_dev(mle_ctx ctx, mle_lattice lat, mle_dists dists, const double *shifts_dev, int R, uint64_t k0,
                              uint64_t n, int m, uint64_t mc_seed, double *out_dev) {
    MLE_API_SCOPE;
    if (!ctx || !out_dev) return MLE_ERR_ARG;
    GenArgs g;
    int rc = make_gen_args(ctx, lat, dists, shifts_dev, R, k0, n, m, mc_seed, &g);
    if (rc) return rc;
    if (n == 0) return MLE_OK;
    if (m > POINTS_TILE_CAP) return fail(ctx, MLE_ERR_ARG, "mle_points: at most 4096 dimensions");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const int P = POINTS_TILE_CAP / m > 0 ? POINTS_TILE_CAP / m : 1;
    const long long tps = (long long)((n + (uint64_t)P - 1) / (uint64_t)P);
    const long long ntiles = tps * R;
#ifdef MLE_POINTS_DOUBLE_BUF
    const size_t smem = 2 * POINTS_TILE_CAP * sizeof(double) + POINTS_TILE_CAP * sizeof(unsigned short) + (POINTS_THREADS / 32) * 4;
#else
    const size_t smem = 1 * POINTS_TILE_CAP * sizeof(double) + POINTS_TILE_CAP * sizeof(unsigned short) + (POINTS_THREADS / 32) * 4;
#endif
    long long grid = (long long)ctx->sm_count * MLE_OCC_BY_MODE(g.mode, points_k, POINTS_THREADS, smem);
    if (grid > ntiles) grid = ntiles;
    grid = churn_grid_x(grid, ntiles);
    MLE_LAUNCH_BY_MODE(g.mode, points_k, (unsigned)grid, POINTS_THREADS, smem, ctx->stream, g, out_dev, P, tps, ntiles);
    ctx->launches += 1;
    CU_TRY(ctx, cudaGetLastError());
    return MLE_OK;
}

extern "C" int mle_points(mle_ctx ctx, mle_lattice lat, mle_dists dists, const double *shifts, int R, uint64_t k0, uint64_t n,
                          int m, uint64_t mc_seed, double *out) {
    MLE_API_SCOPE;
    if (!ctx || !out) return MLE_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if (R < 1 || m < 1) return fail(ctx, MLE_ERR_ARG, "R and m must be positive");
    const double *sh_dev = nullptr;
    int rc = upload_shifts(ctx, lat, shifts, R, &sh_dev);
    if (rc) return rc;
    if (n == 0) return MLE_OK;
    const size_t cnt = (size_t)R * n * (size_t)m;
    double *d_out = nullptr;
    if (MLE_MALLOC(&d_out, cnt * sizeof(double)) != cudaSuccess) { (void)cudaGetLastError(); return fail(ctx, MLE_ERR_NOMEM, "mle_points: cudaMalloc(out)"); }
    rc = mle_points_dev(ctx, lat, dists, sh_dev, R, k0, n, m, mc_seed, d_out);
    if (rc == MLE_OK) {
        cudaError_t e = cudaMemcpyAsync(out, d_out, cnt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { (void)cudaGetLastError(); rc = fail(ctx, MLE_ERR_CUDA, std::string("mle_points: ") + cudaGetErrorString(e)); }
    }
    cudaFree(d_out);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// the hot path
// ---------------------------------------------------------------------------------------------------------------
static long long grid_x_for(mle_ctx ctx, long long tiles_per_shift, int R, int ctas_per_sm) {
    // Persistent CTAs: never launch more CTAs than are resident at once (a second, nearly empty wave would double the
    // run time), so the grid is the largest gx with gx * R <= SMs * ctas_per_sm.
    long long total = (long long)ctx->sm_count * ctas_per_sm;
    long long gx = total / R;
    if (gx < 1) gx = 1;
    if (gx > tiles_per_shift) gx = tiles_per_shift;
    return gx;
}

// one launch = one parallel_sample! call; `flag_dev` (optional, mapped host memory) receives `seq` once the moments of every
// shift have been written (cta_finalize)
static int launch_batch(mle_ctx ctx, mle_model M, mle_lattice lat, mle_dists dists, const int32_t *index, int d,
                        const double *shifts_dev, int R, uint64_t k0, uint64_t n, int m, uint64_t mc_seed,
                        double *moments_dev, double *samples_dev, double *host_moments_dev = nullptr, unsigned *flag_dev = nullptr,
                        unsigned seq = 0) {
    if (!ctx || !M || !index || !moments_dev) return MLE_ERR_ARG;
    if (d != M->ndims) return fail(ctx, MLE_ERR_ARG, "index dimension does not match the model");
    for (int t = 0; t < d; ++t)
        if (index[t] < 0) return fail(ctx, MLE_ERR_ARG, "index must be non-negative");
    if (n == 0) return fail(ctx, MLE_ERR_ARG, "sample!: number of samples must be positive");
    GenArgs g;
    {   // an error left behind by an earlier unchecked failure must not be blamed on this launch
        cudaError_t stale = cudaGetLastError();
        if (stale != cudaSuccess) return fail(ctx, MLE_ERR_CUDA, std::string("CUDA error pending before mle_sample_batch: ") + cudaGetErrorString(stale));
    }
    int rc = make_gen_args(ctx, lat, dists, shifts_dev, R, k0, n, m, mc_seed, &g);
    if (rc) return rc;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const int V = 2 * M->nqoi;
    long long gx = 0;
    if (R > 65535) return fail(ctx, MLE_ERR_ARG, "at most 65535 shifts per call");
    if (ctx->tickets_cap < (size_t)R + 1) {  // ticket counters of cta_finalize: zeroed once, every launch leaves them at zero
        const size_t before = ctx->tickets_cap;
        rc = ensure_dev(ctx, &ctx->d_tickets, &ctx->tickets_cap, (size_t)R + 1);
        if (rc) return rc;
        (void)before;
        CU_TRY(ctx, cudaMemsetAsync(ctx->d_tickets, 0, ctx->tickets_cap * sizeof(unsigned), ctx->stream));
    }
    FinArgs fa{nullptr, ctx->d_tickets, moments_dev, host_moments_dev, flag_dev, seq};
    if (M->kind == MODEL_EXPSUM || M->kind == MODEL_PROBLEM18) {
        const long long tps = (long long)((n + SIMPLE_THREADS - 1) / SIMPLE_THREADS);
        SimpleModelArgs ma{};
        ma.weights = M->d_weights;
        ma.nweights = M->nweights;
        ma.ndims = M->ndims;
        ma.idx0 = index[0];
        ma.idx1 = d > 1 ? index[1] : 0;
        const int W = SIMPLE_THREADS / 32;
        const size_t smem = (size_t)SIMPLE_THREADS * SIMPLE_MC * 8 + (size_t)3 * V * W * 8 + (size_t)V * sizeof(Partial) +
                            (size_t)SIMPLE_THREADS * SIMPLE_MC * 2 + ((W + 3) & ~3) * 4 + 2 * SIMPLE_MC * 16;
        const int occ = M->kind == MODEL_EXPSUM ? MLE_OCC_BY_MODE(g.mode, expsum_k, SIMPLE_THREADS, smem)
                                                : MLE_OCC_BY_MODE(g.mode, problem18_k, SIMPLE_THREADS, smem);
        gx = churn_grid_x(grid_x_for(ctx, tps, R, occ), tps);
        rc = ensure_dev(ctx, &ctx->d_partials, &ctx->partials_cap, (size_t)R * V * (size_t)gx);
        if (rc) return rc;
        fa.partials = ctx->d_partials;
        dim3 grid((unsigned)gx, (unsigned)R);
        if (M->kind == MODEL_EXPSUM) {
            if (index[0] > 30) return fail(ctx, MLE_ERR_ARG, "expsum: level too large");
            long long ml = (long long)M->n0 << index[0];
            if (ml > m) ml = m;
            if (ml > M->nweights) ml = M->nweights;
            long long mlc = 0;
            if (index[0] > 0) {
                mlc = (long long)M->n0 << (index[0] - 1);
                if (mlc > m) mlc = m;
                if (mlc > M->nweights) mlc = M->nweights;
            }
            ma.ml = (int)ml;
            ma.mlc = (int)mlc;
            MLE_LAUNCH_BY_MODE(g.mode, expsum_k, grid, SIMPLE_THREADS, smem, ctx->stream, g, ma, fa, samples_dev, tps);
        } else {
            if (m < M->ndims) return fail(ctx, MLE_ERR_ARG, "problem18: needs one uncertainty per index dimension");
            MLE_LAUNCH_BY_MODE(g.mode, problem18_k, grid, SIMPLE_THREADS, smem, ctx->stream, g, ma, fa, samples_dev, tps);
        }
    } else if (M->kind == MODEL_LOGNORMAL1D) {
        const int level = index[0];
        if (level >= MLE_MAX_LEVELS || !M->d_basis[level]) return fail(ctx, MLE_ERR_STATE, "lognormal1d: no KL basis set for this level (mle_model_set_basis)");
        // the sample function reads the leading KL terms and ignores surplus coordinates, like a reference closure would
        if (m < M->basis_cols[level]) return fail(ctx, MLE_ERR_ARG, "lognormal1d: nb_of_uncertainties is smaller than the number of KL terms of the level's basis");
        g.m = M->basis_cols[level];
        Lognormal1dArgs la{};
        la.basis = M->d_basis[level];
        la.N = M->n0 << level;
        la.mpad = M->basis_mpad[level];
        la.nchunks = M->basis_nchunks[level];
        la.ntiles = M->frag_tiles[level];
        la.mid = la.basis + (size_t)la.ntiles * (la.mpad / 4) * 32;
        la.level = level;
        la.h2 = 1.0 / ((double)la.N * (double)la.N);
        la.h2c = 1.0 / ((double)(la.N / 2) * (double)(la.N / 2));
        // warp-autonomous kernel (mle_kernel_ln1d.cuh): CTA tiles of LN_BT samples, a warp owns WS of them.  WS = 8 (four
        // warps per CTA, 20 resident warps per SM) wins while the generator dominates (small grids); WS = 16 (4 x 2 DMMA
        // warp tiles, half the basis traffic per sample) wins once the KL contraction dominates.  MLE_LN_WS overrides.
        // Two schedules of the same arithmetic (bit-identical results, same basis layout): the phase-synchronous CTA kernel
        // (mle_kernel_ln1d_block.cuh) and the warp-autonomous one (mle_kernel_ln1d.cuh).  Measured on the C2 step
        // (profiles/r02_ln1d_variants.txt): grids of a single half chunk (N <= 16: the generator is 70 % of the work and the
        // mid-node chain would make one warp of a barrier-synchronised CTA the straggler) run 6 % faster warp-autonomous,
        // every larger grid 7-15 % faster phase-synchronous (its four warps walk the basis together: L1 hits).
        // MLE_LN_BLOCK=0|1 forces one of them.
        const bool block = ctx->ln_block >= 0 ? ctx->ln_block != 0 : la.ntiles > 2;
        // tiny batches (the 16 x (1 .. 100)-sample calls of the adaptive loop, src/run.jl:96-113): the latency variant -- 8-sample
        // tiles, the whole field of a tile in shared memory, every phase but the flux sums spread over the CTA -- while its tiles
        // fit in two waves of CTAs; measured (back-to-back launches of 16 x 1 samples) 16-22 -> 10-12 us at l <= 3, 88 -> 18 us at l = 6
        bool small_done = false;
        if (ctx->ln_small != 0 && g.m <= 1024 && la.N % 2 == 0 && (level == 0 || la.N % 4 == 0)) {
            const long long tps = (long long)((n + LNS_BT - 1) / LNS_BT);
            // fine grids: a thread-block cluster of cs CTAs per sample tile -- every rank contracts, exponentiates and inverts
            // 1/cs of the rows (a single SM takes the level-6 basis in at ~28 bytes per clock: 14 us for its 819 KB) and stores
            // its reciprocals into the leader's shared memory; at least 8 row tiles per rank, all clusters resident at once
            int cs = 1;
            if (ctx->ln_cluster != 0)
                for (int c = ctx->ln_cluster > 1 ? ctx->ln_cluster : 8; c >= 2; c >>= 1) {
                    if (la.ntiles < 8 * c || tps * R * c > ctx->sm_count) continue;
                    const size_t sm_c = lns_smem_bytes(la.mpad, la.ntiles, g.m, c);
                    if (sm_c <= 220 * 1024 && tps * R <= MLE_CLUSTERS_BY_MODE(g.mode, lognormal1d_kc, LNS_THREADS, sm_c, c)) { cs = c; break; }
                }
            const size_t smem_l = lns_smem_bytes(la.mpad, la.ntiles, g.m, cs);
            if (smem_l <= 220 * 1024 && cs > 1) {
                gx = tps;  // one cluster per sample tile, all of them resident at once
                rc = ensure_dev(ctx, &ctx->d_partials, &ctx->partials_cap, (size_t)R * V * (size_t)gx);
                if (rc) return rc;
                fa.partials = ctx->d_partials;
                dim3 grid((unsigned)(gx * cs), (unsigned)R);
                MLE_LAUNCH_CLUSTER_BY_MODE(g.mode, lognormal1d_kc, grid, LNS_THREADS, smem_l, ctx->stream, cs, g, la, fa, samples_dev, tps);
                small_done = true;
            } else if (smem_l <= 220 * 1024) {
                const int occ = MLE_OCC_BY_MODE(g.mode, lognormal1d_kl, LNS_THREADS, smem_l);
                if (ctx->ln_small > 0 || (unsigned long long)tps * (unsigned long long)R <= 2ull * (unsigned long long)ctx->sm_count * (unsigned long long)occ) {
                    gx = grid_x_for(ctx, tps, R, occ);
                    rc = ensure_dev(ctx, &ctx->d_partials, &ctx->partials_cap, (size_t)R * V * (size_t)gx);
                    if (rc) return rc;
                    fa.partials = ctx->d_partials;
                    dim3 grid((unsigned)gx, (unsigned)R);
                    MLE_LAUNCH_BY_MODE(g.mode, lognormal1d_kl, grid, LNS_THREADS, smem_l, ctx->stream, g, la, fa, samples_dev, tps);
                    small_done = true;
                }
            }
        }
        if (small_done) {
        } else if (block) {
            const int bt = 32, threads = lnb_threads(bt), W = threads / 32;
            const long long tps = (long long)((n + bt - 1) / bt);
            const int qdims = g.m < 256 ? ((g.m + 3) & ~3) : 256;  // whole generator iterations: ceil(m / 4) * 4 dims
            const size_t fbytes = (size_t)LN_RC * lnb_ldf(bt) * 8, qbytes = (size_t)bt * qdims * 2;
            const size_t smem = (size_t)la.mpad * lnb_ldx(bt) * 8 + fbytes + (size_t)3 * bt * sizeof(SweepState) +
                                (size_t)bt * 8 + (size_t)la.mpad * 8 + ((W + 3) & ~3) * 4 +
                                (lat ? (size_t)g.m * 16 : 0) + (qbytes <= fbytes ? 0 : qbytes) + 16;
            // small batches (the CTAs of the call do not even fill the SMs twice: every CTA is practically alone on its SM and
            // bound by the L2 latency of the basis fragments): the variant that stages each chunk's A panel with a TMA bulk load
            const bool staged = ctx->ln_staged >= 0 ? ctx->ln_staged != 0 : (unsigned long long)tps * (unsigned long long)R <= 2ull * (unsigned long long)ctx->sm_count;
            const size_t smem_s = smem + 128 + (size_t)2 * 4 * (la.mpad / 4) * 32 * 8 + 16;
            if (staged && smem_s <= 220 * 1024) {
                const int occ = MLE_OCC_BY_MODE(g.mode, lognormal1d_ks, threads, smem_s);
                gx = churn_grid_x(grid_x_for(ctx, tps, R, occ), tps);
                rc = ensure_dev(ctx, &ctx->d_partials, &ctx->partials_cap, (size_t)R * V * (size_t)gx);
                if (rc) return rc;
                fa.partials = ctx->d_partials;
                dim3 grid((unsigned)gx, (unsigned)R);
                MLE_LAUNCH_BY_MODE(g.mode, lognormal1d_ks, grid, threads, smem_s, ctx->stream, g, la, fa, samples_dev, tps);
            } else {
                const int occ = MLE_OCC_BY_MODE(g.mode, lognormal1d_kb, threads, smem);
                gx = churn_grid_x(grid_x_for(ctx, tps, R, occ), tps);
                rc = ensure_dev(ctx, &ctx->d_partials, &ctx->partials_cap, (size_t)R * V * (size_t)gx);
                if (rc) return rc;
                fa.partials = ctx->d_partials;
                dim3 grid((unsigned)gx, (unsigned)R);
                MLE_LAUNCH_BY_MODE(g.mode, lognormal1d_kb, grid, threads, smem, ctx->stream, g, la, fa, samples_dev, tps);
            }
        } else {
        const int ws = 16;  // samples per warp (the 8-sample instantiation never won a level and was dropped)
        const long long tps = (long long)((n + LN_BT - 1) / LN_BT);
        size_t smem = ln1d_smem_bytes(la.mpad, lat ? g.m : 0, ws);
        if (const char *e = getenv("MLE_LN_PAD")) smem += (size_t)atoi(e);  // occupancy experiment: extra bytes per CTA
        const int threads = ln_threads(ws);
        const int occ = MLE_OCC_BY_MODE(g.mode, lognormal1d_k16, threads, smem);
        gx = churn_grid_x(grid_x_for(ctx, tps, R, occ), tps);
        rc = ensure_dev(ctx, &ctx->d_partials, &ctx->partials_cap, (size_t)R * V * (size_t)gx);
        if (rc) return rc;
        fa.partials = ctx->d_partials;
        dim3 grid((unsigned)gx, (unsigned)R);
        MLE_LAUNCH_BY_MODE(g.mode, lognormal1d_k16, grid, threads, smem, ctx->stream, g, la, fa, samples_dev, tps);
        }
    } else if (M->kind == MODEL_LOGNORMAL2D) {
        const int lx = index[0], ly = d == 2 ? index[1] : index[0];
        if (lx > 12 || ly > 12) return fail(ctx, MLE_ERR_ARG, "lognormal2d: index out of range");
        const int key = d == 2 ? lx + 16 * ly : lx;
        if (key >= MLE_MAX_LEVELS || !M->d_basis[key]) return fail(ctx, MLE_ERR_STATE, "lognormal2d: no KL basis set for this index (mle_model_set_basis)");
        if (m < M->basis_cols[key]) return fail(ctx, MLE_ERR_ARG, "lognormal2d: nb_of_uncertainties is smaller than the number of KL terms of the index's basis");
        g.m = M->basis_cols[key];
        // register-window kernel (power-of-two Nx <= 64) when its shared memory fits; else the shared-memory-window kernel
        bool rows_done = false;
        if (M->d_frag[key] && !ctx->ln2d_old) {
            const int Nx = M->n0 << lx, Ny = M->n0 << ly;
            int rows_rc = -1;
#define MLE_ROWS_BODY(WW, KK, CFG)                                                                                     \
            {                                                                                                          \
                using Cfg = CFG;                                                                                       \
                Lognormal2dRowsArgs ra{};                                                                              \
                ra.basis = M->d_frag[key]; ra.Nx = Nx; ra.Ny = Ny; ra.ndims = d; ra.lx = lx; ra.ly = ly;               \
                ra.mpad = M->frag_mpad[key]; ra.ntiles = M->frag_tiles[key];                                           \
                const size_t fbytes = (size_t)ra.ntiles * 8 * Cfg::LDF * 8, qbytes = (size_t)32 * Cfg::THREADS * 2;    \
                const size_t fixed = (size_t)ra.mpad * Cfg::LDX * 8 + (size_t)Cfg::CS * Cfg::SH * 8 +                 \
                                     (size_t)2 * Cfg::S * 8 + (size_t)2 * Cfg::S * sizeof(Partial) +                    \
                                     (((Cfg::THREADS / 32) + 3) & ~3) * 4 + 128;                                       \
                ra.use_scratch = (fixed + fbytes + (qbytes > fbytes ? qbytes : 0) > 200 * 1024) ? 1 : 0;               \
                const size_t smem = fixed + (ra.use_scratch ? qbytes : fbytes + (qbytes > fbytes ? qbytes : 0));       \
                if (smem <= 220 * 1024) {                                                                              \
                const int occ = MLE_OCC_BY_MODE(g.mode, KK, Cfg::THREADS, smem);                                       \
                /* samples per tile: S when the batch fills the resident CTAs, else as few as spread it over all of them */ \
                {                                                                                                      \
                    const long long slots = std::max<long long>(1, (long long)ctx->sm_count * occ / R);               \
                    long long spt = (long long)((n + (uint64_t)slots - 1) / (uint64_t)slots);                          \
                    spt = (spt + Cfg::CS - 1) / Cfg::CS * Cfg::CS;                                                     \
                    ra.spt = (int)std::min<long long>(Cfg::S, std::max<long long>(Cfg::CS, spt));                      \
                }                                                                                                      \
                const long long tps = (long long)((n + (uint64_t)ra.spt - 1) / (uint64_t)ra.spt);                      \
                gx = grid_x_for(ctx, tps, R, occ);                                                                     \
                rc = ensure_dev(ctx, &ctx->d_partials, &ctx->partials_cap, (size_t)R * V * (size_t)gx);                \
                if (rc) return rc;                                                                                     \
                fa.partials = ctx->d_partials;                                                                         \
                {                                                                                                      \
                    const size_t park = (size_t)gx * R * ROWS_PARK * Cfg::THREADS;                                            \
                    rc = ensure_dev(ctx, &ctx->d_scratch, &ctx->scratch_cap, park + (ra.use_scratch ? (size_t)gx * R * (fbytes / 8) : 0)); \
                    if (rc) return rc;                                                                                 \
                    ra.park = ctx->d_scratch;                                                                          \
                    ra.scratch = ctx->d_scratch + park;                                                                \
                }                                                                                                      \
                dim3 grid((unsigned)gx, (unsigned)R);                                                                  \
                MLE_LAUNCH_BY_MODE(g.mode, KK, grid, Cfg::THREADS, smem, ctx->stream, g, ra, fa, samples_dev, tps); \
                rows_rc = 0;                                                                                           \
                }                                                                                                      \
            }
#define MLE_ROWS_CASE(WW, KK) case WW: MLE_ROWS_BODY(WW, KK, Rows2dCfg<WW>) break;
            // blocked DMMA solver: multilevel differences on square grids of 64 / 128 cells
            // (at 64 cells the register-window solver is still faster -- four samples side by side hide its latencies --, so the
            // blocked solver is only taken there when forced with MLE_2D_DMMA=2)
            if (ctx->ln2d_dmma >= 2 && d == 1 && Nx == Ny && Nx == 64) MLE_ROWS_BODY(64, lognormal2d_d64, Dmma2dCfg<64>)
            else if (ctx->ln2d_dmma && d == 1 && Nx == Ny && Nx == 128) MLE_ROWS_BODY(128, lognormal2d_d128, Dmma2dCfg<128>)
            if (rows_rc != 0)
            switch (Nx) {
                MLE_ROWS_CASE(4, lognormal2d_r4)
                MLE_ROWS_CASE(8, lognormal2d_r8)
                MLE_ROWS_CASE(16, lognormal2d_r16)
                MLE_ROWS_CASE(32, lognormal2d_r32)
                MLE_ROWS_CASE(64, lognormal2d_r64)
                MLE_ROWS_CASE(128, lognormal2d_r128)
                default: break;
            }
#undef MLE_ROWS_CASE
#undef MLE_ROWS_BODY
            rows_done = rows_rc == 0;
        }
        if (!rows_done) {
        Lognormal2dArgs la{};
        la.basisT = M->d_basis[key];
        la.Nx = M->n0 << lx;
        la.Ny = M->n0 << ly;
        la.nodes_pad = M->basis_mpad[key];
        la.ndims = d;
        la.lx = lx;
        la.ly = ly;
        const int nn = la.Nx - 1, W = la.Nx, nodes = (la.Nx + 1) * (la.Ny + 1), mpad = (g.m + 3) & ~3;
        la.scratch_per_group = 2ll * nn * nn + nodes;
        // a warp per sample while window, centre blocks and field of 4 samples fit in shared memory; else a CTA per sample
        // with the centre blocks and the field in (L2-resident) global scratch
        const size_t warp_smem = ((size_t)mpad + (size_t)W * W + 5 * (size_t)W + (size_t)la.scratch_per_group) * 4 * sizeof(double);
        const bool big = la.Nx > 32 || warp_smem > 200 * 1024;
        const int threads = big ? 256 : 128, groups = big ? 1 : 4;
        la.use_scratch = big ? 1 : 0;
        const size_t per_group = (size_t)mpad + (size_t)W * W + 5 * (size_t)W + (big ? 0 : (size_t)la.scratch_per_group);
        const size_t smem = per_group * groups * sizeof(double) + 2 * groups * sizeof(Partial) + 16;
        if (smem > 220 * 1024) return fail(ctx, MLE_ERR_ARG, "lognormal2d: KL terms x grid do not fit in shared memory");
        const long long tps = (long long)((n + groups - 1) / groups);
        const int occ = big ? MLE_OCC_BY_MODE(g.mode, lognormal2d_c, threads, smem) : MLE_OCC_BY_MODE(g.mode, lognormal2d_w, threads, smem);
        gx = grid_x_for(ctx, tps, R, occ);
        rc = ensure_dev(ctx, &ctx->d_partials, &ctx->partials_cap, (size_t)R * V * (size_t)gx);
        if (rc) return rc;
        fa.partials = ctx->d_partials;
        if (big) {
            rc = ensure_dev(ctx, &ctx->d_scratch, &ctx->scratch_cap, (size_t)gx * R * groups * (size_t)la.scratch_per_group);
            if (rc) return rc;
            la.scratch = ctx->d_scratch;
        }
        dim3 grid((unsigned)gx, (unsigned)R);
        if (big) MLE_LAUNCH_BY_MODE(g.mode, lognormal2d_c, grid, threads, smem, ctx->stream, g, la, fa, samples_dev);
        else MLE_LAUNCH_BY_MODE(g.mode, lognormal2d_w, grid, threads, smem, ctx->stream, g, la, fa, samples_dev);
        }
    } else {
        return fail(ctx, MLE_ERR_ARG, "unknown model kind");
    }
    ctx->launches += 1;
    CU_TRY(ctx, cudaGetLastError());
    return MLE_OK;
}

extern "C" int mle_sample_batch_dev(mle_ctx ctx, mle_model M, mle_lattice lat, mle_dists dists, const int32_t *index, int d,
                                    const double *shifts_dev, int R, uint64_t k0, uint64_t n, int m, uint64_t mc_seed,
                                    double *moments_dev, double *samples_dev) {
    MLE_API_SCOPE;
    return launch_batch(ctx, M, lat, dists, index, d, shifts_dev, R, k0, n, m, mc_seed, moments_dev, samples_dev);
}

#ifdef MLE_LNS_TIMING
extern "C" int mle_debug_lns_timing(long long *out16) { return cudaMemcpyFromSymbol(out16, g_lns_t, 16 * sizeof(long long)) == cudaSuccess ? 0 : -2; }
#endif

extern "C" int mle_moments_merge(double *acc, const double *part, size_t count);
static int update_samples_group(mle_ctx ctx, mle_model M, mle_lattice lat, mle_dists dists, int count, const int32_t *indices, int d,
                                const double *const *shifts, const int *R, const uint64_t *k0, const uint64_t *n, const int *m,
                                uint64_t mc_seed, double *const *moments, double *gpu_seconds);

// Wait for the launch that carries sequence number `seq` to raise the mapped completion flag.  The host spins on its own
// memory (the GPU writes the moment table and then the flag over PCIe, __threadfence_system in between), which saves the
// stream synchronisation and the device-to-host copy of the classic path: 10-15 us per call, i.e. most of a small call.  The
// stream is queried now and then so that a failed launch is reported instead of spinning forever.
static int wait_flag(mle_ctx ctx, unsigned seq, const char *who) {
    volatile unsigned *flag = ctx->h_flag;
    for (unsigned long long spins = 0;; ++spins) {
        if (*flag == seq) { __sync_synchronize(); return MLE_OK; }
        if ((spins & 0x3fff) == 0x3fff) {
            cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q == cudaSuccess) {  // everything has run: the flag must be there now
                if (*flag == seq) { __sync_synchronize(); return MLE_OK; }
                return fail(ctx, MLE_ERR_CUDA, std::string(who) + ": the kernel finished without raising the completion flag");
            }
            if (q != cudaErrorNotReady) { (void)cudaGetLastError(); return fail(ctx, MLE_ERR_CUDA, std::string(who) + ": " + cudaGetErrorString(q)); }
        }
    }
}

static int ensure_res(mle_ctx ctx, size_t need) {
    if (ctx->res_cap >= need) return MLE_OK;
    if (ctx->h_res) { cudaStreamSynchronize(ctx->stream); cudaFreeHost(ctx->h_res); }
    ctx->h_res = nullptr; ctx->d_res = nullptr; ctx->res_cap = 0;
    const size_t want = need + need / 2 + 512;
    if (cudaHostAlloc(reinterpret_cast<void **>(&ctx->h_res), want * sizeof(double), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void **>(&ctx->d_res), ctx->h_res, 0) != cudaSuccess) {
        (void)cudaGetLastError();
        if (ctx->h_res) cudaFreeHost(ctx->h_res);
        ctx->h_res = nullptr; ctx->d_res = nullptr;
        return fail(ctx, MLE_ERR_NOMEM, "cudaHostAlloc(mapped result table) failed");
    }
    ctx->res_cap = want;
    return MLE_OK;
}

extern "C" int mle_sample_batch(mle_ctx ctx, mle_model M, mle_lattice lat, mle_dists dists, const int32_t *index, int d,
                                const double *shifts, int R, uint64_t k0, uint64_t n, int m, uint64_t mc_seed, double *moments,
                                double *samples, double *gpu_seconds) {
    MLE_API_SCOPE;
    if (!ctx || !M || !moments) return MLE_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if (R < 1) return fail(ctx, MLE_ERR_ARG, "number of shifts must be at least 1");
    if (!ctx->peers.empty() && !samples) {  // multi-GPU context: split over the members (raw samples stay on member 0)
        if (!index) return MLE_ERR_ARG;
        const double *shp = shifts;
        double *mp = moments;
        return update_samples_group(ctx, M, lat, dists, 1, index, d, shifts ? &shp : nullptr, &R, &k0, &n, &m, mc_seed, &mp, gpu_seconds);
    }
    const double *sh_dev = nullptr;
    int rc = upload_shifts(ctx, lat, shifts, R, &sh_dev);
    if (rc) return rc;
    const size_t nm = (size_t)R * 2 * (size_t)M->nqoi * 3;
    const size_t ns = samples ? (size_t)R * 2 * (size_t)M->nqoi * n : 0;
    // zero-copy: the last CTA writes the moment table straight into mapped host memory and raises a flag the host polls
    const bool zc = ctx->zero_copy && !ns;
    rc = ensure_dev(ctx, &ctx->d_moments, &ctx->moments_cap, nm);
    if (rc == MLE_OK && zc) rc = ensure_res(ctx, nm);
    if (rc) return rc;
    if (!zc && nm + 64 > ctx->pinned_cap) {  // staging for the moments (classic path)
        ctx->shifts_staged = 0;
        rc = ensure_pinned(ctx, nm + (shifts && lat ? (size_t)R * (size_t)lat->s : 0) + 4096);
        if (rc) return rc;
        if (shifts && lat) { sh_dev = nullptr; rc = upload_shifts(ctx, lat, shifts, R, &sh_dev); if (rc) return rc; }
    }
    if (ns) {  // raw samples: device buffer kept across calls
        rc = ensure_dev(ctx, &ctx->d_samples, &ctx->samples_cap, ns);
        if (rc) return rc;
    }
    if (gpu_seconds) cudaEventRecord(ctx->ev0, ctx->stream);  // timing is optional: two event records cost ~4 us per call
    const unsigned seq = ++ctx->seq ? ctx->seq : ++ctx->seq;  // never 0 (the flag's initial value)
    rc = launch_batch(ctx, M, lat, dists, index, d, sh_dev, R, k0, n, m, mc_seed, ctx->d_moments, ns ? ctx->d_samples : nullptr,
                      zc ? ctx->d_res : nullptr, zc ? ctx->d_flag : nullptr, seq);
    if (gpu_seconds) cudaEventRecord(ctx->ev1, ctx->stream);
    if (rc != MLE_OK) { cudaStreamSynchronize(ctx->stream); return rc; }
    if (zc) {
        rc = wait_flag(ctx, seq, "mle_sample_batch");
        if (rc) return rc;
        memcpy(moments, ctx->h_res, nm * sizeof(double));
        if (gpu_seconds) {
            cudaError_t e = cudaEventSynchronize(ctx->ev1);
            if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(ctx, MLE_ERR_CUDA, std::string("mle_sample_batch: ") + cudaGetErrorString(e)); }
        }
    } else {
        // the moment table comes back through the pinned staging area (behind the shifts image): a pageable destination
        // would make the copy synchronous inside the driver
        const size_t mom_off = (shifts && lat) ? (size_t)R * (size_t)lat->s : 0;
        double *stage = (ctx->h_pinned && mom_off + nm <= ctx->pinned_cap) ? ctx->h_pinned + mom_off : nullptr;
        cudaError_t e = cudaMemcpyAsync(stage ? stage : moments, ctx->d_moments, nm * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && ns) e = cudaMemcpyAsync(samples, ctx->d_samples, ns * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(ctx, MLE_ERR_CUDA, std::string("mle_sample_batch: ") + cudaGetErrorString(e)); }
        if (stage) memcpy(moments, stage, nm * sizeof(double));
    }
    if (gpu_seconds) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        *gpu_seconds = (double)ms * 1e-3;
    }
    return MLE_OK;
}

// Unbiased estimators: one call evaluates every idx in zero(index):index on the same points and reduces the per-sample sums
// Z = sum_idx inv_prob[idx] * dQ_idx on the device (include/mle_cuda.h).  Every idx is a launch of the model's normal kernel that
// leaves its raw samples in a device buffer the library keeps across calls; nothing per-sample crosses PCIe.
extern "C" int mle_sample_unbiased(mle_ctx ctx, mle_model M, mle_lattice lat, mle_dists dists, const int32_t *index, int d,
                                   const double *shifts, int R, uint64_t k0, uint64_t n, int m, uint64_t mc_seed,
                                   const double *inv_prob, double *moments, double *acc, double *gpu_seconds) {
    MLE_API_SCOPE;
    if (!ctx || !M || !index || !inv_prob || !moments || !acc) return MLE_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    if (R < 1) return fail(ctx, MLE_ERR_ARG, "number of shifts must be at least 1");
    if (d != M->ndims || d < 1 || d > 8) return fail(ctx, MLE_ERR_ARG, "index dimension does not match the model");
    if (n == 0) return fail(ctx, MLE_ERR_ARG, "sample!: number of samples must be positive");
    size_t nidx = 1;
    for (int t = 0; t < d; ++t) {
        if (index[t] < 0) return fail(ctx, MLE_ERR_ARG, "index must be non-negative");
        nidx *= (size_t)index[t] + 1;
    }
    const double *sh_dev = nullptr;
    int rc = upload_shifts(ctx, lat, shifts, R, &sh_dev);
    if (rc) return rc;
    const size_t q = (size_t)M->nqoi, nm = (size_t)R * 2 * q * 3, na = (size_t)R * q * 3;
    const size_t ntab = nidx * nm + na;
    rc = ensure_dev(ctx, &ctx->d_moments, &ctx->moments_cap, ntab);
    if (rc == MLE_OK) rc = ensure_dev(ctx, &ctx->d_samples, &ctx->samples_cap, (size_t)R * 2 * q * n);
    if (rc == MLE_OK) rc = ensure_dev(ctx, &ctx->d_zacc, &ctx->zacc_cap, (size_t)R * q * n);
    if (rc) return rc;
    const size_t sh_off = (shifts && lat) ? (size_t)R * (size_t)lat->s : 0;
    if (sh_off + ntab + 64 > ctx->pinned_cap) {  // staging for the tables behind the shifts image
        ctx->shifts_staged = 0;
        rc = ensure_pinned(ctx, sh_off + ntab + 4096);
        if (rc) return rc;
        if (shifts && lat) { sh_dev = nullptr; rc = upload_shifts(ctx, lat, shifts, R, &sh_dev); if (rc) return rc; }
    }
    if (gpu_seconds) cudaEventRecord(ctx->ev0, ctx->stream);
    const unsigned long long total = (unsigned long long)R * q * n;
    const unsigned zgrid = (unsigned)std::min<unsigned long long>((total + 255) / 256, (unsigned long long)ctx->sm_count * 8);
    int32_t idx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t i = 0; i < nidx; ++i) {  // column-major walk of zero(index):index: the first entry runs fastest
        rc = launch_batch(ctx, M, lat, dists, idx, d, sh_dev, R, k0, n, m, mc_seed, ctx->d_moments + i * nm, ctx->d_samples);
        if (rc != MLE_OK) { cudaStreamSynchronize(ctx->stream); return rc; }
        zacc_kernel<<<zgrid, 256, 0, ctx->stream>>>(ctx->d_zacc, ctx->d_samples, inv_prob[i], i == 0 ? 1 : 0, (unsigned long long)n, total);
        ++ctx->launches;
        for (int t = 0; t < d; ++t) {
            if (++idx[t] <= index[t]) break;
            idx[t] = 0;
        }
    }
    zacc_moments_kernel<<<(unsigned)(R * q), 256, 0, ctx->stream>>>(ctx->d_zacc, (unsigned long long)n, ctx->d_moments + nidx * nm);
    ++ctx->launches;
    if (gpu_seconds) cudaEventRecord(ctx->ev1, ctx->stream);
    double *stage = ctx->h_pinned + sh_off;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(stage, ctx->d_moments, ntab * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(ctx, MLE_ERR_CUDA, std::string("mle_sample_unbiased: ") + cudaGetErrorString(e)); }
    memcpy(moments, stage, nidx * nm * sizeof(double));
    memcpy(acc, stage + nidx * nm, na * sizeof(double));
    if (gpu_seconds) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        *gpu_seconds = (double)ms * 1e-3;
    }
    return MLE_OK;
}

// `update_samples`: several sample! requests, one synchronisation (include/mle_cuda.h).  Split in two halves so that a
// multi-GPU context can issue the work of every member before it waits for any of them.
struct IssueState {
    bool used = false, zc = false, timed = false;
    size_t sh_total = 0, mom_total = 0;
    unsigned seq = 0;
};

static int update_issue(mle_ctx ctx, mle_model M, mle_lattice lat, mle_dists dists, int count, const int32_t *const *indices,
                        int d, const double *const *shifts, const int *R, const uint64_t *k0, const uint64_t *n, const int *m,
                        uint64_t mc_seed, bool timed, IssueState *st) {
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    // one arena for all shifts (device + pinned host) and one for all moment tables (mapped host when zero-copy)
    size_t sh_total = 0, mom_total = 0;
    for (int e = 0; e < count; ++e) {
        if (R[e] < 1) return fail(ctx, MLE_ERR_ARG, "update_samples: every entry needs R >= 1 and a moment table");
        if (shifts && shifts[e] && !lat) return fail(ctx, MLE_ERR_ARG, "MC sampling takes no shifts");
        if (lat && shifts && shifts[e]) sh_total += (size_t)R[e] * (size_t)lat->s;
        mom_total += (size_t)R[e] * 2 * (size_t)M->nqoi * 3;
    }
    const bool zc = ctx->zero_copy != 0;
    const double *batch_before = ctx->d_batch;
    int rc = ensure_dev(ctx, &ctx->d_batch, &ctx->batch_cap, sh_total + mom_total);
    if (rc) return rc;
    if (ctx->d_batch != batch_before) ctx->batch_staged = 0;
    rc = ensure_pinned(ctx, sh_total + (zc ? 0 : mom_total) + 4096);
    if (rc) return rc;
    if (zc) { rc = ensure_res(ctx, mom_total); if (rc) return rc; }
    // unchanged shifts are not uploaded again (the adaptive loop calls with the same shifts of every index for a whole run)
    bool same = ctx->batch_staged == sh_total && sh_total > 0;
    size_t off = 0;
    for (int e = 0; e < count && same; ++e)
        if (lat && shifts && shifts[e]) {
            const size_t cnt = (size_t)R[e] * (size_t)lat->s;
            same = memcmp(ctx->h_pinned + off, shifts[e], cnt * sizeof(double)) == 0;
            off += cnt;
        }
    ctx->shifts_staged = 0;  // the pinned area is this entry point's staging arena now
    if (!same) {
        ctx->batch_staged = 0;
        off = 0;
        for (int e = 0; e < count; ++e)
            if (lat && shifts && shifts[e]) {
                const size_t cnt = (size_t)R[e] * (size_t)lat->s;
                memcpy(ctx->h_pinned + off, shifts[e], cnt * sizeof(double));
                off += cnt;
            }
        if (sh_total) CU_TRY(ctx, cudaMemcpyAsync(ctx->d_batch, ctx->h_pinned, sh_total * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        ctx->batch_staged = sh_total;
    }
    double *mom_base = ctx->d_batch + sh_total;  // device tables; zero-copy: each launch's last CTA mirrors its table to h_res
    const unsigned seq = ++ctx->seq ? ctx->seq : ++ctx->seq;  // never 0 (the flag's initial value)
    if (timed) cudaEventRecord(ctx->ev0, ctx->stream);
    size_t sh_off = 0, mom_off = 0;
    for (int e = 0; e < count && rc == MLE_OK; ++e) {
        const double *sh_dev = nullptr;
        if (lat && shifts && shifts[e]) { sh_dev = ctx->d_batch + sh_off; sh_off += (size_t)R[e] * (size_t)lat->s; }
        const bool last = e == count - 1;
        rc = launch_batch(ctx, M, lat, dists, indices[e], d, sh_dev, R[e], k0[e], n[e], m[e], mc_seed, mom_base + mom_off,
                          nullptr, zc ? ctx->d_res + mom_off : nullptr, (zc && last) ? ctx->d_flag : nullptr, seq);
        mom_off += (size_t)R[e] * 2 * (size_t)M->nqoi * 3;
    }
    if (timed) cudaEventRecord(ctx->ev1, ctx->stream);
    if (rc != MLE_OK) { cudaStreamSynchronize(ctx->stream); return rc; }
    st->used = true; st->zc = zc; st->timed = timed; st->sh_total = sh_total; st->mom_total = mom_total; st->seq = seq;
    return MLE_OK;
}

static int update_collect(mle_ctx ctx, mle_model M, int count, const int *R, double *const *moments, const IssueState &st,
                          double *gpu_seconds) {
    if (!st.used) return MLE_OK;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const double *host_mom;
    if (st.zc) {
        int rc = wait_flag(ctx, st.seq, "mle_update_samples");
        if (rc) return rc;
        host_mom = ctx->h_res;
    } else {
        cudaError_t err = cudaMemcpyAsync(ctx->h_pinned + st.sh_total, ctx->d_batch + st.sh_total, st.mom_total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
        if (err != cudaSuccess) { (void)cudaGetLastError(); return fail(ctx, MLE_ERR_CUDA, std::string("mle_update_samples: ") + cudaGetErrorString(err)); }
        host_mom = ctx->h_pinned + st.sh_total;
    }
    size_t mom_off = 0;
    for (int e = 0; e < count; ++e) {
        const size_t nm = (size_t)R[e] * 2 * (size_t)M->nqoi * 3;
        memcpy(moments[e], host_mom + mom_off, nm * sizeof(double));
        mom_off += nm;
    }
    if (st.timed && gpu_seconds) {
        cudaError_t e = cudaEventSynchronize(ctx->ev1);
        if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(ctx, MLE_ERR_CUDA, std::string("mle_update_samples: ") + cudaGetErrorString(e)); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        *gpu_seconds = (double)ms * 1e-3;
    }
    return MLE_OK;
}

// contiguous balanced slice i of g over `total` items: [lo, lo + len)
static inline void split_range(uint64_t total, int g, int i, uint64_t *lo, uint64_t *len) {
    const uint64_t base = total / (uint64_t)g, rem = total % (uint64_t)g;
    *len = base + ((uint64_t)i < rem ? 1 : 0);
    *lo = (uint64_t)i * base + std::min<uint64_t>((uint64_t)i, rem);
}

// The requests of one update_samples call over the members of a (possibly multi-GPU) context.  Every (shift, point) is
// independent (pmap over the product, src/sample.jl:95), so a request splits without any data-path exchange:
//   - it uses g = clamp(samples / shard_min, 1, G) members: a request smaller than about one wave of CTAs per extra GPU
//     stays on one device (more devices would add launch latency, not throughput);
//   - R >= g: WHOLE SHIFTS per member (contiguous, balanced).  A member's moment triplets are final and land at their place
//     in the caller's table: no merge at all;
//   - R < g (MC: R = 1): every member takes a slice of the point numbers of every shift; the members' triplets are merged on
//     the host in member order (Chan), deterministically.
// All members' kernels are issued before the first wait; each member signals through its own mapped flag.
static int update_samples_group(mle_ctx ctx, mle_model M, mle_lattice lat, mle_dists dists, int count, const int32_t *indices, int d,
                                const double *const *shifts, const int *R, const uint64_t *k0, const uint64_t *n, const int *m,
                                uint64_t mc_seed, double *const *moments, double *gpu_seconds) {
    const int G = group_size(ctx);
    const size_t V3 = (size_t)2 * (size_t)M->nqoi * 3;
    struct Sub { std::vector<const int32_t *> idx; std::vector<const double *> sh; std::vector<int> R, m, entry; std::vector<uint64_t> k0, n; std::vector<double *> mom; };
    std::vector<Sub> sub((size_t)G);
    std::vector<std::vector<double>> tmp;  // member tables of point-split entries
    tmp.reserve((size_t)count * (size_t)G);
    struct Merge { int entry; std::vector<double *> parts; };
    std::vector<Merge> merges;
    for (int e = 0; e < count; ++e) {
        if (R[e] < 1 || !moments[e]) return fail(ctx, MLE_ERR_ARG, "update_samples: every entry needs R >= 1 and a moment table");
        const uint64_t total = n[e] * (uint64_t)R[e];
        int g = (int)std::min<uint64_t>((uint64_t)G, std::max<uint64_t>(1, total / std::max<uint64_t>(1, ctx->shard_min)));
        const bool by_shift = R[e] >= g;
        if (!by_shift && n[e] < (uint64_t)g) g = (int)std::max<uint64_t>(1, n[e]);
        Merge mg; mg.entry = e;
        for (int i = 0; i < g; ++i) {
            Sub &S = sub[(size_t)i];
            uint64_t lo, len;
            S.idx.push_back(indices + (size_t)e * d);
            S.m.push_back(m[e]);
            S.entry.push_back(e);
            if (by_shift) {
                split_range((uint64_t)R[e], g, i, &lo, &len);
                S.sh.push_back(shifts && shifts[e] && lat ? shifts[e] + lo * (size_t)lat->s : nullptr);
                S.R.push_back((int)len); S.k0.push_back(k0[e]); S.n.push_back(n[e]);
                S.mom.push_back(moments[e] + lo * V3);
            } else {
                split_range(n[e], g, i, &lo, &len);
                S.sh.push_back(shifts ? shifts[e] : nullptr);
                S.R.push_back(R[e]); S.k0.push_back(k0[e] + lo); S.n.push_back(len);
                if (i == 0) S.mom.push_back(moments[e]);
                else { tmp.emplace_back((size_t)R[e] * V3); S.mom.push_back(tmp.back().data()); mg.parts.push_back(tmp.back().data()); }
            }
        }
        if (!mg.parts.empty()) merges.push_back(std::move(mg));
    }
    std::vector<IssueState> st((size_t)G);
    int rc = MLE_OK;
    for (int i = 0; i < G && rc == MLE_OK; ++i) {
        Sub &S = sub[(size_t)i];
        if (S.idx.empty()) continue;
        mle_ctx c = member_ctx(ctx, i);
        rc = update_issue(c, member_handle(M, i), member_handle(lat, i), member_handle(dists, i), (int)S.idx.size(), S.idx.data(), d,
                          S.sh.data(), S.R.data(), S.k0.data(), S.n.data(), S.m.data(), mc_seed, gpu_seconds != nullptr, &st[(size_t)i]);
        if (rc != MLE_OK && i > 0) peer_fail(ctx, c, rc);
    }
    double tmax = 0.0;
    for (int i = 0; i < G; ++i) {  // collect from every member that was issued, even after a failure elsewhere
        Sub &S = sub[(size_t)i];
        if (!st[(size_t)i].used) continue;
        mle_ctx c = member_ctx(ctx, i);
        double t = 0.0;
        const int rc2 = update_collect(c, member_handle(M, i), (int)S.idx.size(), S.R.data(), S.mom.data(), st[(size_t)i], gpu_seconds ? &t : nullptr);
        if (rc2 != MLE_OK && rc == MLE_OK) { rc = rc2; if (i > 0) peer_fail(ctx, c, rc2); }
        tmax = std::max(tmax, t);
    }
    if (G > 1) cudaSetDevice(ctx->device);
    if (rc != MLE_OK) return rc;
    for (const Merge &mg : merges)
        for (double *part : mg.parts) mle_moments_merge(moments[mg.entry], part, (size_t)R[mg.entry] * 2 * (size_t)M->nqoi);
    if (gpu_seconds) *gpu_seconds = tmax;
    return MLE_OK;
}

extern "C" int mle_update_samples(mle_ctx ctx, mle_model M, mle_lattice lat, mle_dists dists, int count, const int32_t *indices,
                                  int d, const double *const *shifts, const int *R, const uint64_t *k0, const uint64_t *n,
                                  const int *m, uint64_t mc_seed, double *const *moments) {
    MLE_API_SCOPE;
    if (!ctx || !M || count < 0 || (count && (!indices || !R || !k0 || !n || !m || !moments))) return MLE_ERR_ARG;
    if (count == 0) return MLE_OK;
    return update_samples_group(ctx, M, lat, dists, count, indices, d, shifts, R, k0, n, m, mc_seed, moments, nullptr);
}

// ---------------------------------------------------------------------------------------------------------------
// host-side moment algebra
// ---------------------------------------------------------------------------------------------------------------
extern "C" int mle_moments_merge(double *acc, const double *part, size_t count) {
    MLE_API_SCOPE;
    if (!acc || !part) return MLE_ERR_ARG;
    for (size_t i = 0; i < count; ++i) {
        Moment a{acc[3 * i], acc[3 * i + 1], acc[3 * i + 2]}, b{part[3 * i], part[3 * i + 1], part[3 * i + 2]};
        Moment o = moment_merge(a, b);
        acc[3 * i] = o.n; acc[3 * i + 1] = o.mean; acc[3 * i + 2] = o.m2;
    }
    return MLE_OK;
}

