// fp64_microbench.cu -- measures the FP64 roofs that MEASURED_PEAKS.json does not carry:
//   DFMA pipe, DMMA (mma.sync f64) pipe, both together, IEEE double division, and a store-only HBM stream.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/fp64_microbench tools/fp64_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void k_dfma(double *out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void k_dmma884(double *out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void k_dmma16816(double *out, int iters, double av, double bv) {
    double c[4][4], a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = av + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) { b[i] = bv - i; for (int j = 0; j < 4; ++j) c[i][j] = i + j; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) dmma16816(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 12345.678) out[0] = s;
}

// even warps DFMA, odd warps DMMA: if the two share one pipe the sum of rates stays at the single-pipe peak
__global__ void k_mixed(double *out, int iters, double a, double b) {
    if ((threadIdx.x >> 5) & 1) {
        double c[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
        if (s == 12345.678) out[0] = s;
    } else {
        double x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-9 + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += x[i];
        if (s == 12345.678) out[0] = s;
    }
}

__global__ void k_ddiv(double *out, int iters, double a) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 1.0 + threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __ddiv_rn(a, x[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void k_exp(double *out, int iters, double a) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 0.1 + threadIdx.x * 1e-3 + i * 0.01;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = exp(x[i]) * a;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

// dependent-issue latency of DFMA / DMMA: one warp, one chain, clock64 around it
__global__ void k_lat(double *out, long long *cyc, double a, double b) {
    double x = threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x = fma(x, a, b);
    }
    long long t1 = clock64();
    double c0 = x, c1 = -x;
    long long t2 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma884(c0, c1, a, b);
    }
    long long t3 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; }
    if (c0 + c1 + x == 12345.678) out[0] = x;
}

__global__ void k_store(double2 *out, size_t n2, double v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    double2 w = make_double2(v, v + 1);
    for (; i < n2; i += stride) out[i] = w;
}

__global__ void k_copy(const double2 *in, double2 *out, size_t n2) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) out[i] = in[i];
}

template <typename F>
static float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double *d; CK(cudaMalloc(&d, 1 << 20));
    const int iters = 4096, threads = 256;
    const int blocks = sms * 8;
    const double nthreads = (double)blocks * threads, nwarps = nthreads / 32;
    printf("{\"device\": \"%s\", \"sms\": %d", p.name, sms);
    float ms = time_ms([&] { k_dfma<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9); });
    printf(", \"dfma_tflops\": %.2f", 2.0 * 8 * iters * nthreads / ms * 1e-9);
    ms = time_ms([&] { k_dmma884<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9); });
    printf(", \"dmma_m8n8k4_tflops\": %.2f", 2.0 * 256 * 8 * iters * nwarps / ms * 1e-9);
    ms = time_ms([&] { k_dmma16816<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9); });
    printf(", \"dmma_m16n8k16_tflops\": %.2f", 2.0 * 2048 * 4 * iters * nwarps / ms * 1e-9);
    ms = time_ms([&] { k_mixed<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9); });
    printf(", \"mixed_dfma_plus_dmma_tflops\": %.2f", (2.0 * 8 * iters * nthreads / 2 + 2.0 * 256 * 8 * iters * nwarps / 2) / ms * 1e-9);
    ms = time_ms([&] { k_ddiv<<<blocks, threads>>>(d, iters / 8, 1.5); });
    printf(", \"ddiv_gops\": %.1f", 8.0 * (iters / 8) * nthreads / ms * 1e-6);
    ms = time_ms([&] { k_exp<<<blocks, threads>>>(d, iters / 8, 0.5); });
    printf(", \"exp_gops\": %.1f", 8.0 * (iters / 8) * nthreads / ms * 1e-6);
    {
        long long *cyc, h[2];
        CK(cudaMalloc(&cyc, 16));
        k_lat<<<1, 32>>>(d, cyc, 1.0000001, 1e-9);
        k_lat<<<1, 32>>>(d, cyc, 1.0000001, 1e-9);
        CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
        printf(", \"dfma_dependent_latency_cycles\": %.2f, \"dmma_dependent_latency_cycles\": %.2f", h[0] / 4096.0, h[1] / 4096.0);
    }
    const size_t bytes = (size_t)4 << 30;
    double2 *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    ms = time_ms([&] { k_store<<<sms * 16, 512>>>(a, bytes / 16, 1.0); });
    printf(", \"store_gbs\": %.1f", bytes / ms * 1e-6);
    ms = time_ms([&] { k_copy<<<sms * 16, 512>>>(a, b, bytes / 16); });
    printf(", \"copy_gbs\": %.1f", 2.0 * bytes / ms * 1e-6);
    ms = time_ms([&] { cudaMemsetAsync(a, 0, bytes); });
    printf(", \"memset_gbs\": %.1f", bytes / ms * 1e-6);
    printf("}\n");
    return 0;
}
