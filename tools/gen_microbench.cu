// gen_microbench.cu -- what bounds the inverse-normal generator?  Variants of the pass-1 loop (central approximant evaluated
// for U coordinates in flight) run in isolation at several occupancies; prints G coordinates/s per variant.
//   V0 pure math (lattice coordinate -> u -> central erfinv), constant-bank coefficients, result summed in registers
//   V1 V0 with the 14 coefficients held in registers
//   V2 V0 + integer class test + select + store to a shared-memory tile
//   V3 V2 + per-lane class masks (the library's pass 1)
//   template parameter U = coordinates in flight
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multilevelestimators.jl_b200/csrc -o tools/bin/gen_microbench tools/gen_microbench.cu
#include <cstdio>
#include <cstdlib>
#include "mle_device.cuh"
using namespace mle;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int V, int U>
__global__ void __launch_bounds__(128) k_gen(double *out, int iters, double t0, const double *zz) {
    __shared__ double tile[128 * 8];
    __shared__ double2 zs[128];
    if (threadIdx.x < 128) zs[threadIdx.x] = make_double2(zz[threadIdx.x], zz[128 + threadIdx.x]);
    __syncthreads();
    double p1[7], q1[7];
    if (V == 1) {
#pragma unroll
        for (int i = 0; i < 7; ++i) { p1[i] = ((volatile double *)zz)[256 + i]; q1[i] = ((volatile double *)zz)[263 + i]; }
    }
    const double t = t0 + threadIdx.x * 1e-3 + blockIdx.x * 1e-7;
    double acc = 0.0;
    unsigned m2 = 0, m3 = 0;
    for (int it = 0; it < iters; ++it) {
        double v[U];
        int cls[U];
#pragma unroll
        for (int e = 0; e < U; ++e) {
            const double2 w = zs[(it * U + e) & 127];
            const double x = lattice_coord(t, w.x, w.y);
            const double u = __fma_rn(2.0, x, -1.0);
            double c;
            if (V == 1) {
                const double tt = __dadd_rn(__dmul_rn(u, u), -0.5625);
                double a = p1[6], b = q1[6];
#pragma unroll
                for (int i = 5; i >= 0; --i) { a = __fma_rn(tt, a, p1[i]); b = __fma_rn(tt, b, q1[i]); }
                c = ddiv_normal(__dmul_rn(u, a), b);
            } else {
                c = erfinv_central(u);
            }
            c = __dmul_rn(kMisc[3], c);
            if (V >= 2) {
                const int k = erfinv_class(u);
                cls[e] = k;
                v[e] = k == 1 ? c : u;
            } else {
                v[e] = c;
            }
        }
#pragma unroll
        for (int e = 0; e < U; ++e) {
            if (V >= 2) tile[((it * U + e) & 7) * 128 + threadIdx.x] = v[e];
            else acc += v[e];
            if (V >= 3) { m2 |= (unsigned)(cls[e] == 2) << ((it * U + e) & 31); m3 |= (unsigned)(cls[e] == 3) << ((it * U + e) & 31); }
        }
    }
    if (V >= 2) { __syncthreads(); acc = tile[threadIdx.x] + (double)(m2 ^ m3); }
    if (acc == 12345.678) out[0] = acc;
}

template <int V, int U>
void run(const char *name, double *d_out, const double *d_z, int ctas_per_sm, int sms) {
    const int iters = 4096 / U * 4;
    const int grid = sms * ctas_per_sm;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    k_gen<V, U><<<grid, 128>>>(d_out, iters, 0.123, d_z);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int r = 0; r < 3; ++r) k_gen<V, U><<<grid, 128>>>(d_out, iters, 0.123, d_z);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double coords = 3.0 * grid * 128.0 * iters * U;
    printf("%-28s U=%d CTAs/SM=%d  %.1f G coords/s\n", name, U, ctas_per_sm, coords / (ms * 1e-3) * 1e-9);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double *d_out, *d_z;
    CK(cudaMalloc(&d_out, 64));
    double h[512];
    for (int i = 0; i < 128; ++i) { h[i] = 1234567.0 + 2 * i + 1; h[128 + i] = (i * 0.618033988749895) - (int)(i * 0.618033988749895); }
    const double P1[7] = {0.16030495584406622931e2, -0.9078495926296032665e2, 0.18644914861620987391e3, -0.16900142734642382420e3, 0.654546628479448704e2, -0.86421301158724779e1, 0.17605878213905900};
    const double Q1[7] = {0.14780647071513831611e2, -0.9137416702426031394e2, 0.21015790486205317714e3, -0.22210254121855132366e3, 0.10760453916055123830e3, -0.20601073032826544e2, 1.0};
    for (int i = 0; i < 7; ++i) { h[256 + i] = P1[i]; h[263 + i] = Q1[i]; }
    CK(cudaMalloc(&d_z, sizeof(h)));
    CK(cudaMemcpy(d_z, h, sizeof(h), cudaMemcpyHostToDevice));
    for (int c : {1, 2, 3, 5, 8}) {
        run<0, 4>("V0 math, const coeffs", d_out, d_z, c, sms);
        run<1, 4>("V1 math, register coeffs", d_out, d_z, c, sms);
        run<2, 4>("V2 + class/select/STS", d_out, d_z, c, sms);
        run<3, 4>("V3 + masks (pass 1)", d_out, d_z, c, sms);
    }
    for (int c : {3, 5}) {
        run<0, 1>("V0", d_out, d_z, c, sms);
        run<0, 2>("V0", d_out, d_z, c, sms);
        run<0, 8>("V0", d_out, d_z, c, sms);
        run<3, 2>("V3", d_out, d_z, c, sms);
        run<3, 8>("V3", d_out, d_z, c, sms);
    }
    printf("peak FP64 instruction rate: %d SMs x 64 lanes/clk; at 1.965 GHz = %.1f T/s -> %.0f G coords/s at 29 FP64 instructions per coordinate\n", sms, sms * 64 * 1.965e-3, sms * 64 * 1.965 / 29);
    return 0;
}
