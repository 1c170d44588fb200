// DFMA throughput as a function of resident warps per SM and per-thread ILP (how much parallelism saturates the FP64 pipe?)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8 / (ILP > 8 ? 8 : ILP) ; ++r)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}
template <int ILP>
void run(double *d, int sms) {
    for (int warps : {4, 8, 16, 32, 64}) {
        int blocks = sms * (warps > 32 ? 2 : 1), threads = (warps > 32 ? warps / 2 : warps) * 32;
        const int iters = 20000;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<ILP><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e0);
        k<ILP><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double per_iter = (ILP > 8 ? ILP : 8);
        double fmas = per_iter * iters * (double)blocks * threads;
        printf("ILP %d warps/SM %2d: %.2f TFLOP/s\n", ILP, warps, 2 * fmas / ms * 1e-9);
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *d; cudaMalloc(&d, 1024);
    run<1>(d, p.multiProcessorCount); run<2>(d, p.multiProcessorCount); run<4>(d, p.multiProcessorCount); run<8>(d, p.multiProcessorCount);
    return 0;
}
