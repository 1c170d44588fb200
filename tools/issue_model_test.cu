// Does an FP64 instruction on B200 cost the warp scheduler ONE issue slot (the FP64 pipe then takes it for 2 cycles, and other
// instructions can issue in the shadow) or TWO (the slot is blocked while the half-rate pipe accepts the warp)?  Loop of 32
// independent DFMA (8 chains) with M independent integer instructions (LOP3/IADD3 on private registers) mixed in, 8 warps per
// scheduler; cycles per iteration and scheduler against M.  Also the same with DMMA.8x8x4 instead of DFMA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/issue_model_test tools/issue_model_test.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int M, bool TENSOR>
__global__ void __launch_bounds__(256) k(double *out, unsigned *iout, int iters, double a, double b) {
    double c[8];
    unsigned x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i] = threadIdx.x + i; x[i] = threadIdx.x * 7 + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (TENSOR) {
                    if (i < 4) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[2 * i]), "+d"(c[2 * i + 1]) : "d"(a), "d"(b));
                } else {
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c[i]) : "d"(a), "d"(b));
                }
                // M integer instructions per 32 FP64 instructions, spread evenly
                constexpr int per = (M + 31) / 32;
#pragma unroll
                for (int q = 0; q < per; ++q)
                    if ((r * 8 + i) * per + q < M) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[(i + q) & 7]) : "r"(it), "r"(x[(i + q + 1) & 7]));
            }
        }
    }
    double s = 0; unsigned t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += c[i]; t ^= x[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int M, bool TENSOR>
static void run(double *out, unsigned *iout, int sms) {
    const int iters = 20000;
    k<M, TENSOR><<<sms * 4, 256>>>(out, iout, 10, 1.0000001, 1e-9);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<M, TENSOR><<<sms * 4, 256>>>(out, iout, iters, 1.0000001, 1e-9);  // 4 CTAs x 8 warps = 32 warps per SM = 8 per scheduler
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    // per scheduler: 8 warps x iters iterations
    const double cyc = (double)ms * 1e-3 * clk * 1e3 / ((double)iters * 8.0);
    printf("{\"fp64\": \"%s\", \"fp64_per_iter\": %d, \"int_per_iter\": %d, \"cycles_per_warp_iteration\": %.1f}\n", TENSOR ? "DMMA.8x8x4" : "DFMA", TENSOR ? 16 : 32, M, cyc);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out; unsigned *iout;
    cudaMalloc(&out, sizeof(double) * sms * 4 * 256); cudaMalloc(&iout, 4 * sms * 4 * 256);
    run<0, false>(out, iout, sms); run<8, false>(out, iout, sms); run<16, false>(out, iout, sms); run<32, false>(out, iout, sms);
    run<64, false>(out, iout, sms); run<96, false>(out, iout, sms);
    run<0, true>(out, iout, sms); run<32, true>(out, iout, sms); run<64, true>(out, iout, sms); run<128, true>(out, iout, sms); run<192, true>(out, iout, sms);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
