// How fast can the KL-contraction inner loop run in isolation?  (DMMA m8n8k4, A fragments from global/L1, B fragments from
// shared memory, 2x2 tiles per warp, 4 warps per CTA -- the loop of lognormal1d_kernel without anything else)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int VARIANT>
__global__ void __launch_bounds__(128) k(const double *__restrict__ basis, double *out, int nchunks, int nks, int reps) {
    extern __shared__ double xi[];  // [100][36]
    const int LDX = 36;
    for (int e = threadIdx.x; e < 100 * LDX; e += 128) xi[e] = 1e-3 * (e % 97);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wr = warp >> 1, wc = warp & 1, lr = lane >> 2, lk = lane & 3;
    double acc = 0;
    for (int rep = 0; rep < reps; ++rep)
        for (int c = 0; c < nchunks; ++c) {
            const int rt0 = c * 4 + 2 * wr;
            double c00a = 0, c00b = 0, c01a = 0, c01b = 0, c10a = 0, c10b = 0, c11a = 0, c11b = 0;
            const double *A0 = basis + (size_t)rt0 * nks * 32 + lane, *A1 = A0 + (size_t)nks * 32;
            const double *B0 = xi + lk * LDX + 16 * wc + lr;
            if (VARIANT == 0) {  // 1-step register prefetch (the kernel's loop)
                double a0 = __ldg(A0), a1 = __ldg(A1), b0 = B0[0], b1 = B0[8];
#pragma unroll 5
                for (int ks = 0; ks < nks; ++ks) {
                    const int kn = (ks + 1 < nks) ? ks + 1 : ks;
                    const double na0 = __ldg(A0 + kn * 32), na1 = __ldg(A1 + kn * 32), nb0 = B0[kn * 4 * LDX], nb1 = B0[kn * 4 * LDX + 8];
                    dmma(c00a, c00b, a0, b0); dmma(c01a, c01b, a0, b1); dmma(c10a, c10b, a1, b0); dmma(c11a, c11b, a1, b1);
                    a0 = na0; a1 = na1; b0 = nb0; b1 = nb1;
                }
            } else if (VARIANT == 1) {  // no prefetch, fully unrolled by 5: let the compiler schedule loads early
#pragma unroll 5
                for (int ks = 0; ks < nks; ++ks) {
                    const double a0 = __ldg(A0 + ks * 32), a1 = __ldg(A1 + ks * 32), b0 = B0[ks * 4 * LDX], b1 = B0[ks * 4 * LDX + 8];
                    dmma(c00a, c00b, a0, b0); dmma(c01a, c01b, a0, b1); dmma(c10a, c10b, a1, b0); dmma(c11a, c11b, a1, b1);
                }
            } else {  // operands constant (upper bound: DMMA issue only)
                const double a0 = __ldg(A0), a1 = __ldg(A1), b0 = B0[0], b1 = B0[8];
#pragma unroll 5
                for (int ks = 0; ks < nks; ++ks) {
                    dmma(c00a, c00b, a0, b0); dmma(c01a, c01b, a0, b1); dmma(c10a, c10b, a1, b0); dmma(c11a, c11b, a1, b1);
                }
            }
            acc += c00a + c00b + c01a + c01b + c10a + c10b + c11a + c11b;
        }
    if (acc == 1.2345) out[0] = acc;
}
template <int V>
void run(const double *basis, double *out, int sms, int nchunks) {
    const int nks = 25, reps = 200 / nchunks + 1;
    cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int occ : {1, 2, 3, 4, 5}) {
        size_t smem = occ == 5 ? 40 * 1024 : occ == 4 ? 52 * 1024 : occ == 3 ? 70 * 1024 : occ == 2 ? 100 * 1024 : 100 * 1024;
        int blocks = sms * occ;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<V><<<blocks, 128, smem>>>(basis, out, nchunks, nks, reps);
        cudaEventRecord(e0);
        k<V><<<blocks, 128, smem>>>(basis, out, nchunks, nks, reps);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fma = (double)blocks * 4 /*warps*/ * reps * nchunks * nks * 4 /*dmma*/ * 256.0;
        printf("variant %d chunks %2d CTAs/SM %d: %.2f TFLOP/s\n", V, nchunks, occ, 2 * fma / ms * 1e-9);
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *basis, *out; cudaMalloc(&basis, 34 * 4 * 25 * 32 * 8); cudaMemset(basis, 0, 34 * 4 * 25 * 32 * 8); cudaMalloc(&out, 64);
    for (int nch : {1, 33}) { run<0>(basis, out, p.multiProcessorCount, nch); run<1>(basis, out, p.multiProcessorCount, nch); run<2>(basis, out, p.multiProcessorCount, nch); }
    return 0;
}
