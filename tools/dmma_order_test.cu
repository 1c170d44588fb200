// Is mma.sync.m8n8k4.f64 bit-identical to c = fma(a_k, b_k, c), k = 0..3 in order?  (decides whether the KL contraction can
// use the FP64 tensor instruction and still match the oracle's sequential fused multiply-add order bit for bit)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(const double *A, const double *B, const double *C, double *D, int ntiles) {
    int lane = threadIdx.x & 31;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const double *a = A + t * 32, *b = B + t * 32, *c = C + t * 64;
        // A row-major 8x4: lane holds A[row = lane/4][k = lane%4]; B col: lane holds B[k = lane%4][col = lane/4]
        double av = a[(lane >> 2) * 4 + (lane & 3)], bv = b[(lane & 3) * 8 + (lane >> 2)];
        double c0 = c[(lane >> 2) * 8 + (lane & 3) * 2], c1 = c[(lane >> 2) * 8 + (lane & 3) * 2 + 1];
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(av), "d"(bv));
        D[t * 64 + (lane >> 2) * 8 + (lane & 3) * 2] = c0;
        D[t * 64 + (lane >> 2) * 8 + (lane & 3) * 2 + 1] = c1;
    }
}
int main() {
    const int nt = 4096;
    double *hA = (double *)malloc(nt * 32 * 8), *hB = (double *)malloc(nt * 32 * 8), *hC = (double *)malloc(nt * 64 * 8), *hD = (double *)malloc(nt * 64 * 8);
    srand(1);
    auto rnd = [] { return (rand() / (double)RAND_MAX - 0.5) * exp((rand() % 40 - 20) * 0.5); };
    for (int i = 0; i < nt * 32; ++i) { hA[i] = rnd(); hB[i] = rnd(); }
    for (int i = 0; i < nt * 64; ++i) hC[i] = rnd();
    double *A, *B, *C, *D;
    cudaMalloc(&A, nt * 32 * 8); cudaMalloc(&B, nt * 32 * 8); cudaMalloc(&C, nt * 64 * 8); cudaMalloc(&D, nt * 64 * 8);
    cudaMemcpy(A, hA, nt * 32 * 8, cudaMemcpyHostToDevice); cudaMemcpy(B, hB, nt * 32 * 8, cudaMemcpyHostToDevice); cudaMemcpy(C, hC, nt * 64 * 8, cudaMemcpyHostToDevice);
    k<<<64, 32>>>(A, B, C, D, nt);
    cudaMemcpy(hD, D, nt * 64 * 8, cudaMemcpyDeviceToHost);
    long seq = 0, rev = 0, tot = 0;
    for (int t = 0; t < nt; ++t)
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j) {
                double c = hC[t * 64 + i * 8 + j], cr = c;
                for (int kk = 0; kk < 4; ++kk) c = fma(hA[t * 32 + i * 4 + kk], hB[t * 32 + kk * 8 + j], c);
                for (int kk = 3; kk >= 0; --kk) cr = fma(hA[t * 32 + i * 4 + kk], hB[t * 32 + kk * 8 + j], cr);
                double d = hD[t * 64 + i * 8 + j];
                seq += (d == c); rev += (d == cr); ++tot;
            }
    printf("{\"dmma_equals_sequential_fma_k_ascending\": %ld, \"k_descending\": %ld, \"total\": %ld}\n", seq, rev, tot);
    return 0;
}
