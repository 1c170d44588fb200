// Accumulation order of the wider FP64 tensor shapes on sm_100a: is mma.sync.m16n8k16.f64 (and k8, k4) bit-identical to
// c = fma(a_k, b_k, c) for k = 0, 1, ... in ascending order -- the property that lets mma.sync.m8n8k4.f64 stand for the oracle's
// sequential chain (tools/dmma_order_test.cu)?  Fragment layouts per the PTX ISA (groupID = lane / 4, t = lane % 4):
//   A (16 x K, row): a_i at row groupID + 8 (i % 2), col t + 4 (i / 2);   B (K x 8, col): b_i at row t + 4 i, col groupID;
//   C (16 x 8): c_i at row groupID + 8 (i / 2), col 2 t + (i % 2).
// nvcc -gencode arch=compute_100a,code=sm_100a -o tools/bin/dmma_order_test2 tools/dmma_order_test2.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

template <int K>
__global__ void run(const double *A, const double *B, const double *C, double *D, int tiles) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const double *a = A + (size_t)tile * 16 * K, *b = B + (size_t)tile * K * 8, *c = C + (size_t)tile * 128;
        double *d = D + (size_t)tile * 128;
        double c0 = c[g * 8 + 2 * t], c1 = c[g * 8 + 2 * t + 1], c2 = c[(g + 8) * 8 + 2 * t], c3 = c[(g + 8) * 8 + 2 * t + 1];
        if (K == 16) {
            double av[8], bv[4];
            for (int i = 0; i < 8; ++i) av[i] = a[(g + 8 * (i % 2)) * K + t + 4 * (i / 2)];
            for (int i = 0; i < 4; ++i) bv[i] = b[(t + 4 * i) * 8 + g];
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c0), "+d"(c1), "+d"(c2), "+d"(c3)
                         : "d"(av[0]), "d"(av[1]), "d"(av[2]), "d"(av[3]), "d"(av[4]), "d"(av[5]), "d"(av[6]), "d"(av[7]), "d"(bv[0]), "d"(bv[1]), "d"(bv[2]), "d"(bv[3]));
        } else if (K == 8) {
            double av[4], bv[2];
            for (int i = 0; i < 4; ++i) av[i] = a[(g + 8 * (i % 2)) * K + t + 4 * (i / 2)];
            for (int i = 0; i < 2; ++i) bv[i] = b[(t + 4 * i) * 8 + g];
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c0), "+d"(c1), "+d"(c2), "+d"(c3)
                         : "d"(av[0]), "d"(av[1]), "d"(av[2]), "d"(av[3]), "d"(bv[0]), "d"(bv[1]));
        } else {
            double av[2], bv[1];
            for (int i = 0; i < 2; ++i) av[i] = a[(g + 8 * i) * K + t];
            bv[0] = b[t * 8 + g];
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                         : "+d"(c0), "+d"(c1), "+d"(c2), "+d"(c3)
                         : "d"(av[0]), "d"(av[1]), "d"(bv[0]));
        }
        d[g * 8 + 2 * t] = c0; d[g * 8 + 2 * t + 1] = c1; d[(g + 8) * 8 + 2 * t] = c2; d[(g + 8) * 8 + 2 * t + 1] = c3;
    }
}

template <int K>
static void test() {
    const int tiles = 2048;
    size_t na = (size_t)tiles * 16 * K, nb = (size_t)tiles * K * 8, nc = (size_t)tiles * 128;
    double *hA = (double *)malloc(na * 8), *hB = (double *)malloc(nb * 8), *hC = (double *)malloc(nc * 8), *hD = (double *)malloc(nc * 8);
    srand(7 + K);
    auto rnd = []() { return (rand() / (double)RAND_MAX - 0.5) * exp2((double)(rand() % 20 - 10)); };
    for (size_t i = 0; i < na; ++i) hA[i] = rnd();
    for (size_t i = 0; i < nb; ++i) hB[i] = rnd();
    for (size_t i = 0; i < nc; ++i) hC[i] = rnd();
    double *dA, *dB, *dC, *dD;
    cudaMalloc(&dA, na * 8); cudaMalloc(&dB, nb * 8); cudaMalloc(&dC, nc * 8); cudaMalloc(&dD, nc * 8);
    cudaMemcpy(dA, hA, na * 8, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, nb * 8, cudaMemcpyHostToDevice); cudaMemcpy(dC, hC, nc * 8, cudaMemcpyHostToDevice);
    run<K><<<64, 32>>>(dA, dB, dC, dD, tiles);
    cudaMemcpy(hD, dD, nc * 8, cudaMemcpyDeviceToHost);
    long asc = 0, desc = 0, grp4 = 0, total = 0;
    for (int tile = 0; tile < tiles; ++tile)
        for (int r = 0; r < 16; ++r)
            for (int c = 0; c < 8; ++c) {
                const double *a = hA + (size_t)tile * 16 * K + r * K, *b = hB + (size_t)tile * K * 8 + c;
                double x = hC[(size_t)tile * 128 + r * 8 + c], y = x, z = x;
                for (int k = 0; k < K; ++k) x = fma(a[k], b[k * 8], x);
                for (int k = K - 1; k >= 0; --k) y = fma(a[k], b[k * 8], y);
                for (int k0 = 0; k0 < K; k0 += 4) {  // groups of four summed exactly-in-order inside, groups ascending: same as asc
                    for (int k = k0; k < k0 + 4 && k < K; ++k) z = fma(a[k], b[k * 8], z);
                }
                const double got = hD[(size_t)tile * 128 + r * 8 + c];
                asc += got == x; desc += got == y; grp4 += got == z; ++total;
            }
    printf("{\"shape\": \"m16n8k%d\", \"equals_sequential_fma_k_ascending\": %ld, \"k_descending\": %ld, \"total\": %ld}\n", K, asc, desc, total);
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dD); free(hA); free(hB); free(hC); free(hD);
}

int main() {
    test<4>();
    test<8>();
    test<16>();
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
