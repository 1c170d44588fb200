// Floor of one "launch a kernel, learn on the host that it has finished" round trip on this box: an empty 16-CTA kernel whose
// last thread raises a flag in mapped host memory (system-scope fence in front, like cta_finalize), host spinning on the flag --
// the mechanism of mle_sample_batch -- against cudaStreamSynchronize and against a 5 KB cudaMemcpyAsync + synchronise.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/launch_floor_test tools/launch_floor_test.cu
#include <chrono>
#include <cstdio>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>

__global__ void k_flag(volatile unsigned *flag, unsigned seq, unsigned *ticket) {
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) { *ticket = 0; __threadfence_system(); *flag = seq; }
    }
}
__global__ void k_empty(unsigned *ticket) { if (threadIdx.x == 0 && blockIdx.x == 0xffffffffu) *ticket = 1; }

int main() {
    unsigned *h_flag, *d_flag, *ticket; double *d_buf, *h_buf;
    cudaHostAlloc((void **)&h_flag, 64, cudaHostAllocMapped); cudaHostGetDevicePointer((void **)&d_flag, h_flag, 0);
    cudaMalloc(&ticket, 4); cudaMemset(ticket, 0, 4);
    cudaMalloc(&d_buf, 5376); cudaHostAlloc((void **)&h_buf, 5376, cudaHostAllocDefault);
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    auto med = [](std::vector<double> &v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; };
    const int reps = 2000;
    std::vector<double> a, b, c;
    unsigned seq = 0;
    for (int i = 0; i < reps + 100; ++i) {
        auto t0 = std::chrono::steady_clock::now();
        ++seq;
        k_flag<<<16, 256, 0, s>>>(d_flag, seq, ticket);
        while (*(volatile unsigned *)h_flag != seq) { }
        auto t1 = std::chrono::steady_clock::now();
        if (i >= 100) a.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
    }
    for (int i = 0; i < reps + 100; ++i) {
        auto t0 = std::chrono::steady_clock::now();
        k_empty<<<16, 256, 0, s>>>(ticket);
        cudaStreamSynchronize(s);
        auto t1 = std::chrono::steady_clock::now();
        if (i >= 100) b.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
    }
    for (int i = 0; i < reps + 100; ++i) {
        auto t0 = std::chrono::steady_clock::now();
        k_empty<<<16, 256, 0, s>>>(ticket);
        cudaMemcpyAsync(h_buf, d_buf, 5376, cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        auto t1 = std::chrono::steady_clock::now();
        if (i >= 100) c.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
    }
    printf("{\"empty_kernel_plus_mapped_flag_us\": %.2f, \"empty_kernel_plus_stream_sync_us\": %.2f, \"empty_kernel_plus_5KB_d2h_plus_sync_us\": %.2f}\n", med(a), med(b), med(c));
    return 0;
}
