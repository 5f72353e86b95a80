// l2_bw.cu — measured L2 -> SM read bandwidth on this GPU: the denominator of the traversal kernels' L2 roofline
// (bench.py `rays.roofline.l2_model`).  Every SM streams a buffer that fits the L2 (default 32 MiB, well below the
// 126 MB of a B200) many times with 16-byte loads; the first pass (HBM -> L2) is excluded by a warm-up launch.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2_bw l2_bw.cu && ./l2_bw [MiB] > l2_peak.json
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024) k_read(const uint4* __restrict__ buf, size_t n_vec, int passes, unsigned* sink) {
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; ++p)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
            uint4 v;
            asm volatile("ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + i));
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x12345678u) *sink = acc;
}

int main(int argc, char** argv) {
    const size_t mib = argc > 1 ? (size_t)atoi(argv[1]) : 32;
    const size_t bytes = mib << 20, n_vec = bytes / 16;
    uint4* buf; unsigned* sink;
    cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, bytes);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int blocks = prop.multiProcessorCount * 2, passes = 50;
    k_read<<<blocks, 1024>>>(buf, n_vec, 2, sink);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k_read<<<blocks, 1024>>>(buf, n_vec, passes, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double gbs = (double)bytes * passes / (best * 1e-3) / 1e9;
    printf("{\"l2_read_GBps\": %.1f, \"buffer_MiB\": %zu, \"passes\": %d, \"ms\": %.4f, \"gpu\": \"%s\", \"how\": \"scripts/micro/l2_bw.cu: %d blocks x 1024 threads, 16-byte ld.global.ca over an L2-resident buffer, best of 5\"}\n",
           gbs, mib, passes, best, prop.name, blocks);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
