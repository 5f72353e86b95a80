// dsmem_bench.cu — micro-benchmark behind the cluster-tier design decision:
// how fast can a 16-CTA cluster scatter / gather 4-byte items through distributed shared memory, compared with the
// same pattern through global memory (L2-resident), and what does cluster.sync cost next to it?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_bench dsmem_bench.cu && ./dsmem_bench
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

constexpr int CAP = 16384;   // slots per CTA
constexpr int TH = 1024;
constexpr int REPS = 20;

__device__ __forceinline__ uint32_t dest_of(int pattern, uint32_t j, uint32_t n) {
    switch (pattern) {
        // n is a power of two: no integer division anywhere in the timed loop
        case 0: return (j + 1) & (n - 1);                    // shift by one (back R's)
        case 1: return (j & 1) ? (n - 1 - j / 2) : j / 2;    // compaction: evens to the front, odds reversed to the back
        case 2: return (j * 2654435761u) & (n - 1);          // pseudo-random bijection
        default: return j;
    }
}

// mode 0: DSMEM scatter store, 1: DSMEM gather load, 2: global scatter store, 3: global gather load, 4: only cluster.sync
__global__ void __launch_bounds__(TH, 1) k_bench(int mode, int pattern, uint32_t slots_per_cta, uint32_t* gbuf, unsigned long long* out) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ uint32_t s_arr[];
    const uint32_t C = cluster.num_blocks(), rank = cluster.block_rank();
    const uint32_t n = C * slots_per_cta;
    const uint32_t sh = 31 - __clz(slots_per_cta);
    for (uint32_t i = threadIdx.x; i < CAP; i += TH) s_arr[i] = i;
    cluster.sync();
    unsigned long long t0 = 0, acc_t = 0;
    uint32_t sink = 0;
    for (int rep = 0; rep < REPS; ++rep) {
        cluster.sync();
        if (threadIdx.x == 0) t0 = clock64();
        if (mode != 4) {
            for (uint32_t i = threadIdx.x; i < slots_per_cta; i += TH) {
                const uint32_t j = rank * slots_per_cta + i;
                const uint32_t d = dest_of(pattern, j, n);
                const uint32_t owner = d >> sh, off = d & (slots_per_cta - 1);
                if (mode == 0) { uint32_t* r = cluster.map_shared_rank(s_arr, owner); r[off] = j + rep; }
                else if (mode == 1) { const uint32_t* r = cluster.map_shared_rank(s_arr, owner); sink += r[off]; }
                else if (mode == 2) { gbuf[d] = j + rep; }
                else { sink += __ldcg(&gbuf[d]); }
            }
        }
        cluster.sync();
        if (threadIdx.x == 0) acc_t += clock64() - t0;
    }
    if (sink == 0xDEADBEEF) gbuf[0] = sink;
    if (threadIdx.x == 0 && rank == 0) out[blockIdx.x / C] = acc_t / REPS;
}

int main() {
    uint32_t* gbuf; unsigned long long* out;
    cudaMalloc(&gbuf, 16 * CAP * 4 * 16); cudaMemset(gbuf, 0, 16 * CAP * 4 * 16);
    cudaMallocManaged(&out, 64 * 8);
    cudaFuncSetAttribute(k_bench, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(k_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, CAP * 4);
    const char* mname[] = {"dsmem_store", "dsmem_load", "global_store", "global_load", "sync_only"};
    const char* pname[] = {"shift1", "compact", "random"};
    for (int cs : {2, 4, 8, 16}) {
        for (uint32_t slots : {2048u, 4096u, 16384u}) {
            for (int mode = 0; mode < 5; ++mode) {
                for (int pattern = 0; pattern < 3; ++pattern) {
                    if (mode == 4 && pattern) continue;
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(cs); cfg.blockDim = dim3(TH); cfg.dynamicSmemBytes = CAP * 4;
                    cudaLaunchAttribute attr[1];
                    attr[0].id = cudaLaunchAttributeClusterDimension;
                    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                    cfg.attrs = attr; cfg.numAttrs = 1;
                    out[0] = 0;
                    cudaError_t e = cudaLaunchKernelEx(&cfg, k_bench, mode, pattern, slots, gbuf, out);
                    if (e == cudaSuccess) e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("cs=%d launch failed: %s\n", cs, cudaGetErrorString(e)); cudaGetLastError(); continue; }
                    printf("cs=%2d slots/cta=%5u %-12s %-8s %8llu cycles  (%.2f B/cyc/SM)\n", cs, slots, mname[mode], mode == 4 ? "-" : pname[pattern],
                           out[0], mode == 4 ? 0.0 : 4.0 * slots / (double)out[0]);
                }
            }
        }
    }
    // how many 16-CTA clusters of 1024 threads / 64 KB can be co-resident?
    for (int cs : {16, 8, 4}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(TH); cfg.dynamicSmemBytes = CAP * 4;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int nc = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, k_bench, &cfg);
        printf("max active clusters cs=%d: %d (%s)\n", cs, nc, cudaGetErrorString(e));
    }
    return 0;
}
