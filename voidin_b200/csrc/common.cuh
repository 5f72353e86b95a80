// common.cuh — shared declarations of libbvh_cuda.so (sm_100a only; compiled with -fmad=false so that no
// a*b+c is ever fused: the reference is rustc/x86-64 scalar f32, which never contracts).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>

#include "../../include/bvh_cuda.h"

#define FULL_MASK 0xFFFFFFFFu

// device-side error bits (OR-ed into one word, read back at the end of a build)
#define DERR_BAD_INDEX 1u
#define DERR_DEGENERATE 2u
#define DERR_QUEUE 4u   // task queue overflow or spin limit hit
#define DERR_DEPTH 8u

struct bvh_cuda_ctx {
    int device = 0;
    int sm_count = 0;
    std::string err;
    uint64_t launches = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t own_stream2 = nullptr;  // second compute stream of the host-pointer trace pipeline (chunks alternate)
    // host-pointer trace calls: uploads and read-backs run on their own streams, pipelined against the kernels
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t pipe_ev[2][16] = {};
    // growable device workspace (one allocation, carved per call)
    void* ws = nullptr;
    size_t ws_bytes = 0;
    // pinned host scratch for small read-backs
    uint32_t* h_pin = nullptr;
    // last BLAS build
    BvhCudaBuildStats stats{};
    uint32_t* d_last_order = nullptr;
    size_t last_n = 0;
    uint32_t epoch = 0;
    int t2_blocks_per_sm = 0;
    int t2w_blocks_per_sm = 0;
    int t1_blocks_per_sm = 0;
    int tc_cluster_size = 0;  // CTAs per cluster of the cluster tier (0: tier unavailable on this device)
    int tc_clusters = 0;      // co-resident clusters of that size
    // host-API staging arena (grow-only, so repeated host calls do not cudaMalloc)
    void* stage = nullptr;
    size_t stage_bytes = 0;
    void* trace_counter = nullptr;  // persistent-warp ray counter of trace_blas
    uint32_t* defer_list = nullptr;  // rays the order-free any-hit kernel hands to the exact kernel
    size_t defer_cap = 0;
    uint32_t* defer_list2 = nullptr;  // the same for launches in control slot 1 (two traces of one scene in flight)
    size_t defer_cap2 = 0;
    // optional per-phase timing
    bool profiling = false;
    cudaEvent_t ev[10] = {};
    bool t4_ready = false;  // k_t4's dynamic shared-memory limit has been raised on this context's device
    // a stream-ordered build that has been enqueued but not yet collected (bvh_cuda_blas_build_batch_async_dev / _finish)
    bool pending = false;
    uint32_t pend_n = 0, pend_nm = 0, pend_launches = 0;
    bool pend_prof = false;
    cudaStream_t pend_stream = nullptr;
    uint32_t* pend_ids = nullptr;
};

struct bvh_cuda_scene {
    BvhCudaSceneDesc d{};  // device pointers
    bool owned = false;
    void* block = nullptr;  // single allocation when owned
    void* baked = nullptr;  // 3 x float4 per pooled triangle (always owned)
    void* counter = nullptr;  // persistent-warp ray counter (tail of `baked`)
    void* wbox = nullptr;     // optional: 2 x float4 per instance, tight world box of the instance's BLAS root (owned)
    bool wbox_on = false;     // the exact-order kernels drop instance visits whose world box the ray misses
};

// makes the context's device current for the duration of an entry point
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int ctx_fail(bvh_cuda_ctx* ctx, int code, const char* what);
int ctx_cuda_fail(bvh_cuda_ctx* ctx, cudaError_t e, const char* where);
int ctx_reserve(bvh_cuda_ctx* ctx, size_t bytes);
int ctx_stage_reserve(bvh_cuda_ctx* ctx, size_t bytes);

#define CU_CHECK(ctx, call)                                          \
    do {                                                             \
        cudaError_t _e = (call);                                     \
        if (_e != cudaSuccess) return ctx_cuda_fail(ctx, _e, #call); \
    } while (0)

// implemented in blas_build.cu / tlas.cu / trace.cu
int blas_build_device(bvh_cuda_ctx* ctx, const float* d_vertices, size_t n_vertices, uint32_t* d_indices,
                      size_t n_tris, MeshInfo* d_mesh_info, size_t n_meshes, BvhNode* d_nodes_out, size_t nodes_cap,
                      uint32_t* n_nodes_out, cudaStream_t stream, uint32_t* d_result = nullptr, bool async = false);
int blas_build_finish(bvh_cuda_ctx* ctx, uint32_t* n_nodes_out);
int tlas_build_device(bvh_cuda_ctx* ctx, const Instance* d_instances, size_t n_inst, const MeshInfo* d_meshes,
                      size_t n_mesh, TlasNode* d_nodes_out, uint32_t* d_children_out, cudaStream_t stream);
int scene_bake_device(bvh_cuda_ctx* ctx, bvh_cuda_scene* scene, cudaStream_t stream);
int scene_instance_boxes_device(bvh_cuda_ctx* ctx, bvh_cuda_scene* scene, int enable, cudaStream_t stream);
int trace_blas_device(bvh_cuda_ctx* ctx, const BvhNode* d_nodes, const float* d_vertices,
                      const uint32_t* d_indices, const float* d_ray_o, const float* d_ray_d, size_t n_rays,
                      float* d_t, uint32_t* d_tri, cudaStream_t stream);
int trace_blas_rec_device(bvh_cuda_ctx* ctx, const BvhNode* d_nodes, const float* d_vertices, const uint32_t* d_indices,
                          const float* d_ray_o, const float* d_ray_d, size_t n_rays, uint32_t node_idx, float t0, float* d_t,
                          uint8_t* d_hit, cudaStream_t stream);
int trace_scene_device(bvh_cuda_ctx* ctx, const bvh_cuda_scene* scene, const float* d_ray_o,
                       const float* d_ray_d, size_t n_rays, float tmax, int any_hit, float* d_t, uint32_t* d_tri,
                       uint32_t* d_inst, uint8_t* d_occ, cudaStream_t stream, int slot = 0);

#ifdef __CUDACC__
// Order-preserving float <-> uint map (so min/max can use integer redux / atomics).  -0.0 sorts below +0.0.
__device__ __forceinline__ uint32_t f2o(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float o2f(uint32_t o) {
    uint32_t b = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
    return __uint_as_float(b);
}
#define ENC_POS_INIT 0xF149F2CAu /* f2o(1e30f):  0x7149F2CA | 0x80000000 */
#define ENC_NEG_INIT 0x0EB60D35u /* f2o(-1e30f): ~0xF149F2CA */

// Aabb::area (crates/bvh/src/intersection.rs:16-19): (dx*dy + dx*dz + dy*dz) * 2, left to right, unfused.
__device__ __forceinline__ float aabb_area(float lx, float ly, float lz, float hx, float hy, float hz) {
    float dx = __fsub_rn(hx, lx), dy = __fsub_rn(hy, ly), dz = __fsub_rn(hz, lz);
    float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dy), __fmul_rn(dx, dz)), __fmul_rn(dy, dz));
    return __fmul_rn(s, 2.0f);
}
// Vec3::lerp component (glam 0.24): a + (b - a) * s
__device__ __forceinline__ float lerp1(float a, float b, float s) {
    return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), s));
}
#endif
