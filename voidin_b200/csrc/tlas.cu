// tlas.cu — Tlas::build (crates/bvh/src/tlas.rs:31-85), exact.
//   k_tlas_leaves   leaf boxes at slots 1..=I: fold over the 8 transformed corners seeded with the
//                   UNTRANSFORMED local mesh box (tlas.rs:34-54, the seed at :39 is a reference quirk).
//   k_tlas_chain    the best-match chain (tlas.rs:56-84) with find_best_match (tlas.rs:87-105) as a
//                   block-wide arg-min: strict <, first index wins, threshold 1e30, target skipped.
//                   The chain is inherently sequential (~3.65*I dependent arg-mins); one persistent block
//                   keeps the loop state in registers and the live slot boxes in a slot-indexed SoA mirror.
//                   Reproduced quirks: loop runs until count == 0 so the last cluster merges with itself
//                   and exactly 2I+1 nodes exist (tlas.rs:61); `a` may be a stale slot >= count (tlas.rs:94);
//                   left_right = a + (b << 16) in wrapping u32 (tlas.rs:71).
#include "common.cuh"

namespace {

constexpr int CHAIN_THREADS = 1024;

__device__ __forceinline__ float3 xform_point(const float* m, float x, float y, float z) {
    // glam Mat4::transform_point3: ((X*x + Y*y) + Z*z) + W
    float3 r;
    r.x = ((m[0] * x + m[4] * y) + m[8] * z) + m[12];
    r.y = ((m[1] * x + m[5] * y) + m[9] * z) + m[13];
    r.z = ((m[2] * x + m[6] * y) + m[10] * z) + m[14];
    return r;
}

__global__ void __launch_bounds__(256) k_tlas_leaves(const Instance* __restrict__ instances, uint32_t n_inst,
                                                     const MeshInfo* __restrict__ meshes, uint32_t n_mesh,
                                                     TlasNode* nodes, uint32_t* children, float* slot_box,
                                                     uint32_t* node_indices, uint32_t* err) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_inst) return;
    const Instance* in = instances + i;
    uint32_t mi = in->mesh;
    if (mi >= n_mesh) { atomicOr(err, DERR_BAD_INDEX); mi = 0; }
    const MeshInfo* mesh = meshes + mi;
    const float bx[2] = {mesh->min[0], mesh->max[0]}, by[2] = {mesh->min[1], mesh->max[1]}, bz[2] = {mesh->min[2], mesh->max[2]};
    float mn[3] = {bx[0], by[0], bz[0]}, mx[3] = {bx[1], by[1], bz[1]};
    for (int c = 0; c < 8; ++c) {
        const int ix = (c & 1) == 0, iy = (c & 2) == 0, iz = (c & 4) == 0;
        const float3 q = xform_point(in->transform, bx[ix], by[iy], bz[iz]);
        mn[0] = fminf(mn[0], q.x); mn[1] = fminf(mn[1], q.y); mn[2] = fminf(mn[2], q.z);
        mx[0] = fmaxf(mx[0], q.x); mx[1] = fmaxf(mx[1], q.y); mx[2] = fmaxf(mx[2], q.z);
    }
    TlasNode nd;
    nd.min[0] = mn[0]; nd.min[1] = mn[1]; nd.min[2] = mn[2];
    nd.max[0] = mx[0]; nd.max[1] = mx[1]; nd.max[2] = mx[2];
    nd.left_right = 0;
    nd.instance_idx = i;
    nodes[i + 1] = nd;
    if (children) { children[2 * (size_t)(i + 1)] = 0; children[2 * (size_t)(i + 1) + 1] = 0; }
    for (int k = 0; k < 3; ++k) {
        slot_box[(size_t)k * n_inst + i] = mn[k];
        slot_box[(size_t)(3 + k) * n_inst + i] = mx[k];
    }
    node_indices[i] = i + 1;
}

// Block-wide find_best_match (tlas.rs:87-105).  All threads return the same slot.
__device__ __forceinline__ uint32_t find_best_match(const float* slot_box, uint32_t n_inst, uint32_t count,
                                                    uint32_t target, unsigned long long* s_red) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float t0 = slot_box[target], t1 = slot_box[(size_t)n_inst + target], t2 = slot_box[2 * (size_t)n_inst + target];
    const float t3 = slot_box[3 * (size_t)n_inst + target], t4 = slot_box[4 * (size_t)n_inst + target],
                t5 = slot_box[5 * (size_t)n_inst + target];
    unsigned long long best = 0xFFFFFFFFFFFFFFFFull;
    for (uint32_t i = tid; i < count; i += CHAIN_THREADS) {
        if (i == target) continue;
        const float lx = fminf(t0, slot_box[i]), ly = fminf(t1, slot_box[(size_t)n_inst + i]),
                    lz = fminf(t2, slot_box[2 * (size_t)n_inst + i]);
        const float hx = fmaxf(t3, slot_box[3 * (size_t)n_inst + i]), hy = fmaxf(t4, slot_box[4 * (size_t)n_inst + i]),
                    hz = fmaxf(t5, slot_box[5 * (size_t)n_inst + i]);
        const float sa = aabb_area(lx, ly, lz, hx, hy, hz);
        if (sa < 1e30f) {  // only areas below the initial `smallest` can ever be selected; NaN never is
            // sa >= 0 here, so its bit pattern orders like the float; ties resolve to the lower index
            const unsigned long long key = ((unsigned long long)__float_as_uint(sa) << 32) | i;
            best = key < best ? key : best;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long y = __shfl_xor_sync(FULL_MASK, best, o);
        best = y < best ? y : best;
    }
    if (lane == 0) s_red[warp] = best;
    __syncthreads();
    unsigned long long v = s_red[lane];  // CHAIN_THREADS / 32 == 32 warps
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long y = __shfl_xor_sync(FULL_MASK, v, o);
        v = y < v ? y : v;
    }
    __syncthreads();
    return (v == 0xFFFFFFFFFFFFFFFFull) ? target : (uint32_t)(v & 0xFFFFFFFFull);
}

__global__ void __launch_bounds__(CHAIN_THREADS) k_tlas_chain(uint32_t n_inst, TlasNode* nodes, uint32_t* children,
                                                              float* slot_box, uint32_t* node_indices) {
    __shared__ unsigned long long s_red[32];
    const uint32_t tid = threadIdx.x;
    uint32_t count = n_inst;
    uint32_t used = 1 + n_inst;
    uint32_t a = 0;
    uint32_t b = find_best_match(slot_box, n_inst, count, a, s_red);
    while (count > 0) {
        const uint32_t c = find_best_match(slot_box, n_inst, count, b, s_red);
        if (a == c) {
            if (tid == 0) {
                const uint32_t idx_a = node_indices[a], idx_b = node_indices[b];
                float u[6];
                for (int k = 0; k < 3; ++k) {
                    u[k] = fminf(slot_box[(size_t)k * n_inst + a], slot_box[(size_t)k * n_inst + b]);
                    u[3 + k] = fmaxf(slot_box[(size_t)(3 + k) * n_inst + a], slot_box[(size_t)(3 + k) * n_inst + b]);
                }
                TlasNode nd;
                nd.min[0] = u[0]; nd.min[1] = u[1]; nd.min[2] = u[2];
                nd.max[0] = u[3]; nd.max[1] = u[4]; nd.max[2] = u[5];
                nd.left_right = idx_a + (idx_b << 16);
                nd.instance_idx = 0xFFFFFFFFu;
                nodes[used] = nd;
                if (children) { children[2 * (size_t)used] = idx_a; children[2 * (size_t)used + 1] = idx_b; }
                for (int k = 0; k < 6; ++k) slot_box[(size_t)k * n_inst + a] = u[k];
                node_indices[a] = used;
                node_indices[b] = node_indices[count - 1];
                for (int k = 0; k < 6; ++k) slot_box[(size_t)k * n_inst + b] = slot_box[(size_t)k * n_inst + (count - 1)];
            }
            used += 1;
            count -= 1;
            __syncthreads();
            b = find_best_match(slot_box, n_inst, count, a, s_red);
        } else {
            a = b;
            b = c;
        }
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t ia = node_indices[a];
        nodes[0] = nodes[ia];
        if (children) { children[0] = children[2 * (size_t)ia]; children[1] = children[2 * (size_t)ia + 1]; }
    }
}

}  // namespace

int tlas_build_device(bvh_cuda_ctx* ctx, const Instance* d_instances, size_t n_inst, const MeshInfo* d_meshes,
                      size_t n_mesh, TlasNode* d_nodes_out, uint32_t* d_children_out, cudaStream_t stream) {
    if (!d_instances || !d_meshes || !d_nodes_out || n_inst == 0 || n_mesh == 0)
        return ctx_fail(ctx, BVH_CUDA_EINVAL, "tlas_build: no instances / null pointer (the reference guards this in MeshPool::generate_tlas)");
    if (n_inst > 0x3FFFFFFFull) return ctx_fail(ctx, BVH_CUDA_EINVAL, "tlas_build: too many instances");
    const uint32_t I = (uint32_t)n_inst;
    const size_t need = 256 + sizeof(float) * 6 * (size_t)I + 256 + sizeof(uint32_t) * (size_t)I + 256;
    int rc = ctx_reserve(ctx, need);
    if (rc) return rc;
    char* base = (char*)ctx->ws;
    uint32_t* err = (uint32_t*)base;
    float* slot_box = (float*)(base + 256);
    uint32_t* node_indices = (uint32_t*)(base + 256 + ((sizeof(float) * 6 * (size_t)I + 255) & ~(size_t)255));
    CU_CHECK(ctx, cudaMemsetAsync(err, 0, 256, stream));
    CU_CHECK(ctx, cudaMemsetAsync(d_nodes_out, 0, sizeof(TlasNode), stream));
    k_tlas_leaves<<<(I + 255) / 256, 256, 0, stream>>>(d_instances, I, d_meshes, (uint32_t)n_mesh, d_nodes_out, d_children_out,
                                                      slot_box, node_indices, err);
    k_tlas_chain<<<1, CHAIN_THREADS, 0, stream>>>(I, d_nodes_out, d_children_out, slot_box, node_indices);
    ctx->launches += 2;
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_pin, err, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CU_CHECK(ctx, cudaStreamSynchronize(stream));
    CU_CHECK(ctx, cudaGetLastError());
    if (ctx->h_pin[0] & DERR_BAD_INDEX) return ctx_fail(ctx, BVH_CUDA_EINVAL, "tlas_build: instance.mesh out of range");
    return BVH_CUDA_OK;
}
