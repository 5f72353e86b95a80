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

#include <cooperative_groups.h>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace {

// f32::min / f32::max as the reference's folds use them: the accumulator stays unless the new value is strictly
// smaller / larger, so among zeros of different sign the first one is kept (tlas.rs:43,69-70).
__device__ __forceinline__ float min_rs(float acc, float v) { return (v < acc) ? v : acc; }
__device__ __forceinline__ float max_rs(float acc, float v) { return (v > acc) ? v : acc; }

constexpr int CHAIN_THREADS = 1024;

// warp-wide minimum of 64-bit keys (area bits << 32 | slot) with two redux.sync instead of five 64-bit shuffle rounds
__device__ __forceinline__ unsigned long long warp_min_key(unsigned long long key) {
    const uint32_t hi = (uint32_t)(key >> 32), lo = (uint32_t)key;
    const uint32_t mh = __reduce_min_sync(FULL_MASK, hi);
    const uint32_t ml = __reduce_min_sync(FULL_MASK, hi == mh ? lo : 0xFFFFFFFFu);
    return ((unsigned long long)mh << 32) | ml;
}

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
        mn[0] = min_rs(mn[0], q.x); mn[1] = min_rs(mn[1], q.y); mn[2] = min_rs(mn[2], q.z);
        mx[0] = max_rs(mx[0], q.x); mx[1] = max_rs(mx[1], q.y); mx[2] = max_rs(mx[2], q.z);
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
                                                    uint32_t target, unsigned long long* s_red, uint32_t& par) {
    par += 1;
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
    best = warp_min_key(best);
    // s_red is double buffered by call parity: one block barrier per call is enough (a warp can be at most one call ahead)
    unsigned long long* red = s_red + 32 * (par & 1u);
    if (lane == 0) red[warp] = best;
    __syncthreads();
    const unsigned long long v = warp_min_key(red[lane]);  // CHAIN_THREADS / 32 == 32 warps
    return (v == 0xFFFFFFFFFFFFFFFFull) ? target : (uint32_t)(v & 0xFFFFFFFFull);
}

__global__ void __launch_bounds__(CHAIN_THREADS) k_tlas_chain(uint32_t n_inst, TlasNode* nodes, uint32_t* children,
                                                              float* slot_box, uint32_t* node_indices) {
    __shared__ unsigned long long s_red[64];
    uint32_t par = 0;
    const uint32_t tid = threadIdx.x;
    uint32_t count = n_inst;
    uint32_t used = 1 + n_inst;
    uint32_t a = 0;
    uint32_t b = find_best_match(slot_box, n_inst, count, a, s_red, par);
    while (count > 0) {
        const uint32_t c = find_best_match(slot_box, n_inst, count, b, s_red, par);
        if (a == c) {
            if (tid == 0) {
                const uint32_t idx_a = node_indices[a], idx_b = node_indices[b];
                float u[6];
                for (int k = 0; k < 3; ++k) {
                    u[k] = min_rs(slot_box[(size_t)k * n_inst + a], slot_box[(size_t)k * n_inst + b]);
                    u[3 + k] = max_rs(slot_box[(size_t)(3 + k) * n_inst + a], slot_box[(size_t)(3 + k) * n_inst + b]);
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
            b = find_best_match(slot_box, n_inst, count, a, s_red, par);
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


// ---------------------------------------------------------------------------------------------------------
// Cluster version of the chain for 12 K < I <= ~117 K instances: the live slot boxes are spread over the shared
// memory of a thread-block cluster (16 CTAs on B200, non-portable size), every find_best_match is one local scan,
// one block reduction and ONE cluster barrier (candidates are exchanged through distributed shared memory, double
// buffered), instead of a 1024-thread block streaming all slots from L2.  The walk state (a, b, their boxes and
// node ids, count, nodes_used) is replicated in every thread, so no CTA ever waits to learn what to do next.
// Same sequence of operations as tlas.rs:56-84, same tie rules.
// ---------------------------------------------------------------------------------------------------------
#ifndef CL_THREADS_V
#define CL_THREADS_V 512
#endif
constexpr int CL_THREADS = CL_THREADS_V;
constexpr int CL_MAX = 16;

struct Cand {
    unsigned long long key;  // (area bits << 32) | slot, or ~0 when the CTA has no candidate
    float box[6];
    uint32_t ni;
    uint32_t pad;
};

__global__ void __launch_bounds__(CL_THREADS) k_tlas_chain_cluster(uint32_t n_inst, uint32_t cap, TlasNode* nodes, uint32_t* children,
                                                                  const float* __restrict__ slot_box_init) {
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t C = cluster.num_blocks(), rank = cluster.block_rank();
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    extern __shared__ float s_dyn[];
    float* bx = s_dyn;                                              // [6][cap] boxes of my slots
    uint32_t* ni = reinterpret_cast<uint32_t*>(s_dyn + 6 * (size_t)cap);  // [cap] node index of my slots
    __shared__ Cand s_cl[2][CL_MAX];
    __shared__ unsigned long long s_red[2][CL_THREADS / 32];
    const uint32_t base = rank * cap;

    for (uint32_t j = tid; j < cap; j += CL_THREADS) {
        const uint32_t g = base + j;
        if (g < n_inst) {
#pragma unroll
            for (int k = 0; k < 6; ++k) bx[(size_t)k * cap + j] = slot_box_init[(size_t)k * n_inst + g];
            ni[j] = g + 1;
        }
    }
    cluster.sync();

    // read one slot (box + node id) from whichever CTA owns it
    auto read_slot = [&](uint32_t slot, float* box, uint32_t& node_id) {
        const uint32_t owner = slot / cap, j = slot % cap;
        const float* rb = cluster.map_shared_rank(bx, owner);
        const uint32_t* rn = cluster.map_shared_rank(ni, owner);
#pragma unroll
        for (int k = 0; k < 6; ++k) box[k] = rb[(size_t)k * cap + j];
        node_id = rn[j];
    };

    uint32_t par = 0;
    // find_best_match (tlas.rs:87-105) for a target whose box/node id every thread already knows
    auto fbm = [&](uint32_t count, uint32_t target, const float* tb, uint32_t t_ni, uint32_t& best, float* bbox, uint32_t& b_ni) {
        unsigned long long key = 0xFFFFFFFFFFFFFFFFull;
        for (uint32_t j = tid; j < cap; j += CL_THREADS) {
            const uint32_t g = base + j;
            if (g >= count) break;
            if (g == target) continue;
            const float lx = fminf(tb[0], bx[j]), ly = fminf(tb[1], bx[(size_t)cap + j]), lz = fminf(tb[2], bx[2 * (size_t)cap + j]);
            const float hx = fmaxf(tb[3], bx[3 * (size_t)cap + j]), hy = fmaxf(tb[4], bx[4 * (size_t)cap + j]),
                        hz = fmaxf(tb[5], bx[5 * (size_t)cap + j]);
            const float sa = aabb_area(lx, ly, lz, hx, hy, hz);
            if (sa < 1e30f) {
                const unsigned long long k2 = ((unsigned long long)__float_as_uint(sa) << 32) | g;
                key = k2 < key ? k2 : key;
            }
        }
        key = warp_min_key(key);
        // one block barrier per call: s_red is double buffered by call parity (as is the exchange buffer s_cl)
        if (lane == 0) s_red[par][warp] = key;
        __syncthreads();
        const unsigned long long bkey = warp_min_key((lane < CL_THREADS / 32) ? s_red[par][lane] : 0xFFFFFFFFFFFFFFFFull);
        // publish my CTA's candidate into every CTA's exchange buffer (thread t writes to CTA t)
        if (tid < C) {
            Cand c;
            c.key = bkey;
            c.pad = 0;
            if (bkey != 0xFFFFFFFFFFFFFFFFull) {
                const uint32_t j = (uint32_t)(bkey & 0xFFFFFFFFull) - base;
#pragma unroll
                for (int k = 0; k < 6; ++k) c.box[k] = bx[(size_t)k * cap + j];
                c.ni = ni[j];
            } else {
#pragma unroll
                for (int k = 0; k < 6; ++k) c.box[k] = 0.0f;
                c.ni = 0;
            }
            Cand* dst = cluster.map_shared_rank(&s_cl[par][rank], tid);
            *dst = c;
        }
        cluster.sync();
        unsigned long long gk = 0xFFFFFFFFFFFFFFFFull;
        uint32_t gw = 0;
        for (uint32_t r2 = 0; r2 < C; ++r2) {
            const unsigned long long k2 = s_cl[par][r2].key;
            if (k2 < gk) { gk = k2; gw = r2; }
        }
        if (gk == 0xFFFFFFFFFFFFFFFFull) {
            best = target;
#pragma unroll
            for (int k = 0; k < 6; ++k) bbox[k] = tb[k];
            b_ni = t_ni;
        } else {
            best = (uint32_t)(gk & 0xFFFFFFFFull);
#pragma unroll
            for (int k = 0; k < 6; ++k) bbox[k] = s_cl[par][gw].box[k];
            b_ni = s_cl[par][gw].ni;
        }
        par ^= 1;
    };

    uint32_t count = n_inst, used = 1 + n_inst;
    uint32_t a = 0, b = 0, a_ni = 0, b_ni = 0;
    float a_box[6], b_box[6];
    read_slot(0, a_box, a_ni);
    fbm(count, a, a_box, a_ni, b, b_box, b_ni);
    while (count > 0) {
        uint32_t c = 0, c_ni = 0;
        float c_box[6];
        fbm(count, b, b_box, b_ni, c, c_box, c_ni);
        if (a == c) {
            float u[6];
#pragma unroll
            for (int k = 0; k < 3; ++k) { u[k] = min_rs(a_box[k], b_box[k]); u[3 + k] = max_rs(a_box[3 + k], b_box[3 + k]); }
            const uint32_t last = count - 1;
            if (rank == 0 && tid == 0) {
                TlasNode nd;
                nd.min[0] = u[0]; nd.min[1] = u[1]; nd.min[2] = u[2];
                nd.max[0] = u[3]; nd.max[1] = u[4]; nd.max[2] = u[5];
                nd.left_right = a_ni + (b_ni << 16);
                nd.instance_idx = 0xFFFFFFFFu;
                nodes[used] = nd;
                if (children) { children[2 * (size_t)used] = a_ni; children[2 * (size_t)used + 1] = b_ni; }
            }
            // node_indices[a] = used; node_indices[b] = node_indices[last] (tlas.rs:74-76), in that order
            float lb[6];
            uint32_t l_ni = used;
#pragma unroll
            for (int k = 0; k < 6; ++k) lb[k] = u[k];
            if (last != a && tid == 0 && b / cap == rank) read_slot(last, lb, l_ni);
            if (tid == 0 && a / cap == rank) {
                const uint32_t j = a % cap;
#pragma unroll
                for (int k = 0; k < 6; ++k) bx[(size_t)k * cap + j] = u[k];
                ni[j] = used;
            }
            if (tid == 0 && b / cap == rank) {
                const uint32_t j = b % cap;
#pragma unroll
                for (int k = 0; k < 6; ++k) bx[(size_t)k * cap + j] = lb[k];
                ni[j] = l_ni;
            }
            // A block barrier is enough here: every CTA scans only its own slots, the only remote access of the update is
            // the read of slot `last`, which nobody writes in this step, and the next remote access to the two rewritten
            // slots comes after the cluster barrier inside the find_best_match below.
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 6; ++k) a_box[k] = u[k];
            a_ni = used;
            if (a == b) read_slot(a, a_box, a_ni);  // slot a was overwritten by the swap-remove
            used += 1;
            count -= 1;
            fbm(count, a, a_box, a_ni, b, b_box, b_ni);
        } else {
            a = b; a_ni = b_ni;
            b = c; b_ni = c_ni;
#pragma unroll
            for (int k = 0; k < 6; ++k) { a_box[k] = b_box[k]; b_box[k] = c_box[k]; }
        }
    }
    cluster.sync();
    if (rank == 0 && tid == 0) {
        nodes[0] = nodes[a_ni];
        if (children) { children[0] = children[2 * (size_t)a_ni]; children[1] = children[2 * (size_t)a_ni + 1]; }
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
    // the workspace is shared with the BLAS builder: its "last order" snapshot is overwritten from here on
    ctx->d_last_order = nullptr;
    ctx->last_n = 0;
    char* base = (char*)ctx->ws;
    uint32_t* err = (uint32_t*)base;
    float* slot_box = (float*)(base + 256);
    uint32_t* node_indices = (uint32_t*)(base + 256 + ((sizeof(float) * 6 * (size_t)I + 255) & ~(size_t)255));
    CU_CHECK(ctx, cudaMemsetAsync(err, 0, 256, stream));
    CU_CHECK(ctx, cudaMemsetAsync(d_nodes_out, 0, sizeof(TlasNode), stream));
    k_tlas_leaves<<<(I + 255) / 256, 256, 0, stream>>>(d_instances, I, d_meshes, (uint32_t)n_mesh, d_nodes_out, d_children_out,
                                                      slot_box, node_indices, err);
    bool clustered = false;
    static const bool no_cluster = [] { const char* e = getenv("BVH_CUDA_TLAS"); return e && e[0] == 'b'; }();  // "block": force the one-block kernel
    if (I > 12288 && !no_cluster) {  // measured crossover: 4 096 -> block 16.7 ms vs cluster 25.1 ms; 32 767 -> 594 ms vs 234 ms
        static const uint32_t cl_env = [] { const char* e = getenv("BVH_CUDA_TLAS_CL"); const int v = e ? atoi(e) : 0; return (uint32_t)((v == 2 || v == 4 || v == 8 || v == 16) ? v : 0); }();
        const uint32_t csize = cl_env ? cl_env : (uint32_t)CL_MAX;  // CTAs per cluster (A/B: BVH_CUDA_TLAS_CL=8)
        const uint32_t cap = (I + csize - 1) / csize;
        const size_t smem = sizeof(float) * 6 * (size_t)cap + sizeof(uint32_t) * (size_t)cap;
        if (smem <= 200 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(k_tlas_chain_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tlas_chain_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(csize);
                cfg.blockDim = dim3(CL_THREADS);
                cfg.dynamicSmemBytes = smem;
                cfg.stream = stream;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = csize;
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                uint32_t cap_arg = cap;
                e = cudaLaunchKernelEx(&cfg, k_tlas_chain_cluster, I, cap_arg, d_nodes_out, d_children_out, (const float*)slot_box);
            }
            if (e == cudaSuccess) clustered = true;
            else cudaGetLastError();  // fall through to the one-block kernel (e.g. a 16-CTA cluster cannot be placed)
        }
    }
    if (!clustered) k_tlas_chain<<<1, CHAIN_THREADS, 0, stream>>>(I, d_nodes_out, d_children_out, slot_box, node_indices);
    ctx->launches += 2;
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_pin, err, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CU_CHECK(ctx, cudaStreamSynchronize(stream));
    CU_CHECK(ctx, cudaGetLastError());
    if (ctx->h_pin[0] & DERR_BAD_INDEX) return ctx_fail(ctx, BVH_CUDA_EINVAL, "tlas_build: instance.mesh out of range");
    return BVH_CUDA_OK;
}
