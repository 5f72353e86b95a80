// blas_emit.cuh -- part of blas_build.cu (included there, inside its anonymous namespace; not a stand-alone header):
// numbering scans, node emit, mesh tables, index permutation, roots.
#pragma once

// ------------------------------------------------------------------------------------------------
// Exclusive prefix sum over n u32 (in place), three small kernels.
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_TILE = 4096;  // 1024 threads x 4

__global__ void __launch_bounds__(1024) k_scan_reduce(const uint32_t* x, uint32_t n, uint32_t* sums) {
    __shared__ uint32_t s_w[32];
    const uint32_t base = blockIdx.x * SCAN_TILE;
    uint32_t v = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t k = base + i * 1024 + threadIdx.x;
        if (k < n) v += x[k];
    }
    v = __reduce_add_sync(FULL_MASK, v);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t t = s_w[threadIdx.x];
        t = __reduce_add_sync(FULL_MASK, t);
        if (threadIdx.x == 0) sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) k_scan_top(uint32_t* sums, uint32_t nb, uint32_t* total) {
    __shared__ uint32_t s_part[1024];
    const uint32_t tid = threadIdx.x;
    const uint32_t per = (nb + 1023) / 1024;
    const uint32_t b = min(nb, tid * per), e = min(nb, b + per);
    uint32_t sum = 0;
    for (uint32_t i = b; i < e; ++i) sum += sums[i];
    s_part[tid] = sum;
    __syncthreads();
    if (tid == 0) {
        uint32_t acc = 0;
        for (int i = 0; i < 1024; ++i) { const uint32_t v = s_part[i]; s_part[i] = acc; acc += v; }
        *total = acc;
    }
    __syncthreads();
    uint32_t acc = s_part[tid];
    for (uint32_t i = b; i < e; ++i) { const uint32_t v = sums[i]; sums[i] = acc; acc += v; }
}

__global__ void __launch_bounds__(1024) k_scan_apply(uint32_t* x, uint32_t n, const uint32_t* sums) {
    __shared__ uint32_t s_w[32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t v[4], tot = 0;
    for (int i = 0; i < 4; ++i) { v[i] = (base + i < n) ? x[base + i] : 0; tot += v[i]; }
    uint32_t inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL_MASK, inc, o);
        if ((int)lane >= o) inc += y;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t t = s_w[lane], ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL_MASK, ti, o);
            if ((int)lane >= o) ti += y;
        }
        s_w[lane] = ti - t;
    }
    __syncthreads();
    uint32_t acc = sums[blockIdx.x] + s_w[warp] + inc - tot;
    for (int i = 0; i < 4; ++i) {
        if (base + i < n) x[base + i] = acc;
        acc += v[i];
    }
}

// ------------------------------------------------------------------------------------------------
// Emit: records -> BvhNode[] in DFS pre-order pair numbering (blas.rs:90,110-112,125-126).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_emit(const uint4* __restrict__ recs, uint32_t n_slots,
                                              const uint32_t* __restrict__ P, const uint32_t* __restrict__ tbase,
                                              const uint32_t* __restrict__ node_base, BvhNode* nodes, uint32_t nodes_cap,
                                              BuildState* st) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long s_add = 0;
    uint32_t i_add = 0;
    if (slot < n_slots) {
        const uint4 r1 = recs[3 * (size_t)slot + 1];
        if (r1.w != 0) {
            const uint4 r0 = recs[3 * (size_t)slot], r2 = recs[3 * (size_t)slot + 2];
            const uint32_t start = r0.w, count = r1.w;
            const uint32_t mesh = r2.w >> TF_MESH_SHIFT;
            const uint32_t mb = tbase[mesh], Pb = P[mb], nb = node_base[mesh];  // numbering restarts per mesh
            const bool root = (r2.w & TF_ROOT) != 0;
            const uint32_t pos = nb + (root ? 0u : 2u + 2u * (P[r2.y] - Pb + r2.z) + (r2.w & TF_RIGHT));
            uint32_t lf, cn;
            if (count > 3) { lf = 2u + 2u * (P[start] - Pb + r2.x); cn = 0; s_add = count; i_add = 1; }
            else { lf = start - mb; cn = count; }
            if (pos < nodes_cap && nb + 1 < nodes_cap) {
                uint4* o = reinterpret_cast<uint4*>(nodes + pos);
                o[0] = make_uint4(r0.x, r0.y, r0.z, lf);
                o[1] = make_uint4(r1.x, r1.y, r1.z, cn);
                if (root) {  // node 1 of every mesh is never used (blas.rs:90)
                    uint4* z = reinterpret_cast<uint4*>(nodes + nb + 1);
                    z[0] = make_uint4(0, 0, 0, 0);
                    z[1] = make_uint4(0, 0, 0, 0);
                }
            } else atomicOr(&st->err, DERR_QUEUE);
        }
    }
    // block-level aggregation of the S / interior counters
    __shared__ unsigned long long s_s;
    __shared__ uint32_t s_i;
    if (threadIdx.x == 0) { s_s = 0; s_i = 0; }
    __syncthreads();
    if (s_add) { atomicAdd(&s_s, s_add); atomicAdd(&s_i, i_add); }
    __syncthreads();
    if (threadIdx.x == 0 && s_i) { atomicAdd(&st->sum_interior, s_s); atomicAdd(&st->interior_total, s_i); }
}

// Per-mesh node counts M_m = 2 + 2 * interior_m (blas.rs:93) from the scanned A, written where the scan kernels
// will turn them into node_base; also validates nothing.
__global__ void __launch_bounds__(256) k_mesh_counts(const uint32_t* __restrict__ P, const uint32_t* __restrict__ tbase,
                                                     uint32_t n_meshes, uint32_t* node_base) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < n_meshes) node_base[m] = 2u + 2u * (P[tbase[m + 1]] - P[tbase[m]]);
    if (m == n_meshes) node_base[m] = 0;
}

// Same for a scene of at most 1024 meshes, together with the exclusive scan that turns the counts into node bases and
// the total: one block instead of four launches (a single mesh is the common case).
__global__ void __launch_bounds__(1024) k_mesh_bases_small(const uint32_t* __restrict__ P, const uint32_t* __restrict__ tbase,
                                                          uint32_t n_meshes, uint32_t* node_base, uint32_t* total) {
    __shared__ uint32_t s_w[32];
    const uint32_t m = threadIdx.x, lane = m & 31, warp = m >> 5;
    const uint32_t v = (m < n_meshes) ? 2u + 2u * (P[tbase[m + 1]] - P[tbase[m]]) : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL_MASK, inc, o);
        if ((int)lane >= o) inc += y;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    const uint32_t wv = s_w[lane];
    const uint32_t before = __reduce_add_sync(FULL_MASK, lane < warp ? wv : 0u);
    const uint32_t all = __reduce_add_sync(FULL_MASK, wv);
    if (m < n_meshes) node_base[m] = before + inc - v;
    if (m == n_meshes) node_base[m] = all;
    if (m == 0) *total = all;
}

__global__ void __launch_bounds__(256) k_write_bvh_index(MeshInfo* infos, const uint32_t* __restrict__ node_base, uint32_t n_meshes) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < n_meshes) infos[m].bvh_index = node_base[m];
}

// indices[i] <- indices_in[order[i]]  (blas.rs:95-100)
__global__ void __launch_bounds__(256) k_permute_gather(const uint32_t* __restrict__ I, const uint32_t* __restrict__ order,
                                                        uint32_t N, uint32_t* tmp) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const size_t s = 3 * (size_t)order[i];
    tmp[3 * (size_t)i] = I[s];
    tmp[3 * (size_t)i + 1] = I[s + 1];
    tmp[3 * (size_t)i + 2] = I[s + 2];
}

__global__ void k_init_state(BuildState* st) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { BuildState s{}; *st = s; }
}

// Mesh table of a batched build: triangle base and vertex offset per mesh from the caller's MeshInfo array (the
// meshes must be pooled back to back in order, as MeshPool::add lays them out, mesh/mod.rs:310-331).
__global__ void __launch_bounds__(256) k_mesh_table(const MeshInfo* __restrict__ infos, uint32_t n_meshes, uint32_t n_indices,
                                                    uint32_t* tbase, uint32_t* voff, BuildState* st) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m > n_meshes) return;
    if (m == n_meshes) { tbase[m] = n_indices / 3; return; }
    const MeshInfo mi = infos[m];
    const uint32_t next = (m + 1 < n_meshes) ? infos[m + 1].base_index : n_indices;
    if (mi.base_index % 3 != 0 || mi.index_count % 3 != 0 || mi.index_count == 0 || mi.base_index + mi.index_count != next ||
        (m == 0 && mi.base_index != 0) || mi.vertex_offset < 0)
        atomicOr(&st->err, DERR_BAD_INDEX);
    tbase[m] = mi.base_index / 3;
    voff[m] = (uint32_t)mi.vertex_offset;
}

__global__ void k_single_mesh_table(uint32_t N, uint32_t* tbase, uint32_t* voff) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { tbase[0] = 0; tbase[1] = N; voff[0] = 0; }
}

// One root per mesh, routed to the tier of its size.
__global__ void __launch_bounds__(256) k_roots(const uint32_t* __restrict__ tbase, uint32_t n_meshes, Queues Q, LevelNode* lv0,
                                               uint32_t lv_cap, BuildState* st, uint32_t epoch) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_meshes) return;
    // An inconsistent caller mesh table (k_mesh_table) or an out-of-range vertex index (k_setup) has already been
    // flagged by an earlier launch: post no root at all, so that every tier exits at once and the build returns EINVAL
    // without ever sizing tiles or queues from a bogus triangle range.
    if (ld_vol(&st->err) & DERR_BAD_INDEX) return;
    const uint32_t start = tbase[m], n = tbase[m + 1] - tbase[m];
    const uint32_t flags = TF_ROOT | (m << TF_MESH_SHIFT);
    if (n == 0 || tbase[m + 1] < start) { atomicOr(&st->err, DERR_BAD_INDEX); return; }
    if (n > Q.tc_cap) {
        const uint32_t idx = atomicAdd(&st->lv_count[0], 1u);
        if (idx >= lv_cap) { atomicOr(&st->err, DERR_QUEUE); return; }
        LevelNode l;
        l.start = start; l.n = n; l.leftrun = 0; l.pstart = start; l.pleftrun = 0; l.flags = flags; l.tile_base = 0; l.pad = 0;
        lv0[idx] = l;
    } else {
        push_any(Q, st, epoch, start, n, 0, start, 0, flags);
    }
}
