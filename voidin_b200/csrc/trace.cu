// trace.cu — ray traversal kernels.
//   k_trace_blas   Bvh::traverse_iter (crates/bvh/src/blas.rs:247-295) with the Rust tests
//                  (crates/bvh/src/intersection.rs:47-55,68-92): division slabs, two-sided triangles,
//                  near pushed first => far popped first.
//   k_trace_scene  traverse_tlas / instance_intersect / traverse_bvh (shaders/utils/bvh.wgsl:35-123) with the
//                  WGSL tests (shaders/utils/intersections.wgsl:13-45): reciprocal slabs, back-face culling,
//                  near child popped first; ANY = early exit at the first accepted triangle, which yields
//                  exactly traverse_tlas(ray).hit (src/bin/raytraced_shadows.wgsl:98-102).
//   k_trace_any    the same any-hit answer without the reference's visit order (see the comment above the kernel);
//                  this is what bvh_cuda_trace_any runs, k_trace_scene<true> takes the rays it defers.
// One thread per ray; nodes are fetched as two 16-byte loads (32-byte aligned), the top of the TLAS from a copy in
// shared memory (TlasTop); per-ray stacks live in
// local memory (64 entries; the reference's 32 / 24-entry stacks overflow silently on deeper trees).
// Compiled with -fmad=false: hit ids depend on exact, unfused float arithmetic in the reference's order.
#include "common.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

constexpr int STACK_CAP = 64;        // BLAS stack (reference: 32 in blas.rs:299, 24 in stack.wgsl:1)
// Lanes that must be waiting before the (more expensive, less frequent) triangle / TLAS paths run in an iteration.
// Measured on the dragon-class scene: thresholds of 10 ("postponed leaves") were 14 % SLOWER than 1, because lanes
// parked on a pending leaf thin out the pop path, which is where most instructions are.
__constant__ uint32_t c_votes[2] = {1, 1};  // {triangle path, TLAS path}; BVH_CUDA_TRACE_VOTES=tri,tlas overrides (tuning only)
#define TRI_VOTE c_votes[0]
#define TLAS_VOTE c_votes[1]
constexpr int TLAS_STACK_CAP = 256;  // the reference's agglomerative TLAS can be deep (88 levels on a 32x32 lattice)
constexpr float MAXD = 1e30f;

struct NodeW {  // BvhNode / TlasNode as two float4
    float4 a, b;
};

__device__ __forceinline__ NodeW ld_node(const void* base, uint32_t i) {
    const float4* p = reinterpret_cast<const float4*>(base) + 2 * (size_t)i;
    NodeW n;
    n.a = __ldg(p);
    n.b = __ldg(p + 1);
    return n;
}

// ---- top of the TLAS in shared memory -------------------------------------------------------------------
// Tlas::build appends every merged node (tlas.rs:64-73), so the nodes nearest the root are the LAST ones of the array
// (plus node 0, the copy of the root, tlas.rs:84).  Every ray starts there and every block walks them all the time:
// each block of the two-level kernels stages the last TLAS_TOP nodes (and node 0, and their unpacked child pairs when
// the side buffer exists) in shared memory once and serves those fetches from there; deeper nodes come through L1/L2.
constexpr int TLAS_TOP = 255;  // + node 0 = 256 staged nodes: 8 KB of nodes + 2 KB of child pairs per block
struct TlasTop {
    float4 n[2 * (TLAS_TOP + 1)];
    uint2 kids[TLAS_TOP + 1];
};
__device__ __forceinline__ uint32_t tlas_top_first(size_t n_tlas_nodes) {
    return n_tlas_nodes > (size_t)TLAS_TOP + 1 ? (uint32_t)(n_tlas_nodes - TLAS_TOP) : 1u;
}
__device__ __forceinline__ void tlas_top_load(TlasTop& t, const BvhCudaSceneDesc& sc) {
    const uint32_t first = tlas_top_first(sc.n_tlas_nodes);
    const uint32_t cnt = (uint32_t)sc.n_tlas_nodes - first;  // nodes [first, n) -> slots [1, cnt]; node 0 -> slot 0
    const float4* g = reinterpret_cast<const float4*>(sc.tlas_nodes);
    for (uint32_t i = threadIdx.x; i < 2 * (cnt + 1); i += blockDim.x) {
        const uint32_t slot = i >> 1, node = slot == 0 ? 0u : first + slot - 1;
        t.n[i] = __ldg(g + 2 * (size_t)node + (i & 1));
    }
    if (sc.tlas_children)
        for (uint32_t slot = threadIdx.x; slot <= cnt; slot += blockDim.x)
            t.kids[slot] = __ldg(reinterpret_cast<const uint2*>(sc.tlas_children) + (slot == 0 ? 0u : first + slot - 1));
    __syncthreads();
}
// slot of TLAS node ni in the staged copy, or 0xFFFFFFFF
__device__ __forceinline__ uint32_t tlas_top_slot(uint32_t ni, uint32_t first) {
    return ni >= first ? ni - first + 1 : (ni == 0 ? 0u : 0xFFFFFFFFu);
}
// TOP = false: the instantiation without the staged copy.  MEASURED (B200, profiles/r02_tlas_top_ab.txt): staging LOSES on
// every configuration -- config 2 (5 TLAS nodes) 2 981 -> 2 870 Mrays/s, config 3 (65 535 nodes, 1 Mi closest-hit rays)
// 256 -> 302 ms, config 5 (2 049 nodes, 128 Mi any-hit rays) 574 -> 551 Mrays/s: the nodes next to the root are L1 hits
// already (l1tex hit rate 73 %), so the shared copy only adds the slot arithmetic and a divergent branch to every TLAS
// step.  The staged kernels stay available (BVH_CUDA_TLAS_TOP=1) and are parity-tested; the default launches TOP = false.
template <bool TOP>
__device__ __forceinline__ NodeW tlas_node(const TlasTop& t, const BvhCudaSceneDesc& sc, uint32_t ni, uint32_t first) {
    if (TOP) {
        const uint32_t s = tlas_top_slot(ni, first);
        if (s != 0xFFFFFFFFu) { NodeW n; n.a = t.n[2 * s]; n.b = t.n[2 * s + 1]; return n; }
    }
    return ld_node(sc.tlas_nodes, ni);
}
template <bool TOP>
__device__ __forceinline__ uint2 tlas_kids(const TlasTop& t, const BvhCudaSceneDesc& sc, uint32_t ni, uint32_t first) {
    if (TOP) {
        const uint32_t s = tlas_top_slot(ni, first);
        if (s != 0xFFFFFFFFu) return t.kids[s];
    }
    return __ldg(reinterpret_cast<const uint2*>(sc.tlas_children) + ni);
}

// ---- Rust-mode tests --------------------------------------------------------------------------------
struct RDist {
    bool hit;
    float t;
};
__device__ __forceinline__ bool dist_gt(RDist a, RDist b) {  // derive(PartialOrd) on enum {Hit(f32), Miss}
    if (a.hit && b.hit) return a.t > b.t;
    if (!a.hit && !b.hit) return false;
    return !a.hit;
}
// intersection.rs:47-55
__device__ __forceinline__ RDist aabb_rs(const float* o, const float* d, const float4& mn, const float4& mx, float t) {
    const float ax = __fdiv_rn(mn.x - o[0], d[0]), ay = __fdiv_rn(mn.y - o[1], d[1]), az = __fdiv_rn(mn.z - o[2], d[2]);
    const float bx = __fdiv_rn(mx.x - o[0], d[0]), by = __fdiv_rn(mx.y - o[1], d[1]), bz = __fdiv_rn(mx.z - o[2], d[2]);
    const float tmax = fminf(fmaxf(ax, bx), fminf(fmaxf(ay, by), fmaxf(az, bz)));
    const float tmin = fmaxf(fminf(ax, bx), fmaxf(fminf(ay, by), fminf(az, bz)));
    RDist r;
    r.hit = (tmax >= tmin) && (tmin < t) && (tmax > 0.0f);
    r.t = tmin;
    return r;
}
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return (ax * bx + ay * by) + az * bz;
}
// intersection.rs:68-92
__device__ __forceinline__ RDist tri_rs(const float* o, const float* d, const float* v0, const float* v1,
                                        const float* v2) {
    const float EPS = 0.0001f;
    RDist miss{false, 0.0f};
    const float e1x = v1[0] - v0[0], e1y = v1[1] - v0[1], e1z = v1[2] - v0[2];
    const float e2x = v2[0] - v0[0], e2y = v2[1] - v0[1], e2z = v2[2] - v0[2];
    const float hx = d[1] * e2z - e2y * d[2], hy = d[2] * e2x - e2z * d[0], hz = d[0] * e2y - e2x * d[1];
    const float a = dot3(e1x, e1y, e1z, hx, hy, hz);
    if (-EPS < a && a < EPS) return miss;
    const float f = __fdiv_rn(1.0f, a);
    const float sx = o[0] - v0[0], sy = o[1] - v0[1], sz = o[2] - v0[2];
    const float u = f * dot3(sx, sy, sz, hx, hy, hz);
    if (!(0.0f <= u && u <= 1.0f)) return miss;
    const float qx = sy * e1z - e1y * sz, qy = sz * e1x - e1z * sx, qz = sx * e1y - e1x * sy;
    const float v = f * dot3(d[0], d[1], d[2], qx, qy, qz);
    if (v < 0.0f || u + v > 1.0f) return miss;
    const float t = f * dot3(e2x, e2y, e2z, qx, qy, qz);
    RDist r;
    r.hit = t > EPS;
    r.t = t;
    return r;
}

// Bvh::traverse_iter (blas.rs:247-295) with persistent warps, dynamic ray fetch and one step per lane per iteration
// (same scheme as k_trace_scene below).  Stack entries are node indices as in the reference: the popped node is
// re-read, because a leaf's triangles and an interior node's children both hang off its (left_first, count).
__global__ void __launch_bounds__(128) k_trace_blas(const BvhNode* __restrict__ nodes, const float* __restrict__ V,
                                                    const uint32_t* __restrict__ I, const float* __restrict__ ro,
                                                    const float* __restrict__ rd, size_t R, float* t_out,
                                                    uint32_t* tri_out, unsigned long long* next_ray) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t stack[STACK_CAP];
    int head = 0;
    bool active = false, exhausted = false;
    size_t r = 0;
    float o[3] = {0, 0, 0}, d[3] = {0, 0, 0};
    bool hit = false;
    float t = 0.0f;
    uint32_t tri = BVH_CUDA_NO_HIT, leaf_first = 0, leaf_cnt = 0;
    for (;;) {
        const uint32_t idle = __ballot_sync(FULL_MASK, !active);
        if (idle != 0 && !exhausted && (__popc(idle) >= 6 || idle == FULL_MASK)) {
            const uint32_t n_idle = __popc(idle);
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(next_ray, (unsigned long long)n_idle);
            base = __shfl_sync(FULL_MASK, base, 0);
            if (base + n_idle >= R) exhausted = true;
            const unsigned long long mine = base + __popc(idle & lt_mask);
            if (!active && mine < R) {
                r = (size_t)mine;
                o[0] = ro[3 * r]; o[1] = ro[3 * r + 1]; o[2] = ro[3 * r + 2];
                d[0] = rd[3 * r]; d[1] = rd[3 * r + 1]; d[2] = rd[3 * r + 2];
                head = 0;
                stack[head++] = 0;
                hit = false; t = 0.0f; tri = BVH_CUDA_NO_HIT; leaf_cnt = 0;
                active = true;
            }
        }
        if (__ballot_sync(FULL_MASK, active) == 0) {
            if (exhausted) break;
            continue;
        }
        if (!active) continue;
        bool finished = false;
        if (leaf_cnt == 0) {
            if (head > 0) {
                const NodeW node = ld_node(nodes, stack[--head]);
                const uint32_t left_first = __float_as_uint(node.a.w), count = __float_as_uint(node.b.w);
                if (count > 0) {
                    leaf_first = left_first;
                    leaf_cnt = count;
                } else {
                    uint32_t min_index = left_first, max_index = left_first + 1;
                    const NodeW ca = ld_node(nodes, min_index), cb = ld_node(nodes, max_index);
                    const float lim = hit ? t : MAXD;
                    RDist min_dist = aabb_rs(o, d, ca.a, ca.b, lim);
                    RDist max_dist = aabb_rs(o, d, cb.a, cb.b, lim);
                    if (dist_gt(min_dist, max_dist)) {
                        const uint32_t ti = min_index; min_index = max_index; max_index = ti;
                        const RDist td = min_dist; min_dist = max_dist; max_dist = td;
                    }
                    if (min_dist.hit) {
                        if (head < STACK_CAP) stack[head++] = min_index;
                        if (max_dist.hit && head < STACK_CAP) stack[head++] = max_index;
                    }
                }
            } else {
                finished = true;
            }
        }
        if (leaf_cnt > 0) {
            const uint32_t ti = leaf_first;
            leaf_first++;
            leaf_cnt--;
            const uint32_t* idx = I + 3 * (size_t)ti;
            const float* p0 = V + 3 * (size_t)idx[0];
            const float* p1 = V + 3 * (size_t)idx[1];
            const float* p2 = V + 3 * (size_t)idx[2];
            const float v0[3] = {p0[0], p0[1], p0[2]}, v1[3] = {p1[0], p1[1], p1[2]}, v2[3] = {p2[0], p2[1], p2[2]};
            const RDist h = tri_rs(o, d, v0, v1, v2);
            if (h.hit) {
                if (!hit) { hit = true; t = h.t; tri = ti; }
                else if (h.t < t) { t = h.t; tri = ti; }
            }
        }
        if (finished) {
            t_out[r] = hit ? t : MAXD;
            tri_out[r] = tri;
            active = false;
        }
    }
}

// Bvh::traverse (blas.rs:211-245), the recursive variant, as an explicit DFS: left before right, each node's box is
// tested against the running t when it is visited, Miss only if the start node's box is missed.
__global__ void __launch_bounds__(128) k_trace_blas_rec(const BvhNode* __restrict__ nodes, const float* __restrict__ V,
                                                        const uint32_t* __restrict__ I, const float* __restrict__ ro,
                                                        const float* __restrict__ rd, size_t R, uint32_t node_idx, float t0,
                                                        float* t_out, uint8_t* hit_out) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float o[3] = {ro[3 * r], ro[3 * r + 1], ro[3 * r + 2]};
    const float d[3] = {rd[3 * r], rd[3 * r + 1], rd[3 * r + 2]};
    uint32_t stack[STACK_CAP];
    int head = 0;
    stack[head++] = node_idx;
    float t = t0;
    bool root_hit = false, first = true;
    while (head > 0) {
        const NodeW node = ld_node(nodes, stack[--head]);
        const bool bh = aabb_rs(o, d, node.a, node.b, t).hit;
        if (first) { root_hit = bh; first = false; }
        if (!bh) continue;
        const uint32_t left_first = __float_as_uint(node.a.w), count = __float_as_uint(node.b.w);
        if (count > 0) {
            for (uint32_t i = 0; i < count; ++i) {
                const uint32_t* idx = I + 3 * (size_t)(left_first + i);
                const float* p0 = V + 3 * (size_t)idx[0];
                const float* p1 = V + 3 * (size_t)idx[1];
                const float* p2 = V + 3 * (size_t)idx[2];
                const float v0[3] = {p0[0], p0[1], p0[2]}, v1[3] = {p1[0], p1[1], p1[2]}, v2[3] = {p2[0], p2[1], p2[2]};
                const RDist h = tri_rs(o, d, v0, v1, v2);
                if (h.hit && h.t < t) t = h.t;  // t.min(dist)
            }
        } else {
            if (head + 2 <= STACK_CAP) { stack[head++] = left_first + 1; stack[head++] = left_first; }
        }
    }
    hit_out[r] = root_hit ? 1 : 0;
    t_out[r] = root_hit ? t : MAXD;
}

// ---- WGSL-mode tests --------------------------------------------------------------------------------
// intersections.wgsl:13-23
__device__ __forceinline__ float aabb_w(const float* eye, const float* inv, const float4& mn, const float4& mx, float t) {
    const float ax = (mn.x - eye[0]) * inv[0], ay = (mn.y - eye[1]) * inv[1], az = (mn.z - eye[2]) * inv[2];
    const float bx = (mx.x - eye[0]) * inv[0], by = (mx.y - eye[1]) * inv[1], bz = (mx.z - eye[2]) * inv[2];
    const float tmax = fminf(fmaxf(ax, bx), fminf(fmaxf(ay, by), fmaxf(az, bz)));
    const float tmin = fmaxf(fminf(ax, bx), fmaxf(fminf(ay, by), fminf(az, bz)));
    return ((tmax >= tmin) && (tmin < t) && (tmax > 0.0f)) ? tmin : MAXD;
}
// intersections.wgsl:25-45
__device__ __forceinline__ bool trig_w(const float* eye, const float* dir, const float* v0, const float* v1,
                                       const float* v2, float* hit) {
    const float e1x = v1[0] - v0[0], e1y = v1[1] - v0[1], e1z = v1[2] - v0[2];
    const float e2x = v2[0] - v0[0], e2y = v2[1] - v0[1], e2z = v2[2] - v0[2];
    const float ux = dir[1] * e2z - e2y * dir[2], uy = dir[2] * e2x - e2z * dir[0], uz = dir[0] * e2y - e2x * dir[1];
    const float det = dot3(e1x, e1y, e1z, ux, uy, uz);
    if (det < 1e-10f) return false;
    const float inv_det = __fdiv_rn(1.0f, det);
    const float ox = eye[0] - v0[0], oy = eye[1] - v0[1], oz = eye[2] - v0[2];
    const float u = inv_det * dot3(ox, oy, oz, ux, uy, uz);
    if (u < 0.0f || 1.0f < u) return false;
    const float vx = oy * e1z - e1y * oz, vy = oz * e1x - e1z * ox, vz = ox * e1y - e1x * oy;
    const float v = inv_det * dot3(dir[0], dir[1], dir[2], vx, vy, vz);
    if (v < 0.0f || u + v > 1.0f) return false;
    const float t = inv_det * dot3(e2x, e2y, e2z, vx, vy, vz);
    if (t > 0.0f && t < *hit) { *hit = t; return true; }
    return false;
}

// mat4 * vec4 as column sums, left to right (bvh.wgsl:82-83)
__device__ __forceinline__ void mat_mul(const float4* m, const float* p, float w, float* out) {
    const float4 c0 = __ldg(m), c1 = __ldg(m + 1), c2 = __ldg(m + 2), c3 = __ldg(m + 3);
    out[0] = ((c0.x * p[0] + c1.x * p[1]) + c2.x * p[2]) + c3.x * w;
    out[1] = ((c0.y * p[0] + c1.y * p[1]) + c2.y * p[2]) + c3.y * w;
    out[2] = ((c0.z * p[0] + c1.z * p[1]) + c2.z * p[2]) + c3.z * w;
}

// Bakes the pooled (vertices, indices, vertex_offset) triple of every triangle into 3 x float4 in pooled triangle
// order, so a leaf test is 3 independent 16-byte loads instead of 3 index loads + 9 dependent scalar loads
// (same vertex values => same arithmetic as fetch_vertex, bvh.wgsl:30-33).
__global__ void __launch_bounds__(256) k_bake_tris(BvhCudaSceneDesc sc, float4* tris, uint32_t n_tris) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tris) return;
    // mesh that owns pooled triangle t: last mesh with base_index <= 3t (meshes are pooled in order, mesh/mod.rs:327)
    uint32_t lo = 0, hi = (uint32_t)sc.n_meshes;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sc.meshes[mid].base_index <= 3 * t) lo = mid; else hi = mid;
    }
    const uint32_t voff = (uint32_t)sc.meshes[lo].vertex_offset;
    const uint32_t* ip = sc.indices + 3 * (size_t)t;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float* p = sc.vertices + 3 * (size_t)(voff + ip[k]);
        tris[3 * (size_t)t + k] = make_float4(p[0], p[1], p[2], 0.0f);
    }
}

// ---- tight world boxes of the instances (optional, exact-order kernels only) ----------------------------------------
// Tlas::build seeds every TLAS leaf with the untransformed local box (tlas.rs:39), so an instance far from the origin has a
// leaf box stretched all the way back to it and most rays "enter" it: instance_intersect (bvh.wgsl:78-87) then transforms
// the ray, fetches the BLAS root's children and finds that it misses both.  With the box below the kernel drops such a
// visit before it loads the 144-byte instance.  Results are identical: a triangle can only be accepted inside an instance
// if some box of its BLAS passes the float slab test (or, for a root that is a leaf, the triangle test itself), which
// puts the ray within rounding distance of the BLAS root box; the box here is that root box mapped to world space through
// the inverse of the linear part of `inv_transform` (in double; the reference maps rays with inv_transform and never
// looks at `transform`), grown by 1e-3 of its own size, and the test adds 1e-4 of the coordinates involved -- both far above
// the error of the object-space arithmetic for the conditioning this accepts (cond <= 100; anything else, or any
// non-finite number, gets an infinite box, i.e. no culling).  w of the low corner = largest |coordinate| of the box.
__global__ void __launch_bounds__(128) k_instance_wbox(BvhCudaSceneDesc sc, float4* wbox) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sc.n_instances) return;
    const float inf = __int_as_float(0x7F800000);
    float4 lo = make_float4(-inf, -inf, -inf, inf), hi = make_float4(inf, inf, inf, 0.0f);
    const Instance* in = sc.instances + i;
    if (in->mesh < sc.n_meshes && sc.meshes[in->mesh].bvh_index < sc.n_bvh_nodes) {
        const BvhNode root = sc.bvh_nodes[sc.meshes[in->mesh].bvh_index];
        const float* A = in->inv_transform;  // column-major; object = L * world + t
        const double l00 = A[0], l10 = A[1], l20 = A[2], l01 = A[4], l11 = A[5], l21 = A[6], l02 = A[8], l12 = A[9], l22 = A[10];
        const double t0 = A[12], t1 = A[13], t2 = A[14];
        const double c00 = l11 * l22 - l12 * l21, c01 = l02 * l21 - l01 * l22, c02 = l01 * l12 - l02 * l11;
        const double c10 = l12 * l20 - l10 * l22, c11 = l00 * l22 - l02 * l20, c12 = l02 * l10 - l00 * l12;
        const double c20 = l10 * l21 - l11 * l20, c21 = l01 * l20 - l00 * l21, c22 = l00 * l11 - l01 * l10;
        const double det = l00 * c00 + l01 * c10 + l02 * c20;
        const double nl = sqrt(l00 * l00 + l10 * l10 + l20 * l20 + l01 * l01 + l11 * l11 + l21 * l21 + l02 * l02 + l12 * l12 + l22 * l22);
        const double na = sqrt(c00 * c00 + c01 * c01 + c02 * c02 + c10 * c10 + c11 * c11 + c12 * c12 + c20 * c20 + c21 * c21 + c22 * c22);
        const double cond = nl * na / fabs(det);  // |L|_F * |L^-1|_F
        if (isfinite(det) && det != 0.0 && isfinite(cond) && cond <= 100.0 && isfinite(t0) && isfinite(t1) && isfinite(t2)) {
            const double r = 1.0 / det;
            double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
            bool ok = true;
            for (int c = 0; c < 8; ++c) {
                const double x = ((c & 1) ? root.max[0] : root.min[0]) - t0, y = ((c & 2) ? root.max[1] : root.min[1]) - t1,
                             z = ((c & 4) ? root.max[2] : root.min[2]) - t2;
                const double w[3] = {(c00 * x + c01 * y + c02 * z) * r, (c10 * x + c11 * y + c12 * z) * r, (c20 * x + c21 * y + c22 * z) * r};
                for (int k = 0; k < 3; ++k) {
                    ok = ok && isfinite(w[k]);
                    mn[k] = fmin(mn[k], w[k]);
                    mx[k] = fmax(mx[k], w[k]);
                }
            }
            if (ok) {
                double amax = 0.0, ext = 0.0;
                for (int k = 0; k < 3; ++k) { amax = fmax(amax, fmax(fabs(mn[k]), fabs(mx[k]))); ext = fmax(ext, mx[k] - mn[k]); }
                const double g = 1e-3 * (amax + ext) + 1e-30;
                if (amax < 1e30) {
                    lo = make_float4(__double2float_rd(mn[0] - g), __double2float_rd(mn[1] - g), __double2float_rd(mn[2] - g),
                                     __double2float_ru(amax + g));
                    hi = make_float4(__double2float_ru(mx[0] + g), __double2float_ru(mx[1] + g), __double2float_ru(mx[2] + g), 0.0f);
                }
            }
        }
    }
    wbox[2 * i] = lo;
    wbox[2 * i + 1] = hi;
}

// true when the ray certainly misses the (grown) world box of an instance; NaNs and unusable reciprocals never say "miss"
__device__ __forceinline__ bool wbox_miss(const float4* __restrict__ wbox, uint32_t ii, const float* eye, const float* inv) {
    const float4 lo = __ldg(wbox + 2 * (size_t)ii), hi = __ldg(wbox + 2 * (size_t)ii + 1);
    const float m = 1e-4f * (fmaxf(fmaxf(fabsf(eye[0]), fabsf(eye[1])), fabsf(eye[2])) + lo.w);
    const float ax = (lo.x - m - eye[0]) * inv[0], ay = (lo.y - m - eye[1]) * inv[1], az = (lo.z - m - eye[2]) * inv[2];
    const float bx = (hi.x + m - eye[0]) * inv[0], by = (hi.y + m - eye[1]) * inv[1], bz = (hi.z + m - eye[2]) * inv[2];
    const float tmx = fminf(fmaxf(ax, bx), fminf(fmaxf(ay, by), fmaxf(az, bz)));
    const float tmn = fmaxf(fminf(ax, bx), fmaxf(fminf(ay, by), fminf(az, bz)));
    const bool usable = isfinite(inv[0]) && isfinite(inv[1]) && isfinite(inv[2]) && inv[0] != 0.0f && inv[1] != 0.0f && inv[2] != 0.0f;
    return usable && ((tmx < tmn) || (tmx < 0.0f));
}

// Stack entries carry what the pop needs (left_first | count << 30), taken from the child node at the time its box
// is tested, so a node is fetched once (as a child) instead of twice.
__device__ __forceinline__ uint32_t pack_meta(const NodeW& n) {
    return (__float_as_uint(n.a.w) & 0x3FFFFFFFu) | (__float_as_uint(n.b.w) << 30);
}

// Persistent-warp traversal with dynamic ray fetch, one traversal STEP per lane per iteration.
// Per-ray work has a heavy tail (mean ~35 node visits, some rays > 1000).  With one ray per thread for the life of a
// warp only ~4.5 of 32 lanes were busy (ncu r01b); a while-while loop ("advance every ray to its next leaf") still ran
// the interior-node code with 5.4 lanes (ncu r01c), because the number of interior pops before the next leaf varies
// a lot between rays.  Here every lane keeps its own small state machine and each iteration performs exactly one step
// of whatever its ray needs next — a TLAS pop, a BLAS pop (interior node: two child boxes; leaf: unpack) or one
// triangle test — so no lane waits for another ray's loop to end.  At the top of each iteration (a converged point)
// idle lanes grab the next unprocessed rays from a global counter (one warp-aggregated atomicAdd).  The per-ray
// visit order is exactly the reference's; only the lane/ray assignment is dynamic.
template <bool ANY, bool TOP>
__global__ void __launch_bounds__(128) k_trace_scene(BvhCudaSceneDesc sc, const float4* __restrict__ tris,
                                                     const float* __restrict__ ro, const float* __restrict__ rd, size_t R,
                                                     float tmax, float* t_out, uint32_t* tri_out, uint32_t* inst_out,
                                                     uint8_t* occ_out, unsigned long long* next_ray,
                                                     const uint32_t* __restrict__ list = nullptr,
                                                     const unsigned long long* list_ctl = nullptr,
                                                     const float4* __restrict__ wbox = nullptr) {
    // Optional indirection: trace only the rays the wide any-hit kernel deferred (list_ctl[0] = how many,
    // list_ctl[2] != 0 = all of them, in which case the list is not used).
    if (list_ctl) {
        if (list_ctl[2] != 0) list = nullptr;
        else R = (size_t)list_ctl[0];
    }
    __shared__ TlasTop s_top;
    if (TOP) tlas_top_load(s_top, sc);
    const uint32_t top_first = tlas_top_first(sc.n_tlas_nodes);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t tstack[TLAS_STACK_CAP], bstack[STACK_CAP];
    int th = 0, bh = 0;
    bool active = false;
    size_t r = 0;
    float eye[3] = {0, 0, 0}, dir[3] = {0, 0, 0}, inv[3] = {0, 0, 0};
    float e2[3] = {0, 0, 0}, d2[3] = {0, 0, 0}, inv2[3] = {0, 0, 0};
    float dist = tmax, hit = tmax;
    uint32_t tri = BVH_CUDA_NO_HIT, inst = BVH_CUDA_NO_HIT, ii = 0, tri_base = 0, bvh_index = 0;
    uint32_t leaf_first = 0, leaf_cnt = 0;  // triangles of the current leaf still to test
    bool res_hit = false;
    bool exhausted = false;  // warp-uniform: the ray counter ran past R

    for (;;) {
        // ---- refill idle lanes (converged) ----
        const uint32_t idle = __ballot_sync(FULL_MASK, !active);
        if (idle != 0 && !exhausted && (__popc(idle) >= 6 || idle == FULL_MASK)) {
            const uint32_t n_idle = __popc(idle);
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(next_ray, (unsigned long long)n_idle);
            base = __shfl_sync(FULL_MASK, base, 0);
            if (base + n_idle >= R) exhausted = true;
            const unsigned long long mine = base + __popc(idle & lt_mask);
            if (!active && mine < R) {
                r = list ? (size_t)list[mine] : (size_t)mine;
                eye[0] = ro[3 * r]; eye[1] = ro[3 * r + 1]; eye[2] = ro[3 * r + 2];
                dir[0] = rd[3 * r]; dir[1] = rd[3 * r + 1]; dir[2] = rd[3 * r + 2];
                inv[0] = __fdiv_rn(1.0f, dir[0]); inv[1] = __fdiv_rn(1.0f, dir[1]); inv[2] = __fdiv_rn(1.0f, dir[2]);
                th = 0; bh = 0; leaf_cnt = 0;
                tstack[th++] = 0;
                dist = tmax; hit = tmax;
                tri = BVH_CUDA_NO_HIT; inst = BVH_CUDA_NO_HIT;
                res_hit = false;
                active = true;
            }
        }
        if (__ballot_sync(FULL_MASK, active) == 0) {
            if (exhausted) break;
            continue;
        }
        // ---- vote which step kinds run this iteration ----
        // The pop path always runs.  The triangle and TLAS paths are more expensive and needed less often, so lanes
        // that need them wait until enough lanes do (or nothing else can run): "postponed leaves".  A lane's own
        // sequence of steps is unchanged, so results are identical.
        const bool want_pop = active && leaf_cnt == 0 && bh > 0;
        const bool want_tlas = active && leaf_cnt == 0 && bh == 0 && th > 0;
        const uint32_t n_pop = __popc(__ballot_sync(FULL_MASK, want_pop));
        const uint32_t n_tlas = __popc(__ballot_sync(FULL_MASK, want_tlas));
        bool finished = active && leaf_cnt == 0 && bh == 0 && th == 0;
        if (want_pop) {
            // traverse_bvh (bvh.wgsl:35-76): one pop
            const uint32_t m = bstack[--bh];
            if (m >> 30) {
                leaf_cnt = m >> 30;
                leaf_first = m & 0x3FFFFFFFu;
            } else {
                const uint32_t c0 = bvh_index + (m & 0x3FFFFFFFu);
                const NodeW ca = ld_node(sc.bvh_nodes, c0), cb = ld_node(sc.bvh_nodes, c0 + 1);
                uint32_t min_meta = pack_meta(ca), max_meta = pack_meta(cb);
                float min_dist = aabb_w(e2, inv2, ca.a, ca.b, hit);
                float max_dist = aabb_w(e2, inv2, cb.a, cb.b, hit);
                if (min_dist > max_dist) {
                    const uint32_t ti = min_meta; min_meta = max_meta; max_meta = ti;
                    const float td = min_dist; min_dist = max_dist; max_dist = td;
                }
                if (!(min_dist >= hit)) {
                    if (max_dist <= hit && bh < STACK_CAP) bstack[bh++] = max_meta;
                    if (bh < STACK_CAP) bstack[bh++] = min_meta;
                }
            }
        }
        const uint32_t n_tri = __popc(__ballot_sync(FULL_MASK, active && leaf_cnt > 0));
        if (n_tlas != 0 && (n_tlas >= TLAS_VOTE || (n_pop == 0 && n_tri < TRI_VOTE))) {
            if (want_tlas) {
                // traverse_tlas (bvh.wgsl:89-123): one pop
                const uint32_t ni = tstack[--th];
                const NodeW node = tlas_node<TOP>(s_top, sc, ni, top_first);
                const uint32_t left_right = __float_as_uint(node.a.w);
                if (left_right == 0) {
                    // instance_intersect (bvh.wgsl:78-87); a visit that cannot accept a triangle is dropped (k_instance_wbox)
                    const uint32_t iv = __float_as_uint(node.b.w);
                    if (!(wbox && wbox_miss(wbox, iv, eye, inv))) {
                        ii = iv;
                        const Instance* in = sc.instances + ii;
                        const MeshInfo* mesh = sc.meshes + in->mesh;
                        tri_base = mesh->base_index / 3u;
                        bvh_index = mesh->bvh_index;
                        const float4* im = reinterpret_cast<const float4*>(in->inv_transform);
                        mat_mul(im, eye, 1.0f, e2);
                        mat_mul(im, dir, 0.0f, d2);
                        inv2[0] = __fdiv_rn(1.0f, d2[0]); inv2[1] = __fdiv_rn(1.0f, d2[1]); inv2[2] = __fdiv_rn(1.0f, d2[2]);
                        bstack[bh++] = pack_meta(ld_node(sc.bvh_nodes, bvh_index));
                        hit = dist;
                    }
                } else {
                    uint32_t min_index, max_index;
                    if (sc.tlas_children) {
                        const uint2 k = tlas_kids<TOP>(s_top, sc, ni, top_first);
                        min_index = k.x; max_index = k.y;
                    } else {
                        min_index = left_right & 0xFFFFu; max_index = left_right >> 16;
                    }
                    // The reference's root is merged with itself (tlas.rs:61), so both children can be the same
                    // node.  Its second traversal can never accept a triangle (same boxes, hit only shrinks,
                    // acceptance is strict t < hit), so it is skipped: results are identical.
                    const bool twin = (min_index == max_index);
                    const NodeW ca = tlas_node<TOP>(s_top, sc, min_index, top_first), cb = tlas_node<TOP>(s_top, sc, max_index, top_first);
                    float min_dist = aabb_w(eye, inv, ca.a, ca.b, dist);
                    float max_dist = aabb_w(eye, inv, cb.a, cb.b, dist);
                    if (min_dist > max_dist) {
                        const uint32_t ti = min_index; min_index = max_index; max_index = ti;
                        const float td = min_dist; min_dist = max_dist; max_dist = td;
                    }
                    if (!(min_dist >= dist)) {
                        if (!twin && max_dist < dist && th < TLAS_STACK_CAP) tstack[th++] = max_index;
                        if (th < TLAS_STACK_CAP) tstack[th++] = min_index;
                    }
                }
            }
        }
        if (n_tri != 0 && (n_tri >= TRI_VOTE || n_pop == 0)) {
            if (active && leaf_cnt > 0) {
                // one triangle of the current leaf (bvh.wgsl:43-51)
                const uint32_t idx = leaf_first;
                leaf_first++;
                leaf_cnt--;
                const float4* tp = tris + 3 * (size_t)(tri_base + idx);
                const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                const float v0[3] = {a.x, a.y, a.z}, v1[3] = {b.x, b.y, b.z}, v2[3] = {c.x, c.y, c.z};
                if (trig_w(e2, d2, v0, v1, v2, &hit)) {
                    dist = hit; tri = idx; inst = ii; res_hit = true;
                    if (ANY) finished = true;
                }
            }
        }
        if (finished) {
            if (ANY) {
                occ_out[r] = res_hit ? 1 : 0;
            } else {
                t_out[r] = res_hit ? dist : MAXD;
                tri_out[r] = tri;
                inst_out[r] = inst;
            }
            active = false;
        }
    }
}


// =====================================================================================================
// Any-hit without the reference's visit order.
//
// traverse_tlas(ray).hit (raytraced_shadows.wgsl:98-102) does not depend on the order in which nodes are visited:
// until the first triangle is accepted `hit` stays at tmax, so which leaves are reached is a pure function of box
// tests made with t = tmax.  Written out for traverse_bvh (bvh.wgsl:35-76) with tmax = 1e30: at an interior node
// both children are pushed as soon as ONE of their boxes is hit (the far one because `max_dist <= hit` also holds
// for the miss value 1e30); a popped interior child whose own box was missed pushes nothing, because child boxes
// are contained in the parent box and every operation of the slab test is monotone in the box planes (no NaN when
// 1/dir is finite and non-zero), so its children miss as well; a popped LEAF has its triangles tested without
// looking at its box again.  Hence
//     leaf X is tested      <=>  box(X) is hit  or  box(sibling(X)) is hit     (tmax = 1e30)
//     leaf X is tested      <=>  box(X) is hit                                  (tmax < 1e30: a missed far child is not pushed)
//     interior X matters    <=>  box(X) is hit
// With the order free, interior nodes and leaves go to two separate per-lane stacks and every iteration a lane
// does one node step AND one triangle test, instead of one or the other; missed interior children are never
// visited.  Rays for which the argument does not hold — a non-finite or zero reciprocal direction (world or object
// space), a stack overflow — are handed to the exact kernel (k_trace_scene<true>) through a list; tmax > 1e30 uses
// the exact kernel for everything.  The per-ray answer is bit-identical to the reference order's.
constexpr int A_LSTACK = 24;
// layout of the scene's 256-byte control block (scene->counter), in u64 words
constexpr int CTL_RAY = 0, CTL_DEFER_N = 1, CTL_DEFER_RAY = 2;  // word 3: "defer everything" (unused, stays 0)

__device__ __forceinline__ bool slab_hit_w(const float* e, const float* inv, const float4& mn, const float4& mx, float t, float& tnear) {
    const float ax = (mn.x - e[0]) * inv[0], ay = (mn.y - e[1]) * inv[1], az = (mn.z - e[2]) * inv[2];
    const float bx = (mx.x - e[0]) * inv[0], by = (mx.y - e[1]) * inv[1], bz = (mx.z - e[2]) * inv[2];
    const float tmax = fminf(fmaxf(ax, bx), fminf(fmaxf(ay, by), fmaxf(az, bz)));
    const float tmin = fmaxf(fminf(ax, bx), fmaxf(fminf(ay, by), fminf(az, bz)));
    tnear = tmin;
    return (tmax >= tmin) && (tmin < t) && (tmax > 0.0f);
}
__device__ __forceinline__ bool inv_usable(const float* inv) {
    return isfinite(inv[0]) && isfinite(inv[1]) && isfinite(inv[2]) && inv[0] != 0.0f && inv[1] != 0.0f && inv[2] != 0.0f;
}

template <bool TOP>
__global__ void __launch_bounds__(128, 9) k_trace_any(BvhCudaSceneDesc sc, const float4* __restrict__ tris,
                                                   const float* __restrict__ ro, const float* __restrict__ rd, size_t R,
                                                   float tmax, uint8_t* occ_out, unsigned long long* ctl,
                                                   uint32_t* defer_list, uint32_t steal) {
    __shared__ TlasTop s_top;
    if (TOP) tlas_top_load(s_top, sc);
    const uint32_t top_first = tlas_top_first(sc.n_tlas_nodes);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const bool pair_rule = tmax >= MAXD;  // a missed far child is pushed iff 1e30 <= hit (bvh.wgsl:71)
    uint32_t tstack[TLAS_STACK_CAP], nstack[STACK_CAP], lstack[A_LSTACK];
    int th = 0, nh = 0, lh = 0;
    bool active = false;
    size_t r = 0;
    float eye[3] = {0, 0, 0}, dir[3] = {0, 0, 0}, inv[3] = {0, 0, 0};
    float e2[3] = {0, 0, 0}, d2[3] = {0, 0, 0}, inv2[3] = {0, 0, 0};
    uint32_t tri_base = 0, bvh_index = 0, leaf_first = 0, leaf_cnt = 0;
    constexpr uint32_t NO_NODE = 0xFFFFFFFFu;
    uint32_t top = NO_NODE;  // top of the interior-node stack; nstack[0..nh) holds the rest
    bool exhausted = false;

    for (;;) {
        bool defer = false;
        const uint32_t idle = __ballot_sync(FULL_MASK, !active);
        if (idle != 0 && !exhausted && (__popc(idle) >= 6 || idle == FULL_MASK)) {
            const uint32_t n_idle = __popc(idle);
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&ctl[CTL_RAY], (unsigned long long)n_idle);
            base = __shfl_sync(FULL_MASK, base, 0);
            if (base + n_idle >= R) exhausted = true;
            const unsigned long long mine = base + __popc(idle & lt_mask);
            if (!active && mine < R) {
                r = (size_t)mine;
                eye[0] = ro[3 * r]; eye[1] = ro[3 * r + 1]; eye[2] = ro[3 * r + 2];
                dir[0] = rd[3 * r]; dir[1] = rd[3 * r + 1]; dir[2] = rd[3 * r + 2];
                inv[0] = __fdiv_rn(1.0f, dir[0]); inv[1] = __fdiv_rn(1.0f, dir[1]); inv[2] = __fdiv_rn(1.0f, dir[2]);
                th = 0; nh = 0; lh = 0; leaf_cnt = 0; top = NO_NODE;
                tstack[th++] = 0;
                active = true;
                if (!inv_usable(inv)) defer = true;
            }
        }
        // ---- tail: no rays are left to fetch.  The answer does not depend on the order in which the entries of the
        // interior stack are processed, so an idle lane takes the top entry of a busy lane's stack together with that
        // ray's object-space state and works on it as if it were its own ray (it has no TLAS stack, so it stops when
        // the sub-tree is done).  Only hits are written (the output is zeroed before the launch), so whoever finds
        // one reports it.  Without this a launch ends with a few lanes walking >1000-node rays alone (0.3 ms floor).
        if (exhausted && steal && idle != 0) {
            const bool donor = active && !defer && nh >= 1;
            const uint32_t D = __ballot_sync(FULL_MASK, donor);
            const uint32_t I = __ballot_sync(FULL_MASK, !active);
            if (D != 0 && I != 0) {
                const uint32_t pairs = min((uint32_t)__popc(I), (uint32_t)__popc(D));
                uint32_t given = 0;
                if (donor && (uint32_t)__popc(D & lt_mask) < pairs) given = nstack[--nh];
                const uint32_t my_rank = __popc(I & lt_mask);
                const bool take = !active && my_rank < pairs;
                const uint32_t src = take ? __fns(D, 0, my_rank + 1) : lane;
                const uint32_t g_top = __shfl_sync(FULL_MASK, given, src);
                const uint32_t g_tb = __shfl_sync(FULL_MASK, tri_base, src), g_bi = __shfl_sync(FULL_MASK, bvh_index, src);
                const unsigned long long g_r = __shfl_sync(FULL_MASK, (unsigned long long)r, src);
                float ge[3], gd[3], gi[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    ge[k] = __shfl_sync(FULL_MASK, e2[k], src);
                    gd[k] = __shfl_sync(FULL_MASK, d2[k], src);
                    gi[k] = __shfl_sync(FULL_MASK, inv2[k], src);
                }
                if (take) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) { e2[k] = ge[k]; d2[k] = gd[k]; inv2[k] = gi[k]; }
                    tri_base = g_tb; bvh_index = g_bi; r = (size_t)g_r;
                    top = g_top; th = 0; nh = 0; lh = 0; leaf_cnt = 0;
                    active = true;
                }
            }
        }
        if (__ballot_sync(FULL_MASK, active) == 0) {
            if (exhausted) break;
            continue;
        }
        bool finished = false, occluded = false;
        // ---- node step: one interior node, both child boxes ----
        // The top of the interior stack lives in a register (`top`), so the common "exactly one interior child is
        // hit" step touches no local memory and the child-pair load does not wait for a stack load.
        if (active && !defer && top != NO_NODE && lh <= A_LSTACK - 2) {
            if (nh >= STACK_CAP) {
                defer = true;
            } else {
                const uint32_t c0 = bvh_index + top;
                const NodeW ca = ld_node(sc.bvh_nodes, c0), cb = ld_node(sc.bvh_nodes, c0 + 1);
                uint32_t m0 = pack_meta(ca), m1 = pack_meta(cb);
                float d0, d1;
                bool h0 = slab_hit_w(e2, inv2, ca.a, ca.b, tmax, d0);
                bool h1 = slab_hit_w(e2, inv2, cb.a, cb.b, tmax, d1);
                // a leaf is visited when the pair is hit (tmax = 1e30), an interior node only when its own box is
                const bool any = h0 | h1;
                bool v0 = h0 || (any && pair_rule && (m0 >> 30) != 0);
                bool v1 = h1 || (any && pair_rule && (m1 >> 30) != 0);
                // nearer child last, so that it ends up on top (only a heuristic here)
                if (h0 && h1 && d0 < d1) {
                    const uint32_t tm = m0; m0 = m1; m1 = tm;
                    const bool tv = v0; v0 = v1; v1 = tv;
                }
                const bool i0 = v0 && (m0 >> 30) == 0, i1 = v1 && (m1 >> 30) == 0;
                if (v0 && !i0) lstack[lh++] = m0;
                if (v1 && !i1) lstack[lh++] = m1;
                if (i0 && i1) { nstack[nh++] = m0; top = m1; }
                else if (i0) top = m0;
                else if (i1) top = m1;
                else top = (nh > 0) ? nstack[--nh] : NO_NODE;
            }
        }
        // ---- triangle step: one triangle of the current leaf (bvh.wgsl:43-51) ----
        // A lane with pending leaves can keep doing node steps, so the triangle path waits until enough lanes have
        // a triangle to test, or until some lane has nothing else to do (or no room for more leaves).
        const bool has_tri = active && !defer && (leaf_cnt > 0 || lh > 0);
        const bool must_tri = has_tri && (top == NO_NODE || lh > A_LSTACK - 2);
        const uint32_t tri_lanes = __ballot_sync(FULL_MASK, has_tri);
        const bool run_tri = (__ballot_sync(FULL_MASK, must_tri) != 0) || ((uint32_t)__popc(tri_lanes) >= TRI_VOTE);
        if (run_tri && has_tri) {
            if (leaf_cnt == 0) {
                const uint32_t m = lstack[--lh];
                leaf_cnt = m >> 30;
                leaf_first = m & 0x3FFFFFFFu;
            }
            const uint32_t idx = leaf_first;
            leaf_first++;
            leaf_cnt--;
            const float4* tp = tris + 3 * (size_t)(tri_base + idx);
            const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
            const float v0[3] = {a.x, a.y, a.z}, v1[3] = {b.x, b.y, b.z}, v2[3] = {c.x, c.y, c.z};
            float h = tmax;
            if (trig_w(e2, d2, v0, v1, v2, &h)) { finished = true; occluded = true; }
        }
        // ---- TLAS step (bvh.wgsl:89-123), only when the current instance is exhausted ----
        if (active && !defer && !finished && top == NO_NODE && lh == 0 && leaf_cnt == 0) {
            if (th == 0) {
                finished = true;
            } else {
                const uint32_t ni = tstack[--th];
                const NodeW node = tlas_node<TOP>(s_top, sc, ni, top_first);
                const uint32_t left_right = __float_as_uint(node.a.w);
                if (left_right == 0) {
                    // instance_intersect (bvh.wgsl:78-87): the BLAS root is visited without a box test
                    const uint32_t ii = __float_as_uint(node.b.w);
                    const Instance* in = sc.instances + ii;
                    const MeshInfo* mesh = sc.meshes + in->mesh;
                    tri_base = mesh->base_index / 3u;
                    bvh_index = mesh->bvh_index;
                    const float4* im = reinterpret_cast<const float4*>(in->inv_transform);
                    mat_mul(im, eye, 1.0f, e2);
                    mat_mul(im, dir, 0.0f, d2);
                    inv2[0] = __fdiv_rn(1.0f, d2[0]); inv2[1] = __fdiv_rn(1.0f, d2[1]); inv2[2] = __fdiv_rn(1.0f, d2[2]);
                    if (!inv_usable(inv2)) {
                        defer = true;
                    } else {
                        const uint32_t m = pack_meta(ld_node(sc.bvh_nodes, bvh_index));
                        if (m >> 30) lstack[lh++] = m; else top = m;
                    }
                } else {
                    uint32_t min_index, max_index;
                    if (sc.tlas_children) {
                        const uint2 k = tlas_kids<TOP>(s_top, sc, ni, top_first);
                        min_index = k.x; max_index = k.y;
                    } else {
                        min_index = left_right & 0xFFFFu; max_index = left_right >> 16;
                    }
                    const bool twin = (min_index == max_index);
                    const NodeW ca = tlas_node<TOP>(s_top, sc, min_index, top_first), cb = tlas_node<TOP>(s_top, sc, max_index, top_first);
                    const float d0 = aabb_w(eye, inv, ca.a, ca.b, tmax);
                    const float d1 = aabb_w(eye, inv, cb.a, cb.b, tmax);
                    // traverse_tlas pushes a child iff its distance is below `dist` (bvh.wgsl:116-119), == tmax here
                    if (d0 < tmax && th < TLAS_STACK_CAP) tstack[th++] = min_index;
                    if (d1 < tmax && !twin && th < TLAS_STACK_CAP) tstack[th++] = max_index;
                }
            }
        }
        if (defer) {
            const uint32_t dm = __activemask();
            const uint32_t dl = __ffs(dm) - 1;
            unsigned long long slot = 0;
            if (lane == dl) slot = atomicAdd(&ctl[CTL_DEFER_N], (unsigned long long)__popc(dm));
            slot = __shfl_sync(dm, slot, dl);
            defer_list[slot + __popc(dm & lt_mask)] = (uint32_t)r;
            active = false;
        } else if (finished) {
            if (occluded) occ_out[r] = 1;  // the output is zeroed before the launch; lanes working on stolen sub-trees report hits only
            active = false;
        }
    }
}

}  // namespace

// tuning hook (not part of the ABI): lanes that must be waiting before the triangle / TLAS paths run
extern "C" int bvh_cuda_debug_set_votes(unsigned tri, unsigned tlas) {
    const uint32_t hv[2] = {tri ? tri : 1u, tlas ? tlas : 1u};
    return (int)cudaMemcpyToSymbol(c_votes, hv, sizeof(hv));
}

int trace_blas_device(bvh_cuda_ctx* ctx, const BvhNode* d_nodes, const float* d_vertices, const uint32_t* d_indices,
                      const float* d_ray_o, const float* d_ray_d, size_t n_rays, float* d_t, uint32_t* d_tri,
                      cudaStream_t stream) {
    if (!d_nodes || !d_vertices || !d_indices || !d_ray_o || !d_ray_d || !d_t || !d_tri)
        return ctx_fail(ctx, BVH_CUDA_EINVAL, "trace_blas: null pointer");
    if (n_rays == 0) return BVH_CUDA_OK;
    if (!ctx->trace_counter) CU_CHECK(ctx, cudaMalloc(&ctx->trace_counter, 256));
    unsigned long long* counter = reinterpret_cast<unsigned long long*>(ctx->trace_counter);
    CU_CHECK(ctx, cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream));
    const size_t want = (n_rays + 127) / 128;
    const size_t cap = (size_t)ctx->sm_count * 16;
    k_trace_blas<<<(unsigned)(want < cap ? want : cap), 128, 0, stream>>>(d_nodes, d_vertices, d_indices, d_ray_o, d_ray_d, n_rays, d_t, d_tri, counter);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return BVH_CUDA_OK;
}

int trace_blas_rec_device(bvh_cuda_ctx* ctx, const BvhNode* d_nodes, const float* d_vertices, const uint32_t* d_indices,
                          const float* d_ray_o, const float* d_ray_d, size_t n_rays, uint32_t node_idx, float t0, float* d_t,
                          uint8_t* d_hit, cudaStream_t stream) {
    if (!d_nodes || !d_vertices || !d_indices || !d_ray_o || !d_ray_d || !d_t || !d_hit)
        return ctx_fail(ctx, BVH_CUDA_EINVAL, "trace_blas_recursive: null pointer");
    if (n_rays == 0) return BVH_CUDA_OK;
    const size_t blocks = (n_rays + 127) / 128;
    if (blocks > 0x7FFFFFFFull) return ctx_fail(ctx, BVH_CUDA_EINVAL, "trace_blas_recursive: too many rays for one call");
    k_trace_blas_rec<<<(unsigned)blocks, 128, 0, stream>>>(d_nodes, d_vertices, d_indices, d_ray_o, d_ray_d, n_rays, node_idx, t0, d_t, d_hit);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return BVH_CUDA_OK;
}

int trace_scene_device(bvh_cuda_ctx* ctx, const bvh_cuda_scene* scene, const float* d_ray_o, const float* d_ray_d,
                       size_t n_rays, float tmax, int any_hit, float* d_t, uint32_t* d_tri, uint32_t* d_inst,
                       uint8_t* d_occ, cudaStream_t stream, int slot) {
    // slot 0 / 1: which of the scene's two control blocks (ray counter, deferral counters) and of the context's two deferral
    // lists this launch uses, so that two launches on one scene can be in flight on two streams (host-pointer pipeline)
    if (!scene || !d_ray_o || !d_ray_d) return ctx_fail(ctx, BVH_CUDA_EINVAL, "trace: null pointer");
    if (any_hit ? !d_occ : (!d_t || !d_tri || !d_inst)) return ctx_fail(ctx, BVH_CUDA_EINVAL, "trace: null output");
    if (n_rays == 0) return BVH_CUDA_OK;
    const float4* tris = reinterpret_cast<const float4*>(scene->baked);
    if (!tris) return ctx_fail(ctx, BVH_CUDA_EINVAL, "trace: scene has no baked triangles");
    static const bool votes_set = [] {
        const char* e = getenv("BVH_CUDA_TRACE_VOTES");
        unsigned v[2] = {1, 1};
        if (e && sscanf(e, "%u,%u", &v[0], &v[1]) == 2) {
            uint32_t hv[2] = {v[0] ? v[0] : 1u, v[1] ? v[1] : 1u};
            cudaMemcpyToSymbol(c_votes, hv, sizeof(hv));
        }
        return true;
    }();
    (void)votes_set;
    // persistent warps: the ray counters live in the scene's control block, reset per launch
    unsigned long long* counter = reinterpret_cast<unsigned long long*>(scene->counter) + 4 * (slot ? 1 : 0);
    CU_CHECK(ctx, cudaMemsetAsync(counter, 0, 4 * sizeof(unsigned long long), stream));
    uint32_t*& dlist = slot ? ctx->defer_list2 : ctx->defer_list;
    size_t& dcap = slot ? ctx->defer_cap2 : ctx->defer_cap;
    size_t want = (n_rays + 127) / 128;
    const size_t cap = (size_t)ctx->sm_count * 16;  // 16 blocks x 4 warps per SM is the register-limited maximum
    const unsigned blocks = (unsigned)(want < cap ? want : cap);
    static const bool wide_off = [] { const char* e = getenv("BVH_CUDA_ANYHIT"); return e && !strcmp(e, "exact"); }();
    // instance culling by tight world boxes (exact-order kernels): BVH_CUDA_INSTANCE_CULL=0 switches it off (A/B)
    static const bool cull_off = [] { const char* e = getenv("BVH_CUDA_INSTANCE_CULL"); return e && atoi(e) == 0; }();
    const float4* wb = (scene->wbox_on && scene->wbox && !cull_off) ? reinterpret_cast<const float4*>(scene->wbox) : nullptr;
    // top of the TLAS in shared memory (TlasTop): off by default (measured slower, see tlas_node above); BVH_CUDA_TLAS_TOP=1 enables it
    static const bool top = [] { const char* e = getenv("BVH_CUDA_TLAS_TOP"); return e && atoi(e) != 0; }();
    if (any_hit && !wide_off && tmax <= MAXD && n_rays < 0xFFFFFFFFull) {
        // order-free kernel first; whatever it defers (normally nothing) goes through the exact kernel
        if (dcap < n_rays) {
            if (dlist) cudaFree(dlist);
            dlist = nullptr; dcap = 0;
            cudaError_t e = cudaMalloc(&dlist, n_rays * sizeof(uint32_t));
            if (e != cudaSuccess) return ctx_cuda_fail(ctx, e, "cudaMalloc(defer list)");
            dcap = n_rays;
        }
        // persistent grid: exactly the resident blocks (more warps than that only start when the rays are gone)
        static const int any_bps = [] {
            int occ = 0, occ_top = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trace_any<false>, 128, 0) != cudaSuccess || occ < 1) occ = 8;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_top, k_trace_any<true>, 128, 0) != cudaSuccess || occ_top < 1) occ_top = 8;
            return occ < occ_top ? occ : occ_top;
        }();
        CU_CHECK(ctx, cudaMemsetAsync(d_occ, 0, n_rays, stream));  // k_trace_any writes hits only
        const size_t cap_any = (size_t)ctx->sm_count * any_bps;
        const unsigned blocks_any = (unsigned)(want < cap_any ? want : cap_any);
        // BVH_CUDA_TRACE_STEAL=0 switches the tail's intra-warp work stealing off (A/B)
        static const uint32_t steal = [] { const char* e = getenv("BVH_CUDA_TRACE_STEAL"); return (e && atoi(e) == 0) ? 0u : 1u; }();
        if (top) k_trace_any<true><<<blocks_any, 128, 0, stream>>>(scene->d, tris, d_ray_o, d_ray_d, n_rays, tmax, d_occ, counter, dlist, steal);
        else k_trace_any<false><<<blocks_any, 128, 0, stream>>>(scene->d, tris, d_ray_o, d_ray_d, n_rays, tmax, d_occ, counter, dlist, steal);
        ctx->launches++;
        // (the deferral pass normally finds an empty list: it runs without staging)
        k_trace_scene<true, false><<<blocks, 128, 0, stream>>>(scene->d, tris, d_ray_o, d_ray_d, n_rays, tmax, nullptr, nullptr, nullptr, d_occ,
                                                                counter + CTL_DEFER_RAY, dlist, counter + CTL_DEFER_N, wb);
    } else if (any_hit) {
        if (top) k_trace_scene<true, true><<<blocks, 128, 0, stream>>>(scene->d, tris, d_ray_o, d_ray_d, n_rays, tmax, nullptr, nullptr, nullptr, d_occ, counter, nullptr, nullptr, wb);
        else k_trace_scene<true, false><<<blocks, 128, 0, stream>>>(scene->d, tris, d_ray_o, d_ray_d, n_rays, tmax, nullptr, nullptr, nullptr, d_occ, counter, nullptr, nullptr, wb);
    } else {
        if (top) k_trace_scene<false, true><<<blocks, 128, 0, stream>>>(scene->d, tris, d_ray_o, d_ray_d, n_rays, tmax, d_t, d_tri, d_inst, nullptr, counter, nullptr, nullptr, wb);
        else k_trace_scene<false, false><<<blocks, 128, 0, stream>>>(scene->d, tris, d_ray_o, d_ray_d, n_rays, tmax, d_t, d_tri, d_inst, nullptr, counter, nullptr, nullptr, wb);
    }
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return BVH_CUDA_OK;
}

// (Re)computes the per-instance world boxes from the scene's CURRENT instance / mesh / node buffers and switches the
// culling of the exact-order kernels on (enable = 0: off).  Stream-ordered.
int scene_instance_boxes_device(bvh_cuda_ctx* ctx, bvh_cuda_scene* scene, int enable, cudaStream_t stream) {
    if (!enable) { scene->wbox_on = false; return BVH_CUDA_OK; }
    const size_t n = scene->d.n_instances;
    if (n == 0) return BVH_CUDA_OK;
    if (!scene->wbox) {
        cudaError_t e = cudaMalloc(&scene->wbox, sizeof(float4) * 2 * n);
        if (e != cudaSuccess) return ctx_cuda_fail(ctx, e, "cudaMalloc(instance boxes)");
    }
    k_instance_wbox<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(scene->d, reinterpret_cast<float4*>(scene->wbox));
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    scene->wbox_on = true;
    return BVH_CUDA_OK;
}

// Builds the scene's baked triangle array (device, owned by the scene).  Stream-ordered; the scene's source buffers
// must not change afterwards.
int scene_bake_device(bvh_cuda_ctx* ctx, bvh_cuda_scene* scene, cudaStream_t stream) {
    const size_t n_tris = scene->d.n_indices / 3;
    if (n_tris == 0 || n_tris >= (1ull << 30) || scene->d.n_bvh_nodes >= (1ull << 30))
        return ctx_fail(ctx, BVH_CUDA_EINVAL, "scene: triangle / node count must be in [1, 2^30)");
    if (!scene->baked) {
        cudaError_t e = cudaMalloc(&scene->baked, sizeof(float4) * 3 * n_tris + 256);
        if (e != cudaSuccess) return ctx_cuda_fail(ctx, e, "cudaMalloc(baked triangles)");
        scene->counter = (char*)scene->baked + sizeof(float4) * 3 * n_tris;  // 16-byte aligned tail
    }
    k_bake_tris<<<(unsigned)((n_tris + 255) / 256), 256, 0, stream>>>(scene->d, reinterpret_cast<float4*>(scene->baked), (uint32_t)n_tris);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return BVH_CUDA_OK;
}
