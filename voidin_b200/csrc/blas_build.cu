// blas_build.cu — exact, order-faithful GPU restatement of BvhBuilder::build (crates/bvh/src/blas.rs:69-204).
//
// The reference is NOT a binned builder: for each of 21 candidate planes it physically re-partitions the
// node's primitive range with an unstable two-cursor swap (partition_shuffle, blas.rs:168-182) and costs
// the two halves with exact vertex bounds; the final order, the pivots and therefore the whole topology
// depend on the running array order.  To be bit-exact every node replays all 22 shuffles, each as a
// closed-form scan + scatter (see DESIGN.md "shuffle in scan form"), and evaluates the candidates from
// 3x8 exact bins plus the <=21 "unexamined" primitives that the shuffles single out.
//
// Six tiers, by node size (boundaries measured):
//   k_t1_coop   n > 24576      grid-wide, level-synchronous phases over tiles of 256..2048 slots chosen per level
//                              (global-memory ping-pong, one cooperative launch for all levels)
//   k_t2 (big)  2049..24576    one 1024-thread block per node from a device task queue (payload in shared memory)
//   k_t2        257..2048      one 256-thread block per node, second queue
//   k_t2w       33..256        one warp per node, third queue
//   k_t3        9..32          one warp per sub-tree, explicit DFS stack spread over the lanes
//   k_t4        <= 8           one thread per sub-tree running the reference's sequential loops (t4_seq.cuh)
// Every node writes one 48-byte record at a collision-free slot (leaf: 2*start, interior: 2*split+1);
// DFS pre-order pair numbering (blas.rs:110-112) is recovered afterwards from
//   rank(X) = #interior nodes with start < X.start  +  #ancestors of X sharing X.start
// with one prefix sum over N counters, and a final kernel emits the 32-byte BvhNodes.
#include "common.cuh"
#include "t4_seq.cuh"

#include <cooperative_groups.h>
#include <cstdlib>
#include <atomic>

namespace {

constexpr int T3_MAX = 32;
// Thread-per-sub-tree tier (k_t4, t4_seq.cuh): sub-trees of <= T4_MAX primitives are built by ONE thread each with the
// plain sequential algorithm.  32: k_t4 replaces the warp-per-sub-tree kernel k_t3 altogether; 4..16: k_t3 keeps the
// larger nodes and hands every child of <= T4_MAX to k_t4; 0: tier off.
#ifndef T4_MAX_V
// measured on the dragon-class build (k_t3 + k_t4, ms): 4: 0.94+0.12, 6: 0.67+0.22, 8: 0.49+0.38, 10: 0.39+0.54,
// 12: 0.32+0.68, 16: 0.20+0.99, 32 (no warp tier): 2.41, 0 (no thread tier): 1.34.  k_t4 is bound by its longest task
// (one thread, ~85 K dependent instructions for 16 primitives), so smaller tasks and more threads per SM win.
#define T4_MAX_V 8
#endif
constexpr int T4_MAX = T4_MAX_V;
static_assert(T4_MAX == 0 || (T4_MAX >= 4 && T4_MAX <= 16) || T4_MAX == 32, "T4_MAX_V must be 0, 4..16 or 32");
// Other forms of the small-sub-tree tiers that were built, verified bit-exact and measured slower on B200 (dragon-class,
// commit ac13e1f): all nodes of one depth as lane segments of one warp (1.42 ms vs 1.34 ms: SAH sub-trees are
// deep, not bushy, and a pass costs as much as a node visit); sub-trees through k_t2w's queue (2.94 ms vs 2.46 ms for the
// two tiers); thread tier limited to whole waves of sm_count x T4_THREADS tasks (1.22 ms vs 1.24 ms); thread-tier task
// list counting-sorted by size first (+0.1 ms for the sort, k_t4 unchanged: it is bound by its longest task).
constexpr int T4_CAP = T4_MAX ? T4_MAX : 16;
// 8 words per slot per thread: <= 229 376 B of shared memory per block; 768 threads is the register limit (78 regs)
constexpr int T4_THREADS_SMEM = (229376 / (32 * T4_CAP)) / 32 * 32;
constexpr int T4_THREADS = T4_THREADS_SMEM < 768 ? T4_THREADS_SMEM : 768;
#ifndef T2W_CAP_V
#define T2W_CAP_V 256
#endif
constexpr int T2W_CAP = T2W_CAP_V;  // warp-per-node tier: 33..T2W_CAP
// 1: PB takes the tile's ballots and warp prefix bases from PA (288 B per tile) instead of recomputing them
#ifndef T1_PB_REUSE
#define T1_PB_REUSE 1
#endif
// 1: the grid tier picks the tile size per level (p_t1_nextlevel); 0: always T1_TILE
// grid tier, resident-tile pull form: 0 off, 1 for levels with 256-/512-slot tiles, 2 for all levels (see blas_grid.cuh)
#ifndef T1_PULL_DEFAULT
#define T1_PULL_DEFAULT 0  // measured: mode 1 gains 54 us on the small-tile levels and loses 80 us on the 2048-slot levels of the same kernel (register allocation), profiles/r02_build_variants_ab.txt
#endif
#ifndef T1_VAR_TILE
#define T1_VAR_TILE 1
#endif
#ifndef T2_CAP_V
#define T2_CAP_V 2048
#endif
constexpr int T2_CAP = T2_CAP_V;  // block-per-node tier: 257..T2_CAP (256 threads)
constexpr int T2_THREADS = 256;
#ifndef T2_MIN_BLOCKS
#define T2_MIN_BLOCKS 3  // 64 registers, no spills; 0.85 -> 0.61 ms for the tier on the dragon-class mesh (4 gives no more)
#endif
#ifndef T2B_CAP_V
#define T2B_CAP_V 24576  // measured on the dragon-class build: 16384 -> 6.03 ms, 20480 -> 5.91, 24576 -> 5.86, 28672 -> 6.25 (spills), 32768 -> 6.24 (profiles/r02_build_variants_ab.txt)
#endif
constexpr int T2B_CAP = T2B_CAP_V;  // big-block tier: 2049..T2B_CAP (1024 threads, one block per SM; 6 B of shared memory per slot)
constexpr int T2B_THREADS = 1024;
// Grid tier block shape: threads per block and blocks per SM.  Measured on the dragon-class build (grid tier ms):
// see scripts/variants.py rows t1_512x2 / t1_1024x1 in profiles/r02_build_variants_ab.txt.
#ifndef T1_THREADS_V
#define T1_THREADS_V 256
#endif
#ifndef T1_MIN_BLOCKS
#define T1_MIN_BLOCKS 3
#endif
constexpr int T1_THREADS = T1_THREADS_V;
constexpr int T1_TILE = T1_THREADS * 8;  // largest tile (8 slots per thread); levels that fit the grid with smaller tiles use them
constexpr uint32_t SPIN_LIMIT = 1u << 22;

#define TF_RIGHT 1u
#define TF_ROOT 2u
#define TF_MESH_SHIFT 2  // task / record flags: bit0 right child, bit1 root, bits 2.. mesh id (batched builds)

struct Task {  // 32 B
    uint32_t start, n, leftrun, pstart, pleftrun, flags, ready, pad;
};

struct LevelNode {  // 32 B
    uint32_t start, n, leftrun, pstart, pleftrun, flags, tile_base, pad;
};

struct NodeScratch {
    uint32_t bnd[12];  // ordered-uint: vlo[3], vhi[3], cmin[3], cmax[3]
    uint4 sh[22];      // per shuffle {nL, f, pivot, -}: written by the one thread that owns the boundary element
    uint32_t nL[22];   // per shuffle #L of the node (tile-scan path only)
    uint32_t best, pad;
    uint32_t piv[21], uid[21];
    uint32_t bins[3][8][6];
    unsigned long long zkey[6];  // rare -0.0 path: (first slot with a zero on this face << 32) | its triangle id
};

struct BuildState {
    uint32_t err;
    uint32_t q_head, q_tail, q_pending;
    uint32_t t3_count;
    uint32_t lv_count[2];
    uint32_t lv_tiles[2];
    uint32_t lv_maxtiles[2];
    uint32_t lv_ept[2];  // slots per thread of the level's tiles (tile = T1_THREADS * ept slots)
    uint32_t interior_total;
    unsigned long long sum_interior;
    uint32_t t2_done;
    uint32_t w_head, w_tail, w_pending;
    uint32_t b_head, b_tail, b_pending;
    uint32_t t2b_done;
    uint32_t t2w_done;
    uint32_t levels_done;
    uint32_t t3_inline;
    uint32_t neg_zero;  // some referenced vertex coordinate is -0.0: box zeros need the reference's first-encounter sign
    uint32_t t4_count;
    uint32_t c_head, c_tail, c_pending;  // cluster-tier queue
    uint32_t tc_done;
    uint32_t grid_nodes;            // interior nodes split by the grid and cluster tiers ...
    unsigned long long sum_grid;    // ... and the sum of their primitive counts (statistics for the roofline)
};

static_assert(sizeof(BuildState) <= 256, "BuildState is read back into the first 64 words of the pinned scratch");

__device__ __forceinline__ uint32_t ld_vol(const uint32_t* p) { return *(const volatile uint32_t*)p; }

__device__ __forceinline__ void emit_rec(uint4* recs, uint32_t slot, const float* lo, const float* hi,
                                         uint32_t start, uint32_t count, uint32_t leftrun, uint32_t pstart,
                                         uint32_t pleftrun, uint32_t flags) {
    uint4* r = recs + 3 * (size_t)slot;
    r[0] = make_uint4(__float_as_uint(lo[0]), __float_as_uint(lo[1]), __float_as_uint(lo[2]), start);
    r[1] = make_uint4(__float_as_uint(hi[0]), __float_as_uint(hi[1]), __float_as_uint(hi[2]), count);
    r[2] = make_uint4(leftrun, pstart, pleftrun, flags);
}

// 3-bit plane counts of one centroid: k_a = #{b in 1..7 : !(c_a < pos_ab)}, pos_ab = lerp(cmin,cmax,b/8)[a]
// (blas.rs:145-146,173).  Planes are monotone in b, so "c_a < pos_ab" <=> k_a < b.
__device__ __forceinline__ uint32_t plane_counts(float cx, float cy, float cz, const float* cmin,
                                                 const float* cmax) {
    uint32_t kb = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float c = (a == 0) ? cx : ((a == 1) ? cy : cz);
        uint32_t k = 0;
#pragma unroll
        for (int b = 1; b < 8; ++b) {
            const float pos = lerp1(cmin[a], cmax[a], (float)b * 0.125f);
            k += (c < pos) ? 0u : 1u;
        }
        kb |= k << (3 * a);
    }
    return kb;
}

// SAH cost of one candidate (blas.rs:155): area(bb1) * n1 as f32 + area(bb2) * n2 as f32
__device__ __forceinline__ float sah_cost(const float* L, const float* R, uint32_t n1, uint32_t n2) {
    const float a1 = aabb_area(L[0], L[1], L[2], L[3], L[4], L[5]);
    const float a2 = aabb_area(R[0], R[1], R[2], R[3], R[4], R[5]);
    return __fadd_rn(__fmul_rn(a1, __uint2float_rn(n1)), __fmul_rn(a2, __uint2float_rn(n2)));
}

// ------------------------------------------------------------------------------------------------
// K1 setup: centroid ((v0+v1)+v2)/3 (blas.rs:70-81), per-triangle AABB folded from +-1e30 (blas.rs:185-186),
// identity triangle_indices (blas.rs:83), index validation.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mesh_of_tri(const uint32_t* __restrict__ tbase, uint32_t n_meshes, uint32_t tri) {
    uint32_t lo = 0, hi = n_meshes;  // last mesh with tbase <= tri
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (tbase[mid] <= tri) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_setup(const float* __restrict__ V, uint32_t nV,
                                               const uint32_t* __restrict__ I, uint32_t N,
                                               const uint32_t* __restrict__ tbase, const uint32_t* __restrict__ voff,
                                               uint32_t n_meshes, float4* cent, float4* box, uint32_t* ids,
                                               BuildState* st) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint32_t vo = (n_meshes > 1) ? voff[mesh_of_tri(tbase, n_meshes, i)] : voff[0];
    uint32_t i0 = I[3 * (size_t)i], i1 = I[3 * (size_t)i + 1], i2 = I[3 * (size_t)i + 2];
    const uint32_t lim = vo < nV ? nV - vo : 0u;
    if (i0 >= lim || i1 >= lim || i2 >= lim) {
        atomicOr(&st->err, DERR_BAD_INDEX);
        i0 = i1 = i2 = 0;
    }
    const uint32_t vb = lim ? vo : 0u;
    const float* a = V + 3 * (size_t)(vb + i0);
    const float* b = V + 3 * (size_t)(vb + i1);
    const float* c = V + 3 * (size_t)(vb + i2);
    float lo[3], hi[3], ce[3];
    bool nz = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float x = a[k], y = b[k], z = c[k];
        ce[k] = __fdiv_rn(__fadd_rn(__fadd_rn(x, y), z), 3.0f);
        // f32::min / f32::max keep the accumulator unless the new value is strictly smaller / larger
        // (blas.rs:190-198): among equal zeros of different sign the first one met stays.
        float l = 1e30f, h = -1e30f;
        l = (x < l) ? x : l; l = (y < l) ? y : l; l = (z < l) ? z : l;
        h = (x > h) ? x : h; h = (y > h) ? y : h; h = (z > h) ? z : h;
        lo[k] = l;
        hi[k] = h;
        nz = nz || __float_as_uint(x) == 0x80000000u || __float_as_uint(y) == 0x80000000u || __float_as_uint(z) == 0x80000000u;
    }
    if (nz) atomicOr(&st->neg_zero, 1u);
    cent[i] = make_float4(ce[0], ce[1], ce[2], 0.0f);
    box[2 * (size_t)i] = make_float4(lo[0], lo[1], lo[2], 0.0f);
    box[2 * (size_t)i + 1] = make_float4(hi[0], hi[1], hi[2], 0.0f);
    ids[i] = i;
}

// Device task lists, by node size (see the tier table in DESIGN.md).
struct Queues {
    Task* qc;   // cluster tasks (T2B_CAP < n <= tc_cap)
    Task* qb;   // big-block tasks (T2_CAP < n <= T2B_CAP)
    Task* q;    // block-per-node tasks (T2W_CAP < n <= T2_CAP)
    Task* qw;   // warp-per-node tasks  (T3_MAX < n <= T2W_CAP)
    Task* t3;   // warp-per-sub-tree tasks (T4_MAX < n <= T3_MAX)
    Task* t4;   // thread-per-sub-tree tasks (n <= T4_MAX)
    uint32_t qc_cap, qb_cap, q_cap, qw_cap, t3_cap, t4_cap;
    uint32_t tc_cap;  // largest node the cluster tier takes (T2B_CAP when the tier is off: then the grid tier keeps them)
};

__device__ __forceinline__ void push_t4(const Queues& Q, BuildState* st, uint32_t start, uint32_t n, uint32_t leftrun,
                                        uint32_t pstart, uint32_t pleftrun, uint32_t flags) {
    const uint32_t idx = atomicAdd(&st->t4_count, 1u);
    if (idx >= Q.t4_cap) { atomicOr(&st->err, DERR_QUEUE); return; }
    Task* d = Q.t4 + idx;
    d->start = start; d->n = n; d->leftrun = leftrun; d->pstart = pstart; d->pleftrun = pleftrun;
    d->flags = flags; d->ready = 0; d->pad = 0;
}

#include "blas_small.cuh"
#include "blas_block.cuh"
#include "blas_cluster.cuh"
#include "blas_grid.cuh"
#include "blas_emit.cuh"

}  // namespace

constexpr size_t T2_SMEM = (size_t)T2_CAP * 6;    // u32 payload (shuffled in place) + u16 table per slot
constexpr size_t T2B_SMEM = (size_t)T2B_CAP * 6;
constexpr size_t T4_SMEM = (size_t)8 * T4_CAP * T4_THREADS * 4;  // 6 float + 2 u32 words per slot per thread

int blas_t2_occupancy() {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_t2<T2_CAP, T2_THREADS, false>, T2_THREADS, T2_SMEM) != cudaSuccess) occ = 1;
    return occ < 1 ? 1 : occ;
}

int blas_t2b_setup() {
    if (cudaFuncSetAttribute(k_t2<T2B_CAP, T2B_THREADS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2B_SMEM) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_t2<T2B_CAP, T2B_THREADS, true>, T2B_THREADS, T2B_SMEM) != cudaSuccess) occ = 0;
    return occ;
}

int blas_t1_timing(unsigned long long* out32) {
#ifdef BVH_T1_TIMING
    unsigned long long z[32] = {};
    cudaMemcpyFromSymbol(out32, g_t1_time, sizeof(z));
    cudaMemcpyToSymbol(g_t1_time, z, sizeof(z));
    return 1;
#else
    (void)out32;
    return 0;
#endif
}

// {total nodes, device error bits (DERR_*), interior nodes, 0} for callers that stay on the stream (async builds)
namespace {
__global__ void k_publish_result(const BuildState* st, const uint32_t* scan_total, uint32_t* out) {
    if (threadIdx.x == 0) { out[0] = scan_total[1]; out[1] = st->err; out[2] = scan_total[0]; out[3] = 0; }
}
}  // namespace

int blas_t1_pull(unsigned long long* out4096) {
#ifdef BVH_T1_TIMING
    static unsigned long long z[8 * 512];
    cudaMemcpyFromSymbol(out4096, g_t1_pull, sizeof(z));
    cudaMemcpyToSymbol(g_t1_pull, z, sizeof(z));
    return 1;
#else
    (void)out4096;
    return 0;
#endif
}

int blas_t1_blocks(unsigned long long* out2048) {
#ifdef BVH_T1_TIMING
    static unsigned long long z[2048];
    cudaMemcpyFromSymbol(out2048, g_t1_blk, sizeof(z));
    cudaMemcpyToSymbol(g_t1_blk, z, sizeof(z));
    return 1;
#else
    (void)out2048;
    return 0;
#endif
}

// debug (library built with -DBVH_TC_TIMING): copies and clears the cluster tier's per-node time stamps
int blas_tc_log(unsigned long long* out, unsigned int cap_rows) {
#ifdef BVH_TC_TIMING
    unsigned int n = 0;
    cudaMemcpyFromSymbol(&n, g_tc_logn, sizeof(n));
    if (n > 4096) n = 4096;
    if (n > cap_rows) n = cap_rows;
    if (n) cudaMemcpyFromSymbol(out, g_tc_log, sizeof(unsigned long long) * 8 * n);
    const unsigned int z = 0;
    cudaMemcpyToSymbol(g_tc_logn, &z, sizeof(z));
    if (cap_rows > n + 1) {  // one extra row: the per-phase cycle sums of the shuffles
        cudaMemcpyFromSymbol(out + 8 * (size_t)n, g_tc_phase, sizeof(unsigned long long) * 8);
        const unsigned long long zz[8] = {};
        cudaMemcpyToSymbol(g_tc_phase, zz, sizeof(zz));
    }
    return (int)n;
#else
    (void)out; (void)cap_rows;
    return -1;
#endif
}

int blas_t1_coop_occupancy() {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_t1_coop<0>, T1_THREADS, 0) != cudaSuccess) occ = 1;
    int o1 = 0, o2 = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, k_t1_coop<1>, T1_THREADS, 0) != cudaSuccess) o1 = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, k_t1_coop<2>, T1_THREADS, 0) != cudaSuccess) o2 = 1;
    if (o1 < occ) occ = o1;  // one grid size for all instantiations (BVH_CUDA_T1_PULL selects at run time)
    if (o2 < occ) occ = o2;
    return occ < 1 ? 1 : occ;
}

// Cluster tier: co-resident clusters of `cluster_size` CTAs (0 when such a cluster cannot be placed on this device).
int blas_tc_setup(int cluster_size) {
    static_assert(TS_SMEM >= TC_SMEM, "the two cluster-tier kernels are launched with the same configuration (the larger one)");
    if (cudaFuncSetAttribute(k_tc, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaFuncSetAttribute(k_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_tcs, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaFuncSetAttribute(k_tcs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_SMEM) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster_size * 64);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = TS_SMEM;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, k_tc, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ncs = 0;
    if (cudaOccupancyMaxActiveClusters(&ncs, k_tcs, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return nc < ncs ? nc : ncs;
}

int blas_t2w_occupancy() {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_t2w<T2W_CAP>, 256, 0) != cudaSuccess) occ = 1;
    return occ < 1 ? 1 : occ;
}

namespace {

struct Carver {
    char* base;
    size_t off = 0;
    template <class T>
    T* take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

}  // namespace

int blas_build_device(bvh_cuda_ctx* ctx, const float* d_vertices, size_t n_vertices, uint32_t* d_indices,
                      size_t n_tris, MeshInfo* d_mesh_info, size_t n_meshes, BvhNode* d_nodes_out, size_t nodes_cap,
                      uint32_t* n_nodes_out, cudaStream_t stream, uint32_t* d_result, bool async) {
    if (ctx->pending) return ctx_fail(ctx, BVH_CUDA_EINVAL, "blas_build: an asynchronous build of this context has not been finished (bvh_cuda_blas_build_finish)");
    if (!d_vertices || !d_indices || !d_nodes_out || n_tris == 0 || n_vertices == 0 || n_meshes == 0)
        return ctx_fail(ctx, BVH_CUDA_EINVAL, "blas_build: empty mesh or null pointer");
    if (n_tris > 0x7FFFFFFFull / 2 || n_vertices > 0xFFFFFFFFull || n_meshes > (1ull << 29) || n_meshes > n_tris)
        return ctx_fail(ctx, BVH_CUDA_EINVAL, "blas_build: mesh too large (2*n_tris must fit in 31 bits)");
    if (nodes_cap < 2 * n_tris)
        return ctx_fail(ctx, BVH_CUDA_EINVAL, "blas_build: nodes_cap must be >= 2*n_tris");
    const uint32_t N = (uint32_t)n_tris;
    const uint32_t NM = (uint32_t)n_meshes;
    const uint32_t max_large = N / T2B_CAP + 2;
    // (a level uses tiles smaller than T1_TILE only when it then has no more tiles than the grid has blocks)
    const uint32_t max_tiles = N / T1_TILE + max_large + 2 + (uint32_t)ctx->sm_count * 16;
    // Cluster tier (one node of 16K-262K triangles per thread-block cluster): built, bit-exact and measured SLOWER than
    // the level-synchronous grid tier on every workload tried (dragon-class: 6.12 / 6.32 ms against 6.00 ms; a forest of many
    // mid-size meshes is far worse, 7 clusters take 7 nodes at a time), see DESIGN.md section 4 and profiles/r02_cluster_tier_*.
    // It therefore stays opt-in: BVH_CUDA_TC=smem (order resident in distributed shared memory, k_tcs) or =global (k_tc).
    static const int tc_mode = [] { const char* e = getenv("BVH_CUDA_TC"); return !e ? 0 : (e[0] == 's' ? 1 : (e[0] == 'g' ? 2 : 0)); }();
    const bool use_tc = tc_mode != 0 && ctx->tc_clusters > 0 && ctx->tc_cluster_size > 0;
    const uint32_t tc_cap = use_tc ? (uint32_t)ctx->tc_cluster_size * (uint32_t)TC_SLOTS : (uint32_t)T2B_CAP;
    const uint32_t qc_cap = N / 4096 + NM + 64;   // nodes with 16385..tc_cap primitives
    const uint32_t qb_cap = N / 256 + NM + 4096;  // nodes with 2049..T2B_CAP primitives
    const uint32_t q_cap = N / 32 + NM + 4096;    // nodes with 257..2048 primitives (typically ~N/100)
    const uint32_t qw_cap = N / 4 + NM + 4096;    // nodes with 33..256 primitives (typically ~N/28)
    // sub-trees of <= 32 primitives: thread tasks (<= T4_MAX) and warp tasks (the rest).  Sibling ranges are disjoint,
    // but with 0 < T4_MAX < 32 the warp kernel re-posts children of its own tasks, hence two full-size lists.
    const uint32_t t3_cap = (T4_MAX >= T3_MAX) ? 16u : (T4_MAX > 0 ? N / (uint32_t)(T4_MAX + 1) + NM + 16 : N + NM + 16);
    const uint32_t t4_cap = (T4_MAX > 0) ? N + NM + 16 : 16u;
    const uint32_t scan_n = N + 1;
    const uint32_t scan_blocks = (scan_n + SCAN_TILE - 1) / SCAN_TILE;
    const uint32_t mscan_n = NM + 1;
    const uint32_t mscan_blocks = (mscan_n + SCAN_TILE - 1) / SCAN_TILE;

    // carve the workspace (first pass sizes, second pass assigns)
    float4 *cent = nullptr, *box = nullptr;
    uint4* tile_desc = nullptr;
    uint32_t *ids0 = nullptr, *ids1 = nullptr, *table = nullptr, *A = nullptr, *tileL = nullptr, *tileLF = nullptr, *pbal = nullptr, *barrier = nullptr,
             *scan_sums = nullptr, *scan_total = nullptr, *tbase = nullptr, *voff = nullptr, *node_base = nullptr, *mscan_sums = nullptr;
    uint16_t *fl0 = nullptr, *fl1 = nullptr;
    uint4* recs = nullptr;
    Task *qc = nullptr, *qb = nullptr, *q = nullptr, *qw = nullptr, *t3 = nullptr, *t4 = nullptr;
    unsigned long long* items = nullptr;
    TcScratch* tcs = nullptr;
    LevelNode* lv[2] = {nullptr, nullptr};
    NodeScratch* sc = nullptr;
    BuildState* st = nullptr;
    for (int pass = 0; pass < 2; ++pass) {
        Carver c{pass ? (char*)ctx->ws : nullptr};
        st = c.take<BuildState>(1);
        cent = c.take<float4>(N);
        box = c.take<float4>(2 * (size_t)N);
        ids0 = c.take<uint32_t>(N);
        ids1 = c.take<uint32_t>(N);
        fl0 = c.take<uint16_t>(N);
        fl1 = c.take<uint16_t>(N);
        table = c.take<uint32_t>(N);
        A = c.take<uint32_t>(scan_n);
        recs = c.take<uint4>(3 * 2 * (size_t)N);
        items = c.take<unsigned long long>(use_tc ? N : 1);
        tcs = c.take<TcScratch>(use_tc ? (size_t)ctx->tc_clusters : 1);
        qc = c.take<Task>(qc_cap);
        qb = c.take<Task>(qb_cap);
        q = c.take<Task>(q_cap);
        qw = c.take<Task>(qw_cap);
        t3 = c.take<Task>(t3_cap);
        t4 = c.take<Task>(t4_cap);
        lv[0] = c.take<LevelNode>(max_large);
        lv[1] = c.take<LevelNode>(max_large);
        sc = c.take<NodeScratch>(max_large);
        tileL = c.take<uint32_t>(22 * (size_t)max_tiles);
        tileLF = c.take<uint32_t>(max_tiles);
        pbal = c.take<uint32_t>(2 * (size_t)max_tiles * T1_META);  // PA -> PB ballots (two-phase form) / tile meta, double buffered (pull form)
        tile_desc = c.take<uint4>(max_tiles);
        barrier = c.take<uint32_t>(T1_BARRIER_WORDS);
        scan_sums = c.take<uint32_t>(scan_blocks + 1);
        scan_total = c.take<uint32_t>(4);
        tbase = c.take<uint32_t>(NM + 1);
        voff = c.take<uint32_t>(NM + 1);
        node_base = c.take<uint32_t>(mscan_n);
        mscan_sums = c.take<uint32_t>(mscan_blocks + 1);
        if (pass == 0) {
            int rc = ctx_reserve(ctx, c.off + 256);
            if (rc) return rc;
        }
    }
    // Task slots are marked ready with a build-unique number; the counter is process-wide because a freed workspace
    // of one context can be handed to another by cudaMalloc.
    static std::atomic<uint32_t> g_epoch{0};
    const uint32_t epoch = 0x80000000u | (g_epoch.fetch_add(1) + 1);
    ctx->epoch = epoch;
    uint32_t launches = 0;


    // The three task queues are contiguous; clearing them makes every `ready` word differ from the epoch even when
    // the workspace still holds other data of an earlier, differently sized build.
    CU_CHECK(ctx, cudaMemsetAsync(qc, 0, (size_t)((char*)(qw + qw_cap) - (char*)qc), stream));
    CU_CHECK(ctx, cudaMemsetAsync(A, 0, sizeof(uint32_t) * scan_n, stream));
    CU_CHECK(ctx, cudaMemsetAsync(recs, 0, sizeof(uint4) * 3 * 2 * (size_t)N, stream));
    const bool prof = ctx->profiling;
    if (prof) cudaEventRecord(ctx->ev[0], stream);
    Queues Q{qc, qb, q, qw, t3, t4, qc_cap, qb_cap, q_cap, qw_cap, t3_cap, t4_cap, tc_cap};
    k_init_state<<<1, 32, 0, stream>>>(st);
    if (d_mesh_info) k_mesh_table<<<(NM + 1 + 255) / 256, 256, 0, stream>>>(d_mesh_info, NM, 3 * N, tbase, voff, st);
    else k_single_mesh_table<<<1, 32, 0, stream>>>(N, tbase, voff);
    k_setup<<<(N + 255) / 256, 256, 0, stream>>>(d_vertices, (uint32_t)n_vertices, d_indices, N, tbase, voff, NM, cent, box, ids0, st);
    k_roots<<<(NM + 255) / 256, 256, 0, stream>>>(tbase, NM, Q, lv[0], max_large, st, epoch);
    // grid of the cooperative kernel: enough blocks for one 256-slot tile each, at most what is co-resident
    uint32_t t1_grid = (uint32_t)ctx->sm_count * (uint32_t)(ctx->t1_blocks_per_sm > 0 ? ctx->t1_blocks_per_sm : 1);
    if (t1_grid > (N + T1_THREADS - 1) / T1_THREADS + 8) t1_grid = (N + T1_THREADS - 1) / T1_THREADS + 8;
    if (const char* e = getenv("BVH_CUDA_T1_GRID")) {  // experiments: fewer blocks than are co-resident
        const uint32_t v = (uint32_t)strtoul(e, nullptr, 10);
        if (v >= 1 && v < t1_grid) t1_grid = v;
    }
    k_t1_level0<<<1, 1024, 0, stream>>>(lv[0], st, t1_grid);
    launches += 5;
    if (prof) cudaEventRecord(ctx->ev[1], stream);

    // ---- T1: grid-wide tier, one cooperative persistent launch (exits at once when no node is that large) ----
    if (N > tc_cap) {
        CU_CHECK(ctx, cudaMemsetAsync(barrier, 0, sizeof(uint32_t) * T1_BARRIER_WORDS, stream));
        T1Args g;
        g.nodes = lv[0]; g.sc = sc; g.n_nodes = 0; g.n_tiles = 0;
        g.ids0 = ids0; g.ids1 = ids1; g.fl0 = fl0; g.fl1 = fl1;
        g.table = table; g.tileL = tileL; g.tileLF = tileLF; g.pbal = pbal; g.tile_desc = tile_desc; g.tile_stride = max_tiles;
        g.cent = cent; g.box = box; g.st = st; g.barrier = barrier;
        LevelNode* lv0 = lv[0];
        LevelNode* lv1 = lv[1];
        uint32_t lv_cap = max_large, ep = epoch, max_levels = 4096;
        uint4* recs_p = recs;
        uint32_t* A_p = A;
        static const uint32_t pull_env = [] { const char* e = getenv("BVH_CUDA_T1_PULL"); return e ? (uint32_t)atoi(e) : (uint32_t)T1_PULL_DEFAULT; }();
        void* args[] = {&g, &lv0, &lv1, &lv_cap, &Q, &recs_p, &A_p, &ep, &max_levels};
        const uint32_t grid = t1_grid;
        const void* kfn = pull_env == 1 ? (const void*)k_t1_coop<1> : (pull_env >= 2 ? (const void*)k_t1_coop<2> : (const void*)k_t1_coop<0>);
        CU_CHECK(ctx, cudaLaunchCooperativeKernel(kfn, dim3(grid), dim3(T1_THREADS), args, 0, stream));
        launches += 1;
    }

    if (prof) cudaEventRecord(ctx->ev[9], stream);
    // ---- TC: one node per thread-block cluster, persistent clusters on their own task queue ----
    if (use_tc && N > (uint32_t)T2B_CAP) {
        cudaLaunchConfig_t cfg = {};
        uint32_t n_cl = (uint32_t)ctx->tc_clusters;
        if (n_cl > N / (uint32_t)T2B_CAP + NM) n_cl = N / (uint32_t)T2B_CAP + NM;  // never more clusters than possible tasks
        cfg.gridDim = dim3(n_cl * (uint32_t)ctx->tc_cluster_size);
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = TS_SMEM;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)ctx->tc_cluster_size;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (tc_mode == 2)
            CU_CHECK(ctx, cudaLaunchKernelEx(&cfg, k_tc, Q, ids0, items, table, (const float4*)cent, (const float4*)box, recs, A, tcs, st, epoch));
        else
            CU_CHECK(ctx, cudaLaunchKernelEx(&cfg, k_tcs, Q, ids0, ids1, (const float4*)cent, (const float4*)box, recs, A, st, epoch));
        launches++;
    }
    if (prof) cudaEventRecord(ctx->ev[2], stream);
    // ---- T2: persistent blocks on the device task queues ----
    {
        const int blocks = ctx->sm_count * (ctx->t2_blocks_per_sm > 0 ? ctx->t2_blocks_per_sm : 1);
        if (N > (uint32_t)T2_CAP) {
            k_t2<T2B_CAP, T2B_THREADS, true><<<ctx->sm_count, T2B_THREADS, T2B_SMEM, stream>>>(Q, ids0, ids1, cent, box, recs, A, st, epoch);
            launches++;
        }
        if (prof) cudaEventRecord(ctx->ev[7], stream);
        k_t2<T2_CAP, T2_THREADS, false><<<blocks, T2_THREADS, T2_SMEM, stream>>>(Q, ids0, ids1, cent, box, recs, A, st, epoch);
        launches++;
    }
    if (prof) cudaEventRecord(ctx->ev[6], stream);
    // ---- T2w: persistent warps on the third task queue ----
    {
        const int blocks = ctx->sm_count * (ctx->t2w_blocks_per_sm > 0 ? ctx->t2w_blocks_per_sm : 1);
        k_t2w<T2W_CAP><<<blocks, 256, 0, stream>>>(Q, ids0, cent, box, recs, A, st, epoch);
        launches++;
    }
    if (prof) cudaEventRecord(ctx->ev[3], stream);
    // ---- T3: one warp per small sub-tree ----
    if (T4_MAX < T3_MAX) {
        const int blocks = ctx->sm_count * 8;
        k_t3<<<blocks, 256, 0, stream>>>(Q, ids0, cent, box, recs, A, st);
        launches++;
    }
    if (prof) cudaEventRecord(ctx->ev[8], stream);
    // ---- T4: one thread per small sub-tree ----
    if (T4_MAX > 0) {
        if (!ctx->t4_ready) {
            CU_CHECK(ctx, cudaFuncSetAttribute(k_t4<T4_CAP, T4_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T4_SMEM));
            ctx->t4_ready = true;
        }
        k_t4<T4_CAP, T4_THREADS><<<ctx->sm_count, T4_THREADS, T4_SMEM, stream>>>(Q, t4, ids0, cent, box, recs, A, st);
        launches++;
    }
    if (prof) cudaEventRecord(ctx->ev[4], stream);
    // ---- numbering + emit ----
    k_scan_reduce<<<scan_blocks, 1024, 0, stream>>>(A, scan_n, scan_sums);
    k_scan_top<<<1, 1024, 0, stream>>>(scan_sums, scan_blocks, scan_total);
    k_scan_apply<<<scan_blocks, 1024, 0, stream>>>(A, scan_n, scan_sums);
    // per-mesh node bases (pooled bvh_index, mesh/mod.rs:322-325): exclusive scan of M_m over the meshes
    if (NM < 1024) {
        k_mesh_bases_small<<<1, 1024, 0, stream>>>(A, tbase, NM, node_base, scan_total + 1);
    } else {
        k_mesh_counts<<<(NM + 1 + 255) / 256, 256, 0, stream>>>(A, tbase, NM, node_base);
        k_scan_reduce<<<mscan_blocks, 1024, 0, stream>>>(node_base, mscan_n, mscan_sums);
        k_scan_top<<<1, 1024, 0, stream>>>(mscan_sums, mscan_blocks, scan_total + 1);
        k_scan_apply<<<mscan_blocks, 1024, 0, stream>>>(node_base, mscan_n, mscan_sums);
        launches += 3;
    }
    k_emit<<<(2 * N + 255) / 256, 256, 0, stream>>>(recs, 2 * N, A, tbase, node_base, d_nodes_out,
                                                    (uint32_t)(nodes_cap > 0xFFFFFFFFull ? 0xFFFFFFFFull : nodes_cap), st);
    if (d_mesh_info) { k_write_bvh_index<<<(NM + 255) / 256, 256, 0, stream>>>(d_mesh_info, node_base, NM); launches++; }
    // permute the caller's index buffer in place (box[] is dead by now and is reused as the staging copy)
    uint32_t* tmp = reinterpret_cast<uint32_t*>(box);
    k_permute_gather<<<(N + 255) / 256, 256, 0, stream>>>(d_indices, ids0, N, tmp);
    CU_CHECK(ctx, cudaMemcpyAsync(d_indices, tmp, sizeof(uint32_t) * 3 * (size_t)N, cudaMemcpyDeviceToDevice, stream));
    launches += 6;
    if (prof) cudaEventRecord(ctx->ev[5], stream);
    if (d_result) { k_publish_result<<<1, 32, 0, stream>>>(st, scan_total, d_result); launches++; }
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_pin, st, sizeof(BuildState), cudaMemcpyDeviceToHost, stream));
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_pin + 64, scan_total, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CU_CHECK(ctx, cudaGetLastError());
    ctx->pending = true;
    ctx->pend_n = N; ctx->pend_nm = NM; ctx->pend_launches = launches; ctx->pend_prof = prof; ctx->pend_stream = stream; ctx->pend_ids = ids0;
    if (async) return BVH_CUDA_OK;  // everything is enqueued; status, node count and statistics are collected by blas_build_finish
    return blas_build_finish(ctx, n_nodes_out);
}

// Second half of a build: waits for the stream, reads the device status back and fills the context's statistics.
int blas_build_finish(bvh_cuda_ctx* ctx, uint32_t* n_nodes_out) {
    if (!ctx->pending) return ctx_fail(ctx, BVH_CUDA_EINVAL, "blas_build_finish: no build is pending on this context");
    ctx->pending = false;
    const uint32_t N = ctx->pend_n, NM = ctx->pend_nm, launches = ctx->pend_launches;
    const bool prof = ctx->pend_prof;
    BvhCudaBuildStats stats{};
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->pend_stream));
    CU_CHECK(ctx, cudaGetLastError());
    const BuildState* hs = reinterpret_cast<const BuildState*>(ctx->h_pin);
    ctx->launches += launches;
    ctx->d_last_order = ctx->pend_ids;
    ctx->last_n = N;
    if (hs->err & DERR_BAD_INDEX) return ctx_fail(ctx, BVH_CUDA_EINVAL, "blas_build: vertex index out of range or inconsistent mesh table");
    if (hs->err & DERR_DEGENERATE)
        return ctx_fail(ctx, BVH_CUDA_EDEGENERATE, "blas_build: a node with >3 triangles has no finite-cost split (the reference would not terminate)");
    if (hs->err) return ctx_fail(ctx, BVH_CUDA_ECUDA, "blas_build: device task queue overflow or stall");
    const uint32_t interior = ctx->h_pin[64];
    stats.interior_nodes = interior;
    stats.n_nodes = ctx->h_pin[65];  // sum over meshes of 2 + 2 * interior_m
    stats.sum_interior_prims = hs->sum_interior;
    stats.grid_levels = hs->levels_done;
    stats.big_block_tasks = hs->t2b_done;
    stats.block_tasks = hs->t2_done;
    stats.warp_node_tasks = hs->t2w_done;
    stats.warp_tasks = hs->t3_count;
    stats.thread_tasks = hs->t4_count;
    stats.kernel_launches = launches;
    stats.cluster_tasks = hs->tc_done;
    stats.grid_nodes = hs->grid_nodes;
    stats.grid_interior_prims = hs->sum_grid;
    if (prof) {
        cudaEventElapsedTime(&stats.ms_setup, ctx->ev[0], ctx->ev[1]);
        cudaEventElapsedTime(&stats.ms_grid, ctx->ev[1], ctx->ev[9]);
        cudaEventElapsedTime(&stats.ms_cluster, ctx->ev[9], ctx->ev[2]);
        cudaEventElapsedTime(&stats.ms_big_block, ctx->ev[2], ctx->ev[7]);
        cudaEventElapsedTime(&stats.ms_block, ctx->ev[7], ctx->ev[6]);
        cudaEventElapsedTime(&stats.ms_warp_node, ctx->ev[6], ctx->ev[3]);
        cudaEventElapsedTime(&stats.ms_warp, ctx->ev[3], ctx->ev[8]);
        cudaEventElapsedTime(&stats.ms_thread, ctx->ev[8], ctx->ev[4]);
        cudaEventElapsedTime(&stats.ms_emit, ctx->ev[4], ctx->ev[5]);
        cudaEventElapsedTime(&stats.ms_total, ctx->ev[0], ctx->ev[5]);
    }
    ctx->stats = stats;
    if (n_nodes_out) *n_nodes_out = stats.n_nodes;
    if (hs->interior_total != interior) return ctx_fail(ctx, BVH_CUDA_ECUDA, "blas_build: internal node count mismatch");
    if (stats.n_nodes != 2 * NM + 2 * interior) return ctx_fail(ctx, BVH_CUDA_ECUDA, "blas_build: per-mesh node count mismatch");
    return BVH_CUDA_OK;
}
