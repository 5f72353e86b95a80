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
//   k_t1_coop   n > 16384      grid-wide, level-synchronous phases over tiles of 256..2048 slots chosen per level
//                              (global-memory ping-pong, one cooperative launch for all levels)
//   k_t2 (big)  2049..16384    one 1024-thread block per node from a device task queue (payload in shared memory)
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
constexpr int T2B_CAP = 16384;  // big-block tier: 2049..16384 (1024 threads, one block per SM)
constexpr int T2B_THREADS = 1024;
constexpr int T1_TILE = 2048;  // largest tile (8 slots per thread); levels that fit the grid with smaller tiles use them
constexpr int T1_THREADS = 256;
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
    uint32_t grid_nodes;            // interior nodes split by the grid tier ...
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
    Task* qb;   // big-block tasks (T2_CAP < n <= T2B_CAP)
    Task* q;    // block-per-node tasks (T2W_CAP < n <= T2_CAP)
    Task* qw;   // warp-per-node tasks  (T3_MAX < n <= T2W_CAP)
    Task* t3;   // warp-per-sub-tree tasks (T4_MAX < n <= T3_MAX)
    Task* t4;   // thread-per-sub-tree tasks (n <= T4_MAX)
    uint32_t qb_cap, q_cap, qw_cap, t3_cap, t4_cap;
};

__device__ __forceinline__ void push_t4(const Queues& Q, BuildState* st, uint32_t start, uint32_t n, uint32_t leftrun,
                                        uint32_t pstart, uint32_t pleftrun, uint32_t flags) {
    const uint32_t idx = atomicAdd(&st->t4_count, 1u);
    if (idx >= Q.t4_cap) { atomicOr(&st->err, DERR_QUEUE); return; }
    Task* d = Q.t4 + idx;
    d->start = start; d->n = n; d->leftrun = leftrun; d->pstart = pstart; d->pleftrun = pleftrun;
    d->flags = flags; d->ready = 0; d->pad = 0;
}

// ------------------------------------------------------------------------------------------------
// T3: one warp builds a whole sub-tree of <= 32 primitives.  Lane j owns slot j of the range.
// ------------------------------------------------------------------------------------------------
// Shared-memory scratch of one warp for a <=32-primitive sub-tree.
struct T3Smem {
    float (*box)[32];   // [6][32]
    float (*cent)[32];  // [3][32]
    uint32_t* gid;      // [32]
    uint8_t* tab;       // [32]
    uint16_t* pay;      // [32]
};

// One warp builds the whole sub-tree of task `t` (<= 32 primitives).  Lane j owns slot j of the range.
__device__ __forceinline__ void t3_subtree(const Task& t, const T3Smem& sm, uint32_t lane, uint32_t* ids,
                                           const float4* __restrict__ cent, const float4* __restrict__ box, uint4* recs,
                                           uint32_t* A, BuildState* st, const Queues& Q) {
    const bool nz = st->neg_zero != 0;
    float (*sm_box)[32] = sm.box;
    float (*sm_cent)[32] = sm.cent;
    uint32_t* sm_gid = sm.gid;
    uint8_t* sm_tab = sm.tab;
    uint16_t* sm_pay = sm.pay;
    __syncwarp();
    if (lane < t.n) {
        const uint32_t g = __ldcg(&ids[t.start + lane]);
        const float4 c = cent[g];
        const float4 b0 = box[2 * (size_t)g], b1 = box[2 * (size_t)g + 1];
        sm_box[0][lane] = b0.x; sm_box[1][lane] = b0.y; sm_box[2][lane] = b0.z;
        sm_box[3][lane] = b1.x; sm_box[4][lane] = b1.y; sm_box[5][lane] = b1.z;
        sm_cent[0][lane] = c.x; sm_cent[1][lane] = c.y; sm_cent[2][lane] = c.z;
        sm_gid[lane] = g;
    }
    __syncwarp();
    uint32_t pay = lane;  // bits 0-4: local primitive, 5-13: plane counts of the current node
    uint32_t s = 0, n = t.n, leftrun = t.leftrun, pstart = t.pstart, pleftrun = t.pleftrun, fl = t.flags;
    uint32_t stk_a = 0, stk_b = 0, stk_c = 0;  // lane i holds stack entry i
    int sp = 0;

    for (;;) {
        const bool active = lane >= s && lane < s + n;
        const uint32_t e = pay & 31u;
        const uint32_t abs_start = t.start + s;
        if (T4_MAX > 0 && T4_MAX < T3_MAX && (int)n <= T4_MAX) {
            // hand the whole child sub-tree to the thread-per-sub-tree kernel (it runs after this one and reads the
            // range in the order this warp writes back at the end; nothing below touches these slots again)
            if (lane == 0) push_t4(Q, st, abs_start, n, leftrun, pstart, pleftrun, fl);
            if (sp == 0) break;
            sp--;
            const uint32_t a2 = __shfl_sync(FULL_MASK, stk_a, sp);
            pstart = __shfl_sync(FULL_MASK, stk_b, sp);
            pleftrun = __shfl_sync(FULL_MASK, stk_c, sp);
            s = a2 & 0xFFu; n = a2 >> 8; leftrun = 0; fl = TF_RIGHT | (t.flags & ~3u);
            continue;
        }
        // own vertex box (blas.rs:87-88,117-123), folded from +-1e30
        float lo[3], hi[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            uint32_t mn = active ? f2o(sm_box[c][e]) : ENC_POS_INIT;
            uint32_t mx = active ? f2o(sm_box[3 + c][e]) : ENC_NEG_INIT;
            mn = min(__reduce_min_sync(FULL_MASK, mn), ENC_POS_INIT);
            mx = max(__reduce_max_sync(FULL_MASK, mx), ENC_NEG_INIT);
            lo[c] = o2f(mn);
            hi[c] = o2f(mx);
        }
        if (nz) {
            // rare path (-0.0 in the input): a zero face takes the sign of the first zero in slot order, as the
            // reference's sequential fold does (lane order == slot order here)
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const float cur = (c < 3) ? lo[c] : hi[c - 3];
                if (cur == 0.0f) {
                    const float mine = sm_box[c][e];
                    const uint32_t p = __reduce_min_sync(FULL_MASK, (active && mine == 0.0f) ? lane : 32u);
                    const float z = __shfl_sync(FULL_MASK, mine, p & 31u);
                    if (c < 3) lo[c] = z; else hi[c - 3] = z;
                }
            }
        }
        bool descend = false;
        if (n <= 3) {  // leaf (blas.rs:106-109)
            if (lane == 0) emit_rec(recs, 2 * abs_start, lo, hi, abs_start, n, leftrun, pstart, pleftrun, fl);
        } else {
            // centroid bounds (blas.rs:142)
            float cmin[3], cmax[3], cc[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                cc[c] = active ? sm_cent[c][e] : 0.0f;
                uint32_t mn = active ? f2o(cc[c]) : ENC_POS_INIT;
                uint32_t mx = active ? f2o(cc[c]) : ENC_NEG_INIT;
                mn = min(__reduce_min_sync(FULL_MASK, mn), ENC_POS_INIT);
                mx = max(__reduce_max_sync(FULL_MASK, mx), ENC_NEG_INIT);
                cmin[c] = o2f(mn);
                cmax[c] = o2f(mx);
            }
            pay = e | (plane_counts(cc[0], cc[1], cc[2], cmin, cmax) << 5);

            const uint32_t j = lane - s;
            const uint32_t nmask = (n >= 32) ? 0xFFFFFFFFu : ((1u << n) - 1u);
            // closed form of partition_shuffle (blas.rs:168-182) on the current order
            auto do_shuffle = [&](uint32_t a, uint32_t b) -> uint32_t {
                const bool L = active && (((pay >> (5 + 3 * a)) & 7u) < b);
                const uint32_t Lm = __ballot_sync(FULL_MASK, L) >> s;
                const uint32_t Rm = ~Lm & nmask;
                const uint32_t below = active ? ((1u << j) - 1u) : 0u;
                const uint32_t RF = __popc(Rm & below), LF = j - RF;
                const uint32_t nL = __popc(Lm);
                const uint32_t LBB = (j + 2 < 32) ? __popc(Lm >> (j + 2)) : 0u;
                const bool pred = active && (j + 2 <= n) && (LBB >= RF);
                const uint32_t f = __popc(__ballot_sync(FULL_MASK, pred));
                const uint32_t pivot = nL - ((Lm >> f) & 1u);
                const uint32_t LB = nL - LF - (L ? 1u : 0u);
                if (active) {
                    if (L) sm_tab[n - 1 - LB] = (uint8_t)j;
                    else sm_tab[RF] = (uint8_t)j;
                }
                __syncwarp();
                if (active) {
                    uint32_t dest;
                    if (j < f) dest = L ? j : (RF == 0 ? n - 1 : (uint32_t)sm_tab[n - RF] - 1u);
                    else if (j == f) dest = pivot;
                    else dest = L ? (uint32_t)sm_tab[LB] : j - 1;
                    sm_pay[dest] = (uint16_t)pay;
                }
                __syncwarp();
                if (active) pay = sm_pay[j];
                return pivot;
            };

            uint32_t my_u = 0xFFu, my_piv = 0;
            for (uint32_t c = 0; c < 21; ++c) {
                const uint32_t pivot = do_shuffle(c / 7, c % 7 + 1);
                const uint32_t up = __shfl_sync(FULL_MASK, pay, s + pivot);
                if (lane == c) { my_u = up & 31u; my_piv = pivot; }
            }
            // candidate `lane` (< 21): exact boxes of {L}\{u} and {R}+{u} (blas.rs:149-155)
            const uint32_t ca = (lane < 21) ? lane / 7 : 0, cb = lane % 7 + 1;
            float Lb[6] = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
            float Rb[6] = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
            for (uint32_t tt = 0; tt < n; ++tt) {
                const uint32_t p = __shfl_sync(FULL_MASK, pay, s + tt);
                const uint32_t et = p & 31u;
                const bool left = (((p >> (5 + 3 * ca)) & 7u) < cb) && (et != my_u);
                const float x0 = sm_box[0][et], x1 = sm_box[1][et], x2 = sm_box[2][et];
                const float x3 = sm_box[3][et], x4 = sm_box[4][et], x5 = sm_box[5][et];
                if (left) {
                    Lb[0] = fminf(Lb[0], x0); Lb[1] = fminf(Lb[1], x1); Lb[2] = fminf(Lb[2], x2);
                    Lb[3] = fmaxf(Lb[3], x3); Lb[4] = fmaxf(Lb[4], x4); Lb[5] = fmaxf(Lb[5], x5);
                } else {
                    Rb[0] = fminf(Rb[0], x0); Rb[1] = fminf(Rb[1], x1); Rb[2] = fminf(Rb[2], x2);
                    Rb[3] = fmaxf(Rb[3], x3); Rb[4] = fmaxf(Rb[4], x4); Rb[5] = fmaxf(Rb[5], x5);
                }
            }
            const float cost = sah_cost(Lb, Rb, my_piv, n - my_piv);
            // strict <, first candidate wins, NaN/inf never win (blas.rs:140,156)
            const uint32_t key = (lane < 21 && cost < 3.402823466e+38f) ? __float_as_uint(cost) : 0xFFFFFFFFu;
            const uint32_t mk = __reduce_min_sync(FULL_MASK, key);
            if (mk == 0xFFFFFFFFu) {
                if (lane == 0) atomicOr(&st->err, DERR_DEGENERATE);
            } else {
                const uint32_t win = __ffs(__ballot_sync(FULL_MASK, key == mk)) - 1;
                const uint32_t p = __shfl_sync(FULL_MASK, my_piv, win);  // recorded pivot (blas.rs:159,165)
                do_shuffle(win / 7, win % 7 + 1);                       // blas.rs:164
                if (lane == 0) {
                    emit_rec(recs, 2 * (abs_start + p) + 1, lo, hi, abs_start, n, leftrun, pstart, pleftrun, fl);
                    if (p <= 3) A[abs_start] = leftrun + 1;
                }
                if ((int)lane == sp) {
                    stk_a = (s + p) | ((n - p) << 8);
                    stk_b = abs_start;
                    stk_c = leftrun;
                }
                sp++;
                pstart = abs_start; pleftrun = leftrun; leftrun = leftrun + 1; n = p; fl = t.flags & ~3u;
                descend = true;
            }
        }
        if (descend) continue;
        if (sp == 0) break;
        sp--;
        const uint32_t a = __shfl_sync(FULL_MASK, stk_a, sp);
        pstart = __shfl_sync(FULL_MASK, stk_b, sp);
        pleftrun = __shfl_sync(FULL_MASK, stk_c, sp);
        s = a & 0xFFu; n = a >> 8; leftrun = 0; fl = TF_RIGHT | (t.flags & ~3u);
    }
    if (lane < t.n) ids[t.start + lane] = sm_gid[pay & 31u];
}

__global__ void __launch_bounds__(256, 4) k_t3(Queues Q, uint32_t* ids,
                                               const float4* __restrict__ cent, const float4* __restrict__ box,
                                               uint4* recs, uint32_t* A, BuildState* st) {
    const Task* __restrict__ tasks = Q.t3;
    __shared__ float s_box[8][6][32];
    __shared__ float s_cent[8][3][32];
    __shared__ uint32_t s_gid[8][32];
    __shared__ uint8_t s_tab[8][32];
    __shared__ uint16_t s_pay[8][32];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t n_tasks = st->t3_count;
    const T3Smem sm{s_box[w], s_cent[w], s_gid[w], s_tab[w], s_pay[w]};
    for (uint32_t ti = blockIdx.x * 8 + w; ti < n_tasks; ti += gridDim.x * 8) {
        const Task t = tasks[ti];
        t3_subtree(t, sm, lane, ids, cent, box, recs, A, st, Q);
    }
}

// ------------------------------------------------------------------------------------------------
// T4: one THREAD per sub-tree of <= CAP primitives (t4_seq.cuh).  Working set in shared memory as [word][thread].
// ------------------------------------------------------------------------------------------------
template <int CAP, int BD>
__global__ void __launch_bounds__(BD, 1) k_t4(Queues Q, const Task* __restrict__ tasks, uint32_t* ids, const float4* __restrict__ cent,
                                             const float4* __restrict__ box, uint4* recs, uint32_t* A, BuildState* st) {
    extern __shared__ uint32_t s_t4[];
    const T4Mem<CAP> m{reinterpret_cast<float*>(s_t4) + threadIdx.x, s_t4 + 6 * CAP * BD + threadIdx.x, (uint32_t)BD};
    const uint32_t n_tasks = min(st->t4_count, Q.t4_cap);
    for (uint32_t ti = blockIdx.x * BD + threadIdx.x; ti < n_tasks; ti += gridDim.x * BD) {
        const Task tk = tasks[ti];
        const T4Task t{tk.start, tk.n, tk.leftrun, tk.pstart, tk.pleftrun, tk.flags};
        for (uint32_t j = 0; j < t.n; ++j) {
            const uint32_t g = __ldcg(&ids[t.start + j]);
            const float4 b0 = box[2 * (size_t)g], b1 = box[2 * (size_t)g + 1];
            m.gid(j) = g;
            m.box(0, j) = b0.x; m.box(1, j) = b0.y; m.box(2, j) = b0.z;
            m.box(3, j) = b1.x; m.box(4, j) = b1.y; m.box(5, j) = b1.z;
        }
        const uint32_t err = t4_core<CAP>(t, m, reinterpret_cast<const T4Cent*>(cent), ids, reinterpret_cast<T4Rec*>(recs), A);
        if (err) atomicOr(&st->err, DERR_DEGENERATE);
    }
}

// ------------------------------------------------------------------------------------------------
// T2: one block per node (33..CAP primitives), tasks from a device queue; children go back to the
// queue (> 32) or to the T3 list.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void push_child(const Queues& Q, BuildState* st, uint32_t epoch, uint32_t start, uint32_t n,
                                           uint32_t leftrun, uint32_t pstart, uint32_t pleftrun, uint32_t flags) {
    if (n > T3_MAX) {
        const int tier = n > T2_CAP ? 2 : (n > T2W_CAP ? 1 : 0);
        uint32_t* pending = tier == 2 ? &st->b_pending : (tier == 1 ? &st->q_pending : &st->w_pending);
        uint32_t* tail = tier == 2 ? &st->b_tail : (tier == 1 ? &st->q_tail : &st->w_tail);
        const uint32_t cap = tier == 2 ? Q.qb_cap : (tier == 1 ? Q.q_cap : Q.qw_cap);
        atomicAdd(pending, 1u);
        const uint32_t idx = atomicAdd(tail, 1u);
        if (idx >= cap) {
            atomicOr(&st->err, DERR_QUEUE);
            atomicSub(pending, 1u);
            return;
        }
        Task* d = (tier == 2 ? Q.qb : (tier == 1 ? Q.q : Q.qw)) + idx;
        d->start = start; d->n = n; d->leftrun = leftrun; d->pstart = pstart; d->pleftrun = pleftrun;
        d->flags = flags; d->pad = 0;
        __threadfence();
        *(volatile uint32_t*)&d->ready = epoch;
    } else if ((int)n <= T4_MAX) {
        push_t4(Q, st, start, n, leftrun, pstart, pleftrun, flags);
    } else {
        const uint32_t idx = atomicAdd(&st->t3_count, 1u);
        if (idx >= Q.t3_cap) { atomicOr(&st->err, DERR_QUEUE); return; }
        Task* d = Q.t3 + idx;
        d->start = start; d->n = n; d->leftrun = leftrun; d->pstart = pstart; d->pleftrun = pleftrun;
        d->flags = flags; d->ready = epoch; d->pad = 0;
    }
}

// Pop one task from a device queue (called by one thread per consumer).  Ticket scheme: every consumer takes
// the next slot number with one atomicAdd (no CAS retries under contention) and then waits for that slot to be
// published, or for the queue to drain: `pending` counts tasks pushed but not yet finished, and a finished task
// has already pushed its children, so pending == 0 with the ticket still unpublished means no task will ever
// land in it.  Returns false when the queue has drained.
__device__ __forceinline__ bool queue_pop(Task* q, uint32_t cap, uint32_t* head, uint32_t* tail, uint32_t* pending,
                                          BuildState* st, uint32_t epoch, uint32_t* out_idx) {
    (void)tail;
    const uint32_t idx = atomicAdd(head, 1u);
    *out_idx = idx;
    if (idx >= cap) return false;
    uint32_t ns = 32;
    for (uint32_t spins = 0;; ++spins) {
        if (ld_vol(&q[idx].ready) == epoch) { __threadfence(); return true; }
        if (ld_vol(pending) == 0) {
            // re-check: the producer publishes the slot before it decrements `pending`
            if (ld_vol(&q[idx].ready) == epoch) { __threadfence(); return true; }
            return false;
        }
        __nanosleep(ns);
        if (ns < 1024) ns <<= 1;
        if (spins > SPIN_LIMIT) { atomicOr(&st->err, DERR_QUEUE); return false; }
    }
}

template <int CAP, int THREADS, bool BIG>
__global__ void __launch_bounds__(THREADS, BIG ? 1 : T2_MIN_BLOCKS) k_t2(Queues Q, uint32_t* ids, uint32_t* ids_snap,
                                                const float4* __restrict__ cent, const float4* __restrict__ box,
                                                uint4* recs, uint32_t* A, BuildState* st, uint32_t epoch) {
    constexpr int NW = THREADS / 32;
    constexpr int EPT = CAP / THREADS;
    Task* const q = BIG ? Q.qb : Q.q;
    const uint32_t q_cap = BIG ? Q.qb_cap : Q.q_cap;
    uint32_t* const q_head = BIG ? &st->b_head : &st->q_head;
    uint32_t* const q_tail = BIG ? &st->b_tail : &st->q_tail;
    uint32_t* const q_pending = BIG ? &st->b_pending : &st->q_pending;
    // dynamic shared memory: payload ping-pong (bits 0-15 local primitive, 16-24 plane counts, 31 special) and the
    // rank -> position table.  The local-primitive -> triangle-id map lives in global memory (ids_snap).
    extern __shared__ uint32_t s_dyn[];
    uint32_t* const s_pay0 = s_dyn;
    uint32_t* const s_pay1 = s_dyn + CAP;
    uint16_t* const s_tab = reinterpret_cast<uint16_t*>(s_dyn + 2 * CAP);
    __shared__ uint32_t s_wtot[NW];
    __shared__ uint32_t s_red[NW][12];
    __shared__ uint32_t s_node[12];
    __shared__ uint32_t s_bins[3][8][6];
    __shared__ uint32_t s_u[21], s_piv[21], s_uk[21];
    __shared__ float s_ubox[21][6];
    __shared__ Task s_task;
    __shared__ int s_have;
    __shared__ uint32_t s_best;
    __shared__ uint32_t s_zpos[6];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;

    for (;;) {
        // ---- pop ----
        if (tid == 0) {
            uint32_t idx = 0;
            const bool have = queue_pop(q, q_cap, q_head, q_tail, q_pending, st, epoch, &idx);
            if (have) {
                const volatile Task* vq = q + idx;
                s_task.start = vq->start; s_task.n = vq->n; s_task.leftrun = vq->leftrun;
                s_task.pstart = vq->pstart; s_task.pleftrun = vq->pleftrun; s_task.flags = vq->flags;
            }
            s_have = have ? 1 : 0;
        }
        __syncthreads();
        if (!s_have) break;
        const Task t = s_task;
        const uint32_t n = t.n, start = t.start;
        // balanced layout: every warp owns E*32 consecutive slots, E = ceil(n / THREADS) <= EPT
        const uint32_t E = (n + THREADS - 1) / THREADS;
        const uint32_t CHUNK = 32 * E;

        // ---- 1. snapshot the order, own vertex box, centroid bounds ----
        {
            float acc[12] = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f, 1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll 4
            for (int i = 0; i < EPT; ++i) {
                const uint32_t j = warp * CHUNK + i * 32 + lane;
                if (i < (int)E && j < n) {
                    const uint32_t g = __ldcg(&ids[start + j]);
                    ids_snap[start + j] = g;
                    const float4 c = cent[g];
                    const float4 b0 = box[2 * (size_t)g], b1 = box[2 * (size_t)g + 1];
                    acc[0] = fminf(acc[0], b0.x); acc[1] = fminf(acc[1], b0.y); acc[2] = fminf(acc[2], b0.z);
                    acc[3] = fmaxf(acc[3], b1.x); acc[4] = fmaxf(acc[4], b1.y); acc[5] = fmaxf(acc[5], b1.z);
                    acc[6] = fminf(acc[6], c.x); acc[7] = fminf(acc[7], c.y); acc[8] = fminf(acc[8], c.z);
                    acc[9] = fmaxf(acc[9], c.x); acc[10] = fmaxf(acc[10], c.y); acc[11] = fmaxf(acc[11], c.z);
                }
            }
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const bool is_min = (k < 3) || (k >= 6 && k < 9);
                const uint32_t v = f2o(acc[k]);
                const uint32_t r = is_min ? __reduce_min_sync(FULL_MASK, v) : __reduce_max_sync(FULL_MASK, v);
                if (lane == 0) s_red[warp][k] = r;
            }
        }
        __syncthreads();
        if (tid < 12) {
            const bool is_min = (tid < 3) || (tid >= 6 && tid < 9);
            uint32_t r = s_red[0][tid];
            for (int w2 = 1; w2 < NW; ++w2) r = is_min ? min(r, s_red[w2][tid]) : max(r, s_red[w2][tid]);
            s_node[tid] = r;
        }
        for (uint32_t k = tid; k < 144; k += THREADS) (&s_bins[0][0][0])[k] = ((k % 6) < 3) ? ENC_POS_INIT : ENC_NEG_INIT;
        if (tid < 6) s_zpos[tid] = 0xFFFFFFFFu;
        __syncthreads();
        if (st->neg_zero) {
            // rare path (-0.0 in the input): remember, per zero-valued face, the first slot that holds a zero there
            bool zero_face[6];
            bool any_zero = false;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                zero_face[c] = o2f(c < 3 ? min(s_node[c], ENC_POS_INIT) : max(s_node[c], ENC_NEG_INIT)) == 0.0f;
                any_zero = any_zero || zero_face[c];
            }
            if (any_zero) {
                for (uint32_t i = 0; i < E; ++i) {
                    const uint32_t j = warp * CHUNK + i * 32 + lane;
                    if (j < n) {
                        const uint32_t g = __ldcg(&ids_snap[start + j]);
                        const float4 b0 = box[2 * (size_t)g], b1 = box[2 * (size_t)g + 1];
                        const float vals[6] = {b0.x, b0.y, b0.z, b1.x, b1.y, b1.z};
#pragma unroll
                        for (int c = 0; c < 6; ++c)
                            if (zero_face[c] && vals[c] == 0.0f) atomicMin(&s_zpos[c], j);
                    }
                }
            }
            __syncthreads();
        }

        // ---- 2. plane counts ----
        {
            float cmin[3], cmax[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) { cmin[c] = o2f(s_node[6 + c]); cmax[c] = o2f(s_node[9 + c]); }
#pragma unroll 4
            for (int i = 0; i < EPT; ++i) {
                const uint32_t j = warp * CHUNK + i * 32 + lane;
                if (i < (int)E && j < n) {
                    const float4 c = cent[__ldcg(&ids_snap[start + j])];
                    s_pay0[j] = j | (plane_counts(c.x, c.y, c.z, cmin, cmax) << 16);
                }
            }
        }
        __syncthreads();

        // ---- 3. shuffles ----
        auto shuffle = [&](int cur, uint32_t a, uint32_t b, int cidx) {
            const uint32_t* pin = cur ? s_pay1 : s_pay0;
            uint32_t* pout = cur ? s_pay0 : s_pay1;
            const uint32_t sh = 16 + 3 * a;
            uint32_t bal[EPT];
            uint32_t cnt = 0;
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                bal[i] = 0;
                if (i < (int)E) {
                    const uint32_t j = warp * CHUNK + i * 32 + lane;
                    const bool L = (j < n) && (((pin[j] >> sh) & 7u) < b);
                    bal[i] = __ballot_sync(FULL_MASK, L);
                    cnt += __popc(bal[i]);
                }
            }
            if (lane == 0) s_wtot[warp] = cnt;
            __syncthreads();  // S1
            // lane w2 reads warp w2's count: total and the sum over the warps before this one by two warp reductions
            // (a serial walk over 32 counts was a third of a big-block shuffle of a small node)
            const uint32_t wv = (lane < (uint32_t)NW) ? s_wtot[lane] : 0u;
            const uint32_t nL = __reduce_add_sync(FULL_MASK, wv);
            const uint32_t wpre = __reduce_add_sync(FULL_MASK, lane < warp ? wv : 0u);
            // boundary element in closed form (see p_t1_table): f is nL-1, nL or nL+1, decided by three flags
            uint32_t f, Lf;
            {
                const uint32_t l0 = nL ? ((((pin[nL - 1] >> sh) & 7u) < b) ? 1u : 0u) : 0u;
                const uint32_t l1 = (nL < n && (((pin[nL < n ? nL : 0] >> sh) & 7u) < b)) ? 1u : 0u;
                const uint32_t l2 = (nL + 1 < n && (((pin[nL + 1 < n ? nL + 1 : 0] >> sh) & 7u) < b)) ? 1u : 0u;
                if (nL >= 1 && !(nL + 1 <= n && l0 + l1 <= 1)) { f = nL - 1; Lf = l0; }
                else if (!(nL + 2 <= n && l1 + l2 == 0)) { f = nL; Lf = l1; }
                else { f = nL + 1; Lf = l2; }
            }
            const uint32_t pivot = nL - Lf;
            uint32_t running = wpre;
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                if (i >= (int)E) break;
                const uint32_t j = warp * CHUNK + i * 32 + lane;
                const uint32_t LF = running + __popc(bal[i] & lt_mask);
                if (j < n) {
                    // only front R's (j < f) and back L's (j > f) are looked up
                    if ((bal[i] >> lane) & 1u) { if (j >= nL) s_tab[n - 1 - (nL - LF - 1)] = (uint16_t)j; }
                    else if (j <= nL) s_tab[j - LF] = (uint16_t)j;
                }
                running += __popc(bal[i]);
            }
            __syncthreads();  // S2
            running = wpre;
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                if (i >= (int)E) break;
                const uint32_t j = warp * CHUNK + i * 32 + lane;
                const uint32_t LF = running + __popc(bal[i] & lt_mask);
                running += __popc(bal[i]);
                if (j < n) {
                    uint32_t pay = pin[j];
                    const uint32_t Lbit = (bal[i] >> lane) & 1u;
                    const uint32_t RF = j - LF;
                    uint32_t dest;
                    if (j < f) dest = Lbit ? j : (RF == 0 ? n - 1 : (uint32_t)s_tab[n - RF] - 1u);
                    else if (j == f) {
                        dest = pivot;
                        pay |= 0x80000000u;
                        if (cidx >= 0) { s_u[cidx] = pay & 0xFFFFu; s_uk[cidx] = (pay >> 16) & 0x1FFu; s_piv[cidx] = pivot; }
                    } else dest = Lbit ? (uint32_t)s_tab[nL - LF - 1] : j - 1;
                    pout[dest] = pay;
                }
            }
            __syncthreads();  // S3
        };

        int cur = 0;
        for (uint32_t c = 0; c < 21; ++c) { shuffle(cur, c / 7, c % 7 + 1, (int)c); cur ^= 1; }

        // ---- 4. exact bins over the non-special primitives (4 slots per thread at a time) ----
        // Boxes are mapped to ordered uints once per slot, so the 24 per-bin reductions below are integer min / max
        // feeding redux.sync directly.
        {
            const uint32_t* pin = cur ? s_pay1 : s_pay0;
            for (uint32_t i0 = 0; i0 < E; i0 += 4) {
                uint32_t lo[4][3], hi[4][3];
                uint32_t kk[4];
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    const uint32_t i = i0 + ii;
                    const uint32_t j = warp * CHUNK + i * 32 + lane;
                    kk[ii] = 0xFFFFFFFFu;
                    lo[ii][0] = lo[ii][1] = lo[ii][2] = ENC_POS_INIT;
                    hi[ii][0] = hi[ii][1] = hi[ii][2] = ENC_NEG_INIT;
                    if (i < E && j < n) {
                        const uint32_t pay = pin[j];
                        if (!(pay & 0x80000000u)) {
                            const uint32_t g = __ldcg(&ids_snap[start + (pay & 0xFFFFu)]);
                            const float4 b0 = box[2 * (size_t)g], b1 = box[2 * (size_t)g + 1];
                            lo[ii][0] = f2o(b0.x); lo[ii][1] = f2o(b0.y); lo[ii][2] = f2o(b0.z);
                            hi[ii][0] = f2o(b1.x); hi[ii][1] = f2o(b1.y); hi[ii][2] = f2o(b1.z);
                            kk[ii] = (pay >> 16) & 0x1FFu;
                        }
                    }
                }
                for (uint32_t a = 0; a < 3; ++a) {
                    for (uint32_t k = 0; k < 8; ++k) {
                        uint32_t m[6] = {ENC_POS_INIT, ENC_POS_INIT, ENC_POS_INIT, ENC_NEG_INIT, ENC_NEG_INIT, ENC_NEG_INIT};
                        bool any = false;
#pragma unroll
                        for (int ii = 0; ii < 4; ++ii) {
                            const bool in = (kk[ii] != 0xFFFFFFFFu) && (((kk[ii] >> (3 * a)) & 7u) == k);
                            if (in) {
                                any = true;
                                m[0] = min(m[0], lo[ii][0]); m[1] = min(m[1], lo[ii][1]); m[2] = min(m[2], lo[ii][2]);
                                m[3] = max(m[3], hi[ii][0]); m[4] = max(m[4], hi[ii][1]); m[5] = max(m[5], hi[ii][2]);
                            }
                        }
                        if (!__any_sync(FULL_MASK, any)) continue;
#pragma unroll
                        for (int c = 0; c < 6; ++c) {
                            const uint32_t r = (c < 3) ? __reduce_min_sync(FULL_MASK, m[c]) : __reduce_max_sync(FULL_MASK, m[c]);
                            if (lane == 0) {
                                if (c < 3) atomicMin(&s_bins[a][k][c], r);
                                else atomicMax(&s_bins[a][k][c], r);
                            }
                        }
                    }
                }
            }
        }
        if (tid < 21) {
            const uint32_t g = __ldcg(&ids_snap[start + s_u[tid]]);
            const float4 b0 = box[2 * (size_t)g], b1 = box[2 * (size_t)g + 1];
            s_ubox[tid][0] = b0.x; s_ubox[tid][1] = b0.y; s_ubox[tid][2] = b0.z;
            s_ubox[tid][3] = b1.x; s_ubox[tid][4] = b1.y; s_ubox[tid][5] = b1.z;
        }
        __syncthreads();

        // ---- 5. candidate costs and selection (warp 0) ----
        if (warp == 0) {
            const uint32_t c = lane;
            const uint32_t a = (c < 21) ? c / 7 : 0, b = c % 7 + 1;
            float Lb[6] = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
            float Rb[6] = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
            for (uint32_t k = 0; k < 8; ++k) {
                float* side = (k < b) ? Lb : Rb;
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    side[x] = fminf(side[x], o2f(s_bins[a][k][x]));
                    side[3 + x] = fmaxf(side[3 + x], o2f(s_bins[a][k][3 + x]));
                }
            }
            const uint32_t myu = s_u[(c < 21) ? c : 0];
            for (uint32_t s2 = 0; s2 < 21; ++s2) {
                const bool left = (s_u[s2] != myu) && (((s_uk[s2] >> (3 * a)) & 7u) < b);
                float* side = left ? Lb : Rb;
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    side[x] = fminf(side[x], s_ubox[s2][x]);
                    side[3 + x] = fmaxf(side[3 + x], s_ubox[s2][3 + x]);
                }
            }
            const uint32_t n1 = s_piv[(c < 21) ? c : 0];
            const float cost = sah_cost(Lb, Rb, n1, n - n1);
            const uint32_t key = (c < 21 && cost < 3.402823466e+38f) ? __float_as_uint(cost) : 0xFFFFFFFFu;
            const uint32_t mk = __reduce_min_sync(FULL_MASK, key);
            const uint32_t win = (mk == 0xFFFFFFFFu) ? 0xFFFFFFFFu : (uint32_t)(__ffs(__ballot_sync(FULL_MASK, key == mk)) - 1);
            if (lane == 0) s_best = win;
        }
        __syncthreads();
        const uint32_t best = s_best;
        if (best == 0xFFFFFFFFu) {
            if (tid == 0) {
                atomicOr(&st->err, DERR_DEGENERATE);
                atomicSub(q_pending, 1u);
            }
            __syncthreads();
            continue;
        }
        // ---- 6. final shuffle (blas.rs:164), write the order back ----
        shuffle(cur, best / 7, best % 7 + 1, -1);
        cur ^= 1;
        {
            const uint32_t* pin = cur ? s_pay1 : s_pay0;
#pragma unroll 4
            for (int i = 0; i < EPT; ++i) {
                const uint32_t j = warp * CHUNK + i * 32 + lane;
                if (i < (int)E && j < n) ids[start + j] = __ldcg(&ids_snap[start + (pin[j] & 0xFFFFu)]);
            }
        }
        __threadfence();
        __syncthreads();
        // ---- 7. record + children ----
        if (tid == 0) {
            const uint32_t p = s_piv[best];
            float lo[3], hi[3];
            for (int c = 0; c < 3; ++c) {
                lo[c] = o2f(min(s_node[c], ENC_POS_INIT));
                hi[c] = o2f(max(s_node[3 + c], ENC_NEG_INIT));
            }
            for (int c = 0; c < 6; ++c)
                if (s_zpos[c] != 0xFFFFFFFFu) {  // only set on the rare -0.0 path
                    const uint32_t g = __ldcg(&ids_snap[start + s_zpos[c]]);
                    const float4 bb = box[2 * (size_t)g + (c < 3 ? 0 : 1)];
                    const float z = (c % 3 == 0) ? bb.x : ((c % 3 == 1) ? bb.y : bb.z);
                    if (c < 3) lo[c] = z; else hi[c - 3] = z;
                }
            emit_rec(recs, 2 * (start + p) + 1, lo, hi, start, n, t.leftrun, t.pstart, t.pleftrun, t.flags);
            if (p <= 3) A[start] = t.leftrun + 1;
            push_child(Q, st, epoch, start, p, t.leftrun + 1, start, t.leftrun, t.flags & ~3u);
            push_child(Q, st, epoch, start + p, n - p, 0, start, t.leftrun, TF_RIGHT | (t.flags & ~3u));
            atomicAdd(BIG ? &st->t2b_done : &st->t2_done, 1u);
            __threadfence();
            atomicSub(q_pending, 1u);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// T2w: one WARP per node (33..WCAP primitives), tasks from a second device queue.  Same algorithm as k_t2,
// but warp-synchronous: ballots and popcounts replace the block scan, bins and specials live in registers
// (lane a*8+k owns bin (a,k); lane c owns candidate c and special c).  No block barriers.
// ------------------------------------------------------------------------------------------------
#ifndef T2W_MIN_BLOCKS
#define T2W_MIN_BLOCKS 3
#endif
template <int WCAP>
__global__ void __launch_bounds__(256, T2W_MIN_BLOCKS) k_t2w(Queues Q, uint32_t* ids, const float4* __restrict__ cent,
                                             const float4* __restrict__ box, uint4* recs, uint32_t* A, BuildState* st,
                                             uint32_t epoch) {
    constexpr int EPL = WCAP / 32;
    constexpr int NWB = 8;
    __shared__ uint32_t s_pay[NWB][2][WCAP];  // bits 0-15 local primitive, 16-24 plane counts, 31 special
    __shared__ uint32_t s_gid[NWB][WCAP];
    __shared__ uint16_t s_tab[NWB][WCAP];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;

    for (;;) {
        // ---- pop (lane 0) ----
        uint32_t have = 0, t_start = 0, t_n = 0, t_leftrun = 0, t_pstart = 0, t_pleftrun = 0, t_flags = 0;
        if (lane == 0) {
            uint32_t idx = 0;
            if (queue_pop(Q.qw, Q.qw_cap, &st->w_head, &st->w_tail, &st->w_pending, st, epoch, &idx)) {
                const volatile Task* vq = Q.qw + idx;
                t_start = vq->start; t_n = vq->n; t_leftrun = vq->leftrun; t_pstart = vq->pstart;
                t_pleftrun = vq->pleftrun; t_flags = vq->flags;
                have = 1;
            }
        }
        have = __shfl_sync(FULL_MASK, have, 0);
        if (!have) break;
        const uint32_t start = __shfl_sync(FULL_MASK, t_start, 0), n = __shfl_sync(FULL_MASK, t_n, 0);
        const uint32_t leftrun = __shfl_sync(FULL_MASK, t_leftrun, 0), pstart = __shfl_sync(FULL_MASK, t_pstart, 0);
        const uint32_t pleftrun = __shfl_sync(FULL_MASK, t_pleftrun, 0), tflags = __shfl_sync(FULL_MASK, t_flags, 0);
        const uint32_t E = (n + 31) >> 5;  // chunks in use, <= EPL

        // ---- 1. load, own vertex box, centroid bounds ----
        float ccx[EPL], ccy[EPL], ccz[EPL];
        float nlo[3], nhi[3], cmin[3], cmax[3];
        {
            float acc[12] = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f, 1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
            for (int i = 0; i < EPL; ++i) {
                const uint32_t j = i * 32 + lane;
                ccx[i] = ccy[i] = ccz[i] = 0.0f;
                if (i < (int)E && j < n) {
                    const uint32_t g = __ldcg(&ids[start + j]);
                    s_gid[w][j] = g;
                    const float4 c = cent[g];
                    const float4 b0 = box[2 * (size_t)g], b1 = box[2 * (size_t)g + 1];
                    ccx[i] = c.x; ccy[i] = c.y; ccz[i] = c.z;
                    acc[0] = fminf(acc[0], b0.x); acc[1] = fminf(acc[1], b0.y); acc[2] = fminf(acc[2], b0.z);
                    acc[3] = fmaxf(acc[3], b1.x); acc[4] = fmaxf(acc[4], b1.y); acc[5] = fmaxf(acc[5], b1.z);
                    acc[6] = fminf(acc[6], c.x); acc[7] = fminf(acc[7], c.y); acc[8] = fminf(acc[8], c.z);
                    acc[9] = fmaxf(acc[9], c.x); acc[10] = fmaxf(acc[10], c.y); acc[11] = fmaxf(acc[11], c.z);
                }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                nlo[k] = o2f(min(__reduce_min_sync(FULL_MASK, f2o(acc[k])), ENC_POS_INIT));
                nhi[k] = o2f(max(__reduce_max_sync(FULL_MASK, f2o(acc[3 + k])), ENC_NEG_INIT));
                cmin[k] = o2f(min(__reduce_min_sync(FULL_MASK, f2o(acc[6 + k])), ENC_POS_INIT));
                cmax[k] = o2f(max(__reduce_max_sync(FULL_MASK, f2o(acc[9 + k])), ENC_NEG_INIT));
            }
        }
        if (st->neg_zero) {
            // rare path (-0.0 in the input): sign of a zero face = first zero in slot order (see k_setup)
            __syncwarp();
            for (int c = 0; c < 6; ++c) {
                const float cur = (c < 3) ? nlo[c] : nhi[c - 3];
                if (cur != 0.0f) continue;
                uint32_t pos = 0xFFFFFFFFu;
                for (uint32_t i = 0; i < E; ++i) {
                    const uint32_t j = i * 32 + lane;
                    if (j < n) {
                        const float4 bb = box[2 * (size_t)s_gid[w][j] + (c < 3 ? 0 : 1)];
                        const float val = (c % 3 == 0) ? bb.x : ((c % 3 == 1) ? bb.y : bb.z);
                        if (val == 0.0f) pos = min(pos, j);
                    }
                }
                pos = __reduce_min_sync(FULL_MASK, pos);
                const float4 bb = box[2 * (size_t)s_gid[w][pos] + (c < 3 ? 0 : 1)];
                const float z = (c % 3 == 0) ? bb.x : ((c % 3 == 1) ? bb.y : bb.z);
                if (c < 3) nlo[c] = z; else nhi[c - 3] = z;
            }
        }
        // ---- 2. plane counts ----
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            const uint32_t j = i * 32 + lane;
            if (i < (int)E && j < n) s_pay[w][0][j] = j | (plane_counts(ccx[i], ccy[i], ccz[i], cmin, cmax) << 16);
        }
        __syncwarp();

        // ---- 3. shuffles ----
        uint32_t last_up = 0;
        auto shuffle = [&](int cur, uint32_t a, uint32_t b) -> uint32_t {
            const uint32_t sh = 16 + 3 * a;
            uint32_t bal[EPL], LFv[EPL];
            uint32_t nL = 0;
#pragma unroll
            for (int i = 0; i < EPL; ++i) {
                bal[i] = 0;
                LFv[i] = 0;
                if (i < (int)E) {
                    const uint32_t j = i * 32 + lane;
                    const bool L = (j < n) && (((s_pay[w][cur][j] >> sh) & 7u) < b);
                    bal[i] = __ballot_sync(FULL_MASK, L);
                    nL += __popc(bal[i]);
                }
            }
            // boundary element in closed form (see p_t1_table): f is nL-1, nL or nL+1, decided by three flags
            auto l_at = [&](uint32_t j) -> uint32_t {  // one broadcast shared-memory read
                return (j < n && (((s_pay[w][cur][j < n ? j : 0] >> sh) & 7u) < b)) ? 1u : 0u;
            };
            uint32_t f, Lf;
            {
                const uint32_t l0 = nL ? l_at(nL - 1) : 0u, l1 = l_at(nL), l2 = l_at(nL + 1);
                if (nL >= 1 && !(nL + 1 <= n && l0 + l1 <= 1)) { f = nL - 1; Lf = l0; }
                else if (!(nL + 2 <= n && l1 + l2 == 0)) { f = nL; Lf = l1; }
                else { f = nL + 1; Lf = l2; }
            }
            const uint32_t pivot = nL - Lf;
            uint32_t running = 0;
#pragma unroll
            for (int i = 0; i < EPL; ++i) {
                if (i >= (int)E) break;
                const uint32_t j = i * 32 + lane;
                const uint32_t LF = running + __popc(bal[i] & lt_mask);
                LFv[i] = LF;
                if (j < n) {
                    if ((bal[i] >> lane) & 1u) { if (j >= nL) s_tab[w][n - 1 - (nL - LF - 1)] = (uint16_t)j; }
                    else if (j <= nL) s_tab[w][j - LF] = (uint16_t)j;
                }
                running += __popc(bal[i]);
            }
            __syncwarp();
            uint32_t upay = 0;
#pragma unroll
            for (int i = 0; i < EPL; ++i) {
                if (i >= (int)E) break;
                const uint32_t j = i * 32 + lane;
                if (j < n) {
                    uint32_t pay = s_pay[w][cur][j];
                    const uint32_t Lbit = (bal[i] >> lane) & 1u;
                    const uint32_t LF = LFv[i], RF = j - LF;
                    uint32_t dest;
                    if (j < f) dest = Lbit ? j : (RF == 0 ? n - 1 : (uint32_t)s_tab[w][n - RF] - 1u);
                    else if (j == f) { dest = pivot; pay |= 0x80000000u; upay = pay; }
                    else dest = Lbit ? (uint32_t)s_tab[w][nL - LF - 1] : j - 1;
                    s_pay[w][cur ^ 1][dest] = pay;
                }
            }
            __syncwarp();
            last_up = __shfl_sync(FULL_MASK, upay, f & 31u);  // payload of the unexamined element
            return pivot;
        };

        int cur = 0;
        uint32_t my_u = 0xFFFFFFFFu, my_kb = 0, my_piv = 0;
        for (uint32_t c = 0; c < 21; ++c) {
            const uint32_t pivot = shuffle(cur, c / 7, c % 7 + 1);
            cur ^= 1;
            if (lane == c) { my_u = last_up & 0xFFFFu; my_kb = (last_up >> 16) & 0x1FFu; my_piv = pivot; }
        }

        // ---- 4. exact bins over the non-special primitives; lane a*8+k keeps bin (a,k) ----
        // (ordered uints from the load to the end of the reductions: one f2o per value, one o2f per bin)
        float mybin[6];
        {
            uint32_t mb[6] = {ENC_POS_INIT, ENC_POS_INIT, ENC_POS_INIT, ENC_NEG_INIT, ENC_NEG_INIT, ENC_NEG_INIT};
            uint32_t lo[EPL][3], hi[EPL][3];
            uint32_t kk[EPL];
#pragma unroll
            for (int i = 0; i < EPL; ++i) {
                const uint32_t j = i * 32 + lane;
                kk[i] = 0xFFFFFFFFu;
                lo[i][0] = lo[i][1] = lo[i][2] = ENC_POS_INIT;
                hi[i][0] = hi[i][1] = hi[i][2] = ENC_NEG_INIT;
                if (i < (int)E && j < n) {
                    const uint32_t pay = s_pay[w][cur][j];
                    if (!(pay & 0x80000000u)) {
                        const uint32_t g = s_gid[w][pay & 0xFFFFu];
                        const float4 b0 = box[2 * (size_t)g], b1 = box[2 * (size_t)g + 1];
                        lo[i][0] = f2o(b0.x); lo[i][1] = f2o(b0.y); lo[i][2] = f2o(b0.z);
                        hi[i][0] = f2o(b1.x); hi[i][1] = f2o(b1.y); hi[i][2] = f2o(b1.z);
                        kk[i] = (pay >> 16) & 0x1FFu;
                    }
                }
            }
            for (uint32_t a = 0; a < 3; ++a) {
                for (uint32_t k = 0; k < 8; ++k) {
                    uint32_t m[6] = {ENC_POS_INIT, ENC_POS_INIT, ENC_POS_INIT, ENC_NEG_INIT, ENC_NEG_INIT, ENC_NEG_INIT};
                    bool any = false;
#pragma unroll
                    for (int i = 0; i < EPL; ++i) {
                        if (i >= (int)E) break;  // most nodes of this tier fill two or three chunks, not eight
                        const bool in = (kk[i] != 0xFFFFFFFFu) && (((kk[i] >> (3 * a)) & 7u) == k);
                        if (in) {
                            any = true;
                            m[0] = min(m[0], lo[i][0]); m[1] = min(m[1], lo[i][1]); m[2] = min(m[2], lo[i][2]);
                            m[3] = max(m[3], hi[i][0]); m[4] = max(m[4], hi[i][1]); m[5] = max(m[5], hi[i][2]);
                        }
                    }
                    if (!__any_sync(FULL_MASK, any)) continue;
#pragma unroll
                    for (int c = 0; c < 6; ++c) {
                        const uint32_t r = (c < 3) ? __reduce_min_sync(FULL_MASK, m[c]) : __reduce_max_sync(FULL_MASK, m[c]);
                        if (lane == a * 8 + k) mb[c] = r;
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 6; ++c) mybin[c] = o2f(mb[c]);
        }
        // ---- 5. candidate costs and selection ----
        float ub[6] = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};  // box of special `lane`
        if (lane < 21) {
            const uint32_t g = s_gid[w][my_u];
            const float4 b0 = box[2 * (size_t)g], b1 = box[2 * (size_t)g + 1];
            ub[0] = b0.x; ub[1] = b0.y; ub[2] = b0.z; ub[3] = b1.x; ub[4] = b1.y; ub[5] = b1.z;
        }
        const uint32_t ca = (lane < 21) ? lane / 7 : 0, cb = lane % 7 + 1;
        float Lb[6] = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
        float Rb[6] = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) {
            float v[6];
#pragma unroll
            for (int x = 0; x < 6; ++x) v[x] = __shfl_sync(FULL_MASK, mybin[x], ca * 8 + k);
            if (k < cb) {
                Lb[0] = fminf(Lb[0], v[0]); Lb[1] = fminf(Lb[1], v[1]); Lb[2] = fminf(Lb[2], v[2]);
                Lb[3] = fmaxf(Lb[3], v[3]); Lb[4] = fmaxf(Lb[4], v[4]); Lb[5] = fmaxf(Lb[5], v[5]);
            } else {
                Rb[0] = fminf(Rb[0], v[0]); Rb[1] = fminf(Rb[1], v[1]); Rb[2] = fminf(Rb[2], v[2]);
                Rb[3] = fmaxf(Rb[3], v[3]); Rb[4] = fmaxf(Rb[4], v[4]); Rb[5] = fmaxf(Rb[5], v[5]);
            }
        }
        for (uint32_t s2 = 0; s2 < 21; ++s2) {
            const uint32_t u2 = __shfl_sync(FULL_MASK, my_u, s2), kb2 = __shfl_sync(FULL_MASK, my_kb, s2);
            float v[6];
#pragma unroll
            for (int x = 0; x < 6; ++x) v[x] = __shfl_sync(FULL_MASK, ub[x], s2);
            const bool left = (u2 != my_u) && (((kb2 >> (3 * ca)) & 7u) < cb);
            if (left) {
                Lb[0] = fminf(Lb[0], v[0]); Lb[1] = fminf(Lb[1], v[1]); Lb[2] = fminf(Lb[2], v[2]);
                Lb[3] = fmaxf(Lb[3], v[3]); Lb[4] = fmaxf(Lb[4], v[4]); Lb[5] = fmaxf(Lb[5], v[5]);
            } else {
                Rb[0] = fminf(Rb[0], v[0]); Rb[1] = fminf(Rb[1], v[1]); Rb[2] = fminf(Rb[2], v[2]);
                Rb[3] = fmaxf(Rb[3], v[3]); Rb[4] = fmaxf(Rb[4], v[4]); Rb[5] = fmaxf(Rb[5], v[5]);
            }
        }
        const float cost = sah_cost(Lb, Rb, my_piv, n - my_piv);
        const uint32_t key = (lane < 21 && cost < 3.402823466e+38f) ? __float_as_uint(cost) : 0xFFFFFFFFu;
        const uint32_t mk = __reduce_min_sync(FULL_MASK, key);
        if (mk == 0xFFFFFFFFu) {
            if (lane == 0) {
                atomicOr(&st->err, DERR_DEGENERATE);
                atomicSub(&st->w_pending, 1u);
            }
            __syncwarp();
            continue;
        }
        const uint32_t win = __ffs(__ballot_sync(FULL_MASK, key == mk)) - 1;
        const uint32_t p = __shfl_sync(FULL_MASK, my_piv, win);
        // ---- 6. final shuffle (blas.rs:164), write the order back ----
        shuffle(cur, win / 7, win % 7 + 1);
        cur ^= 1;
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            const uint32_t j = i * 32 + lane;
            if (i < (int)E && j < n) ids[start + j] = s_gid[w][s_pay[w][cur][j] & 0xFFFFu];
        }
        __threadfence();
        __syncwarp();
        // ---- 7. record + children ----
        // (finishing <=32-primitive children inline on this warp, right after their parent, was measured slower than
        //  handing them on: 3.30 ms vs 2.43 ms for the two tiers on the dragon-class mesh)
        if (lane == 0) {
            emit_rec(recs, 2 * (start + p) + 1, nlo, nhi, start, n, leftrun, pstart, pleftrun, tflags);
            if (p <= 3) A[start] = leftrun + 1;
            push_child(Q, st, epoch, start, p, leftrun + 1, start, leftrun, tflags & ~3u);
            push_child(Q, st, epoch, start + p, n - p, 0, start, leftrun, TF_RIGHT | (tflags & ~3u));
            atomicAdd(&st->t2w_done, 1u);
            __threadfence();
            atomicSub(&st->w_pending, 1u);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// T1: grid-wide phases over tiles of nodes with more than T2_CAP primitives.
// ------------------------------------------------------------------------------------------------
struct T1Args {
    const LevelNode* nodes;
    NodeScratch* sc;
    uint32_t n_nodes, n_tiles;
    uint32_t ept, tile_sz;  // this level: slots per thread (1, 2, 4 or 8) and slots per tile (T1_THREADS * ept)
    uint32_t* ids0;
    uint32_t* ids1;
    uint16_t* fl0;
    uint16_t* fl1;
    uint32_t* table;
    uint32_t* tileL;     // [22][tile_stride] per-candidate L count of every tile (current order at that shuffle)
    uint32_t* tileLF;    // [tile_stride] #L in the node before this tile, for the shuffle in flight
    uint32_t* pbal;      // [tile_stride][8 warps][9] PA -> PB: #L before the warp's first slot, then its <= 8 ballot words
    uint4* tile_desc;    // [tile_stride] {node, start, n, tile index inside the node}
    uint32_t tile_stride;
    const float4* cent;
    const float4* box;
    BuildState* st;
    uint32_t* barrier;   // monotonically increasing arrival counter
};

// Grid-wide barrier for the cooperative (co-resident) persistent kernel: one arrival counter that only ever
// grows, so there is no reset race; generation g completes when it reaches g * gridDim.x.
#ifdef BVH_T1_TIMING
__device__ unsigned long long g_t1_time[32];  // [2*kind] work ns, [2*kind+1] barrier wait ns (block 0)
__device__ unsigned long long g_t1_blk[2][1024];  // level 0: per-block work ns of the table / scatter phases
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define T1_PHASE(kind, call)                                                             \
    do {                                                                                 \
        unsigned long long _t0 = gtimer();                                               \
        call;                                                                            \
        __syncthreads();                                                                 \
        unsigned long long _t1 = gtimer();                                               \
        grid_barrier(g.barrier, gen);                                                    \
        unsigned long long _t2 = gtimer();                                               \
        if (blockIdx.x == 0 && threadIdx.x == 0) {                                       \
            g_t1_time[2 * (kind)] += _t1 - _t0;                                          \
            g_t1_time[2 * (kind) + 1] += _t2 - _t1;                                      \
        }                                                                                \
        if (threadIdx.x == 0 && level == 0 && ((kind) == 7 || (kind) == 8) && blockIdx.x < 1024) \
            g_t1_blk[(kind) - 7][blockIdx.x] += _t1 - _t0;                               \
    } while (0)
#else
#define T1_PHASE(kind, call)          \
    do {                              \
        call;                         \
        grid_barrier(g.barrier, gen); \
    } while (0)
#endif

__device__ __forceinline__ void grid_barrier(uint32_t* counter, uint32_t& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        gen += 1;
        const uint32_t target = gen * gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        while (ld_vol(counter) < target) { }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ void cand_of(const NodeScratch* sc, uint32_t node, int cand, uint32_t& a, uint32_t& b) {
    const uint32_t c = (cand < 21) ? (uint32_t)cand : sc[node].best;
    a = c / 7; b = c % 7 + 1;
}

// L0: per node — scratch init, tile descriptors, zero the per-candidate tile counters of the node's tiles.
__device__ __forceinline__ void p_t1_init(const T1Args& g) {
    for (uint32_t node = blockIdx.x; node < g.n_nodes; node += gridDim.x) {
        NodeScratch* s = g.sc + node;
        const LevelNode nd = g.nodes[node];
        const uint32_t nt = (nd.n + g.tile_sz - 1) / g.tile_sz;
        const uint32_t tid = threadIdx.x;
        if (tid < 12) s->bnd[tid] = ((tid % 6) < 3) ? ENC_POS_INIT : ENC_NEG_INIT;
        if (tid == 22) s->best = 0xFFFFFFFFu;
        if (tid >= 32 && tid < 38) s->zkey[tid - 32] = 0xFFFFFFFFFFFFFFFFull;
        if (tid < 144) (&s->bins[0][0][0])[tid] = ((tid % 6) < 3) ? ENC_POS_INIT : ENC_NEG_INIT;
        for (uint32_t t = tid; t < nt; t += blockDim.x) g.tile_desc[nd.tile_base + t] = make_uint4(node, nd.start, nd.n, t);
        for (uint32_t k = tid; k < 22 * nt; k += blockDim.x) g.tileL[(size_t)(k / nt) * g.tile_stride + nd.tile_base + (k % nt)] = 0;
    }
}

// L1: per tile — vertex box and centroid bounds of the node (blas.rs:87-88,117-123,142).
template <int EPT>
__device__ __forceinline__ void p_t1_bounds(const T1Args& g) {
    __shared__ uint32_t s_red[T1_THREADS / 32][12];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        const uint4 td = g.tile_desc[tile];
        const uint32_t node = td.x;
        struct { uint32_t start, n, tile_base; } nd = {td.y, td.z, tile - td.w};
        const uint32_t j0 = td.w * g.tile_sz;
        float acc[12] = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f, 1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            if (i >= EPT) break;
            const uint32_t j = j0 + i * T1_THREADS + tid;
            if (j < nd.n) {
                const uint32_t id = g.ids0[nd.start + j];
                const float4 c = g.cent[id];
                const float4 b0 = g.box[2 * (size_t)id], b1 = g.box[2 * (size_t)id + 1];
                acc[0] = fminf(acc[0], b0.x); acc[1] = fminf(acc[1], b0.y); acc[2] = fminf(acc[2], b0.z);
                acc[3] = fmaxf(acc[3], b1.x); acc[4] = fmaxf(acc[4], b1.y); acc[5] = fmaxf(acc[5], b1.z);
                acc[6] = fminf(acc[6], c.x); acc[7] = fminf(acc[7], c.y); acc[8] = fminf(acc[8], c.z);
                acc[9] = fmaxf(acc[9], c.x); acc[10] = fmaxf(acc[10], c.y); acc[11] = fmaxf(acc[11], c.z);
            }
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const bool is_min = (k < 3) || (k >= 6 && k < 9);
            const uint32_t v = f2o(acc[k]);
            const uint32_t r = is_min ? __reduce_min_sync(FULL_MASK, v) : __reduce_max_sync(FULL_MASK, v);
            if (lane == 0) s_red[warp][k] = r;
        }
        __syncthreads();
        if (tid < 12) {
            const bool is_min = (tid < 3) || (tid >= 6 && tid < 9);
            uint32_t r = s_red[0][tid];
            for (int w2 = 1; w2 < T1_THREADS / 32; ++w2) r = is_min ? min(r, s_red[w2][tid]) : max(r, s_red[w2][tid]);
            if (is_min) atomicMin(&g.sc[node].bnd[tid], r);
            else atomicMax(&g.sc[node].bnd[tid], r);
        }
        __syncthreads();
    }
}

// L2: per tile — plane counts of every primitive, and the tile's L count for candidate 0 (x axis, b = 1).
template <int EPT>
__device__ __forceinline__ void p_t1_flags(const T1Args& g) {
    __shared__ uint32_t s_w[T1_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        const uint4 td = g.tile_desc[tile];
        const uint32_t node = td.x;
        struct { uint32_t start, n, tile_base; } nd = {td.y, td.z, tile - td.w};
        const uint32_t j0 = td.w * g.tile_sz;
        float cmin[3], cmax[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { cmin[c] = o2f(g.sc[node].bnd[6 + c]); cmax[c] = o2f(g.sc[node].bnd[9 + c]); }
        uint32_t cnt = 0;
        bool zero_face[6];
        bool any_zero = false;
        if (g.st->neg_zero) {
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const uint32_t e = g.sc[node].bnd[c];
                zero_face[c] = o2f(c < 3 ? min(e, ENC_POS_INIT) : max(e, ENC_NEG_INIT)) == 0.0f;
                any_zero = any_zero || zero_face[c];
            }
        }
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            if (i >= EPT) break;
            const uint32_t j = j0 + i * T1_THREADS + tid;
            bool L = false;
            if (j < nd.n) {
                const uint32_t id = g.ids0[nd.start + j];
                const float4 c = g.cent[id];
                const uint32_t kb = plane_counts(c.x, c.y, c.z, cmin, cmax);
                g.fl0[nd.start + j] = (uint16_t)kb;
                L = (kb & 7u) < 1u;
                if (any_zero) {  // rare path (-0.0 in the input): first slot with a zero on each zero-valued face
                    const float4 b0 = g.box[2 * (size_t)id], b1 = g.box[2 * (size_t)id + 1];
                    const float vals[6] = {b0.x, b0.y, b0.z, b1.x, b1.y, b1.z};
#pragma unroll
                    for (int cc = 0; cc < 6; ++cc)
                        if (zero_face[cc] && vals[cc] == 0.0f) atomicMin(&g.sc[node].zkey[cc], ((unsigned long long)j << 32) | id);
                }
            }
            cnt += __popc(__ballot_sync(FULL_MASK, L));
        }
        if (lane == 0) s_w[warp] = cnt;
        __syncthreads();
        if (tid == 0) {
            uint32_t tot = 0;
            for (int w2 = 0; w2 < T1_THREADS / 32; ++w2) tot += s_w[w2];
            g.tileL[tile] = tot;  // candidate 0
        }
        __syncthreads();
    }
}

// Per-tile L count for the final (winning) candidate, whose identity is only known after select.
template <int EPT>
__device__ __forceinline__ void p_t1_count_final(const T1Args& g, const uint16_t* fl) {
    __shared__ uint32_t s_w[T1_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        const uint4 td = g.tile_desc[tile];
        const uint32_t node = td.x;
        struct { uint32_t start, n, tile_base; } nd = {td.y, td.z, tile - td.w};
        const uint32_t j0 = td.w * g.tile_sz;
        uint32_t a, b;
        cand_of(g.sc, node, 21, a, b);
        uint32_t cnt = 0;
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            if (i >= EPT) break;
            const uint32_t j = j0 + i * T1_THREADS + tid;
            const bool L = (j < nd.n) && ((((uint32_t)fl[nd.start + j] >> (3 * a)) & 7u) < b);
            cnt += __popc(__ballot_sync(FULL_MASK, L));
        }
        if (lane == 0) s_w[warp] = cnt;
        __syncthreads();
        if (tid == 0) {
            uint32_t tot = 0;
            for (int w2 = 0; w2 < T1_THREADS / 32; ++w2) tot += s_w[w2];
            g.tileL[(size_t)21 * g.tile_stride + tile] = tot;
        }
        __syncthreads();
    }
}

// Optional (levels whose nodes span many tiles): one warp per node scans its tiles' L counts for shuffle c.
__device__ __forceinline__ void p_t1_tilescan(const T1Args& g, int c) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const uint32_t* tl = g.tileL + (size_t)c * g.tile_stride;
    for (uint32_t node = gw; node < g.n_nodes; node += nw) {
        const LevelNode nd = g.nodes[node];
        const uint32_t nt = (nd.n + g.tile_sz - 1) / g.tile_sz;
        uint32_t carry = 0;
        for (uint32_t base = 0; base < nt; base += 32) {
            const uint32_t i = base + lane;
            const uint32_t v = (i < nt) ? tl[nd.tile_base + i] : 0;
            uint32_t x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(FULL_MASK, x, o);
                if ((int)lane >= o) x += y;
            }
            if (i < nt) g.tileLF[nd.tile_base + i] = carry + x - v;
            carry += __shfl_sync(FULL_MASK, x, 31);
        }
        if (lane == 0) g.sc[node].nL[c] = carry;
    }
}

// Ballots + per-element #L-before (LF) of one tile.  Layout: j = j0 + warp*256 + i*32 + lane.
// Issues the tile's flag loads early (before anything that waits on other loads or barriers).
template <int EPT>
__device__ __forceinline__ void t1_load_flags(const uint16_t* fl, uint32_t start, uint32_t n, uint32_t j0, uint16_t* fw,
                                              const uint32_t ept) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
        if (i >= (int)ept) break;
        const uint32_t j = j0 + warp * (32 * ept) + i * 32 + lane;
        fw[i] = (j < n) ? fl[start + j] : (uint16_t)0;
    }
}

template <int EPT>
__device__ __forceinline__ void t1_prefix(uint32_t n, uint32_t j0, uint32_t a, uint32_t b, uint32_t tile_lf, uint32_t* s_w,
                                          uint32_t* bal, uint32_t* LFv, const uint16_t* fw, const uint32_t ept) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t cnt = 0;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
        if (i >= (int)ept) break;
        const uint32_t j = j0 + warp * (32 * ept) + i * 32 + lane;
        const bool L = (j < n) && ((((uint32_t)fw[i] >> (3 * a)) & 7u) < b);
        bal[i] = __ballot_sync(FULL_MASK, L);
        cnt += __popc(bal[i]);
    }
    if (lane == 0) s_w[warp] = cnt;
    __syncthreads();
    const uint32_t wv = (lane < warp) ? s_w[lane & (T1_THREADS / 32 - 1)] : 0u;
    uint32_t running = tile_lf + __reduce_add_sync(FULL_MASK, wv);
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
        if (i >= (int)ept) break;
        LFv[i] = running + __popc(bal[i] & lt_mask);
        running += __popc(bal[i]);
    }
    __syncthreads();
}

// PA(c): per tile — tile prefix, rank->position table (Appendix B), and the boundary element f: the first
// element that the front cursor does not examine.  "front-examined" is a prefix of the node, so exactly one
// thread of the whole grid sees the true->false transition; it publishes {nL, f, pivot} for the scatter phase.
template <int EPT>
__device__ __forceinline__ void p_t1_table(const T1Args& g, int c, const uint16_t* fl, bool scanned) {
    // 16-byte aligned: the compiler reads s_pre / s_tot with LDS.128, and an unaligned array made the first of those
    // loads cover the last word of its neighbour (harmless, but compute-sanitizer racecheck reports it)
    __shared__ __align__(16) uint32_t s_w[T1_THREADS / 32];
    __shared__ __align__(16) uint32_t s_pre[T1_THREADS / 32];
    __shared__ __align__(16) uint32_t s_tot[T1_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* tl = g.tileL + (size_t)c * g.tile_stride;
    for (uint32_t tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        const uint4 td = g.tile_desc[tile];
        constexpr uint32_t ept = EPT;
        const uint32_t node = td.x, start = td.y, n = td.z, lt = td.w, tile_base = tile - lt, j0 = lt * g.tile_sz;
        uint16_t fw[EPT];
        t1_load_flags<EPT>(fl, start, n, j0, fw, ept);
        uint32_t a, b;
        cand_of(g.sc, node, c, a, b);
        const uint32_t jw = j0 + warp * (32 * ept);
        uint32_t tile_lf, nL;
        if (scanned) {
            tile_lf = g.tileLF[tile];
            nL = g.sc[node].nL[c];
        } else {
            const uint32_t nt = (n + g.tile_sz - 1) / g.tile_sz;
            uint32_t pre = 0, tot = 0;
            for (uint32_t t = tid; t < nt; t += T1_THREADS) {
                const uint32_t v = tl[tile_base + t];
                tot += v;
                if (t < lt) pre += v;
            }
            pre = __reduce_add_sync(FULL_MASK, pre);
            tot = __reduce_add_sync(FULL_MASK, tot);
            if (lane == 0) { s_pre[warp] = pre; s_tot[warp] = tot; }
            __syncthreads();
            tile_lf = 0; nL = 0;
#pragma unroll
            for (int w2 = 0; w2 < T1_THREADS / 32; ++w2) { tile_lf += s_pre[w2]; nL += s_tot[w2]; }
            if (tid == 0) g.tileLF[tile] = tile_lf;
        }
        // The boundary element needs no search: pred(j) <=> j + 2 <= n && j + L(j) + L(j+1) <= nL (the #L-before terms
        // cancel), true for every j <= nL - 2 and false from nL + 1 on, so f is one of nL-1, nL, nL+1 and follows from nL
        // and three flags.  One thread per node publishes {nL, f, pivot}; nobody else evaluates pred.
        if (lt == 0 && tid == 0) {
            auto l_at = [&](uint32_t j) -> uint32_t {
                return (j < n && ((((uint32_t)fl[start + j] >> (3 * a)) & 7u) < b)) ? 1u : 0u;
            };
            const uint32_t l0 = nL ? l_at(nL - 1) : 0u, l1 = l_at(nL), l2 = l_at(nL + 1);
            uint32_t f, lf;
            if (nL >= 1 && !(nL + 1 <= n && l0 + l1 <= 1)) { f = nL - 1; lf = l0; }
            else if (!(nL + 2 <= n && l1 + l2 == 0)) { f = nL; lf = l1; }
            else { f = nL + 1; lf = l2; }
            g.sc[node].sh[c] = make_uint4(nL, f, nL - lf, 0);
        }
        uint32_t bal[EPT], LFv[EPT];
        t1_prefix<EPT>(n, j0, a, b, tile_lf, s_w, bal, LFv, fw, ept);
#if T1_PB_REUSE
        if (lane == 0) {  // PB of this shuffle runs on the same tile: leave it the ballots and the warp's prefix base
            uint32_t* pb = g.pbal + ((size_t)tile * (T1_THREADS / 32) + warp) * 9;
            pb[0] = LFv[0];
#pragma unroll
            for (int i = 0; i < EPT; ++i) pb[1 + i] = bal[i];
        }
#endif
        // Only front R's (j < f) and back L's (j > f) are ever looked up; with f in [nL-1, nL+1] that is j <= nL / j >= nL.
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            const uint32_t j = jw + i * 32 + lane;
            if (j < n) {
                const uint32_t LF = LFv[i];
                if ((bal[i] >> lane) & 1u) { if (j >= nL) g.table[start + n - 1 - (nL - LF - 1)] = j; }
                else if (j <= nL) g.table[start + (j - LF)] = j;
            }
        }
        __syncthreads();
    }
}

// PB(c): per tile — destinations and scatter into the other buffer; also accumulates, per destination tile,
// the L count of the NEXT candidate (so shuffle c+1 needs no separate counting pass).
template <int EPT>
__device__ __forceinline__ void p_t1_scatter(const T1Args& g, int c, const uint32_t* ids_in, const uint16_t* fl,
                                             uint32_t* ids_out, uint16_t* fl_out) {
    __shared__ uint32_t s_w[T1_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool count_next = c < 20;  // candidates 1..20 are known in advance; the final one is not
    const uint32_t na = (uint32_t)(c + 1) / 7, nb = (uint32_t)(c + 1) % 7 + 1;
    uint32_t* tl_next = g.tileL + (size_t)(c + 1) * g.tile_stride;
    for (uint32_t tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        const uint4 td = g.tile_desc[tile];
        const uint32_t node = td.x;
        struct { uint32_t start, n, tile_base; } nd = {td.y, td.z, tile - td.w};
        const uint32_t j0 = td.w * g.tile_sz;
        const uint32_t n = nd.n;
        constexpr uint32_t ept = EPT;
        const uint32_t tshift = 31 - __clz(g.tile_sz);  // tiles are powers of two
        uint16_t fwv[EPT];
        uint32_t idv[EPT];
        t1_load_flags<EPT>(fl, nd.start, n, j0, fwv, ept);
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            if (i >= (int)ept) break;
            const uint32_t j = j0 + warp * (32 * ept) + i * 32 + lane;
            idv[i] = (j < n) ? ids_in[nd.start + j] : 0u;
        }
        uint32_t a, b;
        cand_of(g.sc, node, c, a, b);
        const uint4 sh = g.sc[node].sh[c];
        const uint32_t nL = sh.x, f = sh.y, pivot = sh.z;
        uint32_t bal[EPT], LFv[EPT];
#if T1_PB_REUSE
        {
            const uint32_t* pb = g.pbal + ((size_t)tile * (T1_THREADS / 32) + warp) * 9;
            uint32_t running = pb[0];
            const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                bal[i] = pb[1 + i];
                LFv[i] = running + __popc(bal[i] & lt_mask);
                running += __popc(bal[i]);
            }
            (void)a; (void)b; (void)s_w;
        }
#else
        t1_prefix<EPT>(n, j0, a, b, g.tileLF[tile], s_w, bal, LFv, fwv, ept);
#endif
        uint32_t own_cnt = 0;
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            if (i >= (int)ept) break;
            const uint32_t j = j0 + warp * (32 * ept) + i * 32 + lane;
            uint32_t dtile = 0xFFFFFFFFu;
            bool Lnx = false;
            if (j < n) {
                const uint32_t Lbit = (bal[i] >> lane) & 1u;
                const uint32_t LF = LFv[i], RF = j - LF;
                const uint32_t id = idv[i];
                uint32_t fw = fwv[i];
                uint32_t dest;
                if (j < f) dest = Lbit ? j : (RF == 0 ? n - 1 : g.table[nd.start + n - RF] - 1u);
                else if (j == f) {
                    dest = pivot;
                    fw |= 0x8000u;
                    if (c < 21) { g.sc[node].piv[c] = pivot; g.sc[node].uid[c] = id; }
                } else dest = Lbit ? g.table[nd.start + (nL - LF - 1)] : j - 1;
                ids_out[nd.start + dest] = id;
                fl_out[nd.start + dest] = (uint16_t)fw;
                dtile = nd.tile_base + (dest >> tshift);
                Lnx = ((fw >> (3 * na)) & 7u) < nb;
            }
            if (count_next) {
                // Most elements stay inside their own tile (front L's do not move, back R's shift by one), so those
                // are counted with one ballot into a per-warp register; only elements that change tile use an atomic.
                const bool own = (dtile == tile);
                own_cnt += __popc(__ballot_sync(FULL_MASK, own && Lnx));
                // the rest: warp-aggregated per destination tile (32 consecutive slots land in very few tiles;
                // one atomic per lane was measured 1.7x slower for the whole tier)
                uint32_t todo = __ballot_sync(FULL_MASK, dtile != 0xFFFFFFFFu && !own && Lnx);
                while (todo) {
                    const uint32_t leader = __ffs(todo) - 1;
                    const uint32_t lt = __shfl_sync(FULL_MASK, dtile, leader);
                    const uint32_t same = __ballot_sync(FULL_MASK, dtile == lt) & todo;
                    if (lane == leader) atomicAdd(&tl_next[lt], (uint32_t)__popc(same));
                    todo &= ~same;
                }
            }
        }
        if (count_next && lane == 0 && own_cnt) atomicAdd(&tl_next[tile], own_cnt);
    }
}

template <int EPT>
__device__ __forceinline__ void p_t1_bins(const T1Args& g, const uint32_t* ids, const uint16_t* fl) {
    __shared__ uint32_t s_bins[3][8][6];
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    for (uint32_t tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        const uint4 td = g.tile_desc[tile];
        const uint32_t node = td.x;
        struct { uint32_t start, n, tile_base; } nd = {td.y, td.z, tile - td.w};
        const uint32_t j0 = td.w * g.tile_sz;
        if (tid < 144) (&s_bins[0][0][0])[tid] = ((tid % 6) < 3) ? ENC_POS_INIT : ENC_NEG_INIT;
        __syncthreads();
        uint32_t lo[EPT][3], hi[EPT][3];  // ordered uints (f2o): the per-bin reductions below are integer min / max
        uint32_t kk[EPT];
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            const uint32_t j = j0 + i * T1_THREADS + tid;
            kk[i] = 0xFFFFFFFFu;
            lo[i][0] = lo[i][1] = lo[i][2] = ENC_POS_INIT;
            hi[i][0] = hi[i][1] = hi[i][2] = ENC_NEG_INIT;
            if (j < nd.n) {
                const uint32_t fw = fl[nd.start + j];
                if (!(fw & 0x8000u)) {
                    const uint32_t id = ids[nd.start + j];
                    const float4 b0 = g.box[2 * (size_t)id], b1 = g.box[2 * (size_t)id + 1];
                    lo[i][0] = f2o(b0.x); lo[i][1] = f2o(b0.y); lo[i][2] = f2o(b0.z);
                    hi[i][0] = f2o(b1.x); hi[i][1] = f2o(b1.y); hi[i][2] = f2o(b1.z);
                    kk[i] = fw & 0x1FFu;
                }
            }
        }
        for (uint32_t a = 0; a < 3; ++a) {
            for (uint32_t k = 0; k < 8; ++k) {
                uint32_t m[6] = {ENC_POS_INIT, ENC_POS_INIT, ENC_POS_INIT, ENC_NEG_INIT, ENC_NEG_INIT, ENC_NEG_INIT};
                bool any = false;
#pragma unroll
                for (int i = 0; i < EPT; ++i) {
                    if (i >= EPT) break;
                    const bool in = (kk[i] != 0xFFFFFFFFu) && (((kk[i] >> (3 * a)) & 7u) == k);
                    if (in) {
                        any = true;
                        m[0] = min(m[0], lo[i][0]); m[1] = min(m[1], lo[i][1]); m[2] = min(m[2], lo[i][2]);
                        m[3] = max(m[3], hi[i][0]); m[4] = max(m[4], hi[i][1]); m[5] = max(m[5], hi[i][2]);
                    }
                }
                if (!__any_sync(FULL_MASK, any)) continue;
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const uint32_t r = (c < 3) ? __reduce_min_sync(FULL_MASK, m[c]) : __reduce_max_sync(FULL_MASK, m[c]);
                    if (lane == 0) {
                        if (c < 3) atomicMin(&s_bins[a][k][c], r);
                        else atomicMax(&s_bins[a][k][c], r);
                    }
                }
            }
        }
        __syncthreads();
        if (tid < 144) {
            const uint32_t v = (&s_bins[0][0][0])[tid];
            const bool is_min = (tid % 6) < 3;
            if (is_min) { if (v != ENC_POS_INIT) atomicMin(&(&g.sc[node].bins[0][0][0])[tid], v); }
            else { if (v != ENC_NEG_INIT) atomicMax(&(&g.sc[node].bins[0][0][0])[tid], v); }
        }
        __syncthreads();
    }
}

// One warp per node: evaluate the 21 candidates from bins + specials, pick the winner (blas.rs:155-161).
__device__ __forceinline__ void p_t1_select(const T1Args& g) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t node = gw; node < g.n_nodes; node += nw) {
        const LevelNode nd = g.nodes[node];
        NodeScratch* s = g.sc + node;
        float cmin[3], cmax[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { cmin[c] = o2f(s->bnd[6 + c]); cmax[c] = o2f(s->bnd[9 + c]); }
        const uint32_t c = lane;
        const uint32_t cs = (c < 21) ? c : 0;
        const uint32_t my_uid = s->uid[cs];  // lane c also owns special c
        const float4 ce = g.cent[my_uid];
        const float4 b0 = g.box[2 * (size_t)my_uid], b1 = g.box[2 * (size_t)my_uid + 1];
        const uint32_t my_kb = plane_counts(ce.x, ce.y, ce.z, cmin, cmax);
        const uint32_t a = cs / 7, b = cs % 7 + 1;
        float Lb[6] = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
        float Rb[6] = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
        for (uint32_t k = 0; k < 8; ++k) {
            float* side = (k < b) ? Lb : Rb;
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                side[x] = fminf(side[x], o2f(s->bins[a][k][x]));
                side[3 + x] = fmaxf(side[3 + x], o2f(s->bins[a][k][3 + x]));
            }
        }
        for (uint32_t s2 = 0; s2 < 21; ++s2) {
            const uint32_t uid2 = __shfl_sync(FULL_MASK, my_uid, s2);
            const uint32_t kb2 = __shfl_sync(FULL_MASK, my_kb, s2);
            float bx[6];
            bx[0] = __shfl_sync(FULL_MASK, b0.x, s2); bx[1] = __shfl_sync(FULL_MASK, b0.y, s2);
            bx[2] = __shfl_sync(FULL_MASK, b0.z, s2); bx[3] = __shfl_sync(FULL_MASK, b1.x, s2);
            bx[4] = __shfl_sync(FULL_MASK, b1.y, s2); bx[5] = __shfl_sync(FULL_MASK, b1.z, s2);
            const bool left = (uid2 != my_uid) && (((kb2 >> (3 * a)) & 7u) < b);
            float* side = left ? Lb : Rb;
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                side[x] = fminf(side[x], bx[x]);
                side[3 + x] = fmaxf(side[3 + x], bx[3 + x]);
            }
        }
        const uint32_t n1 = s->piv[cs];
        const float cost = sah_cost(Lb, Rb, n1, nd.n - n1);
        const uint32_t key = (c < 21 && cost < 3.402823466e+38f) ? __float_as_uint(cost) : 0xFFFFFFFFu;
        const uint32_t mk = __reduce_min_sync(FULL_MASK, key);
        const uint32_t bal = __ballot_sync(FULL_MASK, key == mk);
        if (lane == 0) {
            if (mk == 0xFFFFFFFFu) {
                atomicOr(&g.st->err, DERR_DEGENERATE);
                s->best = 0;  // keep the remaining phases well-defined; the build is reported as failed
            } else s->best = __ffs(bal) - 1;
        }
    }
}

// One thread per node: record, A counter, children to the next level / block queue / warp queue / T3 list.
__device__ __forceinline__ void p_t1_children(const T1Args& g, LevelNode* next_nodes, uint32_t next_cap, int next_slot,
                                              const Queues& Q, uint4* recs, uint32_t* A, uint32_t epoch) {
    for (uint32_t node = blockIdx.x * blockDim.x + threadIdx.x; node < g.n_nodes; node += gridDim.x * blockDim.x) {
        const LevelNode nd = g.nodes[node];
        const NodeScratch* s = g.sc + node;
        const uint32_t p = s->piv[s->best];
        float lo[3], hi[3];
        for (int c = 0; c < 3; ++c) { lo[c] = o2f(min(s->bnd[c], ENC_POS_INIT)); hi[c] = o2f(max(s->bnd[3 + c], ENC_NEG_INIT)); }
        if (p == 0 || p >= nd.n) { atomicOr(&g.st->err, DERR_DEGENERATE); continue; }
        for (int c = 0; c < 6; ++c)
            if (s->zkey[c] != 0xFFFFFFFFFFFFFFFFull) {  // only set on the rare -0.0 path
                const uint32_t id = (uint32_t)(s->zkey[c] & 0xFFFFFFFFull);
                const float4 bb = g.box[2 * (size_t)id + (c < 3 ? 0 : 1)];
                const float z = (c % 3 == 0) ? bb.x : ((c % 3 == 1) ? bb.y : bb.z);
                if (c < 3) lo[c] = z; else hi[c - 3] = z;
            }
        emit_rec(recs, 2 * (nd.start + p) + 1, lo, hi, nd.start, nd.n, nd.leftrun, nd.pstart, nd.pleftrun, nd.flags);
        if (p <= 3) A[nd.start] = nd.leftrun + 1;
        atomicAdd(&g.st->grid_nodes, 1u);
        atomicAdd(&g.st->sum_grid, (unsigned long long)nd.n);
        for (int side = 0; side < 2; ++side) {
            const uint32_t cs = side ? nd.start + p : nd.start;
            const uint32_t cn = side ? nd.n - p : p;
            const uint32_t clr = side ? 0 : nd.leftrun + 1;
            const uint32_t cfl = (side ? TF_RIGHT : 0u) | (nd.flags & ~3u);
            if (cn > T2B_CAP) {
                const uint32_t idx = atomicAdd(&g.st->lv_count[next_slot], 1u);
                if (idx >= next_cap) { atomicOr(&g.st->err, DERR_QUEUE); continue; }
                LevelNode c;
                c.start = cs; c.n = cn; c.leftrun = clr; c.pstart = nd.start; c.pleftrun = nd.leftrun; c.flags = cfl;
                c.tile_base = 0; c.pad = 0;
                next_nodes[idx] = c;
            } else {
                push_child(Q, g.st, epoch, cs, cn, clr, nd.start, nd.leftrun, cfl);
            }
        }
    }
}

// One block: picks the level's tile size, then the tile_base prefix of its node list; decides whether the level needs
// the tile scan.  Tile size: a phase is a fixed chain of ~60 dependent instructions per slot a thread owns, so a level
// whose nodes fit the grid with fewer slots per thread (256-, 512- or 1024-slot tiles) takes them.
__device__ __forceinline__ void p_t1_nextlevel(LevelNode* nodes, BuildState* st, int slot, int other, uint32_t grid_blocks,
                                               bool count_level = true) {
    __shared__ uint32_t s_part[1024];
    __shared__ uint32_t s_max;
    __shared__ uint32_t s_tot[4];
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint32_t n = ld_vol(&st->lv_count[slot]);
    const uint32_t per = (n + nt - 1) / nt;
    const uint32_t b = min(n, tid * per), e = min(n, b + per);
    if (tid == 0) s_max = 0;
    if (tid < 4) s_tot[tid] = 0;
    __syncthreads();
    {
        uint32_t t[4] = {0, 0, 0, 0};
        for (uint32_t i = b; i < e; ++i) {
            const uint32_t nn = nodes[i].n;
#pragma unroll
            for (int k = 0; k < 4; ++k) t[k] += (nn + (T1_THREADS << k) - 1) / (T1_THREADS << k);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (t[k]) atomicAdd(&s_tot[k], t[k]);
    }
    __syncthreads();
    uint32_t lg = 3;
#if T1_VAR_TILE
    for (int k = 2; k >= 0; --k)
        if (s_tot[k] <= grid_blocks) lg = (uint32_t)k;
#endif
    const uint32_t ts = (uint32_t)T1_THREADS << lg;
    uint32_t sum = 0, mx = 0;
    for (uint32_t i = b; i < e; ++i) {
        const uint32_t t = (nodes[i].n + ts - 1) / ts;
        sum += t;
        mx = max(mx, t);
    }
    s_part[tid] = sum;
    if (mx) atomicMax(&s_max, mx);
    __syncthreads();
    if (tid == 0) {
        uint32_t acc = 0;
        for (uint32_t i = 0; i < nt; ++i) { const uint32_t v = s_part[i]; s_part[i] = acc; acc += v; }
        st->lv_tiles[slot] = acc;
        st->lv_maxtiles[slot] = s_max;
        st->lv_ept[slot] = 1u << lg;
        st->lv_count[other] = 0;
        if (count_level) st->levels_done += 1;
    }
    __syncthreads();
    uint32_t acc = s_part[tid];
    for (uint32_t i = b; i < e; ++i) {
        nodes[i].tile_base = acc;
        acc += (nodes[i].n + ts - 1) / ts;
    }
    __syncthreads();
}

// Tile prefix of the root level (one block).
__global__ void __launch_bounds__(1024) k_t1_level0(LevelNode* nodes, BuildState* st, uint32_t grid_blocks) {
    p_t1_nextlevel(nodes, st, 0, 1, grid_blocks, false);
}

// All shuffles of one level with EPT slots per thread (tile = T1_THREADS * EPT slots).
template <int EPT>
__device__ __forceinline__ void t1_level(const T1Args& g, const bool scanned, uint32_t& gen, const uint32_t level) {
    (void)level;
    T1_PHASE(0, p_t1_init(g));
    T1_PHASE(1, p_t1_bounds<EPT>(g));
    T1_PHASE(2, p_t1_flags<EPT>(g));
    for (int c = 0; c < 22; ++c) {
        const uint32_t* ids_in = (c & 1) ? g.ids1 : g.ids0;
        uint32_t* ids_out = (c & 1) ? g.ids0 : g.ids1;
        const uint16_t* fl_in = (c & 1) ? g.fl1 : g.fl0;
        uint16_t* fl_out = (c & 1) ? g.fl0 : g.fl1;
        if (c == 21) {
            T1_PHASE(3, p_t1_bins<EPT>(g, ids_in, fl_in));
            T1_PHASE(4, p_t1_select(g));
            T1_PHASE(5, p_t1_count_final<EPT>(g, fl_in));
        }
        if (scanned) T1_PHASE(6, p_t1_tilescan(g, c));
        T1_PHASE(7, p_t1_table<EPT>(g, c, fl_in, scanned));
        T1_PHASE(8, p_t1_scatter<EPT>(g, c, ids_in, fl_in, ids_out, fl_out));
    }
}

// The whole grid-wide tier as ONE cooperative persistent kernel: every phase boundary is a grid barrier
// instead of a kernel launch, and the level loop never returns to the host.  ~52 barriers per level.
// Measured on B200 (dragon-class, -DBVH_T1_TIMING): every level costs 315-440 us whether it has 426 tiles or 9
// (390, 379, 389, 385, 440, 385, 362, 367, 355, 314 us for 426, 426, 427, 430, 435, 396, 221, 65, 17, 9 tiles): a phase is
// ~480 dependent warp-instructions per warp (8 slots per thread) plus a barrier, ~6.5 us, not a matter of bandwidth,
// of L2 round trips or of how many blocks arrive at the barrier.  Built, verified bit-exact and not faster (first form in
// commit ac13e1f): PA + barrier + PB as one function with the tile's ballots kept in registers and the tile
// descriptor read once per level (3.87 ms vs 3.70 ms); barriers restricted to the blocks that own a tile (3.67 vs 3.69).
__global__ void __launch_bounds__(T1_THREADS, 3) k_t1_coop(T1Args g, LevelNode* lv0, LevelNode* lv1, uint32_t lv_cap, Queues Q,
                                                        uint4* recs, uint32_t* A, uint32_t epoch, uint32_t max_levels) {
    LevelNode* lv[2] = {lv0, lv1};
    uint32_t gen = 0;
    int slot = 0;
    for (uint32_t level = 0; level < max_levels; ++level) {
        const uint32_t n_nodes = ld_vol(&g.st->lv_count[slot]);
        const uint32_t n_tiles = ld_vol(&g.st->lv_tiles[slot]);
        const bool scanned = ld_vol(&g.st->lv_maxtiles[slot]) > 512u;
        if (n_nodes == 0) break;
        g.ept = ld_vol(&g.st->lv_ept[slot]);
        g.tile_sz = (uint32_t)T1_THREADS * g.ept;
#ifdef BVH_T1_TIMING
        if (blockIdx.x == 0 && threadIdx.x == 0 && level < 100) {
            g_t1_blk[0][900 + level] = gtimer();  // level start; [1][900 + level] = {nodes, tiles}
            g_t1_blk[1][900 + level] = ((unsigned long long)n_nodes << 32) | n_tiles;
        }
#endif
        g.nodes = lv[slot];
        g.n_nodes = n_nodes;
        g.n_tiles = n_tiles;
        switch (g.ept) {
            case 1: t1_level<1>(g, scanned, gen, level); break;
            case 2: t1_level<2>(g, scanned, gen, level); break;
            case 4: t1_level<4>(g, scanned, gen, level); break;
            default: t1_level<8>(g, scanned, gen, level); break;
        }
        const int next = slot ^ 1;
        T1_PHASE(9, p_t1_children(g, lv[next], lv_cap, next, Q, recs, A, epoch));
        T1_PHASE(10, if (blockIdx.x == 0) p_t1_nextlevel(lv[next], g.st, next, slot, gridDim.x));
        slot = next;
#ifdef BVH_T1_TIMING
        if (blockIdx.x == 0 && threadIdx.x == 0 && level < 100) g_t1_blk[0][900 + level + 1] = gtimer();
#endif
    }
}

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
    const uint32_t start = tbase[m], n = tbase[m + 1] - tbase[m];
    const uint32_t flags = TF_ROOT | (m << TF_MESH_SHIFT);
    if (n == 0) { atomicOr(&st->err, DERR_BAD_INDEX); return; }
    if (n > T2B_CAP) {
        const uint32_t idx = atomicAdd(&st->lv_count[0], 1u);
        if (idx >= lv_cap) { atomicOr(&st->err, DERR_QUEUE); return; }
        LevelNode l;
        l.start = start; l.n = n; l.leftrun = 0; l.pstart = start; l.pleftrun = 0; l.flags = flags; l.tile_base = 0; l.pad = 0;
        lv0[idx] = l;
    } else {
        push_child(Q, st, epoch, start, n, 0, start, 0, flags);
    }
}

}  // namespace

constexpr size_t T2_SMEM = (size_t)T2_CAP * 10;    // 2 x u32 payload + u16 table per slot
constexpr size_t T2B_SMEM = (size_t)T2B_CAP * 10;
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

int blas_t1_coop_occupancy() {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_t1_coop, T1_THREADS, 0) != cudaSuccess) occ = 1;
    return occ < 1 ? 1 : occ;
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
                      uint32_t* n_nodes_out, cudaStream_t stream) {
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
    const uint32_t qb_cap = N / 256 + NM + 4096;  // nodes with 2049..16384 primitives
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
    Task *qb = nullptr, *q = nullptr, *qw = nullptr, *t3 = nullptr, *t4 = nullptr;
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
        pbal = c.take<uint32_t>((size_t)max_tiles * (T1_THREADS / 32) * 9);
        tile_desc = c.take<uint4>(max_tiles);
        barrier = c.take<uint32_t>(64);
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
    BvhCudaBuildStats stats{};

    // The three task queues are contiguous; clearing them makes every `ready` word differ from the epoch even when
    // the workspace still holds other data of an earlier, differently sized build.
    CU_CHECK(ctx, cudaMemsetAsync(qb, 0, (size_t)((char*)(qw + qw_cap) - (char*)qb), stream));
    CU_CHECK(ctx, cudaMemsetAsync(A, 0, sizeof(uint32_t) * scan_n, stream));
    CU_CHECK(ctx, cudaMemsetAsync(recs, 0, sizeof(uint4) * 3 * 2 * (size_t)N, stream));
    const bool prof = ctx->profiling;
    if (prof) cudaEventRecord(ctx->ev[0], stream);
    Queues Q{qb, q, qw, t3, t4, qb_cap, q_cap, qw_cap, t3_cap, t4_cap};
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
    if (N > (uint32_t)T2B_CAP) {
        CU_CHECK(ctx, cudaMemsetAsync(barrier, 0, 256, stream));
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
        void* args[] = {&g, &lv0, &lv1, &lv_cap, &Q, &recs_p, &A_p, &ep, &max_levels};
        const uint32_t grid = t1_grid;
        CU_CHECK(ctx, cudaLaunchCooperativeKernel((const void*)k_t1_coop, dim3(grid), dim3(T1_THREADS), args, 0, stream));
        launches += 1;
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
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_pin, st, sizeof(BuildState), cudaMemcpyDeviceToHost, stream));
    CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_pin + 64, scan_total, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CU_CHECK(ctx, cudaStreamSynchronize(stream));
    CU_CHECK(ctx, cudaGetLastError());
    const BuildState* hs = reinterpret_cast<const BuildState*>(ctx->h_pin);
    ctx->launches += launches;
    ctx->d_last_order = ids0;
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
    stats.grid_nodes = hs->grid_nodes;
    stats.grid_interior_prims = hs->sum_grid;
    if (prof) {
        cudaEventElapsedTime(&stats.ms_setup, ctx->ev[0], ctx->ev[1]);
        cudaEventElapsedTime(&stats.ms_grid, ctx->ev[1], ctx->ev[2]);
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
