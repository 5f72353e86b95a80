// blas_cluster.cuh -- part of blas_build.cu (included there, inside its anonymous namespace; not a stand-alone header):
// the cluster tier k_tc: one node of 16 385 .. 262 144 primitives per thread-block CLUSTER, tasks from a device queue.
//
// Why it exists.  The grid tier walks the tree level by level with ~52 software grid barriers per level (444 blocks on
// one counter, 2-4 us each including the wait for the slowest block), so every level costs 300-400 us however few nodes
// it holds, and ten of the dragon-class build's levels live there.  A node of this size does not need the whole grid:
// a cluster of 16 CTAs x 1024 threads gives it 16 K threads, a HARDWARE barrier (barrier.cluster, ~0.2 us) between the
// phases of a shuffle, and independence from every other node (no level synchronisation: a cluster pops the next task
// as soon as its own node is done).
//
// Data path.  The node's current order lives in global memory (L2-resident) as packed 8-byte items
// {triangle id, plane counts | special << 15}, shuffled IN PLACE: phase A stages the CTA's slice in shared memory while
// it counts, so phase C can overwrite the global range directly.  In place makes the "suffix only" rule free: plane b of
// an axis leaves [0, pivot_{b-1}) untouched (the front cursor of partition_shuffle walks over an all-left prefix without
// a swap, blas.rs:172-174), so shuffle b runs on [pivot_{b-1}, n) alone and the active range is re-spread over all
// threads of the cluster — fewer slots per thread on every later plane of an axis.
// Distributed shared memory carries only the small exchanges (per-CTA counts, partial bounds); the 4-byte scatter
// itself goes through L2, which moves scattered sectors far faster than the ~17 B/cycle/SM of DSMEM
// (scripts/micro/dsmem_bench.cu).
#pragma once

namespace cg = cooperative_groups;  // <cooperative_groups.h> is included by blas_build.cu, outside its anonymous namespace

constexpr int TC_THREADS = 1024;
constexpr int TC_EMAX = 16;                                  // slots per thread at the largest node of the largest cluster
constexpr int TC_SLOTS = TC_THREADS * TC_EMAX;               // slots one CTA can stage (128 KB of shared memory)
constexpr int TC_CLUSTER = 16;                               // CTAs per cluster (non-portable size, as k_tlas_chain_cluster)
constexpr uint32_t TC_CAP = (uint32_t)TC_SLOTS * TC_CLUSTER; // 262 144
constexpr size_t TC_SMEM = sizeof(unsigned long long) * TC_SLOTS;
#define TC_SPECIAL (1ull << 47)

#ifdef BVH_TC_TIMING
// debug build only: one row per node {n | cluster << 32, t_pop, t_bounds, t_flags, t_cand, t_select, t_end, start}
__device__ unsigned long long g_tc_log[4096][8];
__device__ unsigned int g_tc_logn;
__device__ __forceinline__ unsigned long long tc_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define TC_STAMP(k) do { if (rank == 0 && tid == 0) tc_t[k] = tc_now(); } while (0)
// per-phase cycle sums inside the shuffles (thread 0 of every rank-0 CTA): [A, sync1, B, sync2, C, sync3, shuffles]
__device__ unsigned long long g_tc_phase[8];
#define TC_PH_BEGIN unsigned long long ph_t = clock64()
#define TC_PH(k) do { if (rank == 0 && tid == 0) { const unsigned long long ph_n = clock64(); atomicAdd(&g_tc_phase[k], ph_n - ph_t); ph_t = ph_n; if ((k) == 5) atomicAdd(&g_tc_phase[6], 1ull); } } while (0)
#else
#define TC_PH_BEGIN do { } while (0)
#define TC_PH(k) do { } while (0)
#define TC_STAMP(k) do { } while (0)
#endif

struct TcScratch {  // one per resident cluster, global memory
    uint32_t bins[3][8][6];
    uint32_t piv[21], uid[21];
    uint32_t best, pad;
    unsigned long long zkey[6];
};

__device__ __forceinline__ void push_cluster(const Queues& Q, BuildState* st, uint32_t epoch, uint32_t start, uint32_t n,
                                             uint32_t leftrun, uint32_t pstart, uint32_t pleftrun, uint32_t flags) {
    atomicAdd(&st->c_pending, 1u);
    const uint32_t idx = atomicAdd(&st->c_tail, 1u);
    if (idx >= Q.qc_cap) {
        atomicOr(&st->err, DERR_QUEUE);
        atomicSub(&st->c_pending, 1u);
        return;
    }
    Task* d = Q.qc + idx;
    d->start = start; d->n = n; d->leftrun = leftrun; d->pstart = pstart; d->pleftrun = pleftrun;
    d->flags = flags; d->pad = 0;
    __threadfence();
    *(volatile uint32_t*)&d->ready = epoch;
}

// Routes a child (or a root) to the tier of its size.  Nodes above Q.tc_cap belong to the grid tier and never come here.
__device__ __forceinline__ void push_any(const Queues& Q, BuildState* st, uint32_t epoch, uint32_t start, uint32_t n,
                                         uint32_t leftrun, uint32_t pstart, uint32_t pleftrun, uint32_t flags) {
    if (n > (uint32_t)T2B_CAP) push_cluster(Q, st, epoch, start, n, leftrun, pstart, pleftrun, flags);
    else push_child(Q, st, epoch, start, n, leftrun, pstart, pleftrun, flags);
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_tc(Queues Q, uint32_t* ids, unsigned long long* items, uint32_t* table,
                                                      const float4* __restrict__ cent, const float4* __restrict__ box,
                                                      uint4* recs, uint32_t* A, TcScratch* scratch, BuildState* st,
                                                      uint32_t epoch) {
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t C = cluster.num_blocks(), rank = cluster.block_rank();
    TcScratch* const sc = scratch + blockIdx.x / C;
    extern __shared__ unsigned long long s_item[];  // [TC_SLOTS] the CTA's slice of the shuffle in flight
    __shared__ uint32_t s_bal[TC_SLOTS / 32];       // L ballots of the slice
    __shared__ uint32_t s_wcnt[TC_THREADS / 32];
    __shared__ uint32_t s_cnt[TC_CLUSTER];          // per-CTA L counts of the shuffle in flight (written by the peers)
    __shared__ uint32_t s_part[TC_CLUSTER][12];     // per-CTA partial bounds (written by the peers)
    __shared__ uint32_t s_red[TC_THREADS / 32][12];
    __shared__ uint32_t s_node[12];                 // ordered-uint: vlo[3], vhi[3], cmin[3], cmax[3]
    __shared__ uint32_t s_bins[3][8][6];
    __shared__ Task s_task;
    __shared__ int s_have;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t CT = C * TC_THREADS;

#ifdef BVH_TC_TIMING
    unsigned long long tc_t[8];
#endif
    for (;;) {
        // ---- pop (one thread of the cluster), scratch reset ----
        TC_STAMP(0);
        if (rank == 0) {
            if (tid == 0) {
                uint32_t idx = 0;
                const bool have = queue_pop(Q.qc, Q.qc_cap, &st->c_head, &st->c_tail, &st->c_pending, st, epoch, &idx);
                if (have) {
                    const volatile Task* vq = Q.qc + idx;
                    s_task.start = vq->start; s_task.n = vq->n; s_task.leftrun = vq->leftrun;
                    s_task.pstart = vq->pstart; s_task.pleftrun = vq->pleftrun; s_task.flags = vq->flags;
                }
                s_have = have ? 1 : 0;
            }
            if (tid < 144) (&sc->bins[0][0][0])[tid] = ((tid % 6) < 3) ? ENC_POS_INIT : ENC_NEG_INIT;
            if (tid >= 160 && tid < 166) sc->zkey[tid - 160] = 0xFFFFFFFFFFFFFFFFull;
            if (tid == 166) sc->best = 0xFFFFFFFFu;
        }
        cluster.sync();
        if (tid == 0 && rank != 0) {
            const int* rh = cluster.map_shared_rank(&s_have, 0);
            const Task* rt = cluster.map_shared_rank(&s_task, 0);
            s_have = *rh;
            s_task = *rt;
        }
        cluster.sync();  // also keeps rank 0 from leaving (or re-popping) while its task is still being read
        if (!s_have) break;
        const Task t = s_task;
        const uint32_t n = t.n, start = t.start;
        TC_STAMP(1);
        const uint32_t E0 = (n + CT - 1) / CT;            // slots per thread when the whole node is active
        const uint32_t base0 = rank * TC_THREADS * E0 + warp * 32 * E0;

        // ---- 1. own vertex box, centroid bounds (blas.rs:87-88,117-123,142) ----
        {
            float acc[12] = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f, 1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll 4
            for (uint32_t i = 0; i < E0; ++i) {
                const uint32_t j = base0 + i * 32 + lane;
                if (j < n) {
                    const uint32_t g = __ldcg(&ids[start + j]);
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
        if (tid < 12 * C) {  // thread (k, peer): reduce value k over my warps, hand it to CTA `peer`
            const uint32_t k = tid % 12, peer = tid / 12;
            const bool is_min = (k < 3) || (k >= 6 && k < 9);
            uint32_t r = s_red[0][k];
            for (int w2 = 1; w2 < TC_THREADS / 32; ++w2) r = is_min ? min(r, s_red[w2][k]) : max(r, s_red[w2][k]);
            *cluster.map_shared_rank(&s_part[rank][k], peer) = r;
        }
        cluster.sync();
        if (tid < 12) {
            const bool is_min = (tid < 3) || (tid >= 6 && tid < 9);
            uint32_t r = s_part[0][tid];
            for (uint32_t r2 = 1; r2 < C; ++r2) r = is_min ? min(r, s_part[r2][tid]) : max(r, s_part[r2][tid]);
            s_node[tid] = r;
        }
        __syncthreads();

        TC_STAMP(2);
        // ---- 2. plane counts -> packed items (in the slots this CTA reads back in the first shuffle) ----
        {
            float cmin[3], cmax[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) { cmin[c] = o2f(s_node[6 + c]); cmax[c] = o2f(s_node[9 + c]); }
            bool zero_face[6];
            bool any_zero = false;
            if (st->neg_zero) {
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const uint32_t e = s_node[c];
                    zero_face[c] = o2f(c < 3 ? min(e, ENC_POS_INIT) : max(e, ENC_NEG_INIT)) == 0.0f;
                    any_zero = any_zero || zero_face[c];
                }
            }
#pragma unroll 4
            for (uint32_t i = 0; i < E0; ++i) {
                const uint32_t j = base0 + i * 32 + lane;
                if (j < n) {
                    const uint32_t g = __ldcg(&ids[start + j]);
                    const float4 c = cent[g];
                    const uint32_t kb = plane_counts(c.x, c.y, c.z, cmin, cmax);
                    items[start + j] = (unsigned long long)g | ((unsigned long long)kb << 32);
                    if (any_zero) {  // rare path (-0.0 in the input): first slot with a zero on each zero-valued face
                        const float4 b0 = box[2 * (size_t)g], b1 = box[2 * (size_t)g + 1];
                        const float vals[6] = {b0.x, b0.y, b0.z, b1.x, b1.y, b1.z};
#pragma unroll
                        for (int cc = 0; cc < 6; ++cc)
                            if (zero_face[cc] && vals[cc] == 0.0f) atomicMin(&sc->zkey[cc], ((unsigned long long)j << 32) | g);
                    }
                }
            }
        }
        __syncthreads();  // shuffle 0 re-reads exactly the slots this CTA wrote (same layout, s0 = 0)

        // ---- 3. partition_shuffle (blas.rs:168-182) in closed form on [s0, n), in place ----
        // cidx >= 0: candidate number (records pivot and unexamined element); fin: the winner's re-shuffle (blas.rs:164)
        auto shuffle = [&](const uint32_t a, const uint32_t b, const uint32_t s0, const int cidx, const bool fin) {
            const uint32_t act = n - s0;
            const uint32_t E = (act + CT - 1) / CT;
            const uint32_t wbase = rank * TC_THREADS * E + warp * 32 * E;  // first index (relative to s0) of my warp
            const uint32_t lbase = warp * 32 * E;                           // same, inside the CTA's staged slice
            const uint32_t sh = 32 + 3 * a;
            unsigned long long* const g_it = items + start + s0;
            uint32_t* const g_tab = table + start + s0;
            // A: stage the slice (asynchronous copies: all of a thread's loads are in flight at once, no registers held),
            //    then ballots and counts from shared memory
            {
                const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(s_item + lbase);
                for (uint32_t i = 0; i < E; ++i) {
                    const uint32_t idx = wbase + i * 32 + lane;
                    if (idx < act)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s_base + 8u * (i * 32 + lane)), "l"(g_it + idx) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();  // a warp only reads what its own lanes copied
            }
            uint32_t cnt = 0;
#pragma unroll 4
            for (uint32_t i = 0; i < E; ++i) {
                const uint32_t idx = wbase + i * 32 + lane;
                const unsigned long long it = (idx < act) ? s_item[lbase + i * 32 + lane] : 0ull;
                const bool L = (idx < act) && ((uint32_t)(it >> sh) & 7u) < b;
                const uint32_t bal = __ballot_sync(FULL_MASK, L);
                if (lane == 0) s_bal[warp * E + i] = bal;
                cnt += __popc(bal);
            }
            if (lane == 0) s_wcnt[warp] = cnt;
            __syncthreads();
            const uint32_t wv = s_wcnt[lane];
            const uint32_t tot = __reduce_add_sync(FULL_MASK, wv);
            const uint32_t wpre = __reduce_add_sync(FULL_MASK, lane < warp ? wv : 0u);
            if (tid < C) *cluster.map_shared_rank(&s_cnt[rank], tid) = tot;
            cluster.sync();  // #1: every CTA's L count is in every CTA's s_cnt
            const uint32_t cv = (lane < C) ? s_cnt[lane] : 0u;
            const uint32_t nL = __reduce_add_sync(FULL_MASK, cv);
            const uint32_t lf0 = __reduce_add_sync(FULL_MASK, lane < rank ? cv : 0u) + wpre;  // #L before my warp's first slot
            // boundary element in closed form (see p_t1_table): f is nL-1, nL or nL+1, decided by three flags
            uint32_t lbit = 0;
            if (lane < 3) {
                const uint32_t j = nL + lane - 1u;  // nL-1 (wraps for nL = 0: then j >= act), nL, nL+1
                if (j < act) lbit = (((uint32_t)(__ldcg(&g_it[j]) >> sh) & 7u) < b) ? 1u : 0u;
            }
            // B: rank -> position table; only front R's (idx <= nL) and back L's (idx >= nL) are ever looked up
            {
                uint32_t running = lf0;
#pragma unroll 4
                for (uint32_t i = 0; i < E; ++i) {
                    const uint32_t idx = wbase + i * 32 + lane;
                    const uint32_t bal = s_bal[warp * E + i];
                    const uint32_t LF = running + __popc(bal & lt_mask);
                    if (idx < act) {
                        if ((bal >> lane) & 1u) { if (idx >= nL) g_tab[act - 1 - (nL - LF - 1)] = idx; }
                        else if (idx <= nL) g_tab[idx - LF] = idx;
                    }
                    running += __popc(bal);
                }
            }
            const uint32_t l0 = __shfl_sync(FULL_MASK, lbit, 0), l1 = __shfl_sync(FULL_MASK, lbit, 1), l2 = __shfl_sync(FULL_MASK, lbit, 2);
            uint32_t f, Lf;
            if (nL >= 1 && !(nL + 1 <= act && l0 + l1 <= 1)) { f = nL - 1; Lf = l0; }
            else if (!(nL + 2 <= act && l1 + l2 == 0)) { f = nL; Lf = l1; }
            else { f = nL + 1; Lf = l2; }
            const uint32_t pivot = nL - Lf;
            cluster.sync();  // #2: table complete; nobody reads the global items of this shuffle any more
            // C: scatter in place.  First every destination (the table look-ups of a thread's slots are independent loads,
            //    issued together), then the stores.
            {
                uint32_t dest[TC_EMAX];
                uint32_t running = lf0;
#pragma unroll
                for (int i = 0; i < TC_EMAX; ++i) {
                    if (i >= (int)E) break;  // (a real loop exit: a guard around the body is if-converted and every
                                             //  thread then issues all 16 bodies whatever E is)
                    dest[i] = 0xFFFFFFFFu;
                    const uint32_t idx = wbase + i * 32 + lane;
                    const uint32_t bal = s_bal[warp * E + i];
                    const uint32_t LF = running + __popc(bal & lt_mask);
                    running += __popc(bal);
                    if (idx < act) {
                        const uint32_t Lb = (bal >> lane) & 1u;
                        const uint32_t RF = idx - LF;
                        if (idx < f) dest[i] = Lb ? idx : (RF == 0 ? act : __ldcg(&g_tab[act - RF])) - (Lb ? 0u : 1u);
                        else if (idx == f) dest[i] = pivot;
                        else dest[i] = Lb ? __ldcg(&g_tab[nL - LF - 1]) : idx - 1;
                    }
                }
#pragma unroll
                for (int i = 0; i < TC_EMAX; ++i) {
                    if (i >= (int)E) break;
                    const uint32_t idx = wbase + i * 32 + lane;
                    if (idx < act) {
                        unsigned long long it = s_item[lbase + i * 32 + lane];
                        if (idx == f) {
                            it |= TC_SPECIAL;
                            if (cidx >= 0) { sc->piv[cidx] = s0 + pivot; sc->uid[cidx] = (uint32_t)it; }
                        }
                        if (dest[i] != idx || idx == f) g_it[dest[i]] = it;
                        if (fin) ids[start + s0 + dest[i]] = (uint32_t)it;
                    }
                }
            }
            if (fin) __threadfence();  // the final order is read by other clusters / later tiers
            cluster.sync();  // #3
            return pivot;
        };

        TC_STAMP(3);
        uint32_t s0 = 0;
        for (uint32_t c = 0; c < 21; ++c) {
            const uint32_t b = c % 7 + 1;
            if (b == 1) s0 = 0;  // a new axis starts on the whole range
            s0 += shuffle(c / 7, b, s0, (int)c, false);
        }
        TC_STAMP(4);

        // ---- 4. exact bins over the non-special primitives ----
        for (uint32_t k = tid; k < 144; k += TC_THREADS) (&s_bins[0][0][0])[k] = ((k % 6) < 3) ? ENC_POS_INIT : ENC_NEG_INIT;
        __syncthreads();
        for (uint32_t i0 = 0; i0 < E0; i0 += 4) {
            uint32_t lo[4][3], hi[4][3];
            uint32_t kk[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                const uint32_t i = i0 + ii;
                const uint32_t j = base0 + i * 32 + lane;
                kk[ii] = 0xFFFFFFFFu;
                lo[ii][0] = lo[ii][1] = lo[ii][2] = ENC_POS_INIT;
                hi[ii][0] = hi[ii][1] = hi[ii][2] = ENC_NEG_INIT;
                if (i < E0 && j < n) {
                    const unsigned long long it = __ldcg(&items[start + j]);
                    if (!(it & TC_SPECIAL)) {
                        const uint32_t g = (uint32_t)it;
                        const float4 b0 = box[2 * (size_t)g], b1 = box[2 * (size_t)g + 1];
                        lo[ii][0] = f2o(b0.x); lo[ii][1] = f2o(b0.y); lo[ii][2] = f2o(b0.z);
                        hi[ii][0] = f2o(b1.x); hi[ii][1] = f2o(b1.y); hi[ii][2] = f2o(b1.z);
                        kk[ii] = (uint32_t)(it >> 32) & 0x1FFu;
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
        __syncthreads();
        if (tid < 144) {
            const uint32_t v = (&s_bins[0][0][0])[tid];
            if ((tid % 6) < 3) { if (v != ENC_POS_INIT) atomicMin(&(&sc->bins[0][0][0])[tid], v); }
            else { if (v != ENC_NEG_INIT) atomicMax(&(&sc->bins[0][0][0])[tid], v); }
        }
        cluster.sync();

        // ---- 5. candidate costs and selection (one warp of the cluster; blas.rs:149-161) ----
        if (rank == 0 && warp == 0) {
            float cmin[3], cmax[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) { cmin[c] = o2f(s_node[6 + c]); cmax[c] = o2f(s_node[9 + c]); }
            const uint32_t cs = (lane < 21) ? lane : 0;
            const uint32_t my_uid = __ldcg(&sc->uid[cs]);  // lane c also owns special c
            const float4 ce = cent[my_uid];
            const float4 b0 = box[2 * (size_t)my_uid], b1 = box[2 * (size_t)my_uid + 1];
            const uint32_t my_kb = plane_counts(ce.x, ce.y, ce.z, cmin, cmax);
            const uint32_t a = cs / 7, b = cs % 7 + 1;
            float Lb[6] = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
            float Rb[6] = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
            for (uint32_t k = 0; k < 8; ++k) {
                float* side = (k < b) ? Lb : Rb;
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    side[x] = fminf(side[x], o2f(__ldcg(&sc->bins[a][k][x])));
                    side[3 + x] = fmaxf(side[3 + x], o2f(__ldcg(&sc->bins[a][k][3 + x])));
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
            const uint32_t n1 = __ldcg(&sc->piv[cs]);
            const float cost = sah_cost(Lb, Rb, n1, n - n1);
            const uint32_t key = (lane < 21 && cost < 3.402823466e+38f) ? __float_as_uint(cost) : 0xFFFFFFFFu;
            const uint32_t mk = __reduce_min_sync(FULL_MASK, key);
            const uint32_t bal = __ballot_sync(FULL_MASK, key == mk);
            if (lane == 0) sc->best = (mk == 0xFFFFFFFFu) ? 0xFFFFFFFFu : (uint32_t)(__ffs(bal) - 1);
        }
        cluster.sync();
        const uint32_t best = __ldcg(&sc->best);
        if (best == 0xFFFFFFFFu) {  // no candidate with a finite cost: the reference would not terminate (blas.rs:115,139)
            if (rank == 0 && tid == 0) {
                atomicOr(&st->err, DERR_DEGENERATE);
                atomicSub(&st->c_pending, 1u);
            }
            continue;
        }

        // ---- 6. final shuffle (blas.rs:164): writes the order back to ids ----
        TC_STAMP(5);
        shuffle(best / 7, best % 7 + 1, 0, -1, true);
        TC_STAMP(6);

        // ---- 7. record + children ----
        if (rank == 0 && tid == 0) {
            const uint32_t p = __ldcg(&sc->piv[best]);  // the pivot recorded when the candidate was evaluated (blas.rs:159,165)
            float lo[3], hi[3];
            for (int c = 0; c < 3; ++c) {
                lo[c] = o2f(min(s_node[c], ENC_POS_INIT));
                hi[c] = o2f(max(s_node[3 + c], ENC_NEG_INIT));
            }
            for (int c = 0; c < 6; ++c) {
                const unsigned long long zk = __ldcg(&sc->zkey[c]);
                if (zk != 0xFFFFFFFFFFFFFFFFull) {  // only set on the rare -0.0 path
                    const uint32_t g = (uint32_t)(zk & 0xFFFFFFFFull);
                    const float4 bb = box[2 * (size_t)g + (c < 3 ? 0 : 1)];
                    const float z = (c % 3 == 0) ? bb.x : ((c % 3 == 1) ? bb.y : bb.z);
                    if (c < 3) lo[c] = z; else hi[c - 3] = z;
                }
            }
            if (p == 0 || p >= n) {
                atomicOr(&st->err, DERR_DEGENERATE);
            } else {
                emit_rec(recs, 2 * (start + p) + 1, lo, hi, start, n, t.leftrun, t.pstart, t.pleftrun, t.flags);
                if (p <= 3) A[start] = t.leftrun + 1;
                push_any(Q, st, epoch, start, p, t.leftrun + 1, start, t.leftrun, t.flags & ~3u);
                push_any(Q, st, epoch, start + p, n - p, 0, start, t.leftrun, TF_RIGHT | (t.flags & ~3u));
                atomicAdd(&st->tc_done, 1u);
                atomicAdd(&st->grid_nodes, 1u);
                atomicAdd(&st->sum_grid, (unsigned long long)n);
            }
            __threadfence();
            atomicSub(&st->c_pending, 1u);
#ifdef BVH_TC_TIMING
            {
                const unsigned int row = atomicAdd(&g_tc_logn, 1u);
                if (row < 4096) {
                    g_tc_log[row][0] = (unsigned long long)n | ((unsigned long long)(blockIdx.x / C) << 32);
                    for (int k = 0; k < 7; ++k) g_tc_log[row][1 + k] = tc_t[k];
                }
            }
#endif
        }
        __syncthreads();  // rank 0 resets the scratch at the top of the next iteration: not before thread 0 has read it
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// k_tcs: the same tier with the node's order resident in the cluster's DISTRIBUTED SHARED MEMORY.
//
// k_tc above moves the 8-byte items through L2, and every phase boundary of a shuffle then waits for global stores to
// drain before the cluster barrier can release (measured: ~4.3 us fixed per shuffle + 0.55 us per slot a thread owns,
// profiles/r02d_tc_timing.log: 135 us for a 16 K node whose 22 shuffles move 64 KB each).  Here slot j of the node lives
// in the shared memory of CTA j / span for the whole node (static ownership), payload = local primitive index | plane
// counts << 18 | special << 31, and the rank -> position table is distributed the same way; the scatter, the table and
// its look-ups are st/ld.shared::cluster, so a phase boundary is a barrier.cluster that only has to wait for DSMEM.
// In place: a thread parks the payloads of its own slots in a staging array from the counting pass to the scatter.
// The suffix rule is kept (a shuffle only touches [s0, n)), but ownership is static, so it saves work, not latency.
// ------------------------------------------------------------------------------------------------------------------------
// distributed-shared-memory accesses as single instructions (cg::cluster_group::map_shared_rank goes through generic
// addresses and costs ~13 SASS instructions per access; these are mapa + st/ld.shared::cluster)
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local_saddr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void dsmem_st(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t dsmem_ld(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

constexpr size_t TS_SMEM = sizeof(uint32_t) * 3 * TC_SLOTS;  // payload + table + staging, 192 KB
#define TS_IDX_MASK 0x3FFFFu /* 18 bits: TC_CAP = 2^18 local primitives */
#define TS_KB_SHIFT 18

__global__ void __launch_bounds__(TC_THREADS, 1) k_tcs(Queues Q, uint32_t* ids, uint32_t* ids_snap, const float4* __restrict__ cent,
                                                       const float4* __restrict__ box, uint4* recs, uint32_t* A, BuildState* st,
                                                       uint32_t epoch) {
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t C = cluster.num_blocks(), rank = cluster.block_rank();
    extern __shared__ uint32_t s_dyn_ts[];
    uint32_t* const s_pay = s_dyn_ts;             // [TC_SLOTS] my slots of the node's current order
    uint32_t* const s_tab = s_dyn_ts + TC_SLOTS;  // [TC_SLOTS] my slice of the rank -> position table
    uint32_t* const s_stage = s_dyn_ts + 2 * TC_SLOTS;  // [TC_SLOTS] payloads of my slots while a shuffle is in flight
    __shared__ uint32_t s_bal[TC_SLOTS / 32];
    __shared__ uint32_t s_wcnt[TC_THREADS / 32];
    __shared__ uint32_t s_cnt[TC_CLUSTER];
    __shared__ uint32_t s_part[TC_CLUSTER][12];
    __shared__ uint32_t s_red[TC_THREADS / 32][12];
    __shared__ uint32_t s_node[12];
    __shared__ uint32_t s_bins[3][8][6];
    __shared__ uint32_t s_binpart[TC_CLUSTER][144];          // rank 0: every CTA's bins
    __shared__ unsigned long long s_zk[6];                    // rare -0.0 path: my CTA's first zero slot per face
    __shared__ unsigned long long s_zpart[TC_CLUSTER][6];     // rank 0: every CTA's
    __shared__ uint32_t s_u[21], s_piv[21], s_ukb[21];        // rank 0: unexamined element / pivot of every candidate
    __shared__ float s_ubox[21][6];
    __shared__ uint32_t s_best;
    __shared__ Task s_task;
    __shared__ int s_have;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t CT = C * TC_THREADS;
#ifdef BVH_TC_TIMING
    unsigned long long tc_t[8];
#endif

    for (;;) {
        TC_STAMP(0);
        if (rank == 0 && tid == 0) {
            uint32_t idx = 0;
            const bool have = queue_pop(Q.qc, Q.qc_cap, &st->c_head, &st->c_tail, &st->c_pending, st, epoch, &idx);
            if (have) {
                const volatile Task* vq = Q.qc + idx;
                s_task.start = vq->start; s_task.n = vq->n; s_task.leftrun = vq->leftrun;
                s_task.pstart = vq->pstart; s_task.pleftrun = vq->pleftrun; s_task.flags = vq->flags;
            }
            s_have = have ? 1 : 0;
        }
        cluster.sync();
        if (tid == 0 && rank != 0) {
            s_have = *cluster.map_shared_rank(&s_have, 0);
            s_task = *cluster.map_shared_rank(&s_task, 0);
        }
        if (tid < 6) s_zk[tid] = 0xFFFFFFFFFFFFFFFFull;
        cluster.sync();  // also keeps rank 0 from leaving (or re-popping) while its task is still being read
        if (!s_have) break;
        const Task t = s_task;
        const uint32_t n = t.n, start = t.start;
        TC_STAMP(1);
        const uint32_t E0 = (n + CT - 1) / CT;        // slots per thread
        const uint32_t span = TC_THREADS * E0;        // slots per CTA: slot j lives in CTA j / span at s_pay[j % span]
        const uint32_t magic = (uint32_t)((0x100000000ull + E0 - 1) / E0);  // (x * magic) >> 32 == x / E0 for x < 2^16
        // Rows (32 slots) per warp: at least 4, so that a small node keeps 8-24 warps busy with 4 rows each instead of 32
        // warps with one row and the same fixed cost per warp and phase; the other warps only join the barriers.
        const uint32_t RWn = E0 < 4 ? 4u : E0;
        const uint32_t RW = (warp * RWn < 32 * E0) ? RWn : 0u;   // rows of my warp (0: idle for this node)
        const uint32_t lbase = warp * 32 * RWn;       // my warp's first slot inside the CTA
        const uint32_t base0 = rank * span + lbase;   // ... and inside the node
        const uint32_t sa_pay = (uint32_t)__cvta_generic_to_shared(s_pay), sa_tab = (uint32_t)__cvta_generic_to_shared(s_tab);
        // owner CTA and local index of node slot a (a < C * span)
        auto locate = [&](uint32_t a, uint32_t& owner, uint32_t& loc) {
            owner = (E0 == 1) ? (a >> 10) : __umulhi(a >> 10, magic);  // (E0 == 1: the magic number would be 2^32)
            loc = a - owner * span;
        };

        // ---- 1. snapshot the order, own vertex box, centroid bounds ----
        {
            float acc[12] = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f, 1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll 4
            for (uint32_t i = 0; i < RW; ++i) {
                const uint32_t j = base0 + i * 32 + lane;
                if (j < n) {
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
        if (tid < 12 * C) {  // thread (k, peer): reduce value k over my warps, hand it to CTA `peer`
            const uint32_t k = tid % 12, peer = tid / 12;
            const bool is_min = (k < 3) || (k >= 6 && k < 9);
            uint32_t r = s_red[0][k];
            for (int w2 = 1; w2 < TC_THREADS / 32; ++w2) r = is_min ? min(r, s_red[w2][k]) : max(r, s_red[w2][k]);
            *cluster.map_shared_rank(&s_part[rank][k], peer) = r;
        }
        cluster.sync();
        if (tid < 12) {
            const bool is_min = (tid < 3) || (tid >= 6 && tid < 9);
            uint32_t r = s_part[0][tid];
            for (uint32_t r2 = 1; r2 < C; ++r2) r = is_min ? min(r, s_part[r2][tid]) : max(r, s_part[r2][tid]);
            s_node[tid] = r;
        }
        __syncthreads();
        TC_STAMP(2);

        // ---- 2. plane counts -> payload of my own slots ----
        {
            float cmin[3], cmax[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) { cmin[c] = o2f(s_node[6 + c]); cmax[c] = o2f(s_node[9 + c]); }
            bool zero_face[6];
            bool any_zero = false;
            if (st->neg_zero) {
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const uint32_t e = s_node[c];
                    zero_face[c] = o2f(c < 3 ? min(e, ENC_POS_INIT) : max(e, ENC_NEG_INIT)) == 0.0f;
                    any_zero = any_zero || zero_face[c];
                }
            }
#pragma unroll 4
            for (uint32_t i = 0; i < RW; ++i) {
                const uint32_t j = base0 + i * 32 + lane;
                if (j < n) {
                    const uint32_t g = ids_snap[start + j];  // written by this very thread
                    const float4 c = cent[g];
                    s_pay[lbase + i * 32 + lane] = j | (plane_counts(c.x, c.y, c.z, cmin, cmax) << TS_KB_SHIFT);
                    if (any_zero) {  // rare path (-0.0 in the input): first slot with a zero on each zero-valued face
                        const float4 b0 = box[2 * (size_t)g], b1 = box[2 * (size_t)g + 1];
                        const float vals[6] = {b0.x, b0.y, b0.z, b1.x, b1.y, b1.z};
#pragma unroll
                        for (int cc = 0; cc < 6; ++cc)
                            if (zero_face[cc] && vals[cc] == 0.0f) atomicMin(&s_zk[cc], ((unsigned long long)j << 32) | g);
                    }
                }
            }
            if (any_zero) {
                __syncthreads();
                if (tid < 6) *cluster.map_shared_rank(&s_zpart[rank][tid], 0) = s_zk[tid];
            } else if (tid < 6) *cluster.map_shared_rank(&s_zpart[rank][tid], 0) = 0xFFFFFFFFFFFFFFFFull;
        }
        TC_STAMP(3);

        // ---- 3. partition_shuffle (blas.rs:168-182) in closed form on [s0, n), in place, through DSMEM ----
        auto shuffle = [&](const uint32_t a, const uint32_t b, const uint32_t s0, const int cidx) -> uint32_t {
            const uint32_t act = n - s0;
            const uint32_t sh = TS_KB_SHIFT + 3 * a;
            uint32_t cnt = 0;
            TC_PH_BEGIN;
            // A: ballots and counts of my slots; their payloads are parked in s_stage (thread-private use) so that the
            //    scatter can overwrite s_pay in place
            for (uint32_t i = 0; i < RW; ++i) {
                const uint32_t j = base0 + i * 32 + lane;
                const bool valid = (j < n) && (j >= s0);
                const uint32_t pv = valid ? s_pay[lbase + i * 32 + lane] : 0u;
                s_stage[lbase + i * 32 + lane] = pv;
                const bool L = valid && (((pv >> sh) & 7u) < b);
                const uint32_t bal = __ballot_sync(FULL_MASK, L);
                if (lane == 0) s_bal[warp * RWn + i] = bal;
                cnt += __popc(bal);
            }
            if (lane == 0) s_wcnt[warp] = cnt;
            __syncthreads();
            const uint32_t wv = s_wcnt[lane];
            const uint32_t tot = __reduce_add_sync(FULL_MASK, wv);
            const uint32_t wpre = __reduce_add_sync(FULL_MASK, lane < warp ? wv : 0u);
            if (tid < C) *cluster.map_shared_rank(&s_cnt[rank], tid) = tot;
            TC_PH(0);
            cluster.sync();  // #1
            TC_PH(1);
            const uint32_t cv = (lane < C) ? s_cnt[lane] : 0u;
            const uint32_t nL = __reduce_add_sync(FULL_MASK, cv);
            const uint32_t lf0 = __reduce_add_sync(FULL_MASK, lane < rank ? cv : 0u) + wpre;  // #L of [s0, n) before my warp's first slot
            // boundary element in closed form (see p_t1_table): f is nL-1, nL or nL+1, decided by three flags
            uint32_t lbit = 0;
            if (lane < 3) {
                const uint32_t x = nL + lane - 1u;  // relative to s0; wraps for nL = 0 and is then >= act
                if (x < act) {
                    uint32_t owner, loc;
                    locate(s0 + x, owner, loc);
                    lbit = (((dsmem_ld(dsmem_addr(sa_pay + 4u * loc, owner)) >> sh) & 7u) < b) ? 1u : 0u;
                }
            }
            // B: rank -> position table (indices relative to s0), entry t kept with slot s0 + t
            {
                uint32_t running = lf0;
                for (uint32_t i = 0; i < RW; ++i) {
                    const uint32_t j = base0 + i * 32 + lane;
                    const uint32_t bal = s_bal[warp * RWn + i];
                    const uint32_t LF = running + __popc(bal & lt_mask);
                    running += __popc(bal);
                    if (j < n && j >= s0) {
                        const uint32_t idx = j - s0;
                        uint32_t tpos = 0xFFFFFFFFu;
                        if ((bal >> lane) & 1u) { if (idx >= nL) tpos = act - 1 - (nL - LF - 1); }
                        else if (idx <= nL) tpos = idx - LF;
                        if (tpos != 0xFFFFFFFFu) {
                            uint32_t owner, loc;
                            locate(s0 + tpos, owner, loc);
                            dsmem_st(dsmem_addr(sa_tab + 4u * loc, owner), idx);
                        }
                    }
                }
            }
            const uint32_t l0 = __shfl_sync(FULL_MASK, lbit, 0), l1 = __shfl_sync(FULL_MASK, lbit, 1), l2 = __shfl_sync(FULL_MASK, lbit, 2);
            uint32_t f, Lf;
            if (nL >= 1 && !(nL + 1 <= act && l0 + l1 <= 1)) { f = nL - 1; Lf = l0; }
            else if (!(nL + 2 <= act && l1 + l2 == 0)) { f = nL; Lf = l1; }
            else { f = nL + 1; Lf = l2; }
            const uint32_t pivot = nL - Lf;
            TC_PH(2);
            cluster.sync();  // #2: table complete; every payload of [s0, n) is in its owner's registers
            TC_PH(3);
            // C: four slots at a time: destinations first (independent remote look-ups), then the stores
            {
                uint32_t running = lf0;
                for (uint32_t i0 = 0; i0 < RW; i0 += 4) {
                    uint32_t dest[4];
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii) {
                        dest[ii] = 0xFFFFFFFFu;
                        const uint32_t i = i0 + ii;
                        if (i < RW) {
                            const uint32_t j = base0 + i * 32 + lane;
                            const uint32_t bal = s_bal[warp * RWn + i];
                            const uint32_t LF = running + __popc(bal & lt_mask);
                            running += __popc(bal);
                            if (j < n && j >= s0) {
                                const uint32_t idx = j - s0;
                                const uint32_t Lb = (bal >> lane) & 1u;
                                const uint32_t RF = idx - LF;
                                uint32_t look = 0xFFFFFFFFu;  // table entry to fetch, if any
                                if (idx < f) { if (Lb) dest[ii] = idx; else if (RF == 0) dest[ii] = act - 1; else look = act - RF; }
                                else if (idx == f) dest[ii] = pivot;
                                else if (Lb) look = nL - LF - 1;
                                else dest[ii] = idx - 1;
                                if (look != 0xFFFFFFFFu) {
                                    uint32_t owner, loc;
                                    locate(s0 + look, owner, loc);
                                    const uint32_t v = dsmem_ld(dsmem_addr(sa_tab + 4u * loc, owner));
                                    dest[ii] = Lb ? v : v - 1u;
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii) {
                        const uint32_t i = i0 + ii;
                        if (i < RW && dest[ii] != 0xFFFFFFFFu) {
                            const uint32_t idx = base0 + i * 32 + lane - s0;
                            uint32_t pv = s_stage[lbase + i * 32 + lane];
                            if (idx == f) {
                                pv |= 0x80000000u;
                                if (cidx >= 0) {  // candidate record, kept by rank 0
                                    *cluster.map_shared_rank(&s_u[cidx], 0) = pv & TS_IDX_MASK;
                                    *cluster.map_shared_rank(&s_ukb[cidx], 0) = (pv >> TS_KB_SHIFT) & 0x1FFu;
                                    *cluster.map_shared_rank(&s_piv[cidx], 0) = s0 + pivot;
                                }
                            }
                            if (dest[ii] != idx || idx == f) {
                                uint32_t owner, loc;
                                locate(s0 + dest[ii], owner, loc);
                                dsmem_st(dsmem_addr(sa_pay + 4u * loc, owner), pv);
                            }
                        }
                    }
                }
            }
            TC_PH(4);
            cluster.sync();  // #3
            TC_PH(5);
            return pivot;
        };

        {
            uint32_t s0 = 0;
            for (uint32_t c = 0; c < 21; ++c) {
                const uint32_t b = c % 7 + 1;
                if (b == 1) s0 = 0;  // a new axis starts on the whole range
                s0 += shuffle(c / 7, b, s0, (int)c);
            }
        }
        TC_STAMP(4);

        // ---- 4. exact bins over the non-special primitives of my slots ----
        for (uint32_t k = tid; k < 144; k += TC_THREADS) (&s_bins[0][0][0])[k] = ((k % 6) < 3) ? ENC_POS_INIT : ENC_NEG_INIT;
        __syncthreads();
        for (uint32_t i0 = 0; i0 < RW; i0 += 4) {
            uint32_t lo[4][3], hi[4][3];
            uint32_t kk[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                const uint32_t i = i0 + ii;
                const uint32_t j = base0 + i * 32 + lane;
                kk[ii] = 0xFFFFFFFFu;
                lo[ii][0] = lo[ii][1] = lo[ii][2] = ENC_POS_INIT;
                hi[ii][0] = hi[ii][1] = hi[ii][2] = ENC_NEG_INIT;
                if (i < RW && j < n) {
                    const uint32_t pv = s_pay[lbase + i * 32 + lane];
                    if (!(pv & 0x80000000u)) {
                        const uint32_t g = __ldcg(&ids_snap[start + (pv & TS_IDX_MASK)]);
                        const float4 b0 = box[2 * (size_t)g], b1 = box[2 * (size_t)g + 1];
                        lo[ii][0] = f2o(b0.x); lo[ii][1] = f2o(b0.y); lo[ii][2] = f2o(b0.z);
                        hi[ii][0] = f2o(b1.x); hi[ii][1] = f2o(b1.y); hi[ii][2] = f2o(b1.z);
                        kk[ii] = (pv >> TS_KB_SHIFT) & 0x1FFu;
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
        __syncthreads();
        if (tid < 144) *cluster.map_shared_rank(&s_binpart[rank][tid], 0) = (&s_bins[0][0][0])[tid];
        cluster.sync();

        // ---- 5. candidate costs and selection (rank 0; blas.rs:149-161) ----
        if (rank == 0) {
            if (tid < 144) {
                const bool is_min = (tid % 6) < 3;
                uint32_t r = s_binpart[0][tid];
                for (uint32_t r2 = 1; r2 < C; ++r2) r = is_min ? min(r, s_binpart[r2][tid]) : max(r, s_binpart[r2][tid]);
                (&s_bins[0][0][0])[tid] = r;
            }
            if (tid >= 160 && tid < 181) {  // boxes of the 21 unexamined elements
                const uint32_t c = tid - 160;
                const uint32_t g = __ldcg(&ids_snap[start + s_u[c]]);
                const float4 b0 = box[2 * (size_t)g], b1 = box[2 * (size_t)g + 1];
                s_ubox[c][0] = b0.x; s_ubox[c][1] = b0.y; s_ubox[c][2] = b0.z;
                s_ubox[c][3] = b1.x; s_ubox[c][4] = b1.y; s_ubox[c][5] = b1.z;
            }
            __syncthreads();
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
                    const bool left = (s_u[s2] != myu) && (((s_ukb[s2] >> (3 * a)) & 7u) < b);
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
                if (lane < C) *cluster.map_shared_rank(&s_best, lane) = win;
            }
        }
        cluster.sync();
        const uint32_t best = s_best;
        TC_STAMP(5);
        if (best == 0xFFFFFFFFu) {  // no candidate with a finite cost: the reference would not terminate (blas.rs:115,139)
            if (rank == 0 && tid == 0) {
                atomicOr(&st->err, DERR_DEGENERATE);
                atomicSub(&st->c_pending, 1u);
            }
            continue;
        }

        // ---- 6. final shuffle (blas.rs:164), write the order back ----
        shuffle(best / 7, best % 7 + 1, 0, -1);
#pragma unroll 4
        for (uint32_t i = 0; i < RW; ++i) {
            const uint32_t j = base0 + i * 32 + lane;
            if (j < n) ids[start + j] = __ldcg(&ids_snap[start + (s_pay[lbase + i * 32 + lane] & TS_IDX_MASK)]);
        }
        __threadfence();  // the final order is read by other clusters / later tiers
        cluster.sync();
        TC_STAMP(6);

        // ---- 7. record + children ----
        if (rank == 0 && tid == 0) {
            const uint32_t p = s_piv[best];  // the pivot recorded when the candidate was evaluated (blas.rs:159,165)
            float lo[3], hi[3];
            for (int c = 0; c < 3; ++c) {
                lo[c] = o2f(min(s_node[c], ENC_POS_INIT));
                hi[c] = o2f(max(s_node[3 + c], ENC_NEG_INIT));
            }
            for (int c = 0; c < 6; ++c) {
                unsigned long long zk = 0xFFFFFFFFFFFFFFFFull;
                for (uint32_t r2 = 0; r2 < C; ++r2) zk = min(zk, s_zpart[r2][c]);
                if (zk != 0xFFFFFFFFFFFFFFFFull) {  // only set on the rare -0.0 path
                    const uint32_t g = (uint32_t)(zk & 0xFFFFFFFFull);
                    const float4 bb = box[2 * (size_t)g + (c < 3 ? 0 : 1)];
                    const float z = (c % 3 == 0) ? bb.x : ((c % 3 == 1) ? bb.y : bb.z);
                    if (c < 3) lo[c] = z; else hi[c - 3] = z;
                }
            }
            if (p == 0 || p >= n) {
                atomicOr(&st->err, DERR_DEGENERATE);
            } else {
                emit_rec(recs, 2 * (start + p) + 1, lo, hi, start, n, t.leftrun, t.pstart, t.pleftrun, t.flags);
                if (p <= 3) A[start] = t.leftrun + 1;
                push_any(Q, st, epoch, start, p, t.leftrun + 1, start, t.leftrun, t.flags & ~3u);
                push_any(Q, st, epoch, start + p, n - p, 0, start, t.leftrun, TF_RIGHT | (t.flags & ~3u));
                atomicAdd(&st->tc_done, 1u);
                atomicAdd(&st->grid_nodes, 1u);
                atomicAdd(&st->sum_grid, (unsigned long long)n);
            }
            __threadfence();
            atomicSub(&st->c_pending, 1u);
#ifdef BVH_TC_TIMING
            {
                const unsigned int row = atomicAdd(&g_tc_logn, 1u);
                if (row < 4096) {
                    g_tc_log[row][0] = (unsigned long long)n | ((unsigned long long)(blockIdx.x / C) << 32);
                    for (int k = 0; k < 7; ++k) g_tc_log[row][1 + k] = tc_t[k];
                }
            }
#endif
        }
        __syncthreads();
    }
}
