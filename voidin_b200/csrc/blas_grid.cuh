// blas_grid.cuh -- part of blas_build.cu (included there, inside its anonymous namespace; not a stand-alone header):
// the grid tier: k_t1_coop and its phases (nodes above T2B_CAP = 24576 primitives).
#pragma once

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
__device__ unsigned long long g_t1_pull[8][512];   // level 0, per block: ns in the sub-phases of the pull step (blas_grid_pull.cuh)
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

// T1_BARRIER_GROUPS > 1: two-level arrival.  Blocks arrive on one of G group counters (own cache line each); the last
// block of a group arrives on the top counter, the last of those releases every group by writing the generation into the
// group's flag line, which is the only line that group's blocks poll.  One counter for all 444 blocks serialises 444
// atomics and 444 pollers on a single L2 line; with 16 groups it is 28 + 16.  MEASURED SLOWER on B200 (dragon-class build,
// grid tier 3.18 ms flat vs 3.87 / 3.85 / 3.82 ms with 8 / 16 / 37 groups, profiles/r02_build_variants_ab.txt): the flat
// barrier is two dependent L2 round trips (atomic, poll), the two-level one is four, and same-address atomics are not the
// bottleneck.  Kept as a compile-time option; the default is the flat counter.
#ifndef T1_BARRIER_GROUPS
#define T1_BARRIER_GROUPS 1
#endif
constexpr uint32_t T1_BARRIER_WORDS = 32u * (1u + 2u * T1_BARRIER_GROUPS);
__device__ __forceinline__ void grid_barrier(uint32_t* counter, uint32_t& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        gen += 1;
#if T1_BARRIER_GROUPS > 1
        constexpr uint32_t NG = T1_BARRIER_GROUPS;
        const uint32_t grp = blockIdx.x % NG;
        const uint32_t ng = min(NG, gridDim.x);
        const uint32_t gsz = (gridDim.x - grp + NG - 1) / NG;
        __threadfence();
        if (atomicAdd(counter + 32 * (1 + grp), 1u) + 1 == gen * gsz) {
            __threadfence();
            if (atomicAdd(counter, 1u) + 1 == gen * ng) {
                __threadfence();
                for (uint32_t k = 0; k < ng; ++k) *(volatile uint32_t*)(counter + 32 * (1 + NG + k)) = gen;
            }
        }
        while (ld_vol(counter + 32 * (1 + NG + grp)) < gen) { }
        __threadfence();
#else
        const uint32_t target = gen * gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        while (ld_vol(counter) < target) { }
        __threadfence();
#endif
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
            if (cn > Q.tc_cap) {
                const uint32_t idx = atomicAdd(&g.st->lv_count[next_slot], 1u);
                if (idx >= next_cap) { atomicOr(&g.st->err, DERR_QUEUE); continue; }
                LevelNode c;
                c.start = cs; c.n = cn; c.leftrun = clr; c.pstart = nd.start; c.pleftrun = nd.leftrun; c.flags = cfl;
                c.tile_base = 0; c.pad = 0;
                next_nodes[idx] = c;
            } else {
                push_any(Q, g.st, epoch, cs, cn, clr, nd.start, nd.leftrun, cfl);
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

#include "blas_grid_pull.cuh"

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
// PULL selects the resident-tile path of blas_grid_pull.cuh: 0 = never (pure two-phase tier), 1 = for the levels whose tiles
// are 256 or 512 slots (per-hole rank-select variant, 10 KB of shared memory; where it measured faster), 2 = for every level
// whose tiles all have a block (table-expansion variant, 45 KB; measured slower at 2048-slot tiles).
template <int PULL>
__global__ void __launch_bounds__(T1_THREADS, T1_MIN_BLOCKS) k_t1_coop(T1Args g, LevelNode* lv0, LevelNode* lv1, uint32_t lv_cap, Queues Q,
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
        bool done = false;
        if constexpr (PULL == 1) {
            __shared__ __align__(16) T1SmemSmall s_pull;
            if (n_tiles <= gridDim.x && gridDim.x <= (uint32_t)T1_MAX_NT && g.ept <= 2) {
                if (g.ept == 1) t1_level_pull<1, T1SmemSmall>(g, s_pull, gen, level);
                else t1_level_pull<2, T1SmemSmall>(g, s_pull, gen, level);
                done = true;
            }
        }
        if constexpr (PULL == 2) {
            // every tile has its own block: the tile stays in shared memory and a shuffle is one barrier (blas_grid_pull.cuh)
            __shared__ __align__(16) T1SmemFull s_pull;
            if (n_tiles <= gridDim.x && gridDim.x <= (uint32_t)T1_MAX_NT) {
                switch (g.ept) {
                    case 1: t1_level_pull<1, T1SmemFull>(g, s_pull, gen, level); break;
                    case 2: t1_level_pull<2, T1SmemFull>(g, s_pull, gen, level); break;
                    case 4: t1_level_pull<4, T1SmemFull>(g, s_pull, gen, level); break;
                    default: t1_level_pull<8, T1SmemFull>(g, s_pull, gen, level); break;
                }
                done = true;
            }
        }
        if (!done) {
            switch (g.ept) {
                case 1: t1_level<1>(g, scanned, gen, level); break;
                case 2: t1_level<2>(g, scanned, gen, level); break;
                case 4: t1_level<4>(g, scanned, gen, level); break;
                default: t1_level<8>(g, scanned, gen, level); break;
            }
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
