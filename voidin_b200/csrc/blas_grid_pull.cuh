// blas_grid_pull.cuh -- part of blas_grid.cuh (included there): the resident-tile "pull" form of the grid tier's shuffles.
//
// Used for every level whose tiles all have a block of their own (n_tiles <= gridDim.x; the dragon-class build: all
// levels).  ONE grid barrier per partition_shuffle (blas.rs:168-182) instead of two, no rank->position table:
//   before the barrier   every tile publishes, for the plane about to be shuffled, its L count, its ballot words and the
//                        exclusive L prefix of each of its 8 warps (72 words per tile, double buffered by shuffle parity);
//                        the tile's ids / plane counts are in the global ping-pong buffer AND in the block's shared memory;
//   after the barrier    every tile builds its own OUTPUT slots: a slot is kept (front L), shifted (back R, one to the left),
//                        the boundary element f (lands at `pivot`), or a hole:
//                            p <  f, R(p)                      <-  the L that has RF(p) L's after it
//                            q = p + 1 > f, L(q) (or p = n-1)  <-  the R that has (#L after q) + 1 R's before it (0 for p = n-1)
//                        found by a rank-select: binary search in the node's tile prefix (shared memory), then in the partner
//                        tile's warp prefixes and ballot words, then one 6-byte gather from the input buffer.
//                        Kept and shifted slots never leave shared memory.  The new tile goes to shared memory, to the other
//                        global buffer (coalesced) and straight into the ballots of the next plane.
// The form is modelled in tests/shuffle_models.py::pull_form and checked against the sequential loop for every flag
// vector up to 12 elements and random ones (tests/test_oracle.py::test_pull_form_equals_sequential).
#pragma once

constexpr int T1_MAX_NT = 1024;        // tiles per node the shared prefix array can hold (>= the cooperative grid)
constexpr int T1_META = 72;            // words per tile: [0,8) warp prefixes, [8 + 8 w + i] ballot word i of warp w

// TSMAX: largest tile the instantiation keeps resident; TAB: with the rank -> position table (expansion variant) or without
// (per-hole rank-select variant, used for the small tiles of the default build).
template <int TSMAX, bool TAB>
struct T1SmemT {
    static constexpr bool has_tab = TAB;
    static constexpr int ts_max = TSMAX;
    uint32_t id[2][TSMAX];
    uint16_t fw[2][TSMAX];
    uint32_t pre[T1_MAX_NT + 4];       // exclusive prefix of the node's per-tile L counts; pre[nt] = nL
    uint32_t wsum[T1_THREADS / 32];
    uint32_t wfirst[T1_THREADS / 32 + 1];  // L bit of the first slot of every warp run; [8]: of the next tile
    uint32_t tab[TAB ? TSMAX + 8 : 4];     // rank -> source position of this tile's holes (front holes, then back holes)
};
using T1SmemSmall = T1SmemT<2 * T1_THREADS, false>;  // tiles of 256 / 512 slots: the default build's pull levels
using T1SmemFull = T1SmemT<T1_TILE, true>;           // every tile size (BVH_CUDA_T1_PULL=2)

// position of the r-th (0-based) set bit of w; r < popc(w)
__device__ __forceinline__ uint32_t select32(uint32_t w, uint32_t r) {
    uint32_t pos = 0, c;
    c = __popc(w & 0xFFFFu); if (r >= c) { r -= c; pos += 16; w >>= 16; }
    c = __popc(w & 0xFFu);   if (r >= c) { r -= c; pos += 8;  w >>= 8; }
    c = __popc(w & 0xFu);    if (r >= c) { r -= c; pos += 4;  w >>= 4; }
    c = __popc(w & 0x3u);    if (r >= c) { r -= c; pos += 2;  w >>= 2; }
    c = w & 1u;              if (r >= c) { pos += 1; }
    return pos;
}

// rank-select inside a tile from its published meta: position (relative to the node) of its r-th L (or R)
template <int EPT, bool WANT_L>
__device__ __forceinline__ uint32_t t1_pull_select(const uint32_t* meta_tile, uint32_t t_in_node, uint32_t r) {
    const uint4* m4 = reinterpret_cast<const uint4*>(meta_tile);
    const uint4 p0 = __ldcg(m4), p1 = __ldcg(m4 + 1);
    uint32_t wp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    uint32_t w = 0, base = 0;
#pragma unroll
    for (int x = 1; x < 8; ++x) {
        if (!WANT_L) wp[x] = (uint32_t)x * (32 * EPT) - wp[x];
        if (wp[x] <= r) { w = x; base = wp[x]; }
    }
    r -= base;
    const uint4 a0 = __ldcg(m4 + 2 + 2 * w);
    uint32_t wd[4] = {a0.x, a0.y, a0.z, a0.w};   // EPT <= 4 words per warp run in this variant
    uint32_t word = 0, wi = 0;
    bool found = false;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
        const uint32_t x = WANT_L ? wd[i] : ~wd[i];
        const uint32_t c = __popc(x);
        if (!found) {
            if (r < c) { found = true; word = x; wi = i; }
            else r -= c;
        }
    }
    return t_in_node * (uint32_t)(T1_THREADS * EPT) + w * (32 * EPT) + wi * 32 + select32(word, r);
}

#ifdef BVH_T1_TIMING
#define PULL_MARK(k)                                                                       \
    do {                                                                                   \
        const unsigned long long _n = gtimer();                                            \
        if (level == 0 && threadIdx.x == 0 && blockIdx.x < 512) g_t1_pull[k][blockIdx.x] += _n - _pm; \
        _pm = _n;                                                                          \
    } while (0)
#else
#define PULL_MARK(k) do { } while (0)
#endif

struct T1Tile {
    uint32_t tile, node, start, n, lt, tile_base, j0;
};

// Ballots of plane `c` over the tile in shared buffer `cur`; publishes count + meta; leaves bal[] / wpre to the caller.
template <int EPT, class SM>
__device__ __forceinline__ void t1_pull_publish(const T1Args& g, SM& sm, const T1Tile& t, int cur, int c, uint32_t* bal,
                                                uint32_t& wpre) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t a, b;
    cand_of(g.sc, t.node, c, a, b);
    uint32_t cnt = 0;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
        const uint32_t s = warp * (32 * EPT) + i * 32 + lane;
        const bool L = (t.j0 + s < t.n) && ((((uint32_t)sm.fw[cur][s] >> (3 * a)) & 7u) < b);
        bal[i] = __ballot_sync(FULL_MASK, L);
        cnt += __popc(bal[i]);
    }
    __syncthreads();  // wsum / wfirst of the previous step are no longer read
    if (lane == 0) sm.wsum[warp] = cnt;
    __syncthreads();
    uint32_t tot = 0;
    wpre = 0;
#pragma unroll
    for (int w2 = 0; w2 < T1_THREADS / 32; ++w2) {
        const uint32_t v = sm.wsum[w2];
        if ((uint32_t)w2 < warp) wpre += v;
        tot += v;
    }
    uint32_t* m = g.pbal + ((size_t)(c & 1) * g.tile_stride + t.tile) * T1_META;
    if (lane == 0) m[warp] = wpre;
#pragma unroll
    for (int i = 0; i < EPT; ++i)
        if (lane == (uint32_t)i) m[8 + warp * 8 + i] = bal[i];
    if (tid == 0) g.tileL[(size_t)c * g.tile_stride + t.tile] = tot;
}

// One shuffle of the block's tile in pull form.  `bal` / `wpre`: the tile's ballots for plane c (from t1_pull_publish).
template <int EPT, class SM>
__device__ __forceinline__ void t1_pull_step(const T1Args& g, SM& sm, const T1Tile& t, int cur, int c, const uint32_t* bal,
                                             uint32_t wpre, const uint32_t* __restrict__ ids_in, const uint16_t* __restrict__ fl_in,
                                             uint32_t* ids_out, uint16_t* fl_out, const uint32_t level) {
    (void)level;
#ifdef BVH_T1_TIMING
    unsigned long long _pm = gtimer();
#endif
    constexpr uint32_t TS = T1_THREADS * EPT;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = t.n, nt = (n + TS - 1) / TS;
    const uint32_t* meta = g.pbal + (size_t)(c & 1) * g.tile_stride * T1_META;
    // ---- exclusive prefix of the node's tile counts (4 tiles per thread) ----
    {
        const uint32_t* tl = g.tileL + (size_t)c * g.tile_stride + t.tile_base;
        uint32_t v[4], sum = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t x = tid * 4 + k;
            v[k] = (x < nt) ? __ldcg(tl + x) : 0u;
            sum += v[k];
        }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL_MASK, inc, o);
            if ((int)lane >= o) inc += y;
        }
        if (lane == 31) sm.wsum[warp] = inc;
        if (lane == 0) sm.wfirst[warp] = bal[0] & 1u;
        if (tid == 0) sm.wfirst[T1_THREADS / 32] = (t.j0 + TS < n) ? (__ldcg(meta + (size_t)(t.tile + 1) * T1_META + 8) & 1u) : 0u;
        __syncthreads();
        uint32_t wb = 0, tot = 0;
#pragma unroll
        for (int w2 = 0; w2 < T1_THREADS / 32; ++w2) {
            const uint32_t x = sm.wsum[w2];
            if ((uint32_t)w2 < warp) wb += x;
            tot += x;
        }
        uint32_t run = wb + inc - sum;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t x = tid * 4 + k;
            if (x < nt) sm.pre[x] = run;
            run += v[k];
        }
        if (tid == 0) sm.pre[nt] = tot;
        __syncthreads();
    }
    PULL_MARK(0);
    const uint32_t nL = sm.pre[nt], tile_lf = sm.pre[t.lt];
    const uint32_t j_end = min(n, t.j0 + TS);
    // ---- boundary element: only the tiles that touch [nL-1, nL+1] need it exactly ----
    uint32_t f, pivot = 0xFFFFFFFFu;
    if (j_end + 1 <= nL) f = 0xFFFFFFFFu;         // every slot and its right neighbour lie before f
    else if (t.j0 >= nL + 2) f = 0;               // every slot lies behind f (and t.j0 >= 2)
    else {
        uint32_t a, b;
        cand_of(g.sc, t.node, c, a, b);
        auto l_at = [&](uint32_t j) -> uint32_t {
            return (j < n && ((((uint32_t)__ldcg(fl_in + t.start + j) >> (3 * a)) & 7u) < b)) ? 1u : 0u;
        };
        const uint32_t l0 = nL ? l_at(nL - 1) : 0u, l1 = l_at(nL), l2 = l_at(nL + 1);
        uint32_t lf;
        if (nL >= 1 && !(nL + 1 <= n && l0 + l1 <= 1)) { f = nL - 1; lf = l0; }
        else if (!(nL + 2 <= n && l1 + l2 == 0)) { f = nL; lf = l1; }
        else { f = nL + 1; lf = l2; }
        pivot = nL - lf;
    }
    // ---- partner tiles: the holes of this tile have consecutive ranks, so their sources are a run of consecutive L's (or
    // R's) in a short run of tiles on the other side of f.  The block expands the ballot words of those tiles into a
    // rank -> position table in shared memory (one thread per 32-slot word, only the ranks this tile can ask for), so a
    // hole costs one shared-memory lookup instead of a rank-select. ----
    const uint32_t cnt_tile = sm.pre[t.lt + 1] - tile_lf, sz_tile = j_end - t.j0;
    uint32_t r0lo = 0, n0 = 0, r1lo = 0, n1 = 0;
    uint32_t Gmin = 1, Gmax = 0, Kmin = 1, Kmax = 0;   // empty ranges
    auto search_pre = [&](uint32_t G, uint32_t lo, uint32_t hi) {  // largest t in [lo, hi) with pre[t] <= G
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (sm.pre[mid] <= G) lo = mid; else hi = mid;
        }
        return lo;
    };
    auto search_rpre = [&](uint32_t K, uint32_t lo, uint32_t hi) {  // largest t in [lo, hi) with (R's before tile t) <= K
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (mid * TS - sm.pre[mid] <= K) lo = mid; else hi = mid;
        }
        return lo;
    };
    if (t.j0 < f && sz_tile > cnt_tile && nL > 0) {
        const uint32_t rf0 = t.j0 - tile_lf, rf1 = rf0 + (sz_tile - cnt_tile) - 1;  // RF of the tile's first / last R
        if (rf0 <= nL - 1) {
            Gmax = nL - 1 - rf0;
            Gmin = rf1 <= nL - 1 ? nL - 1 - rf1 : 0u;
            r0lo = search_pre(Gmin, 0, nt);
            n0 = search_pre(Gmax, r0lo, nt) - r0lo + 1;
        }
    }
    const uint32_t h0 = Gmax + 1 - Gmin;  // table slots of the front holes (0 when the range is empty)
    if (j_end > f && n > nL) {
        Kmin = nL - tile_lf - cnt_tile;
        Kmax = min(nL - tile_lf, n - nL - 1);
        if (Kmin <= Kmax) {
            r1lo = search_rpre(Kmin, 0, nt);
            n1 = search_rpre(Kmax, r1lo, nt) - r1lo + 1;
        }
    }
    if constexpr (SM::has_tab) {
        constexpr uint32_t WPT = 8 * EPT;  // ballot words per tile
        for (uint32_t x = tid; x < (n0 + n1) * WPT; x += T1_THREADS) {
            const uint32_t e = x / WPT, wi = x % WPT, w = wi / EPT, i = wi % EPT;
            const bool wantL = e < n0;
            const uint32_t T = wantL ? r0lo + e : r1lo + (e - n0);
            const uint32_t* m = meta + (size_t)(t.tile_base + T) * T1_META;
            const uint32_t wp = __ldcg(m + w);
            const uint4 a0 = __ldcg(reinterpret_cast<const uint4*>(m + 8 + 8 * w));
            uint32_t wd[8] = {a0.x, a0.y, a0.z, a0.w, 0, 0, 0, 0};
            if (EPT > 4) {
                const uint4 a1 = __ldcg(reinterpret_cast<const uint4*>(m + 12 + 8 * w));
                wd[4] = a1.x; wd[5] = a1.y; wd[6] = a1.z; wd[7] = a1.w;
            }
            uint32_t before = 0, word = 0;  // L's of the run before this word; the word itself
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                if ((uint32_t)k < i) before += __popc(wd[k]);
                if ((uint32_t)k == i) word = wd[k];
            }
            const uint32_t pos0 = T * TS + w * (32 * EPT) + i * 32;  // node position of the word's bit 0
            if (wantL) {
                uint32_t rank = sm.pre[T] + wp + before;  // global L rank of the word's first L
                if (rank <= Gmax && rank + __popc(word) > Gmin)
                    while (word) {
                        const uint32_t b = __ffs(word) - 1;
                        word &= word - 1;
                        if (rank >= Gmin && rank <= Gmax) sm.tab[rank - Gmin] = pos0 + b;
                        ++rank;
                    }
            } else {
                word = ~word;
                uint32_t rank = (T * TS - sm.pre[T]) + (w * (32 * EPT) - wp) + (i * 32 - before);  // global R rank
                if (rank <= Kmax && rank + __popc(word) > Kmin)
                    while (word) {
                        const uint32_t b = __ffs(word) - 1;
                        word &= word - 1;
                        if (rank >= Kmin && rank <= Kmax) sm.tab[h0 + rank - Kmin] = pos0 + b;
                        ++rank;
                    }
            }
        }
        __syncthreads();
    }
    PULL_MARK(1);
    const int nxt = cur ^ 1;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t running = tile_lf + wpre;
    // ---- pass 1: the source position of every output slot (no memory traffic beyond shared memory) ----
    uint32_t srcq[EPT];
    int piv_i = -1;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
        const uint32_t s = warp * (32 * EPT) + i * 32 + lane, p = t.j0 + s;
        const uint32_t word = bal[i];
        const uint32_t Lp = (word >> lane) & 1u;
        const uint32_t LFp = running + __popc(word & lt_mask);
        running += __popc(word);
        uint32_t Lq;
        if (lane < 31) Lq = (word >> (lane + 1)) & 1u;
        else if (i + 1 < EPT) Lq = bal[(i + 1 < EPT) ? i + 1 : i] & 1u;
        else Lq = sm.wfirst[warp + 1];
        uint32_t q = p;
        if (p < n) {
            if (p == pivot) { q = f; piv_i = i; }
            else if (p < f) {
                if (!Lp) {
                    // hole in the front: the L that has k = RF(p) L's after it, i.e. the L of rank G = nL - 1 - k
                    const uint32_t G = nL - 1u - (p - LFp);
                    if constexpr (SM::has_tab) q = sm.tab[G - Gmin];
                    else {
                        const uint32_t T = search_pre(G, r0lo, r0lo + n0);
                        q = t1_pull_select<(EPT > 4 ? 4 : EPT), true>(meta + (size_t)(t.tile_base + T) * T1_META, T, G - sm.pre[T]);
                    }
                }
            } else if (p + 1 < n && !Lq) q = p + 1;
            else {
                // hole in the back: the R that has K R's before it
                const uint32_t K = (p + 1 == n) ? 0u : nL - (LFp + Lp);
                if constexpr (SM::has_tab) q = sm.tab[h0 + K - Kmin];
                else {
                    const uint32_t T = search_rpre(K, r1lo, r1lo + n1);
                    q = t1_pull_select<(EPT > 4 ? 4 : EPT), false>(meta + (size_t)(t.tile_base + T) * T1_META, T, K - (T * TS - sm.pre[T]));
                }
            }
        }
        srcq[i] = q;
    }
    __syncwarp();
    PULL_MARK(2);
    // ---- pass 2: gather (own tile: shared memory; other tiles: the global input buffer), then write ----
    uint32_t idv[EPT], fwv[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
        const uint32_t s = warp * (32 * EPT) + i * 32 + lane, p = t.j0 + s, q = srcq[i];
        idv[i] = 0; fwv[i] = 0;
        if (p < n) {
            if (q - t.j0 < TS) { idv[i] = sm.id[cur][q - t.j0]; fwv[i] = sm.fw[cur][q - t.j0]; }
            else { idv[i] = __ldcg(ids_in + t.start + q); fwv[i] = __ldcg(fl_in + t.start + q); }
        }
    }
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
        const uint32_t s = warp * (32 * EPT) + i * 32 + lane, p = t.j0 + s;
        if (p < n) {
            uint32_t fw = fwv[i];
            if (i == piv_i) {
                fw |= 0x8000u;
                if (c < 21) { g.sc[t.node].piv[c] = pivot; g.sc[t.node].uid[c] = idv[i]; }
            }
            sm.id[nxt][s] = idv[i];
            sm.fw[nxt][s] = (uint16_t)fw;
            ids_out[t.start + p] = idv[i];
            fl_out[t.start + p] = (uint16_t)fw;
        }
    }
    PULL_MARK(3);
}

// All shuffles of one level, every tile resident in the shared memory of its own block.
template <int EPT, class SM>
__device__ __forceinline__ void t1_level_pull(const T1Args& g, SM& sm, uint32_t& gen, const uint32_t level) {
    (void)level;
    constexpr uint32_t TS = T1_THREADS * EPT;
    T1_PHASE(0, p_t1_init(g));
    T1_PHASE(1, p_t1_bounds<EPT>(g));
    T1_PHASE(2, p_t1_flags<EPT>(g));
    const bool has_tile = blockIdx.x < g.n_tiles;
    T1Tile t = {};
    uint32_t bal[EPT], wpre = 0;
    int cur = 0;
    if (has_tile) {
        const uint4 td = g.tile_desc[blockIdx.x];
        t.tile = blockIdx.x; t.node = td.x; t.start = td.y; t.n = td.z; t.lt = td.w;
        t.tile_base = t.tile - t.lt; t.j0 = t.lt * TS;
        const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            const uint32_t s = warp * (32 * EPT) + i * 32 + lane, p = t.j0 + s;
            sm.id[0][s] = (p < t.n) ? __ldcg(g.ids0 + t.start + p) : 0u;
            sm.fw[0][s] = (p < t.n) ? __ldcg(g.fl0 + t.start + p) : (uint16_t)0;
        }
        __syncthreads();
        t1_pull_publish<EPT, SM>(g, sm, t, cur, 0, bal, wpre);
    }
    grid_barrier(g.barrier, gen);
    for (int c = 0; c < 22; ++c) {
        const uint32_t* ids_in = (c & 1) ? g.ids1 : g.ids0;
        uint32_t* ids_out = (c & 1) ? g.ids0 : g.ids1;
        const uint16_t* fl_in = (c & 1) ? g.fl1 : g.fl0;
        uint16_t* fl_out = (c & 1) ? g.fl0 : g.fl1;
        if (c == 21) {
            T1_PHASE(3, p_t1_bins<EPT>(g, ids_in, fl_in));
            T1_PHASE(4, p_t1_select(g));
            if (has_tile) t1_pull_publish<EPT, SM>(g, sm, t, cur, 21, bal, wpre);
            grid_barrier(g.barrier, gen);
        }
#ifdef BVH_T1_TIMING
        const unsigned long long _t0 = gtimer();
#endif
        if (has_tile) {
            t1_pull_step<EPT, SM>(g, sm, t, cur, c, bal, wpre, ids_in, fl_in, ids_out, fl_out, level);
            cur ^= 1;
            if (c < 20) {
                __syncthreads();
                t1_pull_publish<EPT, SM>(g, sm, t, cur, c + 1, bal, wpre);
            }
        }
#ifdef BVH_T1_TIMING
        __syncthreads();
        const unsigned long long _t1 = gtimer();
#endif
        grid_barrier(g.barrier, gen);
#ifdef BVH_T1_TIMING
        if (blockIdx.x == 0 && threadIdx.x == 0) { g_t1_time[2 * 11] += _t1 - _t0; g_t1_time[2 * 11 + 1] += gtimer() - _t1; }
#endif
    }
}
