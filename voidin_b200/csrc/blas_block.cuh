// blas_block.cuh -- part of blas_build.cu (included there, inside its anonymous namespace; not a stand-alone header):
// device task queues and the block / warp tiers: k_t2<24576,1024>, k_t2<2048,256>, k_t2w.
#pragma once

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
    // dynamic shared memory: payload (bits 0-15 local primitive, 16-24 plane counts, 31 special), shuffled in place, and
    // the rank -> position table.  The local-primitive -> triangle-id map lives in global memory (ids_snap).
    extern __shared__ uint32_t s_dyn[];
    uint32_t* const s_pay0 = s_dyn;
    uint16_t* const s_tab = reinterpret_cast<uint16_t*>(s_dyn + CAP);
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

        // ---- 3. shuffles: partition_shuffle (blas.rs:168-182) in closed form, IN PLACE on [s0, n) ----
        // Plane b of an axis leaves [0, pivot_{b-1}) untouched (the front cursor walks over an all-left prefix without a
        // swap), so it is the shuffle of the suffix alone; the active range is re-spread over all threads, which is what
        // shortens the dependent chain: E = ceil(active / THREADS) slots per thread instead of ceil(n / THREADS).
        // In place: every thread keeps its slots' payloads in registers from the counting pass to the scatter.
        auto shuffle = [&](uint32_t a, uint32_t b, uint32_t s0, int cidx) -> uint32_t {
            uint32_t* const p = s_pay0 + s0;
            const uint32_t act = n - s0;
            const uint32_t Ea = (act + THREADS - 1) / THREADS, CH = 32 * Ea;
            const uint32_t sh = 16 + 3 * a;
            uint32_t bal[EPT], pay[EPT];
            uint32_t cnt = 0;
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                bal[i] = 0;
                pay[i] = 0;
                if (i < (int)Ea) {
                    const uint32_t j = warp * CH + i * 32 + lane;
                    if (j < act) pay[i] = p[j];
                    const bool L = (j < act) && (((pay[i] >> sh) & 7u) < b);
                    bal[i] = __ballot_sync(FULL_MASK, L);
                    cnt += __popc(bal[i]);
                }
            }
            if (lane == 0) s_wtot[warp] = cnt;
            __syncthreads();  // S1
            // lane w2 reads warp w2's count: total and the sum over the warps before this one by two warp reductions
            const uint32_t wv = (lane < (uint32_t)NW) ? s_wtot[lane] : 0u;
            const uint32_t nL = __reduce_add_sync(FULL_MASK, wv);
            const uint32_t wpre = __reduce_add_sync(FULL_MASK, lane < warp ? wv : 0u);
            // boundary element in closed form (see p_t1_table): f is nL-1, nL or nL+1, decided by three flags
            uint32_t f, Lf;
            {
                const uint32_t l0 = nL ? ((((p[nL - 1] >> sh) & 7u) < b) ? 1u : 0u) : 0u;
                const uint32_t l1 = (nL < act && (((p[nL < act ? nL : 0] >> sh) & 7u) < b)) ? 1u : 0u;
                const uint32_t l2 = (nL + 1 < act && (((p[nL + 1 < act ? nL + 1 : 0] >> sh) & 7u) < b)) ? 1u : 0u;
                if (nL >= 1 && !(nL + 1 <= act && l0 + l1 <= 1)) { f = nL - 1; Lf = l0; }
                else if (!(nL + 2 <= act && l1 + l2 == 0)) { f = nL; Lf = l1; }
                else { f = nL + 1; Lf = l2; }
            }
            const uint32_t pivot = nL - Lf;
            uint32_t running = wpre;
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                if (i >= (int)Ea) break;
                const uint32_t j = warp * CH + i * 32 + lane;
                const uint32_t LF = running + __popc(bal[i] & lt_mask);
                if (j < act) {
                    // only front R's (j < f) and back L's (j > f) are looked up
                    if ((bal[i] >> lane) & 1u) { if (j >= nL) s_tab[act - 1 - (nL - LF - 1)] = (uint16_t)j; }
                    else if (j <= nL) s_tab[j - LF] = (uint16_t)j;
                }
                running += __popc(bal[i]);
            }
            __syncthreads();  // S2: table complete, and every read of p (counting pass, boundary flags) is done
            running = wpre;
#pragma unroll
            for (int i = 0; i < EPT; ++i) {
                if (i >= (int)Ea) break;
                const uint32_t j = warp * CH + i * 32 + lane;
                const uint32_t LF = running + __popc(bal[i] & lt_mask);
                running += __popc(bal[i]);
                if (j < act) {
                    const uint32_t Lbit = (bal[i] >> lane) & 1u;
                    const uint32_t RF = j - LF;
                    if (j < f) {
                        if (!Lbit) p[RF == 0 ? act - 1 : (uint32_t)s_tab[act - RF] - 1u] = pay[i];  // front L's stay where they are
                    } else if (j == f) {
                        const uint32_t up = pay[i] | 0x80000000u;
                        p[pivot] = up;
                        if (cidx >= 0) { s_u[cidx] = up & 0xFFFFu; s_uk[cidx] = (up >> 16) & 0x1FFu; s_piv[cidx] = s0 + pivot; }
                    } else p[Lbit ? (uint32_t)s_tab[nL - LF - 1] : j - 1] = pay[i];
                }
            }
            __syncthreads();  // S3
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

        // ---- 4. exact bins over the non-special primitives (4 slots per thread at a time) ----
        // Boxes are mapped to ordered uints once per slot, so the 24 per-bin reductions below are integer min / max
        // feeding redux.sync directly.
        {
            const uint32_t* pin = s_pay0;
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
        shuffle(best / 7, best % 7 + 1, 0, -1);
        {
            const uint32_t* pin = s_pay0;
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
    __shared__ uint32_t s_pay[NWB][WCAP];  // bits 0-15 local primitive, 16-24 plane counts, 31 special; shuffled in place
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
            if (i < (int)E && j < n) s_pay[w][j] = j | (plane_counts(ccx[i], ccy[i], ccz[i], cmin, cmax) << 16);
        }
        __syncwarp();

        // ---- 3. shuffles: closed form of partition_shuffle, in place on [s0, n) (see k_t2) ----
        uint32_t last_up = 0;
        auto shuffle = [&](uint32_t a, uint32_t b, uint32_t s0) -> uint32_t {
            uint32_t* const p = s_pay[w] + s0;
            const uint32_t act = n - s0;
            const uint32_t Ea = (act + 31) >> 5;  // chunks in use
            const uint32_t sh = 16 + 3 * a;
            uint32_t bal[EPL], pay[EPL];
            uint32_t nL = 0;
#pragma unroll
            for (int i = 0; i < EPL; ++i) {
                bal[i] = 0;
                pay[i] = 0;
                if (i < (int)Ea) {
                    const uint32_t j = i * 32 + lane;
                    if (j < act) pay[i] = p[j];
                    const bool L = (j < act) && (((pay[i] >> sh) & 7u) < b);
                    bal[i] = __ballot_sync(FULL_MASK, L);
                    nL += __popc(bal[i]);
                }
            }
            // boundary element in closed form (see p_t1_table): f is nL-1, nL or nL+1, decided by three flags
            auto l_at = [&](uint32_t j) -> uint32_t {  // one broadcast shared-memory read
                return (j < act && (((p[j < act ? j : 0] >> sh) & 7u) < b)) ? 1u : 0u;
            };
            uint32_t f, Lf;
            {
                const uint32_t l0 = nL ? l_at(nL - 1) : 0u, l1 = l_at(nL), l2 = l_at(nL + 1);
                if (nL >= 1 && !(nL + 1 <= act && l0 + l1 <= 1)) { f = nL - 1; Lf = l0; }
                else if (!(nL + 2 <= act && l1 + l2 == 0)) { f = nL; Lf = l1; }
                else { f = nL + 1; Lf = l2; }
            }
            const uint32_t pivot = nL - Lf;
            uint32_t running = 0;
#pragma unroll
            for (int i = 0; i < EPL; ++i) {
                if (i >= (int)Ea) break;
                const uint32_t j = i * 32 + lane;
                const uint32_t LF = running + __popc(bal[i] & lt_mask);
                if (j < act) {
                    if ((bal[i] >> lane) & 1u) { if (j >= nL) s_tab[w][act - 1 - (nL - LF - 1)] = (uint16_t)j; }
                    else if (j <= nL) s_tab[w][j - LF] = (uint16_t)j;
                }
                running += __popc(bal[i]);
            }
            __syncwarp();  // table complete; every read of p (counting pass, boundary flags) is done
            uint32_t upay = 0;
            running = 0;
#pragma unroll
            for (int i = 0; i < EPL; ++i) {
                if (i >= (int)Ea) break;
                const uint32_t j = i * 32 + lane;
                const uint32_t LF = running + __popc(bal[i] & lt_mask);
                running += __popc(bal[i]);
                if (j < act) {
                    const uint32_t Lbit = (bal[i] >> lane) & 1u;
                    const uint32_t RF = j - LF;
                    if (j < f) {
                        if (!Lbit) p[RF == 0 ? act - 1 : (uint32_t)s_tab[w][act - RF] - 1u] = pay[i];  // front L's stay
                    } else if (j == f) {
                        upay = pay[i] | 0x80000000u;
                        p[pivot] = upay;
                    } else p[Lbit ? (uint32_t)s_tab[w][nL - LF - 1] : j - 1] = pay[i];
                }
            }
            __syncwarp();
            last_up = __shfl_sync(FULL_MASK, upay, f & 31u);  // payload of the unexamined element
            return pivot;
        };

        uint32_t my_u = 0xFFFFFFFFu, my_kb = 0, my_piv = 0;
        {
            uint32_t s0 = 0;
            for (uint32_t c = 0; c < 21; ++c) {
                const uint32_t b = c % 7 + 1;
                if (b == 1) s0 = 0;  // a new axis starts on the whole range
                const uint32_t pivot = s0 + shuffle(c / 7, b, s0);
                if (lane == c) { my_u = last_up & 0xFFFFu; my_kb = (last_up >> 16) & 0x1FFu; my_piv = pivot; }
                s0 = pivot;
            }
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
                    const uint32_t pay = s_pay[w][j];
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
        shuffle(win / 7, win % 7 + 1, 0);
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            const uint32_t j = i * 32 + lane;
            if (i < (int)E && j < n) ids[start + j] = s_gid[w][s_pay[w][j] & 0xFFFFu];
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
