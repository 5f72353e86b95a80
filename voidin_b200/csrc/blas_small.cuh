// blas_small.cuh -- part of blas_build.cu (included there, inside its anonymous namespace; not a stand-alone header):
// sub-trees of <= 32 primitives: k_t3 (one warp per sub-tree) and k_t4 (one thread per sub-tree, t4_seq.cuh).
#pragma once

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
    const uint32_t n_tasks = min(st->t3_count, Q.t3_cap);  // push_child keeps counting after an overflow (DERR_QUEUE)
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
