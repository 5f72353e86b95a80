// t4_seq.cuh — one THREAD builds a whole sub-tree of <= CAP primitives by running the reference algorithm as it
// is written: the sequential two-cursor partition_shuffle (crates/bvh/src/blas.rs:168-182) for each of the 21
// candidate planes, exact vertex bounds of both halves after every shuffle (blas.rs:149-155), strict-< first-wins
// selection from f32::MAX (blas.rs:140,156), the 22nd shuffle whose own pivot is ignored (blas.rs:164-165), then
// left-first recursion (blas.rs:124-125) with an explicit stack.
//
// Why a thread and not a warp: 61 % of all interior nodes of a triangle mesh hold 4..8 primitives and 90 % hold
// <= 32 (dragon-class mesh: 236 209 / 72 901 / 38 455 nodes of 4-8 / 9-16 / 17-32 primitives out of 387 009).  A warp
// that replays 22 closed-form shuffles for one such node spends ~3 000 warp-instructions on it with most lanes
// idle; 32 threads that each run the plain sequential loops on their own node need no scans, ballots or barriers
// at all, and there are tens of thousands of independent sub-trees to fill the lanes with.
//
// The per-thread working set (triangle boxes, ids, current order + plane counts: 8 words per primitive) lives in
// shared memory as [word][thread], so a lane only ever touches its own bank whatever slot it indexes.
//
// This header is also compiled by g++ (tests/t4_host.cpp) so that the very same code is checked against the CPU
// oracle without a GPU; every float operation that could be contracted is spelled as a helper below.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define T4_HD __host__ __device__ __forceinline__
#else
#define T4_HD inline
#endif

#ifndef T4_TF_RIGHT
#define T4_TF_RIGHT 1u
#endif
#define T4_ERR_DEGENERATE 2u  // == DERR_DEGENERATE

T4_HD float t4_add(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
T4_HD float t4_sub(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
T4_HD float t4_mul(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
T4_HD float t4_u2f(uint32_t n) {
#ifdef __CUDA_ARCH__
    return __uint2float_rn(n);
#else
    return (float)n;
#endif
}
T4_HD uint32_t t4_bits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c;
    c.f = f;
    return c.u;
#endif
}
// f32::min / f32::max with the accumulator on the left (blas.rs:190-198): the new value replaces the accumulator
// only when it compares strictly below / above it, so among zeros of different sign the first one met stays and a
// NaN is never taken.
T4_HD float t4_min(float acc, float x) { return (x < acc) ? x : acc; }
T4_HD float t4_max(float acc, float x) { return (x > acc) ? x : acc; }

// Aabb::area (crates/bvh/src/intersection.rs:16-19): (dx*dy + dx*dz + dy*dz) * 2, left to right, unfused
T4_HD float t4_area(const float* lo, const float* hi) {
    const float dx = t4_sub(hi[0], lo[0]), dy = t4_sub(hi[1], lo[1]), dz = t4_sub(hi[2], lo[2]);
    return t4_mul(t4_add(t4_add(t4_mul(dx, dy), t4_mul(dx, dz)), t4_mul(dy, dz)), 2.0f);
}

struct T4Task {
    uint32_t start, n, leftrun, pstart, pleftrun, flags;
};

struct alignas(16) T4Rec {  // one 48-byte node record, three 16-byte words (same layout as emit_rec in blas_build.cu)
    uint32_t w[12];
};

struct T4Cent {  // layout of one centroid as k_setup stores it (float4, w unused)
    float x, y, z, w;
};

template <int CAP>
struct T4Mem {
    float* f;         // this thread's column of [6 * CAP] floats: box[6][CAP], indexed by local primitive
    uint32_t* u;      // this thread's column of [2 * CAP] words: triangle id [CAP] by local primitive, order [CAP] by slot
    uint32_t stride;  // threads per block (1 in the host harness)
    T4_HD float& box(uint32_t c, uint32_t e) const { return f[(c * CAP + e) * stride]; }
    T4_HD uint32_t& gid(uint32_t e) const { return u[e * stride]; }
    // current order: bits 0-4 local primitive, bits 5-13 its plane counts (3 bits per axis) under the node in progress
    T4_HD uint32_t& ord(uint32_t j) const { return u[(CAP + j) * stride]; }
};

T4_HD void t4_emit(T4Rec* recs, uint32_t slot, const float* lo, const float* hi, uint32_t start, uint32_t count,
                   uint32_t leftrun, uint32_t pstart, uint32_t pleftrun, uint32_t flags) {
#ifdef __CUDA_ARCH__
    uint4* q = reinterpret_cast<uint4*>(recs + slot);
    q[0] = make_uint4(t4_bits(lo[0]), t4_bits(lo[1]), t4_bits(lo[2]), start);
    q[1] = make_uint4(t4_bits(hi[0]), t4_bits(hi[1]), t4_bits(hi[2]), count);
    q[2] = make_uint4(leftrun, pstart, pleftrun, flags);
#else
    uint32_t* r = recs[slot].w;
    r[0] = t4_bits(lo[0]); r[1] = t4_bits(lo[1]); r[2] = t4_bits(lo[2]); r[3] = start;
    r[4] = t4_bits(hi[0]); r[5] = t4_bits(hi[1]); r[6] = t4_bits(hi[2]); r[7] = count;
    r[8] = leftrun; r[9] = pstart; r[10] = pleftrun; r[11] = flags;
#endif
}

T4_HD T4Cent t4_load_cent(const T4Cent* cent, uint32_t g) {
#ifdef __CUDA_ARCH__
    const float4 c = reinterpret_cast<const float4*>(cent)[g];
    return T4Cent{c.x, c.y, c.z, 0.0f};
#else
    return cent[g];
#endif
}

// 3-bit plane count of one centroid coordinate: #{b in 1..7 : !(c < pos_b)}, pos_b = lerp(cmin, cmax, b/8)
// (blas.rs:145-146,173).  The planes are monotone in b, so "c < pos_b" <=> count < b: one count per axis replaces
// the 7 float compares of an axis for the rest of the node.
T4_HD uint32_t t4_plane_count(float c, float cmin, float cmax) {
    uint32_t k = 0;
    for (uint32_t b = 1; b < 8; ++b) {
        const float pos = t4_add(cmin, t4_mul(t4_sub(cmax, cmin), t4_mul((float)b, 0.125f)));
        k += (c < pos) ? 0u : 1u;
    }
    return k;
}

// partition_shuffle (blas.rs:168-182) on slots [s, s+n) of the current order, "centroid[axis] < pos" read from the
// plane counts at bit `sh`; returns i - s.
template <int CAP>
T4_HD uint32_t t4_shuffle(const T4Mem<CAP>& m, uint32_t s, uint32_t n, uint32_t sh, uint32_t b) {
    uint32_t i = s, end = s + n - 1;
    while (i < end) {
        const uint32_t w = m.ord(i);
        if (((w >> sh) & 7u) < b) {
            i += 1;
        } else {
            const uint32_t o = m.ord(end);
            m.ord(end) = w;
            m.ord(i) = o;
            end -= 1;
        }
    }
    return i - s;
}

// Builds the sub-tree of task `t` whose primitives have been loaded into `m` (gid and box filled for local primitives
// 0..t.n-1, in the order of ids[t.start..]).  Writes the node records, the A counters and the final order of
// ids[t.start .. t.start+t.n).  Returns 0 or T4_ERR_DEGENERATE.
template <int CAP>
T4_HD uint32_t t4_core(const T4Task& t, const T4Mem<CAP>& m, const T4Cent* cent, uint32_t* ids, T4Rec* recs, uint32_t* A) {
    // right children waiting on the path: s | n << 8 | parent s << 16 | parent k << 24 | parent all-left << 31,
    // where the parent's leftrun is (all-left ? t.leftrun : 0) + k
    uint32_t stk[CAP];
    int sp = 0;
    uint32_t s = 0, n = t.n, leftrun = t.leftrun, pstart = t.pstart, pleftrun = t.pleftrun, fl = t.flags;
    uint32_t k = 0, allleft = 1, err = 0;
    for (uint32_t j = 0; j < t.n; ++j) m.ord(j) = j;

    for (;;) {
        // Leaves first: walk on until the current node is an interior one (or the sub-tree is finished).  About half of
        // all node visits are leaves; taking them in this short loop keeps the lanes of a warp together in the long
        // interior part below instead of parking a leaf lane for the whole of its neighbours' 22 shuffles
        // (ncu, first form: 8.25 of 32 lanes active).
        uint32_t abs_start;
        float lo[3], hi[3];
        bool done = false;
        for (;;) {
            abs_start = t.start + s;
            // own vertex box, folded in slot order from +-1e30 (blas.rs:87-88,117-123,185-186)
            lo[0] = lo[1] = lo[2] = 1e30f;
            hi[0] = hi[1] = hi[2] = -1e30f;
            for (uint32_t j = s; j < s + n; ++j) {
                const uint32_t e = m.ord(j) & 31u;
                lo[0] = t4_min(lo[0], m.box(0, e)); lo[1] = t4_min(lo[1], m.box(1, e)); lo[2] = t4_min(lo[2], m.box(2, e));
                hi[0] = t4_max(hi[0], m.box(3, e)); hi[1] = t4_max(hi[1], m.box(4, e)); hi[2] = t4_max(hi[2], m.box(5, e));
            }
            if (n > 3) break;
            t4_emit(recs, 2 * abs_start, lo, hi, abs_start, n, leftrun, pstart, pleftrun, fl);  // leaf (blas.rs:106-109)
            if (sp == 0) { done = true; break; }
            const uint32_t x = stk[--sp];
            s = x & 0xFFu; n = (x >> 8) & 0xFFu;
            pstart = t.start + ((x >> 16) & 0xFFu);
            pleftrun = ((x >> 31) ? t.leftrun : 0u) + ((x >> 24) & 0x7Fu);
            leftrun = 0; k = 0; allleft = 0; fl = T4_TF_RIGHT | (t.flags & ~3u);
        }
        if (done) break;
        bool descend = false;
        {
            // centroid bounds (blas.rs:142), then the plane counts of every primitive under this node's planes
            float cmin[3] = {1e30f, 1e30f, 1e30f}, cmax[3] = {-1e30f, -1e30f, -1e30f};
            for (uint32_t j = s; j < s + n; ++j) {
                const T4Cent c = t4_load_cent(cent, m.gid(m.ord(j) & 31u));
                cmin[0] = t4_min(cmin[0], c.x); cmin[1] = t4_min(cmin[1], c.y); cmin[2] = t4_min(cmin[2], c.z);
                cmax[0] = t4_max(cmax[0], c.x); cmax[1] = t4_max(cmax[1], c.y); cmax[2] = t4_max(cmax[2], c.z);
            }
            for (uint32_t j = s; j < s + n; ++j) {
                const uint32_t e = m.ord(j) & 31u;
                const T4Cent c = t4_load_cent(cent, m.gid(e));
                m.ord(j) = e | (t4_plane_count(c.x, cmin[0], cmax[0]) << 5) | (t4_plane_count(c.y, cmin[1], cmax[1]) << 8) |
                           (t4_plane_count(c.z, cmin[2], cmax[2]) << 11);
            }
            float best_cost = 3.402823466e+38f;  // f32::MAX (blas.rs:140)
            uint32_t best = 0xFFFFFFFFu, best_p = 0;
            uint32_t prev_set = 0;  // primitives on the left of the previous candidate (never empty when evaluated)
            uint32_t prev_p = 0;    // pivot of the previous plane of the same axis
            for (uint32_t c = 0; c < 21; ++c) {
                const uint32_t a = c / 7, b = c % 7 + 1;
                // From the second plane of an axis on, everything before the previous pivot is on the left of this plane
                // too (the planes grow with b and partition_shuffle leaves [0, pivot) all-left): the front cursor would
                // walk over that prefix without a swap, so the shuffle starts behind it.
                const uint32_t off = (b == 1) ? 0u : prev_p;
                const uint32_t p = off + t4_shuffle<CAP>(m, s + off, n - off, 5 + 3 * a, b);
                prev_p = p;
                // The cost (blas.rs:149-155) depends only on WHICH primitives ended up left of the pivot.  An empty left
                // side costs NaN (area of the inverted box is +inf, times 0) and a candidate that splits exactly like
                // the previous one costs the same; neither can win under the strict < of blas.rs:156.
                if (p == 0) continue;
                uint32_t set = 0;
                for (uint32_t j = 0; j < p; ++j) set |= 1u << (m.ord(s + j) & 31u);
                if (set == prev_set) continue;
                prev_set = set;
                float L[6] = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
                float R[6] = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
                for (uint32_t j = 0; j < n; ++j) {
                    const uint32_t e = m.ord(s + j) & 31u;
                    const float x0 = m.box(0, e), x1 = m.box(1, e), x2 = m.box(2, e);
                    const float x3 = m.box(3, e), x4 = m.box(4, e), x5 = m.box(5, e);
                    if (j < p) {
                        L[0] = t4_min(L[0], x0); L[1] = t4_min(L[1], x1); L[2] = t4_min(L[2], x2);
                        L[3] = t4_max(L[3], x3); L[4] = t4_max(L[4], x4); L[5] = t4_max(L[5], x5);
                    } else {
                        R[0] = t4_min(R[0], x0); R[1] = t4_min(R[1], x1); R[2] = t4_min(R[2], x2);
                        R[3] = t4_max(R[3], x3); R[4] = t4_max(R[4], x4); R[5] = t4_max(R[5], x5);
                    }
                }
                const float cost = t4_add(t4_mul(t4_area(L, L + 3), t4_u2f(p)), t4_mul(t4_area(R, R + 3), t4_u2f(n - p)));
                if (cost < best_cost) {  // strict, first wins, NaN / inf never win (blas.rs:156)
                    best_cost = cost;
                    best = c;
                    best_p = p;
                }
            }
            if (best == 0xFFFFFFFFu) {
                err |= T4_ERR_DEGENERATE;  // the reference would not terminate (blas.rs:115,139)
            } else {
                (void)t4_shuffle<CAP>(m, s, n, 5 + 3 * (best / 7), best % 7 + 1);  // blas.rs:164: re-shuffle, keep the recorded pivot
                const uint32_t p = best_p;
                t4_emit(recs, 2 * (abs_start + p) + 1, lo, hi, abs_start, n, leftrun, pstart, pleftrun, fl);
                if (p <= 3) A[abs_start] = leftrun + 1;
                stk[sp++] = (s + p) | ((n - p) << 8) | (s << 16) | (k << 24) | (allleft << 31);
                pstart = abs_start; pleftrun = leftrun; leftrun += 1; k += 1; n = p; fl = t.flags & ~3u;
                descend = true;
            }
        }
        if (descend) continue;
        if (sp == 0) break;
        const uint32_t x = stk[--sp];
        s = x & 0xFFu; n = (x >> 8) & 0xFFu;
        pstart = t.start + ((x >> 16) & 0xFFu);
        pleftrun = ((x >> 31) ? t.leftrun : 0u) + ((x >> 24) & 0x7Fu);
        leftrun = 0; k = 0; allleft = 0; fl = T4_TF_RIGHT | (t.flags & ~3u);
    }
    for (uint32_t j = 0; j < t.n; ++j) ids[t.start + j] = m.gid(m.ord(j) & 31u);
    return err;
}
