// bvh_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A C++17 restatement of the acceleration-structure hot path of pudnax/voidin:
//   crates/bvh/src/blas.rs          (BvhBuilder::build / subdivide / partition / partition_shuffle /
//                                    calculate_bounds, Bvh::traverse_iter)
//   crates/bvh/src/tlas.rs          (Tlas::build / find_best_match)
//   crates/bvh/src/intersection.rs  (Aabb::area, intersect_aabb, Ray::intersect)
//   shaders/utils/bvh.wgsl          (traverse_tlas / instance_intersect / traverse_bvh / fetch_vertex)
//   shaders/utils/intersections.wgsl(intersect_aabb / intersect_trig)
// Each function cites the reference file:line it follows (paths relative to /root/reference).
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path, and no Rust
// toolchain exists in the build image, so this restatement cannot be checked against outputs of the
// reference itself.  It is pinned only by (a) structural self-checks in tests/ (permutation, exact
// subtree boxes, node-count identity, brute-force closest hit) and (b) a second, independently written
// formulation (`oracle_blas_build_model`, the scan/bin form the CUDA kernels use) that must agree
// bit-for-bit with the sequential restatement.
//
// Third-party arithmetic not under /root/reference: glam 0.24.1 (Cargo.lock:844-845).  Semantics
// restated from its published source: Vec3 is 3 scalar f32; `/ f32` is a true division; min/max are
// f32::min/max per component; lerp(a,b,s) = a + (b-a)*s; dot = (x*x'+y*y')+z*z'; cross as usual;
// Mat4::transform_point3(p) = ((X*p.x + Y*p.y) + Z*p.z) + W.  The two assumptions that could move
// bits are behind ORACLE_LERP_ONE_MINUS_S and ORACLE_DIV_BY_RECIPROCAL below.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library.  Build: see oracle/Makefile (g++ -O2 -ffp-contract=off; never -ffast-math).

#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <vector>
#include <algorithm>
#include <chrono>

#include <thread>
#include <atomic>
#include <mutex>

namespace {

constexpr float MAX_DIST = 1e30f;  // intersection.rs:3, math.wgsl:4

struct V3 {
    float x, y, z;
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

// f32::min / f32::max as rustc lowers them on x86-64 (minss/maxss + NaN fix-up): the result is `b`
// only when b compares strictly below (above) `a`, or when `a` is NaN; for (-0,+0) pairs `a` (the
// accumulator in every fold of the reference) is kept.
inline float fmin_rs(float a, float b) { return (a != a) ? b : ((b < a) ? b : a); }
inline float fmax_rs(float a, float b) { return (a != a) ? b : ((b > a) ? b : a); }

inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 vmin(V3 a, V3 b) { return {fmin_rs(a.x, b.x), fmin_rs(a.y, b.y), fmin_rs(a.z, b.z)}; }
inline V3 vmax(V3 a, V3 b) { return {fmax_rs(a.x, b.x), fmax_rs(a.y, b.y), fmax_rs(a.z, b.z)}; }
inline V3 splat(float v) { return {v, v, v}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
inline V3 div_scalar(V3 a, float s) {
#ifdef ORACLE_DIV_BY_RECIPROCAL
    float r = 1.0f / s;
    return {a.x * r, a.y * r, a.z * r};
#else
    return {a.x / s, a.y / s, a.z / s};
#endif
}
inline V3 lerp(V3 a, V3 b, float s) {
#ifdef ORACLE_LERP_ONE_MINUS_S
    return a * (1.0f - s) + b * s;
#else
    return a + (b - a) * s;
#endif
}

struct Aabb {
    V3 min, max;
    // intersection.rs:16-19
    float area() const {
        V3 d = max - min;
        return (d.x * d.y + d.x * d.z + d.y * d.z) * 2.0f;
    }
};

}  // namespace

extern "C" {

// blas.rs:10-17 / bvh.wgsl:11-16
struct BvhNode {
    float min[3];
    uint32_t left_first;
    float max[3];
    uint32_t count;
};
// tlas.rs:7-14 / bvh.wgsl:4-9
struct TlasNode {
    float min[3];
    uint32_t left_right;
    float max[3];
    uint32_t instance_idx;
};
// crates/components/src/shared.rs:67-75 (column-major Mat4)
struct Instance {
    float transform[16];
    float inv_transform[16];
    uint32_t mesh, material, junk[2];
};
// crates/components/src/shared.rs:29-39
struct MeshInfo {
    float min[3];
    uint32_t index_count;
    float max[3];
    uint32_t base_index;
    int32_t vertex_offset;
    uint32_t bvh_index;
    uint32_t junk[2];
};

struct OracleBuildStats {
    uint64_t sum_interior_prims;   // S = sum over interior nodes of their primitive count
    uint32_t interior_nodes;
    uint32_t max_depth;
    uint64_t candidates;           // candidate evaluations (21 per interior node)
    uint64_t unexamined_was_left;  // candidate shuffles whose unexamined element carried flag L
    uint64_t nan_candidates;       // candidates whose cost was NaN (empty left side)
    uint32_t final_pivot_differs;  // interior nodes whose 22nd shuffle returned != recorded pivot
    uint32_t reserved;
};

struct OracleRayStats {
    uint64_t pops;             // stack pops (BLAS + TLAS)
    uint64_t interior_visits;  // interior nodes expanded (both children fetched and tested)
    uint64_t triangle_tests;
    uint64_t instance_visits;  // TLAS leaves entered
    uint64_t max_stack;        // deepest stack seen
    uint64_t hits;
};

enum { ORACLE_OK = 0, ORACLE_EINVAL = -1, ORACLE_EDEGENERATE = -2 };

}  // extern "C"

namespace {

static_assert(sizeof(BvhNode) == 32, "BvhNode must be 32 bytes");
static_assert(sizeof(TlasNode) == 32, "TlasNode must be 32 bytes");
static_assert(sizeof(Instance) == 144, "Instance must be 144 bytes");
static_assert(sizeof(MeshInfo) == 48, "MeshInfo must be 48 bytes");

inline V3 ld3(const float* p) { return {p[0], p[1], p[2]}; }
inline void st3(float* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

struct DegenerateInput {};

// ---------------------------------------------------------------------------------------------
// Sequential restatement of BvhBuilder (blas.rs:41-204)
// ---------------------------------------------------------------------------------------------
struct SeqBuilder {
    const float* vertices;
    size_t n_vertices;
    uint32_t* indices;  // 3N
    size_t n;
    std::vector<V3> centroids;
    std::vector<BvhNode> nodes;
    std::vector<size_t> tri;  // triangle_indices: Vec<usize>
    OracleBuildStats st{};

    V3 vert(uint32_t i) const { return ld3(vertices + 3 * (size_t)i); }

    // blas.rs:184-203
    Aabb calculate_bounds(uint32_t first, uint32_t amount, bool cent) const {
        V3 mx = splat(-MAX_DIST), mn = splat(MAX_DIST);
        for (uint32_t k = 0; k < amount; ++k) {
            size_t idx = tri[(size_t)first + k];
            if (cent) {
                V3 v = centroids[idx];
                mx = vmax(mx, v);
                mn = vmin(mn, v);
            } else {
                for (int c = 0; c < 3; ++c) {
                    V3 v = vert(indices[3 * idx + c]);
                    mx = vmax(mx, v);
                    mn = vmin(mn, v);
                }
            }
        }
        return {mn, mx};
    }

    // blas.rs:168-182
    uint32_t partition_shuffle(int axis, float pos, uint32_t start, uint32_t count) {
        size_t end = (size_t)start + count - 1;
        size_t i = start;
        while (i < end) {
            if (centroids[tri[i]][axis] < pos) {
                i += 1;
            } else {
                std::swap(tri[i], tri[end]);
                end -= 1;
            }
        }
        return (uint32_t)i;
    }

    // blas.rs:135-166
    uint32_t partition(uint32_t start, uint32_t count) {
        const int bins = 8;  // blas.rs:136 (num_bins is never read)
        int optimal_axis = 0;
        float optimal_pos = 0.0f;
        uint32_t optimal_pivot = 0;
        float optimal_cost = FLT_MAX;
        bool any = false;

        Aabb aabb = calculate_bounds(start, count, true);
        for (int axis = 0; axis < 3; ++axis) {
            for (int b = 1; b < bins; ++b) {
                float scale = (float)b / (float)bins;
                float pos = lerp(aabb.min, aabb.max, scale)[axis];
                uint32_t pivot = partition_shuffle(axis, pos, start, count);
                // statistics only: flag of the element left unexamined at slot `pivot`
                if (centroids[tri[pivot]][axis] < pos) st.unexamined_was_left++;
                uint32_t bb1_count = pivot - start;
                uint32_t bb2_count = count - bb1_count;
                Aabb bb1 = calculate_bounds(start, bb1_count, false);
                Aabb bb2 = calculate_bounds(pivot, bb2_count, false);
                float cost = bb1.area() * (float)bb1_count + bb2.area() * (float)bb2_count;
                st.candidates++;
                if (cost != cost) st.nan_candidates++;
                if (cost < optimal_cost) {
                    optimal_axis = axis;
                    optimal_pos = pos;
                    optimal_pivot = pivot;
                    optimal_cost = cost;
                    any = true;
                }
            }
        }
        // The reference would use optimal_pivot = 0 here and underflow / recurse forever
        // (blas.rs:115,139); the oracle reports it instead.
        if (!any) throw DegenerateInput{};
        uint32_t final_pivot = partition_shuffle(optimal_axis, optimal_pos, start, count);
        if (final_pivot != optimal_pivot) st.final_pivot_differs++;
        return optimal_pivot;
    }

    void set_bound(size_t i, const Aabb& b) {
        st3(nodes[i].max, b.max);
        st3(nodes[i].min, b.min);
    }

    // blas.rs:105-128 (recursion made explicit only through the C++ call stack, same order)
    void subdivide(size_t cur, uint32_t start, uint32_t& pool_index, uint32_t depth) {
        st.max_depth = std::max(st.max_depth, depth);
        if (nodes[cur].count <= 3) {
            nodes[cur].left_first = start;
            return;
        }
        uint32_t index = pool_index;
        pool_index += 2;
        nodes[cur].left_first = index;
        st.interior_nodes++;
        st.sum_interior_prims += nodes[cur].count;

        uint32_t pivot = partition(start, nodes[cur].count);
        uint32_t left_count = pivot - start;
        nodes[index].count = left_count;
        set_bound(index, calculate_bounds(start, left_count, false));

        uint32_t right_count = nodes[cur].count - left_count;
        nodes[index + 1].count = right_count;
        set_bound(index + 1, calculate_bounds(pivot, right_count, false));

        subdivide(index, start, pool_index, depth + 1);
        subdivide(index + 1, pivot, pool_index, depth + 1);
        nodes[cur].count = 0;
    }

    // blas.rs:69-103
    uint32_t build() {
        centroids.resize(n);
        for (size_t i = 0; i < n; ++i) {
            V3 a = vert(indices[3 * i]), b = vert(indices[3 * i + 1]), c = vert(indices[3 * i + 2]);
            centroids[i] = div_scalar((a + b) + c, 3.0f);
        }
        tri.resize(n);
        for (size_t i = 0; i < n; ++i) tri[i] = i;
        nodes.assign(2 * n, BvhNode{});
        nodes[0].left_first = 0;
        nodes[0].count = (uint32_t)n;
        set_bound(0, calculate_bounds(0, nodes[0].count, false));
        uint32_t new_node_index = 2;
        subdivide(0, 0, new_node_index, 0);
        nodes.resize(new_node_index);
        std::vector<uint32_t> copy(3 * n);
        for (size_t i = 0; i < n; ++i)
            for (int c = 0; c < 3; ++c) copy[3 * i + c] = indices[3 * tri[i] + c];
        std::memcpy(indices, copy.data(), sizeof(uint32_t) * 3 * n);
        return new_node_index;
    }
};

// ---------------------------------------------------------------------------------------------
// "Model" builder: the formulation the CUDA kernels implement,
// written independently of SeqBuilder.  Must produce byte-identical nodes and primitive order.
//   * one 3-bit plane count per axis per primitive instead of 21 float compares,
//   * partition_shuffle in closed (scan) form,
//   * candidate boxes from 3x8 exact bins plus the <=21 unexamined "special" primitives,
//   * every node computes its own vertex box,
//   * DFS pre-order pair numbering from rank = P[start] + leftrun.
// ---------------------------------------------------------------------------------------------
struct ModelBuilder {
    const float* vertices;
    uint32_t* indices;
    size_t n;
    std::vector<V3> cent, tmin, tmax;  // per triangle
    std::vector<uint32_t> ids, tmp_ids;
    std::vector<uint16_t> flags, tmp_flags;  // per slot: kx | ky<<3 | kz<<6 | special<<15
    std::vector<uint32_t> table;

    struct Rec {  // node record emitted in arbitrary order
        Aabb box;
        uint32_t start, count, leftrun, pstart, pleftrun, is_right, is_root;
    };
    std::vector<Rec> recs;
    std::vector<uint32_t> interior_at_start;

    V3 vert(uint32_t i) const { return ld3(vertices + 3 * (size_t)i); }

    // Closed form of blas.rs:168-182 over slots [s, s+cnt): flag L(j) = (k_axis(j) < b).
    // Returns pivot - s and the slot (absolute) where the unexamined element ended up.
    uint32_t shuffle_scan(uint32_t s, uint32_t cnt, int axis, int b, uint32_t* u_slot) {
        auto isL = [&](uint32_t j) { return (int)((flags[s + j] >> (3 * axis)) & 7) < b; };
        uint32_t nL = 0;
        for (uint32_t j = 0; j < cnt; ++j) nL += isL(j);
        // rank -> position tables sharing one array: R's from the front, L's from the back
        // table[s + m]       = position of the (m+1)-th R from the front
        // table[s+cnt-1-m]   = position of the (m+1)-th L from the back
        uint32_t rf = 0;
        for (uint32_t j = 0; j < cnt; ++j) {
            if (isL(j)) {
                uint32_t lf = j - rf;            // #L before j
                uint32_t lb = nL - lf - 1;       // #L after j
                table[s + cnt - 1 - lb] = j;
            } else {
                table[s + rf] = j;
                rf++;
            }
        }
        // front-examined is a prefix [0,f): j is front-examined iff
        //   #L in [j+2,cnt) + [j <= cnt-2]  >=  RF(j) + 1
        uint32_t f = 0;
        rf = 0;
        {
            std::vector<uint32_t> lsuf;  // lsuf[j] = #L in [j, cnt)
            lsuf.assign(cnt + 3, 0);
            for (int64_t j = (int64_t)cnt - 1; j >= 0; --j) lsuf[j] = lsuf[j + 1] + isL((uint32_t)j);
            for (uint32_t j = 0; j < cnt; ++j) {
                uint32_t lbb = lsuf[j + 2] + ((j + 2 <= cnt) ? 1u : 0u);
                if (lbb >= rf + 1) f++;
                if (!isL(j)) rf++;
            }
        }
        // destinations
        uint32_t lf_f = 0;  // #L in [0,f)
        for (uint32_t j = 0; j < f; ++j) lf_f += isL(j);
        uint32_t lb_f = 0;  // #L in (f,cnt)
        for (uint32_t j = f + 1; j < cnt; ++j) lb_f += isL(j);
        uint32_t pivot = lf_f + lb_f;
        rf = 0;
        for (uint32_t j = 0; j < cnt; ++j) {
            bool L = isL(j);
            uint32_t lf = j - rf;
            uint32_t dst;
            if (j < f) {
                if (L) dst = j;
                else {
                    // m-th front R (m = rf+1) -> q[m-1]-1 with q[0] = cnt
                    uint32_t m = rf + 1;
                    uint32_t qprev = (m == 1) ? cnt : table[s + cnt - 1 - (m - 2)];
                    dst = qprev - 1;
                }
            } else if (j == f) {
                dst = pivot;
            } else {
                if (L) {
                    uint32_t lb = nL - lf - 1;
                    dst = table[s + lb];  // m = lb+1 -> p[m] stored at index m-1
                } else dst = j - 1;
            }
            tmp_ids[s + dst] = ids[s + j];
            tmp_flags[s + dst] = flags[s + j];
            if (!L) rf++;
        }
        std::memcpy(&ids[s], &tmp_ids[s], sizeof(uint32_t) * cnt);
        std::memcpy(&flags[s], &tmp_flags[s], sizeof(uint16_t) * cnt);
        *u_slot = s + pivot;
        return pivot;
    }

    void process(uint32_t start, uint32_t cnt, uint32_t leftrun, uint32_t pstart, uint32_t pleftrun,
                 uint32_t is_right, uint32_t is_root) {
        // own boxes
        V3 vmn = splat(MAX_DIST), vmx = splat(-MAX_DIST), cmn = splat(MAX_DIST), cmx = splat(-MAX_DIST);
        for (uint32_t j = 0; j < cnt; ++j) {
            uint32_t t = ids[start + j];
            vmn = vmin(vmn, tmin[t]);
            vmx = vmax(vmx, tmax[t]);
            cmn = vmin(cmn, cent[t]);
            cmx = vmax(cmx, cent[t]);
        }
        recs.push_back({{vmn, vmx}, start, cnt, leftrun, pstart, pleftrun, is_right, is_root});
        if (cnt <= 3) return;
        interior_at_start[start]++;

        // plane positions and per-primitive plane counts
        float pos[3][8];
        for (int b = 1; b < 8; ++b) {
            V3 p = lerp(cmn, cmx, (float)b / 8.0f);
            pos[0][b] = p.x; pos[1][b] = p.y; pos[2][b] = p.z;
        }
        for (uint32_t j = 0; j < cnt; ++j) {
            V3 c = cent[ids[start + j]];
            uint16_t w = 0;
            for (int a = 0; a < 3; ++a) {
                int k = 0;
                for (int b = 1; b < 8; ++b) k += !(c[a] < pos[a][b]);
                // planes are monotone in b, so {b : !(c<pos_b)} is a prefix and L(b) <=> k < b
                w |= (uint16_t)(k << (3 * a));
            }
            flags[start + j] = w;
        }
        // 21 candidate shuffles (order fixed, independent of costs)
        uint32_t piv[21], uid[21];
        for (int c = 0; c < 21; ++c) {
            uint32_t us;
            piv[c] = shuffle_scan(start, cnt, c / 7, c % 7 + 1, &us);
            uid[c] = ids[us];
            flags[us] |= 0x8000;  // special: keep out of the bins
        }
        // bins over non-special primitives
        Aabb bin[3][8];
        for (auto& ax : bin) for (auto& bb : ax) bb = {splat(MAX_DIST), splat(-MAX_DIST)};
        for (uint32_t j = 0; j < cnt; ++j) {
            uint16_t w = flags[start + j];
            if (w & 0x8000) continue;
            uint32_t t = ids[start + j];
            for (int a = 0; a < 3; ++a) {
                Aabb& bb = bin[a][(w >> (3 * a)) & 7];
                bb.min = vmin(bb.min, tmin[t]);
                bb.max = vmax(bb.max, tmax[t]);
            }
        }
        // per-special plane counts (recomputed from centroid; same arithmetic as above)
        auto kof = [&](uint32_t t, int a) {
            int k = 0;
            for (int b = 1; b < 8; ++b) k += !(cent[t][a] < pos[a][b]);
            return k;
        };
        float best_cost = FLT_MAX;
        int best = -1;
        for (int c = 0; c < 21; ++c) {
            int a = c / 7, b = c % 7 + 1;
            Aabb L{splat(MAX_DIST), splat(-MAX_DIST)}, R = L;
            for (int k = 0; k < 8; ++k) {
                Aabb& side = (k < b) ? L : R;
                side.min = vmin(side.min, bin[a][k].min);
                side.max = vmax(side.max, bin[a][k].max);
            }
            for (int s2 = 0; s2 < 21; ++s2) {
                uint32_t t = uid[s2];
                bool left = (t != uid[c]) && (kof(t, a) < b);
                Aabb& side = left ? L : R;
                side.min = vmin(side.min, tmin[t]);
                side.max = vmax(side.max, tmax[t]);
            }
            uint32_t n1 = piv[c], n2 = cnt - n1;
            float cost = L.area() * (float)n1 + R.area() * (float)n2;
            if (cost < best_cost) { best_cost = cost; best = c; }
        }
        if (best < 0) throw DegenerateInput{};
        uint32_t us;
        shuffle_scan(start, cnt, best / 7, best % 7 + 1, &us);
        for (uint32_t j = 0; j < cnt; ++j) flags[start + j] = 0;
        uint32_t p = piv[best];
        process(start, p, leftrun + 1, start, leftrun, 0, 0);
        process(start + p, cnt - p, 0, start, leftrun, 1, 0);
    }

    uint32_t build(BvhNode* out) {
        cent.resize(n); tmin.resize(n); tmax.resize(n);
        for (size_t i = 0; i < n; ++i) {
            V3 a = vert(indices[3 * i]), b = vert(indices[3 * i + 1]), c = vert(indices[3 * i + 2]);
            cent[i] = div_scalar((a + b) + c, 3.0f);
            tmin[i] = vmin(vmin(vmin(splat(MAX_DIST), a), b), c);
            tmax[i] = vmax(vmax(vmax(splat(-MAX_DIST), a), b), c);
        }
        ids.resize(n); tmp_ids.resize(n); flags.assign(n, 0); tmp_flags.resize(n); table.resize(n + 2);
        for (size_t i = 0; i < n; ++i) ids[i] = (uint32_t)i;
        interior_at_start.assign(n + 1, 0);
        process(0, (uint32_t)n, 0, 0, 0, 0, 1);
        // numbering: rank(X) = (#interior nodes with start < X.start) + leftrun(X)
        std::vector<uint32_t> P(n + 1, 0);
        uint32_t acc = 0;
        for (size_t i = 0; i <= n; ++i) { P[i] = acc; acc += interior_at_start[i]; }
        uint32_t M = 2 + 2 * acc;
        std::memset(out, 0, sizeof(BvhNode) * M);
        for (const Rec& r : recs) {
            uint32_t slot = r.is_root ? 0 : 2 + 2 * (P[r.pstart] + r.pleftrun) + r.is_right;
            BvhNode& nd = out[slot];
            st3(nd.min, r.box.min);
            st3(nd.max, r.box.max);
            if (r.count > 3) { nd.left_first = 2 + 2 * (P[r.start] + r.leftrun); nd.count = 0; }
            else { nd.left_first = r.start; nd.count = r.count; }
        }
        std::vector<uint32_t> copy(3 * n);
        for (size_t i = 0; i < n; ++i)
            for (int c = 0; c < 3; ++c) copy[3 * i + c] = indices[3 * (size_t)ids[i] + c];
        std::memcpy(indices, copy.data(), sizeof(uint32_t) * 3 * n);
        return M;
    }
};

// ---------------------------------------------------------------------------------------------
// Traversal
// ---------------------------------------------------------------------------------------------
constexpr int STACK_CAP = 64;  // reference: 32 (blas.rs:299) / 24 (stack.wgsl:1); overflow there is UB
constexpr int TLAS_STACK_CAP = 256;  // the reference's agglomerative TLAS can be deep (88 levels on a 32x32 lattice)

struct RDist {  // Dist enum, intersection.rs:22-26 with derive(PartialOrd): Hit(_) < Miss
    bool hit;
    float t;
};
inline bool dist_gt(RDist a, RDist b) {
    if (a.hit && b.hit) return a.t > b.t;
    if (!a.hit && !b.hit) return false;
    return !a.hit;  // Miss > Hit
}

// intersection.rs:47-55 (division form)
inline RDist intersect_aabb_rs(V3 o, V3 d, V3 bmin, V3 bmax, float t) {
    V3 tx1 = {(bmin.x - o.x) / d.x, (bmin.y - o.y) / d.y, (bmin.z - o.z) / d.z};
    V3 tx2 = {(bmax.x - o.x) / d.x, (bmax.y - o.y) / d.y, (bmax.z - o.z) / d.z};
    V3 hi = vmax(tx1, tx2), lo = vmin(tx1, tx2);
    float tmax = fmin_rs(hi.x, fmin_rs(hi.y, hi.z));   // min_element = x.min(y.min(z))
    float tmin = fmax_rs(lo.x, fmax_rs(lo.y, lo.z));
    if (tmax >= tmin && tmin < t && tmax > 0.0f) return {true, tmin};
    return {false, 0.0f};
}

// intersection.rs:68-92
inline RDist intersect_tri_rs(V3 o, V3 d, V3 v0, V3 v1, V3 v2) {
    const float EPS = 0.0001f;
    V3 e1 = v1 - v0, e2 = v2 - v0;
    V3 h = cross(d, e2);
    float a = dot(e1, h);
    if (-EPS < a && a < EPS) return {false, 0};
    float f = 1.0f / a;
    V3 s = o - v0;
    float u = f * dot(s, h);
    if (!(0.0f <= u && u <= 1.0f)) return {false, 0};
    V3 q = cross(s, e1);
    float v = f * dot(d, q);
    if (v < 0.0f || u + v > 1.0f) return {false, 0};
    float t = f * dot(e2, q);
    if (t > EPS) return {true, t};
    return {false, 0};
}

// blas.rs:247-295.  Ids are an extension (SURVEY.md §8a "Ids"): tri = left_first + i captured at the
// assignment that lowers t.
inline void traverse_iter_rs(const BvhNode* nodes, const float* vertices, const uint32_t* indices, V3 o,
                             V3 d, float* t_out, uint32_t* tri_out, OracleRayStats* st) {
    uint32_t stack[STACK_CAP];
    int head = 0;
    stack[head++] = 0;
    bool hit = false;
    float t = 0.0f;
    uint32_t tri = 0xFFFFFFFFu;
    while (head > 0) {
        const BvhNode node = nodes[stack[--head]];
        if (st) st->pops++;
        if (node.count > 0) {
            for (uint32_t i = 0; i < node.count; ++i) {
                const uint32_t* idx = indices + 3 * (size_t)(node.left_first + i);
                if (st) st->triangle_tests++;
                RDist r = intersect_tri_rs(o, d, ld3(vertices + 3 * (size_t)idx[0]),
                                           ld3(vertices + 3 * (size_t)idx[1]), ld3(vertices + 3 * (size_t)idx[2]));
                if (r.hit) {
                    if (!hit) { hit = true; t = r.t; tri = node.left_first + i; }
                    else if (r.t < t) { t = r.t; tri = node.left_first + i; }
                    // (t.min(dist) with NaN-free operands)
                }
            }
        } else {
            if (st) st->interior_visits++;
            uint32_t min_index = node.left_first, max_index = node.left_first + 1;
            const BvhNode& a = nodes[min_index];
            const BvhNode& b = nodes[max_index];
            float lim = hit ? t : MAX_DIST;
            RDist min_dist = intersect_aabb_rs(o, d, ld3(a.min), ld3(a.max), lim);
            RDist max_dist = intersect_aabb_rs(o, d, ld3(b.min), ld3(b.max), lim);
            if (dist_gt(min_dist, max_dist)) {
                std::swap(min_index, max_index);
                std::swap(min_dist, max_dist);
            }
            if (!min_dist.hit) continue;
            if (head < STACK_CAP) stack[head++] = min_index;
            if (max_dist.hit && head < STACK_CAP) stack[head++] = max_index;
            if (st && (uint64_t)head > st->max_stack) st->max_stack = head;
        }
    }
    *t_out = hit ? t : MAX_DIST;
    *tri_out = tri;
    if (st && hit) st->hits++;
}

// blas.rs:211-245 (Bvh::traverse, the recursive variant; unused by the reference itself, bvh_cpu.rs:86).  The Vec4/UVec4
// inputs are the Vec3/UVec3 data plus an ignored lane (`truncate()`).  Returns Hit(t) as soon as the node's box is
// hit, with t left at the caller's value when no triangle is closer.
inline RDist traverse_rec_rs(const BvhNode* nodes, const float* vertices, const uint32_t* indices, V3 o, V3 d,
                             size_t node_idx, float t) {
    const BvhNode& node = nodes[node_idx];
    if (!intersect_aabb_rs(o, d, ld3(node.min), ld3(node.max), t).hit) return {false, 0.0f};
    if (node.count > 0) {
        for (uint32_t i = 0; i < node.count; ++i) {
            const uint32_t* idx = indices + 3 * (size_t)(node.left_first + i);
            RDist r = intersect_tri_rs(o, d, ld3(vertices + 3 * (size_t)idx[0]), ld3(vertices + 3 * (size_t)idx[1]),
                                       ld3(vertices + 3 * (size_t)idx[2]));
            if (r.hit) t = fmin_rs(t, r.t);
        }
        return {true, t};
    }
    RDist l = traverse_rec_rs(nodes, vertices, indices, o, d, node.left_first, t);
    if (l.hit) t = fmin_rs(t, l.t);
    RDist r = traverse_rec_rs(nodes, vertices, indices, o, d, (size_t)node.left_first + 1, t);
    if (r.hit) t = fmin_rs(t, r.t);
    return {true, t};
}

// WGSL min/max: IEEE minNum/maxNum (what fminf/fmaxf and the CUDA intrinsics implement)
inline float wmin(float a, float b) { return std::fmin(a, b); }
inline float wmax(float a, float b) { return std::fmax(a, b); }

// intersections.wgsl:13-23
inline float intersect_aabb_w(V3 eye, V3 inv, V3 bmin, V3 bmax, float t) {
    V3 tx1 = {(bmin.x - eye.x) * inv.x, (bmin.y - eye.y) * inv.y, (bmin.z - eye.z) * inv.z};
    V3 tx2 = {(bmax.x - eye.x) * inv.x, (bmax.y - eye.y) * inv.y, (bmax.z - eye.z) * inv.z};
    float tmax = wmin(wmax(tx1.x, tx2.x), wmin(wmax(tx1.y, tx2.y), wmax(tx1.z, tx2.z)));
    float tmin = wmax(wmin(tx1.x, tx2.x), wmax(wmin(tx1.y, tx2.y), wmin(tx1.z, tx2.z)));
    if (tmax >= tmin && tmin < t && tmax > 0.0f) return tmin;
    return MAX_DIST;
}

// intersections.wgsl:25-45
inline bool intersect_trig_w(V3 eye, V3 dir, V3 v0, V3 v1, V3 v2, float* hit) {
    V3 e1 = v1 - v0, e2 = v2 - v0;
    V3 uvec = cross(dir, e2);
    float det = dot(e1, uvec);
    if (det < 1e-10f) return false;
    float inv_det = 1.0f / det;
    V3 orig = eye - v0;
    float u = inv_det * dot(orig, uvec);
    if (u < 0.0f || 1.0f < u) return false;
    V3 vvec = cross(orig, e1);
    float v = inv_det * dot(dir, vvec);
    if (v < 0.0f || u + v > 1.0f) return false;
    float t = inv_det * dot(e2, vvec);
    if (t > 0.0f && t < *hit) { *hit = t; return true; }
    return false;
}

struct Scene {
    const TlasNode* tlas;
    const uint32_t* tlas_children;  // optional side buffer [2*node]; NULL -> unpack left_right
    const Instance* instances;
    const MeshInfo* meshes;
    const BvhNode* bvh_nodes;
    const float* vertices;
    const uint32_t* indices;
};

// bvh.wgsl:30-33
inline V3 fetch_vertex(const Scene& sc, uint32_t idx, const MeshInfo& mesh) {
    uint32_t i = (uint32_t)mesh.vertex_offset + sc.indices[mesh.base_index + idx];
    return ld3(sc.vertices + 3 * (size_t)i);
}

// bvh.wgsl:35-76.  Returns true (and stops) on the first accepted triangle when any_hit is set.
inline bool traverse_bvh_w(const Scene& sc, V3 eye, V3 dir, V3 inv, const MeshInfo& mesh, float* dist,
                           uint32_t* tri, uint32_t inst_id, uint32_t* inst, bool* res_hit, bool any_hit,
                           OracleRayStats* st) {
    uint32_t stack[STACK_CAP];
    int head = 0;
    stack[head++] = mesh.bvh_index;
    float hit = *dist;
    while (head > 0) {
        const BvhNode node = sc.bvh_nodes[stack[--head]];
        if (st) st->pops++;
        if (node.count > 0) {
            for (uint32_t i = 0; i < node.count; ++i) {
                uint32_t idx = node.left_first + i;
                V3 v0 = fetch_vertex(sc, 3 * idx + 0, mesh);
                V3 v1 = fetch_vertex(sc, 3 * idx + 1, mesh);
                V3 v2 = fetch_vertex(sc, 3 * idx + 2, mesh);
                if (st) st->triangle_tests++;
                if (intersect_trig_w(eye, dir, v0, v1, v2, &hit)) {
                    *dist = hit;
                    *tri = idx;
                    *inst = inst_id;
                    *res_hit = true;
                    if (any_hit) return true;
                }
            }
        } else {
            if (st) st->interior_visits++;
            uint32_t min_index = mesh.bvh_index + node.left_first;
            uint32_t max_index = mesh.bvh_index + node.left_first + 1;
            const BvhNode& a = sc.bvh_nodes[min_index];
            const BvhNode& b = sc.bvh_nodes[max_index];
            float min_dist = intersect_aabb_w(eye, inv, ld3(a.min), ld3(a.max), hit);
            float max_dist = intersect_aabb_w(eye, inv, ld3(b.min), ld3(b.max), hit);
            if (min_dist > max_dist) {
                std::swap(min_index, max_index);
                std::swap(min_dist, max_dist);
            }
            if (min_dist >= hit) continue;
            if (max_dist <= hit && head < STACK_CAP) stack[head++] = max_index;
            if (head < STACK_CAP) stack[head++] = min_index;
            if (st && (uint64_t)head > st->max_stack) st->max_stack = head;
        }
    }
    return false;
}

// mat4 * vec4 as column sums, left to right (bvh.wgsl:82-83)
inline V3 mat_mul(const float* m, V3 p, float w) {
    V3 c0 = ld3(m), c1 = ld3(m + 4), c2 = ld3(m + 8), c3 = ld3(m + 12);
    return ((c0 * p.x + c1 * p.y) + c2 * p.z) + c3 * w;
}

// bvh.wgsl:89-123 (+ :78-87).  t starts at tmax (reference: MAX_DIST).
inline void traverse_tlas_w(const Scene& sc, V3 eye, V3 dir, float tmax, bool any_hit, float* t_out,
                            uint32_t* tri_out, uint32_t* inst_out, bool* hit_out, OracleRayStats* st) {
    V3 inv = {1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z};
    uint32_t stack[TLAS_STACK_CAP];
    int head = 0;
    stack[head++] = 0;
    float dist = tmax;
    uint32_t tri = 0xFFFFFFFFu, inst = 0xFFFFFFFFu;
    bool res_hit = false;
    while (head > 0) {
        uint32_t ni = stack[--head];
        const TlasNode node = sc.tlas[ni];
        if (st) st->pops++;
        if (node.left_right == 0) {
            if (st) st->instance_visits++;
            const Instance& in = sc.instances[node.instance_idx];
            const MeshInfo& mesh = sc.meshes[in.mesh];
            V3 e2 = mat_mul(in.inv_transform, eye, 1.0f);
            V3 d2 = mat_mul(in.inv_transform, dir, 0.0f);
            V3 inv2 = {1.0f / d2.x, 1.0f / d2.y, 1.0f / d2.z};
            bool stop = traverse_bvh_w(sc, e2, d2, inv2, mesh, &dist, &tri, node.instance_idx, &inst, &res_hit,
                                       any_hit, st);
            if (stop) break;
        } else {
            if (st) st->interior_visits++;
            uint32_t min_index, max_index;
            if (sc.tlas_children) {
                min_index = sc.tlas_children[2 * (size_t)ni];
                max_index = sc.tlas_children[2 * (size_t)ni + 1];
            } else {
                min_index = node.left_right & 0xffffu;
                max_index = node.left_right >> 16;
            }
            const TlasNode& a = sc.tlas[min_index];
            const TlasNode& b = sc.tlas[max_index];
            float min_dist = intersect_aabb_w(eye, inv, ld3(a.min), ld3(a.max), dist);
            float max_dist = intersect_aabb_w(eye, inv, ld3(b.min), ld3(b.max), dist);
            if (min_dist > max_dist) {
                std::swap(min_index, max_index);
                std::swap(min_dist, max_dist);
            }
            if (min_dist >= dist) continue;
            if (max_dist < dist && head < TLAS_STACK_CAP) stack[head++] = max_index;
            if (head < TLAS_STACK_CAP) stack[head++] = min_index;
            if (st && (uint64_t)head > st->max_stack) st->max_stack = head;
        }
    }
    *t_out = res_hit ? dist : MAX_DIST;
    *tri_out = tri;
    *inst_out = inst;
    *hit_out = res_hit;
    if (st && res_hit) st->hits++;
}

// threads <= 1: one serial loop, like the reference (bvh_cpu.rs:73).  threads > 1: rays split over
// std::thread workers in chunks of 4096 (NOT reference behaviour; used only for labelled all-core timings
// and to make large parity samples affordable).
template <class F>
void parallel_rays(size_t n_rays, int threads, OracleRayStats* total, F&& one) {
    if (threads <= 1) {
        for (size_t r = 0; r < n_rays; ++r) one(r, total);
        return;
    }
    std::atomic<size_t> next{0};
    std::mutex mu;
    std::vector<std::thread> pool;
    for (int w = 0; w < threads; ++w)
        pool.emplace_back([&] {
            OracleRayStats loc{};
            for (;;) {
                size_t b = next.fetch_add(4096);
                if (b >= n_rays) break;
                size_t e = std::min(n_rays, b + 4096);
                for (size_t r = b; r < e; ++r) one(r, &loc);
            }
            std::lock_guard<std::mutex> g(mu);
            total->pops += loc.pops; total->interior_visits += loc.interior_visits;
            total->triangle_tests += loc.triangle_tests; total->instance_visits += loc.instance_visits;
            total->hits += loc.hits; total->max_stack = std::max(total->max_stack, loc.max_stack);
        });
    for (auto& t : pool) t.join();
}

bool indices_ok(const uint32_t* indices, size_t n_tris, size_t n_vertices) {
    for (size_t i = 0; i < 3 * n_tris; ++i)
        if (indices[i] >= n_vertices) return false;
    return true;
}

}  // namespace

extern "C" {

// BvhBuilder::new(vertices, indices).build()  — blas.rs:51-103.
// indices are permuted in place; nodes_out needs capacity 2*n_tris; prim_order_out (optional, n_tris)
// receives the final triangle_indices (original triangle id per slot).
int oracle_blas_build(const float* vertices, size_t n_vertices, uint32_t* indices, size_t n_tris,
                      BvhNode* nodes_out, uint32_t* n_nodes_out, uint32_t* prim_order_out,
                      OracleBuildStats* stats) {
    if (n_tris == 0 || !vertices || !indices || !nodes_out) return ORACLE_EINVAL;  // blas.rs:84 would panic
    if (!indices_ok(indices, n_tris, n_vertices)) return ORACLE_EINVAL;
    SeqBuilder b;
    b.vertices = vertices; b.n_vertices = n_vertices; b.indices = indices; b.n = n_tris;
    std::vector<uint32_t> saved(indices, indices + 3 * n_tris);
    uint32_t m;
    try {
        m = b.build();
    } catch (const DegenerateInput&) {
        std::memcpy(indices, saved.data(), sizeof(uint32_t) * 3 * n_tris);
        return ORACLE_EDEGENERATE;
    }
    std::memcpy(nodes_out, b.nodes.data(), sizeof(BvhNode) * m);
    if (n_nodes_out) *n_nodes_out = m;
    if (prim_order_out)
        for (size_t i = 0; i < n_tris; ++i) prim_order_out[i] = (uint32_t)b.tri[i];
    if (stats) *stats = b.st;
    return ORACLE_OK;
}

// Same contract, computed with the scan/bin/rank formulation (the CUDA algorithm restated on the CPU).
int oracle_blas_build_model(const float* vertices, size_t n_vertices, uint32_t* indices, size_t n_tris,
                            BvhNode* nodes_out, uint32_t* n_nodes_out, uint32_t* prim_order_out) {
    if (n_tris == 0 || !vertices || !indices || !nodes_out) return ORACLE_EINVAL;
    if (!indices_ok(indices, n_tris, n_vertices)) return ORACLE_EINVAL;
    ModelBuilder b;
    b.vertices = vertices; b.indices = indices; b.n = n_tris;
    std::vector<uint32_t> saved(indices, indices + 3 * n_tris);
    uint32_t m;
    try {
        m = b.build(nodes_out);
    } catch (const DegenerateInput&) {
        std::memcpy(indices, saved.data(), sizeof(uint32_t) * 3 * n_tris);
        return ORACLE_EDEGENERATE;
    }
    if (n_nodes_out) *n_nodes_out = m;
    if (prim_order_out) std::memcpy(prim_order_out, b.ids.data(), sizeof(uint32_t) * n_tris);
    return ORACLE_OK;
}

// Sequential partition_shuffle (blas.rs:168-182) on a bare flag array, for the scan-form cross-check:
// flags[j] != 0 means "centroid < pos" (L).  Permutes ids in place, returns the pivot.
uint32_t oracle_shuffle_seq(uint32_t* ids, const uint8_t* flags_by_id, uint32_t n) {
    if (n == 0) return 0;
    size_t end = n - 1, i = 0;
    while (i < end) {
        if (flags_by_id[ids[i]]) i++;
        else { std::swap(ids[i], ids[end]); end--; }
    }
    return (uint32_t)i;
}

// Tlas::build — tlas.rs:31-85.  nodes_out: 2I+1; children_out (optional): 2*(2I+1) unpacked child ids.
int oracle_tlas_build(const Instance* instances, size_t n_inst, const MeshInfo* meshes, size_t n_mesh,
                      TlasNode* nodes_out, uint32_t* children_out, uint64_t* stats_calls,
                      uint64_t* stats_pairs) {
    if (n_inst == 0 || !instances || !meshes || !nodes_out) return ORACLE_EINVAL;  // mesh/mod.rs:280-282
    for (size_t i = 0; i < n_inst; ++i)
        if (instances[i].mesh >= n_mesh) return ORACLE_EINVAL;
    const size_t total = 2 * n_inst + 1;
    std::vector<TlasNode> nodes(total, TlasNode{});
    std::vector<uint32_t> kids(2 * total, 0);
    uint64_t calls = 0, pairs = 0;
    // tlas.rs:34-54
    for (size_t i = 0; i < n_inst; ++i) {
        const Instance& inst = instances[i];
        const MeshInfo& mesh = meshes[inst.mesh];
        V3 bound[2] = {ld3(mesh.min), ld3(mesh.max)};
        V3 mn = bound[0], mx = bound[1];  // fold seed = local box (tlas.rs:39)
        for (int c = 0; c < 8; ++c) {
            // [i&1, i&2, i&4].map(|i| i == 0).map(usize::from): bit clear -> index 1 (max)
            int ix = (c & 1) == 0, iy = (c & 2) == 0, iz = (c & 4) == 0;
            V3 p = {bound[ix].x, bound[iy].y, bound[iz].z};
            V3 q = mat_mul(inst.transform, p, 1.0f);  // transform_point3: ((X*x + Y*y) + Z*z) + W
            mn = vmin(mn, q);
            mx = vmax(mx, q);
        }
        TlasNode& nd = nodes[i + 1];
        st3(nd.min, mn); st3(nd.max, mx);
        nd.left_right = 0;
        nd.instance_idx = (uint32_t)i;
    }
    // tlas.rs:87-105
    std::vector<size_t> node_indices(n_inst);
    for (size_t i = 0; i < n_inst; ++i) node_indices[i] = i + 1;
    auto find_best_match = [&](size_t num_unused, size_t target) -> size_t {
        calls++;
        float smallest = 1e30f;
        size_t best_idx = target;
        const TlasNode& tn = nodes[node_indices[target]];
        V3 tmn = ld3(tn.min), tmx = ld3(tn.max);
        for (size_t i = 0; i < num_unused; ++i) {
            if (target == i) continue;
            pairs++;
            const TlasNode& bn = nodes[node_indices[i]];
            Aabb u{vmin(tmn, ld3(bn.min)), vmax(tmx, ld3(bn.max))};
            float sa = u.area();
            if (sa < smallest) { smallest = sa; best_idx = i; }
        }
        return best_idx;
    };
    // tlas.rs:56-84
    size_t instance_count = n_inst;
    size_t nodes_used = 1 + instance_count;
    size_t a = 0;
    size_t b = find_best_match(instance_count, a);
    while (instance_count > 0) {
        size_t c = find_best_match(instance_count, b);
        if (a == c) {
            size_t idx_a = node_indices[a], idx_b = node_indices[b];
            const TlasNode na = nodes[idx_a], nb = nodes[idx_b];
            if (nodes_used >= total) return ORACLE_EINVAL;  // cannot happen (2I+1 identity); guard only
            TlasNode& nn = nodes[nodes_used];
            st3(nn.min, vmin(ld3(na.min), ld3(nb.min)));
            st3(nn.max, vmax(ld3(na.max), ld3(nb.max)));
            nn.left_right = (uint32_t)idx_a + ((uint32_t)idx_b << 16);  // release-mode wrapping (tlas.rs:71)
            nn.instance_idx = 0xFFFFFFFFu;
            kids[2 * nodes_used] = (uint32_t)idx_a;
            kids[2 * nodes_used + 1] = (uint32_t)idx_b;
            node_indices[a] = nodes_used;
            nodes_used += 1;
            node_indices[b] = node_indices[instance_count - 1];
            instance_count -= 1;
            b = find_best_match(instance_count, a);
        } else {
            a = b;
            b = c;
        }
    }
    nodes[0] = nodes[node_indices[a]];
    kids[0] = kids[2 * node_indices[a]];
    kids[1] = kids[2 * node_indices[a] + 1];
    std::memcpy(nodes_out, nodes.data(), sizeof(TlasNode) * total);
    if (children_out) std::memcpy(children_out, kids.data(), sizeof(uint32_t) * 2 * total);
    if (stats_calls) *stats_calls = calls;
    if (stats_pairs) *stats_pairs = pairs;
    return ORACLE_OK;
}

// Bvh::traverse_iter over R rays (blas.rs:247-295); `indices` are the permuted ones.
// threads <= 1: serial like the reference (bvh_cpu.rs:73); >1: OpenMP over rays (NOT reference behaviour).
int oracle_trace_blas(const BvhNode* nodes, const float* vertices, const uint32_t* indices,
                      const float* ray_o, const float* ray_d, size_t n_rays, float* t_out,
                      uint32_t* tri_out, OracleRayStats* stats, int threads) {
    OracleRayStats total{};
    parallel_rays(n_rays, threads, &total, [&](size_t r, OracleRayStats* st) {
        traverse_iter_rs(nodes, vertices, indices, ld3(ray_o + 3 * r), ld3(ray_d + 3 * r), t_out + r, tri_out + r, st);
    });
    if (stats) *stats = total;
    return ORACLE_OK;
}

// Bvh::traverse over R rays (blas.rs:211-245): hit_out[r] = 1 for Hit(t_out[r]), 0 for Miss.
int oracle_trace_blas_recursive(const BvhNode* nodes, const float* vertices, const uint32_t* indices, const float* ray_o,
                                const float* ray_d, size_t n_rays, uint32_t node_idx, float t0, float* t_out,
                                uint8_t* hit_out) {
    for (size_t r = 0; r < n_rays; ++r) {
        RDist d = traverse_rec_rs(nodes, vertices, indices, ld3(ray_o + 3 * r), ld3(ray_d + 3 * r), node_idx, t0);
        hit_out[r] = d.hit ? 1 : 0;
        t_out[r] = d.hit ? d.t : MAX_DIST;
    }
    return ORACLE_OK;
}

// traverse_tlas over R rays (bvh.wgsl:89-123).  any_hit != 0: occluded_out[r] = traverse_tlas(ray).hit
// computed with early exit (raytraced_shadows.wgsl:98-102); t/tri/inst outputs may be NULL then.
int oracle_trace_scene(const TlasNode* tlas, const uint32_t* tlas_children, const Instance* instances,
                       const MeshInfo* meshes, const BvhNode* bvh_nodes, const float* vertices,
                       const uint32_t* indices, const float* ray_o, const float* ray_d, size_t n_rays,
                       float tmax, int any_hit, float* t_out, uint32_t* tri_out, uint32_t* inst_out,
                       uint8_t* occluded_out, OracleRayStats* stats, int threads) {
    Scene sc{tlas, tlas_children, instances, meshes, bvh_nodes, vertices, indices};
    OracleRayStats total{};
    auto one = [&](size_t r, OracleRayStats* st) {
        float t; uint32_t tri, inst; bool hit;
        traverse_tlas_w(sc, ld3(ray_o + 3 * r), ld3(ray_d + 3 * r), tmax, any_hit != 0, &t, &tri, &inst, &hit, st);
        if (t_out) t_out[r] = t;
        if (tri_out) tri_out[r] = tri;
        if (inst_out) inst_out[r] = inst;
        if (occluded_out) occluded_out[r] = hit ? 1 : 0;
    };
    parallel_rays(n_rays, threads, &total, one);
    if (stats) *stats = total;
    return ORACLE_OK;
}

// Brute-force closest hit over all triangles of one mesh (self-check helper, not from the reference).
// mode 0: intersection.rs semantics; mode 1: intersections.wgsl semantics.
int oracle_brute_force(const float* vertices, const uint32_t* indices, size_t n_tris, const float* ray_o,
                       const float* ray_d, size_t n_rays, int mode, float* t_out) {
    for (size_t r = 0; r < n_rays; ++r) {
        V3 o = ld3(ray_o + 3 * r), d = ld3(ray_d + 3 * r);
        float best = MAX_DIST;
        for (size_t i = 0; i < n_tris; ++i) {
            V3 v0 = ld3(vertices + 3 * (size_t)indices[3 * i]);
            V3 v1 = ld3(vertices + 3 * (size_t)indices[3 * i + 1]);
            V3 v2 = ld3(vertices + 3 * (size_t)indices[3 * i + 2]);
            if (mode == 0) {
                RDist h = intersect_tri_rs(o, d, v0, v1, v2);
                if (h.hit && h.t < best) best = h.t;
            } else {
                float h = best;
                if (intersect_trig_w(o, d, v0, v1, v2, &h)) best = h;
            }
        }
        t_out[r] = best;
    }
    return ORACLE_OK;
}

// Primary rays (src/bin/bvh_cpu.rs:72-84).  clip_to_world column-major.
int oracle_gen_primary_rays(const float* m, uint32_t width, uint32_t height, float* ro, float* rd) {
    auto mv = [&](float x, float y, float z, float w, float* o) {
        for (int r = 0; r < 4; ++r) o[r] = ((m[r] * x + m[4 + r] * y) + m[8 + r] * z) + m[12 + r] * w;
    };
    for (size_t i = 0; i < (size_t)width * height; ++i) {
        float x = (float)(i % width) / (float)width;
        float y = (float)(i / height) / (float)height;  // sic: bvh_cpu.rs:75 divides by HEIGHT
        x = (x - 0.5f) * 2.0f;
        y = (y - 0.5f) * -2.0f;
        float vp[4], vt[4];
        mv(x, y, 1.0f, 1.0f, vp);
        mv(x, y, 0.0f, 1.0f, vt);
        ro[3 * i] = vp[0] / vp[3]; ro[3 * i + 1] = vp[1] / vp[3]; ro[3 * i + 2] = vp[2] / vp[3];
        const float rl = 1.0f / std::sqrt((vt[0] * vt[0] + vt[1] * vt[1]) + vt[2] * vt[2]);  // glam normalize = v * length_recip
        rd[3 * i] = vt[0] * rl; rd[3 * i + 1] = vt[1] * rl; rd[3 * i + 2] = vt[2] * rl;
    }
    return ORACLE_OK;
}

// raytraced_shadows.wgsl:93-99: ray_new(pos + nor * 0.0001, light.position - pos)
int oracle_gen_shadow_rays(const float* pos, const float* nor, size_t n, const float* light, float* ro, float* rd) {
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            ro[3 * i + k] = pos[3 * i + k] + nor[3 * i + k] * 0.0001f;
            rd[3 * i + k] = light[k] - pos[3 * i + k];
        }
    return ORACLE_OK;
}

// Rect area light (corner order of crates/pools/src/light.rs:28-52): same origin, direction toward
// (p0 + (p1 - p0) * u) + (p3 - p0) * v.  The reference has no such shader yet (README TODO "raytraced shadows" for area lights);
// this is the CPU twin of csrc/raygen.cu k_gen_area_shadow, same operation order, no contraction.
int oracle_gen_area_shadow_rays(const float* pos, const float* nor, const float* uv, size_t n, const float* corners, float* ro, float* rd) {
    for (size_t i = 0; i < n; ++i) {
        const float u = uv[2 * i], v = uv[2 * i + 1];
        for (int k = 0; k < 3; ++k) {
            const float p = pos[3 * i + k];
            const float o = p + nor[3 * i + k] * 0.0001f;
            const float t = (corners[k] + (corners[3 + k] - corners[k]) * u) + (corners[9 + k] - corners[k]) * v;
            ro[3 * i + k] = o;
            rd[3 * i + k] = t - o;
        }
    }
    return ORACLE_OK;
}

// shaders/compute_update.wgsl:12-27 with (sin, cos) of the angle as inputs (WGSL leaves their precision open);
// matrix product = four-term column sums left to right (math.wgsl from_rotation_z, column-major).  With
// update_inverse the stale inv_transform is multiplied by from_rotation_z(-angle) on the right.
static void mat_mul44(const float* A, const float* B, float* out) {
    for (int j = 0; j < 4; ++j)
        for (int r = 0; r < 4; ++r)
            out[4 * j + r] = ((A[r] * B[4 * j] + A[4 + r] * B[4 * j + 1]) + A[8 + r] * B[4 * j + 2]) + A[12 + r] * B[4 * j + 3];
}
int oracle_instances_rotate_z(Instance* inst, const uint32_t* ids, size_t n, float s, float c, int update_inverse) {
    for (size_t i = 0; i < n; ++i) {
        Instance* in = inst + (ids ? ids[i] : (uint32_t)i);
        float R[16] = {0}, out[16];
        const float sg = (in->transform[14] > -15.0f) ? s : -s;
        R[0] = c; R[1] = sg; R[4] = -sg; R[5] = c; R[10] = 1.0f; R[15] = 1.0f;
        mat_mul44(R, in->transform, out);
        for (int k = 0; k < 16; ++k) in->transform[k] = out[k];
        if (update_inverse) {
            R[1] = -sg; R[4] = sg;
            mat_mul44(in->inv_transform, R, out);
            for (int k = 0; k < 16; ++k) in->inv_transform[k] = out[k];
        }
    }
    return ORACLE_OK;
}

int oracle_max_threads(void) {
    unsigned h = std::thread::hardware_concurrency();
    return h ? (int)h : 1;
}

}  // extern "C"
