// Host harness for voidin_b200/csrc/t4_seq.cuh (TEST INFRASTRUCTURE): compiles the thread-per-sub-tree builder
// with g++ and runs it as the whole build of one small mesh (<= 32 triangles), followed by a host transcription
// of the numbering / emit step (k_scan_* + k_emit in blas_build.cu), so that tests/test_t4_host.py can compare the
// resulting BvhNode array and primitive order with the CPU oracle without a GPU.
// Build: g++ -O2 -std=c++17 -fPIC -ffp-contract=off -shared (see tests/test_t4_host.py).
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../voidin_b200/csrc/t4_seq.cuh"

namespace {
constexpr uint32_t TF_ROOT = 2u;

struct Node { uint32_t w[8]; };  // {min[3], left_first, max[3], count}
}  // namespace

// CAP = capacity the builder is instantiated with: 8 is what libbvh_cuda.so ships (k_t4<8,768>), 32 covers every size.
template <int CAP>
static int build_with_cap(const float* V, const uint32_t* I, uint32_t n, uint32_t stride, uint32_t column,
                          uint32_t* nodes_out /* 8 words x 2n */, uint32_t* n_nodes_out, uint32_t* order_out) {
    if (n == 0 || n > (uint32_t)CAP || stride == 0 || column >= stride) return -1;
    std::vector<float> f((size_t)6 * CAP * stride, 0.0f);
    std::vector<uint32_t> u((size_t)2 * CAP * stride, 0u);
    T4Mem<CAP> m{f.data() + column, u.data() + column, stride};
    std::vector<uint32_t> ids(n);
    std::vector<T4Cent> cent(n);
    // k_setup: centroid ((v0+v1)+v2)/3 (blas.rs:80), per-triangle box folded from +-1e30 (blas.rs:185-186)
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = V + 3 * (size_t)I[3 * i];
        const float* b = V + 3 * (size_t)I[3 * i + 1];
        const float* c = V + 3 * (size_t)I[3 * i + 2];
        for (int k = 0; k < 3; ++k) {
            const float x = a[k], y = b[k], z = c[k];
            (&cent[i].x)[k] = ((x + y) + z) / 3.0f;
            float l = 1e30f, h = -1e30f;
            l = t4_min(l, x); l = t4_min(l, y); l = t4_min(l, z);
            h = t4_max(h, x); h = t4_max(h, y); h = t4_max(h, z);
            m.box(k, i) = l;
            m.box(3 + k, i) = h;
        }
        m.gid(i) = i;
        ids[i] = i;
    }
    std::vector<T4Rec> recs((size_t)2 * n);
    std::memset(recs.data(), 0, sizeof(T4Rec) * recs.size());
    std::vector<uint32_t> A(n + 1, 0u);
    T4Task t{0, n, 0, 0, 0, TF_ROOT};
    const uint32_t err = t4_core<CAP>(t, m, cent.data(), ids.data(), recs.data(), A.data());
    if (err) return -2;
    // numbering: exclusive scan of A (k_scan_*), then k_emit
    std::vector<uint32_t> P(n + 1);
    uint32_t acc = 0;
    for (uint32_t i = 0; i <= n; ++i) { P[i] = acc; acc += A[i]; }
    const uint32_t M = 2 + 2 * P[n];
    std::memset(nodes_out, 0, sizeof(uint32_t) * 8 * 2 * n);
    for (uint32_t slot = 0; slot < 2 * n; ++slot) {
        const uint32_t* r = recs[slot].w;
        const uint32_t count = r[7];
        if (count == 0) continue;
        const uint32_t start = r[3], flags = r[11];
        const bool root = (flags & TF_ROOT) != 0;
        const uint32_t pos = root ? 0u : 2u + 2u * (P[r[9]] + r[10]) + (flags & 1u);
        uint32_t lf, cn;
        if (count > 3) { lf = 2u + 2u * (P[start] + r[8]); cn = 0; }
        else { lf = start; cn = count; }
        if (pos >= 2 * n && !(root && n == 1)) return -3;
        uint32_t* o = nodes_out + 8 * (size_t)pos;
        o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = lf;
        o[4] = r[4]; o[5] = r[5]; o[6] = r[6]; o[7] = cn;
    }
    *n_nodes_out = M;
    for (uint32_t i = 0; i < n; ++i) order_out[i] = ids[i];
    return 0;
}

extern "C" int t4_host_build_cap(const float* V, const uint32_t* I, uint32_t n, uint32_t stride, uint32_t column,
                                 uint32_t* nodes_out, uint32_t* n_nodes_out, uint32_t* order_out, uint32_t cap) {
    switch (cap) {
        case 8: return build_with_cap<8>(V, I, n, stride, column, nodes_out, n_nodes_out, order_out);
        case 16: return build_with_cap<16>(V, I, n, stride, column, nodes_out, n_nodes_out, order_out);
        case 32: return build_with_cap<32>(V, I, n, stride, column, nodes_out, n_nodes_out, order_out);
        default: return -1;
    }
}

extern "C" int t4_host_build(const float* V, const uint32_t* I, uint32_t n, uint32_t stride, uint32_t column,
                             uint32_t* nodes_out, uint32_t* n_nodes_out, uint32_t* order_out) {
    return build_with_cap<32>(V, I, n, stride, column, nodes_out, n_nodes_out, order_out);
}
