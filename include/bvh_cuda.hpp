// bvh_cuda.hpp — header-only C++17 mirror of voidin's crates/bvh interface on top of the C ABI (bvh_cuda.h).
// Same names and ownership as the Rust API (crates/bvh/src/lib.rs:5-7): the builder borrows vertices, permutes
// `indices` in place and returns the node vector; errors (where Rust panics) are thrown as bvh_cuda::Error.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "bvh_cuda.h"

namespace bvh_cuda {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

class Context {
  public:
    explicit Context(int device = 0) {
        int rc = bvh_cuda_create(device, &ctx_);
        if (rc) throw Error(rc, "bvh_cuda_create failed: a CUDA device is required (no CPU fallback)");
    }
    ~Context() { bvh_cuda_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    bvh_cuda_ctx* get() const { return ctx_; }
    void check(int rc) const {
        if (rc) throw Error(rc, bvh_cuda_last_error(ctx_));
    }

  private:
    bvh_cuda_ctx* ctx_ = nullptr;
};

struct Bvh {  // crates/bvh/src/blas.rs:206-208
    std::vector<BvhNode> nodes;
};

class BvhBuilder {  // crates/bvh/src/blas.rs:41-103
  public:
    // vertices: 3*n_vertices floats; indices: 3*n_tris u32, permuted in place by build()
    BvhBuilder(const Context& ctx, const float* vertices, size_t n_vertices, uint32_t* indices, size_t n_tris)
        : ctx_(ctx), v_(vertices), nv_(n_vertices), i_(indices), nt_(n_tris) {}
    BvhBuilder& set_bin_number(size_t n) { num_bins_ = n; return *this; }  // inert, as in the reference
    Bvh build() {
        Bvh out;
        out.nodes.resize(2 * nt_);
        uint32_t used = 0;
        ctx_.check(bvh_cuda_blas_build(ctx_.get(), v_, nv_, i_, nt_, out.nodes.data(), out.nodes.size(), &used));
        out.nodes.resize(used);
        return out;
    }

  private:
    const Context& ctx_;
    const float* v_;
    size_t nv_;
    uint32_t* i_;
    size_t nt_;
    size_t num_bins_ = 8;
};

struct Tlas {  // crates/bvh/src/tlas.rs:22-85
    std::vector<TlasNode> nodes;
    std::vector<uint32_t> children;  // 2 per node
    void build(const Context& ctx, const Instance* instances, size_t n_inst, const MeshInfo* meshes, size_t n_mesh) {
        nodes.assign(2 * n_inst + 1, TlasNode{});
        children.assign(2 * (2 * n_inst + 1), 0u);
        ctx.check(bvh_cuda_tlas_build(ctx.get(), instances, n_inst, meshes, n_mesh, nodes.data(), children.data()));
    }
};

}  // namespace bvh_cuda
