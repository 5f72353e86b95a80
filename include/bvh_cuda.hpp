// bvh_cuda.hpp — header-only C++17 mirror of voidin's crates/bvh interface on top of the C ABI (bvh_cuda.h).
// Same names and ownership as the Rust API (crates/bvh/src/lib.rs:5-7): the builder borrows vertices, permutes
// `indices` in place and returns the node vector; errors (where Rust panics) are thrown as bvh_cuda::Error.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "bvh_cuda.h"
#include "bvh_cuda_models.h"

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

struct Ray {  // crates/bvh/src/intersection.rs:22-45
    float orig[3];
    float dir[3];
};

struct Dist {  // crates/bvh/src/intersection.rs:6-14: Hit(f32) | Miss
    bool hit = false;
    float t = BVH_CUDA_MAX_DIST;
    uint32_t triangle = BVH_CUDA_NO_HIT;  // id in the permuted order (the reference returns none)
};

struct Bvh {  // crates/bvh/src/blas.rs:206-208
    std::vector<BvhNode> nodes;

    // Bvh::traverse_iter (blas.rs:247-295) for a batch of rays; vertices 3 floats each, indices the PERMUTED triples
    std::vector<Dist> traverse_iter(const Context& ctx, const float* vertices, size_t n_vertices, const uint32_t* indices,
                                    size_t n_tris, const std::vector<Ray>& rays) const {
        std::vector<float> o(3 * rays.size()), d(3 * rays.size()), t(rays.size());
        std::vector<uint32_t> tri(rays.size());
        for (size_t r = 0; r < rays.size(); ++r)
            for (int k = 0; k < 3; ++k) { o[3 * r + k] = rays[r].orig[k]; d[3 * r + k] = rays[r].dir[k]; }
        ctx.check(bvh_cuda_trace_blas(ctx.get(), nodes.data(), nodes.size(), vertices, n_vertices, indices, n_tris, o.data(), d.data(),
                                      rays.size(), t.data(), tri.data()));
        std::vector<Dist> out(rays.size());
        for (size_t r = 0; r < rays.size(); ++r) { out[r].hit = tri[r] != BVH_CUDA_NO_HIT; out[r].t = t[r]; out[r].triangle = tri[r]; }
        return out;
    }
    // Bvh::traverse (blas.rs:211-245), the recursive variant: start node and distance bound from the caller
    Dist traverse(const Context& ctx, const float* vertices, size_t n_vertices, const uint32_t* indices, size_t n_tris, const Ray& ray,
                  uint32_t node_idx, float t0) const {
        float t = 0.0f;
        uint8_t hit = 0;
        ctx.check(bvh_cuda_trace_blas_recursive(ctx.get(), nodes.data(), nodes.size(), vertices, n_vertices, indices, n_tris, ray.orig,
                                                ray.dir, 1, node_idx, t0, &t, &hit));
        Dist out;
        out.hit = hit != 0;
        out.t = t;
        return out;
    }
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

struct TraceResult {  // shaders/utils/bvh.wgsl:18-24, plus the ids the shader does not return
    bool hit;
    float dist;
    uint32_t triangle, instance;
};

// The trace bind group (crates/pools/src/mesh/mod.rs:136-238) uploaded once; traverse_tlas / occluded have the semantics of
// shaders/utils/bvh.wgsl:89-123 and src/bin/raytraced_shadows.wgsl:98-102.
class Scene {
  public:
    Scene(const Context& ctx, const Tlas& tlas, const Instance* instances, size_t n_inst, const MeshInfo* meshes, size_t n_mesh,
          const BvhNode* bvh_nodes, size_t n_bvh_nodes, const float* vertices, size_t n_vertices, const uint32_t* indices, size_t n_indices)
        : ctx_(ctx) {
        BvhCudaSceneDesc d{};
        d.tlas_nodes = tlas.nodes.data(); d.n_tlas_nodes = tlas.nodes.size();
        d.tlas_children = tlas.children.empty() ? nullptr : tlas.children.data();
        d.instances = instances; d.n_instances = n_inst;
        d.meshes = meshes; d.n_meshes = n_mesh;
        d.bvh_nodes = bvh_nodes; d.n_bvh_nodes = n_bvh_nodes;
        d.vertices = vertices; d.n_vertices = n_vertices;
        d.indices = indices; d.n_indices = n_indices;
        ctx_.check(bvh_cuda_scene_upload(ctx_.get(), &d, &scene_));
    }
    ~Scene() { bvh_cuda_scene_free(ctx_.get(), scene_); }
    Scene(const Scene&) = delete;
    Scene& operator=(const Scene&) = delete;

    std::vector<TraceResult> traverse_tlas(const std::vector<Ray>& rays, float tmax = BVH_CUDA_MAX_DIST) const {
        std::vector<float> o, d;
        split(rays, o, d);
        std::vector<float> t(rays.size());
        std::vector<uint32_t> tri(rays.size()), inst(rays.size());
        ctx_.check(bvh_cuda_trace_closest(ctx_.get(), scene_, o.data(), d.data(), rays.size(), tmax, t.data(), tri.data(), inst.data()));
        std::vector<TraceResult> out(rays.size());
        for (size_t r = 0; r < rays.size(); ++r) out[r] = TraceResult{tri[r] != BVH_CUDA_NO_HIT, t[r], tri[r], inst[r]};
        return out;
    }
    std::vector<uint8_t> occluded(const std::vector<Ray>& rays, float tmax = BVH_CUDA_MAX_DIST) const {
        std::vector<float> o, d;
        split(rays, o, d);
        std::vector<uint8_t> occ(rays.size());
        ctx_.check(bvh_cuda_trace_any(ctx_.get(), scene_, o.data(), d.data(), rays.size(), tmax, occ.data()));
        return occ;
    }

  private:
    static void split(const std::vector<Ray>& rays, std::vector<float>& o, std::vector<float>& d) {
        o.resize(3 * rays.size());
        d.resize(3 * rays.size());
        for (size_t r = 0; r < rays.size(); ++r)
            for (int k = 0; k < 3; ++k) { o[3 * r + k] = rays[r].orig[k]; d[3 * r + k] = rays[r].dir[k]; }
    }
    const Context& ctx_;
    bvh_cuda_scene* scene_ = nullptr;
};

// ---- asset loaders (host only): ObjModel::import / GltfDocument::import (crates/app/src/models) --------------------------
struct Mesh {  // MeshRef of crates/pools/src/mesh/mod.rs:24-31
    std::vector<float> vertices, normals, tangents, tex_coords;
    std::vector<uint32_t> indices;
    int32_t material = -1;
    std::string name;
};
struct ModelInstance {
    float transform[16];  // column-major
    uint32_t mesh;
    int32_t material;
};
struct Model {
    std::vector<Mesh> meshes;
    std::vector<ModelInstance> instances;  // glTF: get_scene_instances(Mat4::IDENTITY)
};

inline Model import_model(const std::string& path, bool gltf) {
    bvh_cuda_model* h = nullptr;
    const int rc = gltf ? bvh_cuda_model_load_gltf(path.c_str(), &h) : bvh_cuda_model_load_obj(path.c_str(), &h);
    if (rc) throw Error(rc, bvh_cuda_model_last_error());
    Model out;
    for (size_t i = 0; i < bvh_cuda_model_mesh_count(h); ++i) {
        BvhCudaMeshView v;
        bvh_cuda_model_mesh(h, i, &v);
        Mesh m;
        m.vertices.assign(v.positions, v.positions + 3 * v.n_vertices);
        if (v.normals) m.normals.assign(v.normals, v.normals + 3 * v.n_normals);
        if (v.tangents) m.tangents.assign(v.tangents, v.tangents + 4 * v.n_vertices);
        if (v.texcoords) m.tex_coords.assign(v.texcoords, v.texcoords + 2 * v.n_texcoords);
        m.indices.assign(v.indices, v.indices + v.n_indices);
        m.material = v.material;
        m.name = v.name ? v.name : "";
        out.meshes.push_back(std::move(m));
    }
    for (size_t i = 0; i < bvh_cuda_model_instance_count(h); ++i) {
        BvhCudaInstanceView v;
        bvh_cuda_model_instance(h, i, &v);
        ModelInstance mi;
        for (int k = 0; k < 16; ++k) mi.transform[k] = v.transform[k];
        mi.mesh = v.mesh;
        mi.material = v.material;
        out.instances.push_back(mi);
    }
    bvh_cuda_model_free(h);
    return out;
}
struct ObjModel {  // crates/app/src/models/mod.rs:20-57
    static Model import(const std::string& path) { return import_model(path, false); }
};
struct GltfDocument {  // crates/app/src/models/gltf_model/mod.rs
    static Model import(const std::string& path) { return import_model(path, true); }
};

}  // namespace bvh_cuda
