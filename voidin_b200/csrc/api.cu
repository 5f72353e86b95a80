// api.cu — context management and the extern "C" entry points of include/bvh_cuda.h.
// Host-pointer entry points stage through device buffers owned by the call and forward to the `_dev`
// implementations; nothing here computes on the CPU.
#include "common.cuh"

#include <cstdlib>

#include <cstdio>
#include <cstring>
#include <new>

int blas_t2_occupancy();
int blas_t2b_setup();
int blas_t2w_occupancy();
int blas_t1_coop_occupancy();
int blas_tc_setup(int cluster_size);
int blas_tc_log(unsigned long long* out, unsigned int cap_rows);
int blas_t1_timing(unsigned long long* out32);
int blas_t1_blocks(unsigned long long* out2048);
int blas_t1_pull(unsigned long long* out4096);

int ctx_fail(bvh_cuda_ctx* ctx, int code, const char* what) {
    if (ctx) ctx->err = what ? what : "";
    return code;
}

int ctx_cuda_fail(bvh_cuda_ctx* ctx, cudaError_t e, const char* where) {
    if (ctx) {
        ctx->err = std::string(where ? where : "cuda") + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    }
    cudaGetLastError();  // clear the sticky-less error state
    return e == cudaErrorMemoryAllocation ? BVH_CUDA_ENOMEM : BVH_CUDA_ECUDA;
}

int ctx_reserve(bvh_cuda_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->ws_bytes) return BVH_CUDA_OK;
    if (ctx->ws) {
        cudaFree(ctx->ws);
        ctx->ws = nullptr;
        ctx->ws_bytes = 0;
        ctx->d_last_order = nullptr;
        ctx->last_n = 0;
    }
    const size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&ctx->ws, want);
    if (e != cudaSuccess) {
        e = cudaMalloc(&ctx->ws, bytes);
        if (e != cudaSuccess) return ctx_cuda_fail(ctx, e, "cudaMalloc(workspace)");
        ctx->ws_bytes = bytes;
        return BVH_CUDA_OK;
    }
    ctx->ws_bytes = want;
    return BVH_CUDA_OK;
}

int ctx_stage_reserve(bvh_cuda_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->stage_bytes) return BVH_CUDA_OK;
    if (ctx->stage) { cudaFree(ctx->stage); ctx->stage = nullptr; ctx->stage_bytes = 0; }
    cudaError_t e = cudaMalloc(&ctx->stage, bytes);
    if (e != cudaSuccess) return ctx_cuda_fail(ctx, e, "cudaMalloc(staging)");
    ctx->stage_bytes = bytes;
    return BVH_CUDA_OK;
}

namespace {

// Carves the context's staging arena for one host-pointer call.
struct Stage {
    bvh_cuda_ctx* ctx;
    size_t sizes[8];
    int n = 0;
    explicit Stage(bvh_cuda_ctx* c) : ctx(c) {}
    int add(size_t bytes) { sizes[n] = (bytes + 255) & ~(size_t)255; return n++; }
    int commit() {
        size_t tot = 0;
        for (int i = 0; i < n; ++i) tot += sizes[i];
        return ctx_stage_reserve(ctx, tot ? tot : 256);
    }
    template <class T> T* ptr(int i) const {
        size_t off = 0;
        for (int k = 0; k < i; ++k) off += sizes[k];
        return reinterpret_cast<T*>((char*)ctx->stage + off);
    }
};

// RAII device staging buffer for the host-pointer entry points
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

}  // namespace

extern "C" {

int bvh_cuda_abi_version(void) { return 5; }  // 5: stream-ordered builds (blas_build_batch_async_dev / blas_build_finish); 2: BvhCudaBuildStats grew (ms_thread, thread_tasks); 3: again (grid_nodes, grid_interior_prims); 4: cluster tier (cluster_tasks, ms_cluster)

int bvh_cuda_create(int device, bvh_cuda_ctx** out) {
    if (!out) return BVH_CUDA_EINVAL;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        cudaGetLastError();
        return BVH_CUDA_ECUDA;  // no CPU fallback by design
    }
    bvh_cuda_ctx* ctx = new (std::nothrow) bvh_cuda_ctx();
    if (!ctx) return BVH_CUDA_ENOMEM;
    ctx->device = device;
    DeviceGuard g(device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return BVH_CUDA_ECUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost((void**)&ctx->h_pin, 4096) != cudaSuccess) {
        if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
        delete ctx;
        cudaGetLastError();
        return BVH_CUDA_ECUDA;
    }
    ctx->t2_blocks_per_sm = blas_t2_occupancy();
    if (blas_t2b_setup() < 1) { bvh_cuda_destroy(ctx); return BVH_CUDA_ECUDA; }
    ctx->t2w_blocks_per_sm = blas_t2w_occupancy();
    ctx->t1_blocks_per_sm = blas_t1_coop_occupancy();
    for (int cs : {16, 8}) {  // cluster tier: the largest cluster the device can place (16 is a non-portable size)
        const int nc = blas_tc_setup(cs);
        if (nc > 0) { ctx->tc_cluster_size = cs; ctx->tc_clusters = nc; break; }
    }
    for (auto& e : ctx->ev) cudaEventCreate(&e);
    cudaStreamCreateWithFlags(&ctx->own_stream2, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking);
    for (auto& row : ctx->pipe_ev) for (auto& e : row) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    *out = ctx;
    return BVH_CUDA_OK;
}

void bvh_cuda_destroy(bvh_cuda_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->stage) cudaFree(ctx->stage);
    if (ctx->trace_counter) cudaFree(ctx->trace_counter);
    if (ctx->defer_list) cudaFree(ctx->defer_list);
    if (ctx->defer_list2) cudaFree(ctx->defer_list2);
    if (ctx->own_stream2) cudaStreamDestroy(ctx->own_stream2);
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    for (auto& row : ctx->pipe_ev) for (auto& e : row) if (e) cudaEventDestroy(e);
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* bvh_cuda_last_error(const bvh_cuda_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

uint64_t bvh_cuda_launch_count(const bvh_cuda_ctx* ctx) { return ctx ? ctx->launches : 0; }

// Debug only (library built with -DBVH_T1_TIMING): per-phase-kind work / barrier-wait ns of block 0 in the grid tier.
int bvh_cuda_debug_t1_timing(unsigned long long* out32) { return blas_t1_timing(out32); }
int bvh_cuda_debug_t1_blocks(unsigned long long* out2048) { return blas_t1_blocks(out2048); }
int bvh_cuda_debug_t1_pull(unsigned long long* out4096) { return blas_t1_pull(out4096); }
// Debug only (-DBVH_TC_TIMING): per-node time stamps of the cluster tier; returns the number of rows (8 u64 each), -1 if not built in.
int bvh_cuda_debug_tc_log(unsigned long long* out, unsigned int cap_rows) { return blas_tc_log(out, cap_rows); }

int bvh_cuda_set_profiling(bvh_cuda_ctx* ctx, int enable) {
    if (!ctx) return BVH_CUDA_EINVAL;
    ctx->profiling = enable != 0;
    return BVH_CUDA_OK;
}

// ---- BLAS ------------------------------------------------------------------------------------------
int bvh_cuda_blas_build_dev(bvh_cuda_ctx* ctx, const float* d_vertices, size_t n_vertices, uint32_t* d_indices,
                            size_t n_tris, BvhNode* d_nodes_out, size_t nodes_cap, uint32_t* n_nodes_out, void* stream) {
    if (!ctx) return BVH_CUDA_EINVAL;
    DeviceGuard g(ctx->device);
    return blas_build_device(ctx, d_vertices, n_vertices, d_indices, n_tris, nullptr, 1, d_nodes_out, nodes_cap, n_nodes_out,
                             (cudaStream_t)stream);
}

int bvh_cuda_blas_build_batch_dev(bvh_cuda_ctx* ctx, const float* d_vertices, size_t n_vertices, uint32_t* d_indices,
                                  size_t n_indices, MeshInfo* d_mesh_info, size_t n_meshes, BvhNode* d_nodes_out,
                                  size_t nodes_cap, uint32_t* n_nodes_out, void* stream) {
    if (!ctx) return BVH_CUDA_EINVAL;
    if (!d_mesh_info || n_indices % 3 != 0) return ctx_fail(ctx, BVH_CUDA_EINVAL, "blas_build_batch: null mesh table or n_indices % 3 != 0");
    DeviceGuard g(ctx->device);
    return blas_build_device(ctx, d_vertices, n_vertices, d_indices, n_indices / 3, d_mesh_info, n_meshes, d_nodes_out, nodes_cap,
                             n_nodes_out, (cudaStream_t)stream);
}

int bvh_cuda_blas_build_batch_async_dev(bvh_cuda_ctx* ctx, const float* d_vertices, size_t n_vertices, uint32_t* d_indices,
                                        size_t n_indices, MeshInfo* d_mesh_info, size_t n_meshes, BvhNode* d_nodes_out,
                                        size_t nodes_cap, uint32_t* d_result, void* stream) {
    if (!ctx) return BVH_CUDA_EINVAL;
    if (n_indices % 3 != 0) return ctx_fail(ctx, BVH_CUDA_EINVAL, "blas_build_batch_async: n_indices % 3 != 0");
    DeviceGuard g(ctx->device);
    return blas_build_device(ctx, d_vertices, n_vertices, d_indices, n_indices / 3, d_mesh_info, d_mesh_info ? n_meshes : 1, d_nodes_out,
                             nodes_cap, nullptr, (cudaStream_t)stream, d_result, true);
}

int bvh_cuda_blas_build_finish(bvh_cuda_ctx* ctx, uint32_t* n_nodes_out) {
    if (!ctx) return BVH_CUDA_EINVAL;
    DeviceGuard g(ctx->device);
    return blas_build_finish(ctx, n_nodes_out);
}

int bvh_cuda_blas_build(bvh_cuda_ctx* ctx, const float* vertices, size_t n_vertices, uint32_t* indices, size_t n_tris,
                        BvhNode* nodes_out, size_t nodes_cap, uint32_t* n_nodes_out) {
    if (!ctx) return BVH_CUDA_EINVAL;
    if (!vertices || !indices || !nodes_out || n_tris == 0 || n_vertices == 0)
        return ctx_fail(ctx, BVH_CUDA_EINVAL, "blas_build: empty mesh or null pointer");
    if (nodes_cap < 2 * n_tris) return ctx_fail(ctx, BVH_CUDA_EINVAL, "blas_build: nodes_cap must be >= 2*n_tris");
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->own_stream;
    Stage st(ctx);
    const int iv = st.add(sizeof(float) * 3 * n_vertices), ii = st.add(sizeof(uint32_t) * 3 * n_tris),
              in = st.add(sizeof(BvhNode) * 2 * n_tris);
    int rc = st.commit();
    if (rc) return rc;
    float* dv = st.ptr<float>(iv);
    uint32_t* di = st.ptr<uint32_t>(ii);
    BvhNode* dn = st.ptr<BvhNode>(in);
    CU_CHECK(ctx, cudaMemcpyAsync(dv, vertices, sizeof(float) * 3 * n_vertices, cudaMemcpyHostToDevice, s));
    CU_CHECK(ctx, cudaMemcpyAsync(di, indices, sizeof(uint32_t) * 3 * n_tris, cudaMemcpyHostToDevice, s));
    uint32_t m = 0;
    rc = blas_build_device(ctx, dv, n_vertices, di, n_tris, nullptr, 1, dn, 2 * n_tris, &m, s);
    if (rc) { cudaStreamSynchronize(s); return rc; }
    CU_CHECK(ctx, cudaMemcpyAsync(nodes_out, dn, sizeof(BvhNode) * m, cudaMemcpyDeviceToHost, s));
    CU_CHECK(ctx, cudaMemcpyAsync(indices, di, sizeof(uint32_t) * 3 * n_tris, cudaMemcpyDeviceToHost, s));
    CU_CHECK(ctx, cudaStreamSynchronize(s));
    if (n_nodes_out) *n_nodes_out = m;
    return BVH_CUDA_OK;
}

int bvh_cuda_blas_last_order(bvh_cuda_ctx* ctx, uint32_t* order_out, size_t n_tris) {
    if (!ctx || !order_out) return BVH_CUDA_EINVAL;
    if (!ctx->d_last_order || ctx->last_n != n_tris) return ctx_fail(ctx, BVH_CUDA_EINVAL, "blas_last_order: no matching build");
    DeviceGuard g(ctx->device);
    CU_CHECK(ctx, cudaMemcpy(order_out, ctx->d_last_order, sizeof(uint32_t) * n_tris, cudaMemcpyDeviceToHost));
    return BVH_CUDA_OK;
}

int bvh_cuda_blas_last_stats(const bvh_cuda_ctx* ctx, BvhCudaBuildStats* out) {
    if (!ctx || !out) return BVH_CUDA_EINVAL;
    *out = ctx->stats;
    return BVH_CUDA_OK;
}

// ---- TLAS ------------------------------------------------------------------------------------------
int bvh_cuda_tlas_build_dev(bvh_cuda_ctx* ctx, const Instance* d_instances, size_t n_inst, const MeshInfo* d_meshes,
                            size_t n_mesh, TlasNode* d_nodes_out, uint32_t* d_children_out, void* stream) {
    if (!ctx) return BVH_CUDA_EINVAL;
    DeviceGuard g(ctx->device);
    return tlas_build_device(ctx, d_instances, n_inst, d_meshes, n_mesh, d_nodes_out, d_children_out, (cudaStream_t)stream);
}

int bvh_cuda_tlas_build(bvh_cuda_ctx* ctx, const Instance* instances, size_t n_inst, const MeshInfo* meshes, size_t n_mesh,
                        TlasNode* nodes_out, uint32_t* children_out) {
    if (!ctx) return BVH_CUDA_EINVAL;
    if (!instances || !meshes || !nodes_out || n_inst == 0 || n_mesh == 0)
        return ctx_fail(ctx, BVH_CUDA_EINVAL, "tlas_build: no instances / null pointer");
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->own_stream;
    const size_t total = 2 * n_inst + 1;
    DevBuf di, dm, dn, dc;
    CU_CHECK(ctx, di.alloc(sizeof(Instance) * n_inst));
    CU_CHECK(ctx, dm.alloc(sizeof(MeshInfo) * n_mesh));
    CU_CHECK(ctx, dn.alloc(sizeof(TlasNode) * total));
    CU_CHECK(ctx, dc.alloc(sizeof(uint32_t) * 2 * total));
    CU_CHECK(ctx, cudaMemcpyAsync(di.p, instances, sizeof(Instance) * n_inst, cudaMemcpyHostToDevice, s));
    CU_CHECK(ctx, cudaMemcpyAsync(dm.p, meshes, sizeof(MeshInfo) * n_mesh, cudaMemcpyHostToDevice, s));
    int rc = tlas_build_device(ctx, di.as<Instance>(), n_inst, dm.as<MeshInfo>(), n_mesh, dn.as<TlasNode>(), dc.as<uint32_t>(), s);
    if (rc) return rc;
    CU_CHECK(ctx, cudaMemcpyAsync(nodes_out, dn.p, sizeof(TlasNode) * total, cudaMemcpyDeviceToHost, s));
    if (children_out) CU_CHECK(ctx, cudaMemcpyAsync(children_out, dc.p, sizeof(uint32_t) * 2 * total, cudaMemcpyDeviceToHost, s));
    CU_CHECK(ctx, cudaStreamSynchronize(s));
    return BVH_CUDA_OK;
}

// ---- scene -----------------------------------------------------------------------------------------
static int scene_check(bvh_cuda_ctx* ctx, const BvhCudaSceneDesc* d) {
    if (!d || !d->tlas_nodes || !d->instances || !d->meshes || !d->bvh_nodes || !d->vertices || !d->indices ||
        d->n_tlas_nodes == 0 || d->n_instances == 0 || d->n_meshes == 0)
        return ctx_fail(ctx, BVH_CUDA_EINVAL, "scene: missing buffer");
    if (!d->tlas_children && d->n_instances > 32767)
        return ctx_fail(ctx, BVH_CUDA_EINVAL, "scene: more than 32767 instances need tlas_children (left_right packs 16+16 bits, tlas.rs:71)");
    return BVH_CUDA_OK;
}

int bvh_cuda_scene_upload(bvh_cuda_ctx* ctx, const BvhCudaSceneDesc* h, bvh_cuda_scene** out) {
    if (!ctx || !out) return BVH_CUDA_EINVAL;
    *out = nullptr;
    int rc = scene_check(ctx, h);
    if (rc) return rc;
    DeviceGuard g(ctx->device);
    const size_t sz[7] = {sizeof(TlasNode) * h->n_tlas_nodes,
                          h->tlas_children ? sizeof(uint32_t) * 2 * h->n_tlas_nodes : 0,
                          sizeof(Instance) * h->n_instances,
                          sizeof(MeshInfo) * h->n_meshes,
                          sizeof(BvhNode) * h->n_bvh_nodes,
                          sizeof(float) * 3 * h->n_vertices,
                          sizeof(uint32_t) * h->n_indices};
    const void* src[7] = {h->tlas_nodes, h->tlas_children, h->instances, h->meshes, h->bvh_nodes, h->vertices, h->indices};
    size_t off[7], total = 0;
    for (int i = 0; i < 7; ++i) { off[i] = total; total += (sz[i] + 255) & ~(size_t)255; }
    bvh_cuda_scene* sc = new (std::nothrow) bvh_cuda_scene();
    if (!sc) return BVH_CUDA_ENOMEM;
    cudaError_t e = cudaMalloc(&sc->block, total ? total : 256);
    if (e != cudaSuccess) { delete sc; return ctx_cuda_fail(ctx, e, "cudaMalloc(scene)"); }
    sc->owned = true;
    char* b = (char*)sc->block;
    for (int i = 0; i < 7; ++i)
        if (sz[i]) {
            e = cudaMemcpyAsync(b + off[i], src[i], sz[i], cudaMemcpyHostToDevice, ctx->own_stream);
            if (e != cudaSuccess) { cudaFree(sc->block); delete sc; return ctx_cuda_fail(ctx, e, "cudaMemcpy(scene)"); }
        }
    e = cudaStreamSynchronize(ctx->own_stream);
    if (e != cudaSuccess) { cudaFree(sc->block); delete sc; return ctx_cuda_fail(ctx, e, "cudaMemcpy(scene)"); }
    sc->d = *h;
    sc->d.tlas_nodes = (const TlasNode*)(b + off[0]);
    sc->d.tlas_children = h->tlas_children ? (const uint32_t*)(b + off[1]) : nullptr;
    sc->d.instances = (const Instance*)(b + off[2]);
    sc->d.meshes = (const MeshInfo*)(b + off[3]);
    sc->d.bvh_nodes = (const BvhNode*)(b + off[4]);
    sc->d.vertices = (const float*)(b + off[5]);
    sc->d.indices = (const uint32_t*)(b + off[6]);
    rc = scene_bake_device(ctx, sc, ctx->own_stream);
    // an uploaded scene owns immutable copies of its buffers, so the instance boxes can never go stale: always on
    if (rc == BVH_CUDA_OK) rc = scene_instance_boxes_device(ctx, sc, 1, ctx->own_stream);
    if (rc == BVH_CUDA_OK && cudaStreamSynchronize(ctx->own_stream) != cudaSuccess) rc = ctx_cuda_fail(ctx, cudaGetLastError(), "bake");
    if (rc) { bvh_cuda_scene_free(ctx, sc); return rc; }
    *out = sc;
    return BVH_CUDA_OK;
}

int bvh_cuda_scene_wrap_dev(bvh_cuda_ctx* ctx, const BvhCudaSceneDesc* d, void* stream, bvh_cuda_scene** out) {
    if (!ctx || !out) return BVH_CUDA_EINVAL;
    *out = nullptr;
    int rc = scene_check(ctx, d);
    if (rc) return rc;
    DeviceGuard g(ctx->device);
    bvh_cuda_scene* sc = new (std::nothrow) bvh_cuda_scene();
    if (!sc) return BVH_CUDA_ENOMEM;
    sc->d = *d;
    sc->owned = false;
    rc = scene_bake_device(ctx, sc, (cudaStream_t)stream);
    if (rc) { bvh_cuda_scene_free(ctx, sc); return rc; }
    *out = sc;
    return BVH_CUDA_OK;
}

int bvh_cuda_scene_refresh_dev(bvh_cuda_ctx* ctx, bvh_cuda_scene* scene, const BvhCudaSceneDesc* d, void* stream) {
    if (!ctx || !scene) return BVH_CUDA_EINVAL;
    if (scene->owned) return ctx_fail(ctx, BVH_CUDA_EINVAL, "scene_refresh: only wrapped scenes can be refreshed");
    DeviceGuard g(ctx->device);
    if (d) {
        int rc = scene_check(ctx, d);
        if (rc) return rc;
        if (d->n_indices != scene->d.n_indices) return ctx_fail(ctx, BVH_CUDA_EINVAL, "scene_refresh: index count changed");
        scene->d = *d;
    }
    int rc = scene_bake_device(ctx, scene, (cudaStream_t)stream);
    if (rc == BVH_CUDA_OK && scene->wbox_on) rc = scene_instance_boxes_device(ctx, scene, 1, (cudaStream_t)stream);
    return rc;
}

int bvh_cuda_scene_instance_boxes_dev(bvh_cuda_ctx* ctx, bvh_cuda_scene* scene, int enable, void* stream) {
    if (!ctx || !scene) return BVH_CUDA_EINVAL;
    DeviceGuard g(ctx->device);
    return scene_instance_boxes_device(ctx, scene, enable, (cudaStream_t)stream);
}

void bvh_cuda_scene_free(bvh_cuda_ctx* ctx, bvh_cuda_scene* scene) {
    if (!scene) return;
    auto release = [&] {
        if (scene->owned && scene->block) cudaFree(scene->block);
        if (scene->baked) cudaFree(scene->baked);
        if (scene->wbox) cudaFree(scene->wbox);
    };
    if (ctx) { DeviceGuard g(ctx->device); release(); }
    else release();
    delete scene;
}

// ---- traversal -------------------------------------------------------------------------------------
int bvh_cuda_trace_blas_dev(bvh_cuda_ctx* ctx, const BvhNode* d_nodes, const float* d_vertices, const uint32_t* d_indices,
                            const float* d_ray_o, const float* d_ray_d, size_t n_rays, float* d_t_out, uint32_t* d_tri_out,
                            void* stream) {
    if (!ctx) return BVH_CUDA_EINVAL;
    DeviceGuard g(ctx->device);
    return trace_blas_device(ctx, d_nodes, d_vertices, d_indices, d_ray_o, d_ray_d, n_rays, d_t_out, d_tri_out, (cudaStream_t)stream);
}

int bvh_cuda_trace_blas(bvh_cuda_ctx* ctx, const BvhNode* nodes, size_t n_nodes, const float* vertices, size_t n_vertices,
                        const uint32_t* indices, size_t n_tris, const float* ray_o, const float* ray_d, size_t n_rays,
                        float* t_out, uint32_t* tri_out) {
    if (!ctx) return BVH_CUDA_EINVAL;
    if (!nodes || !vertices || !indices || !ray_o || !ray_d || !t_out || !tri_out || n_nodes == 0)
        return ctx_fail(ctx, BVH_CUDA_EINVAL, "trace_blas: null pointer");
    if (n_rays == 0) return BVH_CUDA_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->own_stream;
    DevBuf dn, dv, di, dro, drd, dt, dtri;
    CU_CHECK(ctx, dn.alloc(sizeof(BvhNode) * n_nodes));
    CU_CHECK(ctx, dv.alloc(sizeof(float) * 3 * n_vertices));
    CU_CHECK(ctx, di.alloc(sizeof(uint32_t) * 3 * n_tris));
    CU_CHECK(ctx, dro.alloc(sizeof(float) * 3 * n_rays));
    CU_CHECK(ctx, drd.alloc(sizeof(float) * 3 * n_rays));
    CU_CHECK(ctx, dt.alloc(sizeof(float) * n_rays));
    CU_CHECK(ctx, dtri.alloc(sizeof(uint32_t) * n_rays));
    CU_CHECK(ctx, cudaMemcpyAsync(dn.p, nodes, sizeof(BvhNode) * n_nodes, cudaMemcpyHostToDevice, s));
    CU_CHECK(ctx, cudaMemcpyAsync(dv.p, vertices, sizeof(float) * 3 * n_vertices, cudaMemcpyHostToDevice, s));
    CU_CHECK(ctx, cudaMemcpyAsync(di.p, indices, sizeof(uint32_t) * 3 * n_tris, cudaMemcpyHostToDevice, s));
    CU_CHECK(ctx, cudaMemcpyAsync(dro.p, ray_o, sizeof(float) * 3 * n_rays, cudaMemcpyHostToDevice, s));
    CU_CHECK(ctx, cudaMemcpyAsync(drd.p, ray_d, sizeof(float) * 3 * n_rays, cudaMemcpyHostToDevice, s));
    int rc = trace_blas_device(ctx, dn.as<BvhNode>(), dv.as<float>(), di.as<uint32_t>(), dro.as<float>(), drd.as<float>(), n_rays,
                               dt.as<float>(), dtri.as<uint32_t>(), s);
    if (rc) return rc;
    CU_CHECK(ctx, cudaMemcpyAsync(t_out, dt.p, sizeof(float) * n_rays, cudaMemcpyDeviceToHost, s));
    CU_CHECK(ctx, cudaMemcpyAsync(tri_out, dtri.p, sizeof(uint32_t) * n_rays, cudaMemcpyDeviceToHost, s));
    CU_CHECK(ctx, cudaStreamSynchronize(s));
    return BVH_CUDA_OK;
}

int bvh_cuda_trace_blas_recursive_dev(bvh_cuda_ctx* ctx, const BvhNode* d_nodes, const float* d_vertices, const uint32_t* d_indices,
                                      const float* d_ray_o, const float* d_ray_d, size_t n_rays, uint32_t node_idx, float t0,
                                      float* d_t_out, uint8_t* d_hit_out, void* stream) {
    if (!ctx) return BVH_CUDA_EINVAL;
    DeviceGuard g(ctx->device);
    return trace_blas_rec_device(ctx, d_nodes, d_vertices, d_indices, d_ray_o, d_ray_d, n_rays, node_idx, t0, d_t_out, d_hit_out,
                                 (cudaStream_t)stream);
}

int bvh_cuda_trace_blas_recursive(bvh_cuda_ctx* ctx, const BvhNode* nodes, size_t n_nodes, const float* vertices, size_t n_vertices,
                                  const uint32_t* indices, size_t n_tris, const float* ray_o, const float* ray_d, size_t n_rays,
                                  uint32_t node_idx, float t0, float* t_out, uint8_t* hit_out) {
    if (!ctx) return BVH_CUDA_EINVAL;
    if (!nodes || !vertices || !indices || !ray_o || !ray_d || !t_out || !hit_out || n_nodes == 0 || node_idx >= n_nodes)
        return ctx_fail(ctx, BVH_CUDA_EINVAL, "trace_blas_recursive: null pointer or node_idx out of range");
    if (n_rays == 0) return BVH_CUDA_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->own_stream;
    Stage st(ctx);
    const int in = st.add(sizeof(BvhNode) * n_nodes), iv = st.add(sizeof(float) * 3 * n_vertices), ii = st.add(sizeof(uint32_t) * 3 * n_tris),
              io = st.add(sizeof(float) * 3 * n_rays), id = st.add(sizeof(float) * 3 * n_rays), it = st.add(sizeof(float) * n_rays),
              ih = st.add(n_rays);
    int rc = st.commit();
    if (rc) return rc;
    CU_CHECK(ctx, cudaMemcpyAsync(st.ptr<void>(in), nodes, sizeof(BvhNode) * n_nodes, cudaMemcpyHostToDevice, s));
    CU_CHECK(ctx, cudaMemcpyAsync(st.ptr<void>(iv), vertices, sizeof(float) * 3 * n_vertices, cudaMemcpyHostToDevice, s));
    CU_CHECK(ctx, cudaMemcpyAsync(st.ptr<void>(ii), indices, sizeof(uint32_t) * 3 * n_tris, cudaMemcpyHostToDevice, s));
    CU_CHECK(ctx, cudaMemcpyAsync(st.ptr<void>(io), ray_o, sizeof(float) * 3 * n_rays, cudaMemcpyHostToDevice, s));
    CU_CHECK(ctx, cudaMemcpyAsync(st.ptr<void>(id), ray_d, sizeof(float) * 3 * n_rays, cudaMemcpyHostToDevice, s));
    rc = trace_blas_rec_device(ctx, st.ptr<BvhNode>(in), st.ptr<float>(iv), st.ptr<uint32_t>(ii), st.ptr<float>(io), st.ptr<float>(id),
                               n_rays, node_idx, t0, st.ptr<float>(it), st.ptr<uint8_t>(ih), s);
    if (rc) return rc;
    CU_CHECK(ctx, cudaMemcpyAsync(t_out, st.ptr<void>(it), sizeof(float) * n_rays, cudaMemcpyDeviceToHost, s));
    CU_CHECK(ctx, cudaMemcpyAsync(hit_out, st.ptr<void>(ih), n_rays, cudaMemcpyDeviceToHost, s));
    CU_CHECK(ctx, cudaStreamSynchronize(s));
    return BVH_CUDA_OK;
}

int bvh_cuda_trace_closest_dev(bvh_cuda_ctx* ctx, const bvh_cuda_scene* scene, const float* d_ray_o, const float* d_ray_d,
                               size_t n_rays, float tmax, float* d_t_out, uint32_t* d_tri_out, uint32_t* d_inst_out, void* stream) {
    if (!ctx) return BVH_CUDA_EINVAL;
    DeviceGuard g(ctx->device);
    return trace_scene_device(ctx, scene, d_ray_o, d_ray_d, n_rays, tmax, 0, d_t_out, d_tri_out, d_inst_out, nullptr, (cudaStream_t)stream);
}

int bvh_cuda_trace_any_dev(bvh_cuda_ctx* ctx, const bvh_cuda_scene* scene, const float* d_ray_o, const float* d_ray_d,
                           size_t n_rays, float tmax, uint8_t* d_occluded_out, void* stream) {
    if (!ctx) return BVH_CUDA_EINVAL;
    DeviceGuard g(ctx->device);
    return trace_scene_device(ctx, scene, d_ray_o, d_ray_d, n_rays, tmax, 1, nullptr, nullptr, nullptr, d_occluded_out, (cudaStream_t)stream);
}

// Host-pointer traversal: the rays are cut into up to 16 chunks; chunk k+1 is uploaded (h2d stream) and chunk k-1 read
// back (d2h stream) while chunk k is traced (own stream), so a call costs about max(PCIe in, kernels, PCIe out) instead
// of their sum.  Needs pinned caller buffers to overlap; with pageable memory the copies simply serialise.
static int trace_host_pipelined(bvh_cuda_ctx* ctx, const bvh_cuda_scene* scene, const float* ray_o, const float* ray_d, size_t n_rays,
                                float tmax, int any_hit, float* t_out, uint32_t* tri_out, uint32_t* inst_out, uint8_t* occ_out) {
    DeviceGuard g(ctx->device);
    Stage st(ctx);
    const int io = st.add(sizeof(float) * 3 * n_rays), id = st.add(sizeof(float) * 3 * n_rays);
    const int it = st.add(any_hit ? n_rays : sizeof(float) * n_rays), itr = st.add(any_hit ? 0 : sizeof(uint32_t) * n_rays),
              iin = st.add(any_hit ? 0 : sizeof(uint32_t) * n_rays);
    int rc = st.commit();
    if (rc) return rc;
    float* dro = st.ptr<float>(io);
    float* drd = st.ptr<float>(id);
    float* dt = st.ptr<float>(it);
    uint8_t* docc = st.ptr<uint8_t>(it);
    uint32_t* dtri = st.ptr<uint32_t>(itr);
    uint32_t* dinst = st.ptr<uint32_t>(iin);
    const bool pipe = ctx->h2d_stream && ctx->d2h_stream && ctx->pipe_ev[1][15];
    cudaStream_t s = ctx->own_stream, sin = pipe ? ctx->h2d_stream : s, sout = pipe ? ctx->d2h_stream : s;
    // chunks alternate between two compute streams (and the scene's two control slots): the tail of one chunk's persistent
    // kernel -- a few long rays, ~0.3 ms -- overlaps the start of the next chunk instead of idling the GPU.
    static const bool two = [] { const char* e = getenv("BVH_CUDA_TRACE_STREAMS"); return !(e && atoi(e) == 1); }();
    cudaStream_t cs[2] = {s, (pipe && two && ctx->own_stream2) ? ctx->own_stream2 : s};
    static const size_t n_chunks = [] { const char* e = getenv("BVH_CUDA_TRACE_CHUNKS"); int v = e ? atoi(e) : 8; return (size_t)(v < 1 ? 1 : (v > 16 ? 16 : v)); }();
    size_t chunk = (n_rays + n_chunks - 1) / n_chunks;
    if (chunk < ((size_t)1 << 19)) chunk = (size_t)1 << 19;
    int k = 0;
    for (size_t b0 = 0; b0 < n_rays; b0 += chunk, ++k) {
        const size_t m = (n_rays - b0 < chunk) ? n_rays - b0 : chunk;
        const int slot = (cs[1] != cs[0]) ? (k & 1) : 0;
        cudaStream_t sc = cs[slot];
        CU_CHECK(ctx, cudaMemcpyAsync(dro + 3 * b0, ray_o + 3 * b0, sizeof(float) * 3 * m, cudaMemcpyHostToDevice, sin));
        CU_CHECK(ctx, cudaMemcpyAsync(drd + 3 * b0, ray_d + 3 * b0, sizeof(float) * 3 * m, cudaMemcpyHostToDevice, sin));
        if (pipe) {
            CU_CHECK(ctx, cudaEventRecord(ctx->pipe_ev[0][k], sin));
            CU_CHECK(ctx, cudaStreamWaitEvent(sc, ctx->pipe_ev[0][k], 0));
        }
        rc = any_hit ? trace_scene_device(ctx, scene, dro + 3 * b0, drd + 3 * b0, m, tmax, 1, nullptr, nullptr, nullptr, docc + b0, sc, slot)
                     : trace_scene_device(ctx, scene, dro + 3 * b0, drd + 3 * b0, m, tmax, 0, dt + b0, dtri + b0, dinst + b0, nullptr, sc, slot);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        if (pipe) {
            CU_CHECK(ctx, cudaEventRecord(ctx->pipe_ev[1][k], sc));
            CU_CHECK(ctx, cudaStreamWaitEvent(sout, ctx->pipe_ev[1][k], 0));
        }
        if (any_hit) {
            CU_CHECK(ctx, cudaMemcpyAsync(occ_out + b0, docc + b0, m, cudaMemcpyDeviceToHost, sout));
        } else {
            CU_CHECK(ctx, cudaMemcpyAsync(t_out + b0, dt + b0, sizeof(float) * m, cudaMemcpyDeviceToHost, sout));
            CU_CHECK(ctx, cudaMemcpyAsync(tri_out + b0, dtri + b0, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost, sout));
            CU_CHECK(ctx, cudaMemcpyAsync(inst_out + b0, dinst + b0, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost, sout));
        }
    }
    if (pipe) {
        CU_CHECK(ctx, cudaStreamSynchronize(sin));
        CU_CHECK(ctx, cudaStreamSynchronize(sout));
    }
    CU_CHECK(ctx, cudaStreamSynchronize(s));
    if (cs[1] != s) CU_CHECK(ctx, cudaStreamSynchronize(cs[1]));
    return BVH_CUDA_OK;
}

int bvh_cuda_trace_closest(bvh_cuda_ctx* ctx, const bvh_cuda_scene* scene, const float* ray_o, const float* ray_d, size_t n_rays,
                           float tmax, float* t_out, uint32_t* tri_out, uint32_t* inst_out) {
    if (!ctx) return BVH_CUDA_EINVAL;
    if (!scene || !ray_o || !ray_d || !t_out || !tri_out || !inst_out) return ctx_fail(ctx, BVH_CUDA_EINVAL, "trace_closest: null pointer");
    if (n_rays == 0) return BVH_CUDA_OK;
    return trace_host_pipelined(ctx, scene, ray_o, ray_d, n_rays, tmax, 0, t_out, tri_out, inst_out, nullptr);
}

int bvh_cuda_trace_any(bvh_cuda_ctx* ctx, const bvh_cuda_scene* scene, const float* ray_o, const float* ray_d, size_t n_rays,
                       float tmax, uint8_t* occluded_out) {
    if (!ctx) return BVH_CUDA_EINVAL;
    if (!scene || !ray_o || !ray_d || !occluded_out) return ctx_fail(ctx, BVH_CUDA_EINVAL, "trace_any: null pointer");
    if (n_rays == 0) return BVH_CUDA_OK;
    return trace_host_pipelined(ctx, scene, ray_o, ray_d, n_rays, tmax, 1, nullptr, nullptr, nullptr, occluded_out);
}

}  // extern "C"
