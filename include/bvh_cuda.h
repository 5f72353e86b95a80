/* bvh_cuda.h — C ABI of libbvh_cuda.so: the sm_100a CUDA replacement for voidin's `crates/bvh`.
 *
 * Every entry point names the reference interface it replaces (paths relative to the voidin checkout).
 * All structs are byte-identical to the Rust `repr(C)` / WGSL ones, so the arrays this library writes can be
 * pushed unchanged into voidin's wgpu storage buffers (crates/pools/src/mesh/mod.rs:322-330,285).
 *
 * Conventions
 *   - Plain pointers and sizes only.  "host" entry points take host pointers and do their own H2D/D2H copies;
 *     the `_dev` twins take device pointers plus a `cudaStream_t` (passed as void*) and leave results on the
 *     device.  Work of one context is serialised on the stream it is given.  A context and a scene each own
 *     mutable device scratch (build workspace, persistent-warp ray counter, deferral list): do NOT issue calls on
 *     the same context, or traces of the same scene, concurrently from several streams or threads — use one
 *     context (and one scene wrap) per stream.  Different contexts are independent.
 *   - Every function returns a status: 0 ok, <0 error.  Nothing unwinds, nothing prints.  The reference has
 *     no error channel — it panics (blas.rs:84 on N=0, mesh/mod.rs:321 on len%3!=0, OOB on bad indices) or
 *     never terminates on degenerate input (blas.rs:115,139); those cases map to EINVAL / EDEGENERATE here.
 *   - There is no CPU fallback: without a CUDA device bvh_cuda_create fails with BVH_CUDA_ECUDA.
 */
#ifndef BVH_CUDA_H
#define BVH_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BVH_CUDA_OK 0
#define BVH_CUDA_EINVAL (-1)      /* empty mesh, null pointer, index out of range, capacity too small */
#define BVH_CUDA_EDEGENERATE (-2) /* a node with >3 triangles has no candidate split with finite cost */
#define BVH_CUDA_ECUDA (-3)       /* CUDA runtime error; see bvh_cuda_last_error */
#define BVH_CUDA_ENOMEM (-4)      /* device allocation failed */

#define BVH_CUDA_MAX_DIST 1e30f /* crates/bvh/src/intersection.rs:3, shaders/utils/math.wgsl:4 */
#define BVH_CUDA_NO_HIT 0xFFFFFFFFu

/* crates/bvh/src/blas.rs:10-17; WGSL mirror shaders/utils/bvh.wgsl:11-16.
 * Leaf <=> count > 0: left_first = first triangle (permuted order), count in {1,2,3}.
 * Interior: count == 0, children at left_first and left_first+1 (relative to the mesh's own node array). */
typedef struct BvhNode {
    float min[3];
    uint32_t left_first;
    float max[3];
    uint32_t count;
} BvhNode;

/* crates/bvh/src/tlas.rs:7-14; WGSL mirror shaders/utils/bvh.wgsl:4-9.
 * Leaf <=> left_right == 0; interior left_right = a + (b << 16) in wrapping u32 (tlas.rs:71). */
typedef struct TlasNode {
    float min[3];
    uint32_t left_right;
    float max[3];
    uint32_t instance_idx;
} TlasNode;

/* crates/components/src/shared.rs:67-75 (glam Mat4 is column-major); WGSL shaders/shared.wgsl:53-59. */
typedef struct Instance {
    float transform[16];
    float inv_transform[16];
    uint32_t mesh;
    uint32_t material;
    uint32_t junk[2];
} Instance;

/* crates/components/src/shared.rs:29-39; WGSL shaders/shared.wgsl:43-51. */
typedef struct MeshInfo {
    float min[3];
    uint32_t index_count;
    float max[3];
    uint32_t base_index;
    int32_t vertex_offset;
    uint32_t bvh_index;
    uint32_t junk[2];
} MeshInfo;

typedef struct bvh_cuda_ctx bvh_cuda_ctx;     /* one per host thread and device */
typedef struct bvh_cuda_scene bvh_cuda_scene; /* the six buffers of voidin's trace bind group, on the device */

/* Counters of the most recent BLAS build on a context (all derived on the device). */
typedef struct BvhCudaBuildStats {
    uint64_t sum_interior_prims; /* S: sum over interior nodes of their triangle count */
    uint32_t n_nodes;            /* M = 2 + 2 * interior nodes */
    uint32_t interior_nodes;
    uint32_t grid_levels;        /* levels handled by the grid-wide tier (nodes > 24576 triangles) */
    uint32_t big_block_tasks;    /* nodes handled by one 1024-thread block each (2049..24576) */
    uint32_t block_tasks;        /* nodes handled one block each from the device task queue (257..2048) */
    uint32_t warp_node_tasks;    /* nodes handled one warp each from the second task queue (33..256) */
    uint32_t warp_tasks;         /* sub-trees (<= 32 triangles) handled one warp each; their small children go to the thread tier */
    uint32_t kernel_launches;    /* kernels launched by this build */
    /* Device time per phase in ms (CUDA events on the build's stream); all zero unless profiling is enabled. */
    float ms_setup;              /* k_setup: centroids, triangle boxes */
    float ms_grid;               /* grid-wide tier, all levels */
    float ms_big_block;          /* k_t2<24576,1024>: big-block task-queue kernel (one launch) */
    float ms_block;              /* k_t2<2048,256>: block-per-node task-queue kernel (one launch) */
    float ms_warp_node;          /* k_t2w: warp-per-node task-queue kernel (one launch) */
    float ms_warp;               /* k_t3: warp-per-sub-tree kernel (one launch) */
    float ms_emit;               /* numbering scan + node emit + index permutation */
    float ms_total;
    float ms_thread;             /* k_t4: thread-per-sub-tree kernel (one launch) */
    uint32_t thread_tasks;       /* small sub-trees handled one thread each (k_t4) */
    uint32_t grid_nodes;         /* interior nodes split by the grid-wide and cluster tiers (n > 24576) */
    uint32_t cluster_tasks;      /* nodes handled one thread-block cluster each (16385..262144); also counted in grid_nodes */
    uint64_t grid_interior_prims;/* sum of their triangle counts (the S of the large-node tiers' algorithmic bytes) */
    float ms_cluster;            /* k_tc: cluster-per-node task-queue kernel (one launch) */
    uint32_t reserved0;
} BvhCudaBuildStats;

/* ---- context ------------------------------------------------------------------------------------------- */
int bvh_cuda_create(int device, bvh_cuda_ctx** out);
void bvh_cuda_destroy(bvh_cuda_ctx* ctx);
/* Message of the last failing call on this context ("" if none).  Valid until the next call on the context. */
const char* bvh_cuda_last_error(const bvh_cuda_ctx* ctx);
/* Total kernels launched through this context since creation. */
uint64_t bvh_cuda_launch_count(const bvh_cuda_ctx* ctx);
int bvh_cuda_abi_version(void); /* 5 */
/* enable != 0: later BLAS builds record per-phase CUDA-event timings into BvhCudaBuildStats (a few extra event
 * records per build, no extra synchronisation). */
int bvh_cuda_set_profiling(bvh_cuda_ctx* ctx, int enable);

/* ---- BLAS: BvhBuilder::new(vertices, indices).build()  (crates/bvh/src/blas.rs:51-103) ----------------- *
 * Caller: MeshPool::add, crates/pools/src/mesh/mod.rs:320-321.
 * vertices: 3*n_vertices floats (glam Vec3, 12 B).  indices: 3*n_tris u32, permuted IN PLACE into the builder's
 * final triangle order (blas.rs:95-100).  nodes_out: capacity nodes_cap >= 2*n_tris entries (blas.rs:52);
 * *n_nodes_out = 2 + 2*interior (blas.rs:93), node 1 is all-zero (blas.rs:90).
 * `set_bin_number` (blas.rs:64-67) has no counterpart: the reference never reads the value (blas.rs:136). */
int bvh_cuda_blas_build(bvh_cuda_ctx* ctx, const float* vertices, size_t n_vertices, uint32_t* indices,
                        size_t n_tris, BvhNode* nodes_out, size_t nodes_cap, uint32_t* n_nodes_out);
/* Same, device pointers.  n_nodes_out is a HOST pointer; the call returns after the node count is known
 * (the build is a sequence of dependent launches with a few small read-backs). */
int bvh_cuda_blas_build_dev(bvh_cuda_ctx* ctx, const float* d_vertices, size_t n_vertices, uint32_t* d_indices,
                            size_t n_tris, BvhNode* d_nodes_out, size_t nodes_cap, uint32_t* n_nodes_out,
                            void* stream);
/* MeshPool::add for a whole pooled scene in one call (crates/pools/src/mesh/mod.rs:309-351): every mesh's BLAS is built
 * in the same passes (a forest build: one root per mesh), which is far cheaper than n_meshes separate builds.
 * d_vertices / d_indices: pooled buffers; indices are mesh-local vertex ids (the shader adds vertex_offset,
 * shaders/utils/bvh.wgsl:31) and are permuted in place, mesh by mesh.  d_mesh_info[m] (device, in/out): the caller fills
 * vertex_offset, base_index, index_count (meshes back to back in order: base_index[0]=0, base_index[m+1] =
 * base_index[m]+index_count[m]) and min/max; the call fills bvh_index.  d_nodes_out: pooled nodes (mesh m at
 * bvh_index[m], numbered from 0 inside the mesh), capacity >= 2*n_indices/3; *n_nodes_out = total nodes (host). */
int bvh_cuda_blas_build_batch_dev(bvh_cuda_ctx* ctx, const float* d_vertices, size_t n_vertices, uint32_t* d_indices,
                                  size_t n_indices, MeshInfo* d_mesh_info, size_t n_meshes, BvhNode* d_nodes_out,
                                  size_t nodes_cap, uint32_t* n_nodes_out, void* stream);
/* The same build, STREAM-ORDERED: every kernel of the build is enqueued on `stream` and the call returns without waiting, so
 * the caller can queue the TLAS build and the first trace behind it, or run a second build on ANOTHER context and stream at
 * the same time (a context owns one workspace: one build in flight per context).  d_mesh_info may be NULL for a single
 * mesh.  d_result (device, 4 words, may be NULL) receives {total nodes, device status bits (0 = ok), interior nodes, 0} when
 * the build completes on the stream.  bvh_cuda_blas_build_finish waits for that build, returns its status (EINVAL /
 * EDEGENERATE as for the synchronous calls), fills *n_nodes_out (host, may be NULL) and the statistics; it must be called
 * before the next build on the same context.  Replaces nothing in the reference (its build is a blocking CPU call,
 * crates/pools/src/mesh/mod.rs:320-321); this is what lets a renderer keep the frame's stream busy. */
int bvh_cuda_blas_build_batch_async_dev(bvh_cuda_ctx* ctx, const float* d_vertices, size_t n_vertices, uint32_t* d_indices,
                                        size_t n_indices, MeshInfo* d_mesh_info, size_t n_meshes, BvhNode* d_nodes_out,
                                        size_t nodes_cap, uint32_t* d_result, void* stream);
int bvh_cuda_blas_build_finish(bvh_cuda_ctx* ctx, uint32_t* n_nodes_out);
/* Optional: final triangle_indices of the last build (original triangle id per slot), n_tris entries, device->host. */
int bvh_cuda_blas_last_order(bvh_cuda_ctx* ctx, uint32_t* order_out, size_t n_tris);
int bvh_cuda_blas_last_stats(const bvh_cuda_ctx* ctx, BvhCudaBuildStats* out);

/* ---- TLAS: Tlas::build(&mut self, instances, meshes)  (crates/bvh/src/tlas.rs:31-85) -------------------- *
 * Caller: MeshPool::generate_tlas, crates/pools/src/mesh/mod.rs:279-286 (which guards n_inst == 0).
 * nodes_out: 2*n_inst+1 entries (tlas.rs:32).  children_out (may be NULL): 2*(2*n_inst+1) u32, the unpacked
 * child ids of every node — required by the traversal when n_inst > 32767 because left_right packs 16+16 bits. */
int bvh_cuda_tlas_build(bvh_cuda_ctx* ctx, const Instance* instances, size_t n_inst, const MeshInfo* meshes,
                        size_t n_mesh, TlasNode* nodes_out, uint32_t* children_out);
int bvh_cuda_tlas_build_dev(bvh_cuda_ctx* ctx, const Instance* d_instances, size_t n_inst,
                            const MeshInfo* d_meshes, size_t n_mesh, TlasNode* d_nodes_out,
                            uint32_t* d_children_out, void* stream);

/* ---- scene: the trace bind group (crates/pools/src/mesh/mod.rs:136-238, crates/app/src/app.rs:255-287) -- */
typedef struct BvhCudaSceneDesc {
    const TlasNode* tlas_nodes;    size_t n_tlas_nodes;
    const uint32_t* tlas_children; /* optional, 2*n_tlas_nodes */
    const Instance* instances;     size_t n_instances;
    const MeshInfo* meshes;        size_t n_meshes;
    const BvhNode* bvh_nodes;      size_t n_bvh_nodes;
    const float* vertices;         size_t n_vertices;  /* pooled, 3 floats each */
    const uint32_t* indices;       size_t n_indices;   /* pooled, permuted */
} BvhCudaSceneDesc;
/* Copies host buffers to the device. */
int bvh_cuda_scene_upload(bvh_cuda_ctx* ctx, const BvhCudaSceneDesc* host_desc, bvh_cuda_scene** out);
/* Wraps device buffers that stay owned by the caller (no copy of the six buffers).  Both scene constructors also
 * "bake" an internal, traversal-friendly copy of the triangles (3 x float4 per pooled triangle, gathered through
 * indices + vertex_offset exactly as fetch_vertex does, shaders/utils/bvh.wgsl:30-33) on `stream`. */
int bvh_cuda_scene_wrap_dev(bvh_cuda_ctx* ctx, const BvhCudaSceneDesc* dev_desc, void* stream, bvh_cuda_scene** out);
/* Re-bakes a wrapped scene after its buffers were rewritten in place (e.g. a BLAS rebuilt into the same node /
 * index buffers); dev_desc may be NULL to keep the pointers, or carry new pointers/counts with the same n_indices. */
int bvh_cuda_scene_refresh_dev(bvh_cuda_ctx* ctx, bvh_cuda_scene* scene, const BvhCudaSceneDesc* dev_desc, void* stream);
/* Instance culling for the exact-order kernels (bvh_cuda_trace_closest*, and the any-hit rays that need the reference's
 * order).  Tlas::build seeds every leaf box with the UNtransformed local box (tlas.rs:39), so instances far from the origin get
 * leaf boxes stretched back to it and a ray "enters" many instances only to miss the children of their BLAS root
 * (instance_intersect, bvh.wgsl:78-87).  With enable != 0 the library computes, from the scene's CURRENT instance / mesh / node
 * buffers, a tight world box per instance (BLAS root box mapped through the inverse of inv_transform, grown well beyond the
 * rounding error of the object-space tests) and drops visits whose box the ray misses -- results are identical, the
 * TlasNode bytes are untouched.  Uploaded scenes (immutable copies) have it on from the start.  For wrapped scenes it is
 * off until this call, and the boxes go STALE when the caller rewrites the instance buffer: call again (or
 * bvh_cuda_scene_refresh_dev, which recomputes them when enabled) after every such change.  enable = 0 switches it off. */
int bvh_cuda_scene_instance_boxes_dev(bvh_cuda_ctx* ctx, bvh_cuda_scene* scene, int enable, void* stream);
void bvh_cuda_scene_free(bvh_cuda_ctx* ctx, bvh_cuda_scene* scene);

/* ---- traversal ----------------------------------------------------------------------------------------- *
 * Rays: ray_o / ray_d are 3*n_rays floats each.  Ids are an extension of the reference, which returns none:
 * tri = node.left_first + i in the mesh's permuted order, inst = TlasNode.instance_idx, both captured at the
 * assignment that lowers t, in the reference's visit order.  Miss: t = 1e30, ids = 0xFFFFFFFF.            */

/* Bvh::traverse_iter (crates/bvh/src/blas.rs:247-295) with the Rust intersection tests
 * (crates/bvh/src/intersection.rs:47-55,68-92): division slabs, two-sided triangles, EPS 1e-4, far child first.
 * Caller: src/bin/bvh_cpu.rs:87. */
int bvh_cuda_trace_blas(bvh_cuda_ctx* ctx, const BvhNode* nodes, size_t n_nodes, const float* vertices,
                        size_t n_vertices, const uint32_t* indices, size_t n_tris, const float* ray_o,
                        const float* ray_d, size_t n_rays, float* t_out, uint32_t* tri_out);
int bvh_cuda_trace_blas_dev(bvh_cuda_ctx* ctx, const BvhNode* d_nodes, const float* d_vertices,
                            const uint32_t* d_indices, const float* d_ray_o, const float* d_ray_d,
                            size_t n_rays, float* d_t_out, uint32_t* d_tri_out, void* stream);

/* Bvh::traverse (crates/bvh/src/blas.rs:211-245), the recursive variant (unused by the reference: commented out at
 * src/bin/bvh_cpu.rs:86): starts at node_idx with distance bound t0; hit_out[r] = 1 means Hit(t_out[r]) — which the
 * reference returns as soon as the start node's box is hit, with t_out = t0 when no triangle is closer — 0 means Miss.
 * The reference takes Vec4 / UVec4 slices and truncates them; pass the same Vec3 / UVec3 data here. */
int bvh_cuda_trace_blas_recursive(bvh_cuda_ctx* ctx, const BvhNode* nodes, size_t n_nodes, const float* vertices,
                                  size_t n_vertices, const uint32_t* indices, size_t n_tris, const float* ray_o,
                                  const float* ray_d, size_t n_rays, uint32_t node_idx, float t0, float* t_out,
                                  uint8_t* hit_out);
int bvh_cuda_trace_blas_recursive_dev(bvh_cuda_ctx* ctx, const BvhNode* d_nodes, const float* d_vertices,
                                      const uint32_t* d_indices, const float* d_ray_o, const float* d_ray_d,
                                      size_t n_rays, uint32_t node_idx, float t0, float* d_t_out, uint8_t* d_hit_out,
                                      void* stream);

/* traverse_tlas (shaders/utils/bvh.wgsl:89-123) with the WGSL tests (shaders/utils/intersections.wgsl:13-45):
 * reciprocal slabs, back-face culling, 0 < t < hit, near child first.  tmax: initial res.dist (reference: 1e30). */
int bvh_cuda_trace_closest(bvh_cuda_ctx* ctx, const bvh_cuda_scene* scene, const float* ray_o,
                           const float* ray_d, size_t n_rays, float tmax, float* t_out, uint32_t* tri_out,
                           uint32_t* inst_out);
int bvh_cuda_trace_closest_dev(bvh_cuda_ctx* ctx, const bvh_cuda_scene* scene, const float* d_ray_o,
                               const float* d_ray_d, size_t n_rays, float tmax, float* d_t_out,
                               uint32_t* d_tri_out, uint32_t* d_inst_out, void* stream);

/* Shadow rays: occluded[r] = traverse_tlas(ray).hit (src/bin/raytraced_shadows.wgsl:98-102).  The boolean does not depend
 * on the visit order, so the kernel is free to reorder the traversal (two stacks per lane, lanes of a warp sharing the last
 * long rays); rays for which that argument does not hold (zero / non-finite reciprocal direction, stack overflow) run in
 * the reference's order.  Every entry of the output is written.  The host-pointer calls cut large batches into chunks and
 * overlap upload, kernels (two compute streams) and read-back; give them pinned buffers. */
int bvh_cuda_trace_any(bvh_cuda_ctx* ctx, const bvh_cuda_scene* scene, const float* ray_o, const float* ray_d,
                       size_t n_rays, float tmax, uint8_t* occluded_out);
int bvh_cuda_trace_any_dev(bvh_cuda_ctx* ctx, const bvh_cuda_scene* scene, const float* d_ray_o,
                           const float* d_ray_d, size_t n_rays, float tmax, uint8_t* d_occluded_out,
                           void* stream);

/* ---- ray generation: the step right before traversal ------------------------------------------------------ *
 * Primary rays, one per pixel, from camera.clip_to_world (column-major 4x4, host pointer): src/bin/bvh_cpu.rs:72-84,
 * the CPU twin of src/bin/bvh_trace.wgsl:224-233.  Pixel i: x = (i % W)/W, y = (i / H)/H (sic, bvh_cpu.rs:75),
 * eye = (M*(x',y',1,1)).xyz / w, dir = normalize((M*(x',y',0,1)).xyz). */
int bvh_cuda_gen_primary_rays_dev(bvh_cuda_ctx* ctx, const float* clip_to_world, uint32_t width, uint32_t height,
                                  float* d_ray_o, float* d_ray_d, void* stream);
/* Point-light shadow rays from G-buffer world positions / normals (3 floats each, device):
 * ray_new(pos + nor * 0.0001, light.position - pos), src/bin/raytraced_shadows.wgsl:93-99.  light_pos: host, 3 floats. */
int bvh_cuda_gen_shadow_rays_dev(bvh_cuda_ctx* ctx, const float* d_pos, const float* d_nor, size_t n, const float* light_pos,
                                 float* d_ray_o, float* d_ray_d, void* stream);
/* Rect-area-light shadow rays: same origin, direction toward p0 + (p1-p0)*u + (p3-p0)*v with the corner order of
 * AreaLight::from_transform (crates/pools/src/light.rs:28-52); d_uv: 2 floats per ray; corners: host, 12 floats. */
int bvh_cuda_gen_area_shadow_rays_dev(bvh_cuda_ctx* ctx, const float* d_pos, const float* d_nor, const float* d_uv, size_t n,
                                      const float* corners, float* d_ray_o, float* d_ray_d, void* stream);

/* ---- per-frame instance update: the step that makes a TLAS rebuild necessary -------------------------------- *
 * shaders/compute_update.wgsl:12-27: for every id in d_ids (NULL = all n instances),
 *   instance.transform = from_rotation_z(+-angle) * instance.transform        (shaders/utils/math.wgsl from_rotation_z)
 * with + when transform[3][2] > -15.0 and - otherwise.  sin / cos are implementation-defined in WGSL, so the caller
 * passes sin_a = sin(angle), cos_a = cos(angle) for angle = 2*sin(0.5*time)*dt; the matrix product is evaluated
 * unfused, each component as the four-term column sum left to right.  The reference leaves inv_transform stale (it
 * never rebuilds the TLAS: crates/app/src/app.rs:253 runs once); update_inverse != 0 also sets
 * inv_transform = inv_transform * from_rotation_z(-+angle), so the moved instance can be traced after
 * bvh_cuda_tlas_build_dev.  A wrapped scene (bvh_cuda_scene_wrap_dev) sees the new instances / TLAS in place. */
int bvh_cuda_instances_rotate_z_dev(bvh_cuda_ctx* ctx, Instance* d_instances, const uint32_t* d_ids, size_t n, float sin_a,
                                    float cos_a, int update_inverse, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BVH_CUDA_H */
