/* bvh_cuda_models.h — C ABI of the asset loaders in libbvh_cuda.so: the step right before MeshPool::add.
 *
 * They replace the two importers voidin feeds its BLAS builder from (paths relative to the voidin checkout):
 *   ObjModel::import      crates/app/src/models/mod.rs:20-57          tobj::load_obj(path, &tobj::GPU_LOAD_OPTIONS)
 *   GltfDocument::import  crates/app/src/models/gltf_model/mod.rs     make_meshes :103-155, get_scene_instances /
 *                                                                      gather_instances_recursive :166-207
 * Output layout = what those hand to MeshPool::add (crates/pools/src/mesh/mod.rs:309): per mesh `positions`
 * (3 floats per vertex), optional `normals` (3) and `texcoords` (2), and u32 `indices`, 3 per triangle, mesh-local.
 *
 * tobj 4.0.0 and gltf 1.2.0 are third-party crates that are NOT part of the reference checkout (Cargo.lock only);
 * their published behaviour is restated here:
 *   OBJ  GPU_LOAD_OPTIONS = triangulate + single_index (+ ignore points / lines): one model per `o` / `g` / `usemtl`
 *        run that has faces; polygons are fan-triangulated (0, i-1, i); every distinct (v, vt, vn) index triple becomes
 *        one output vertex, numbered in order of first use inside the model; negative indices are relative.
 *   glTF .gltf (+ external .bin / base64 buffers) and .glb; one mesh per primitive that has POSITION and NORMAL
 *        (make_meshes skips the others, :117-122, and never looks at the primitive's mode); POSITION / NORMAL are the
 *        accessor's count*size CONTIGUOUS bytes (data_of_accessor :209-220 ignores byteStride — reproduced); tangents
 *        padded with (0,1,0,1), texcoord set 0 as f32 padded with 0; indices widened to u32 or 0..n when absent
 *        (:138-141); instances = depth-first walk of every scene, children before the node's own primitives
 *        (:183-205), transform = parent * node (column-major f32, glam order of operations), node matrix from
 *        `matrix` or T*R*S (gltf's own f32 matrix code).
 * Parity with the crates themselves is unpinned (no Rust toolchain here); the tests compare with independent Python
 * readers of the same files and with committed fixtures extracted from the reference's own assets.
 * No GPU is needed for any of these calls.
 */
#ifndef BVH_CUDA_MODELS_H
#define BVH_CUDA_MODELS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bvh_cuda_model bvh_cuda_model;

typedef struct BvhCudaMeshView {
    const float* positions;   /* 3 * n_vertices */
    const float* normals;     /* 3 * n_normals, or NULL */
    const float* texcoords;   /* 2 * n_texcoords, or NULL */
    const float* tangents;    /* 4 * n_vertices (OBJ: zeros, models/mod.rs:44; glTF: TANGENT padded with (0,1,0,1)) */
    const uint32_t* indices;  /* n_indices; 3 per triangle for OBJ and for glTF triangle lists */
    size_t n_vertices;
    size_t n_normals;         /* == n_vertices, except for an OBJ whose faces mix corners with and without vn / vt */
    size_t n_texcoords;
    size_t n_indices;
    int32_t material;         /* OBJ: index into the .mtl list; glTF: material index; -1 = none */
    int32_t gltf_mesh;        /* glTF: (mesh index, primitive index) the mesh came from; -1 for OBJ */
    int32_t gltf_primitive;
    int32_t reserved;
    const char* name;         /* OBJ model name / glTF mesh name, "" if none */
} BvhCudaMeshView;

typedef struct BvhCudaMaterialView {
    float base_color[4]; /* OBJ: (Kd, 0.5) (models/mod.rs:30-33); glTF: pbrMetallicRoughness.baseColorFactor */
    const char* name;
} BvhCudaMaterialView;

typedef struct BvhCudaInstanceView {
    float transform[16]; /* column-major, as glam Mat4 / Instance.transform (crates/components/src/shared.rs:67-75) */
    uint32_t mesh;       /* index into this model's meshes */
    int32_t material;
} BvhCudaInstanceView;

/* 0 ok, <0 error (BVH_CUDA_EINVAL: unreadable or malformed file); *out owns everything the views point to. */
int bvh_cuda_model_load_obj(const char* path, bvh_cuda_model** out);
int bvh_cuda_model_load_gltf(const char* path, bvh_cuda_model** out);
void bvh_cuda_model_free(bvh_cuda_model* model);
/* Message of the last failing load on this thread ("" if none). */
const char* bvh_cuda_model_last_error(void);

size_t bvh_cuda_model_mesh_count(const bvh_cuda_model* model);
int bvh_cuda_model_mesh(const bvh_cuda_model* model, size_t i, BvhCudaMeshView* out);
size_t bvh_cuda_model_material_count(const bvh_cuda_model* model);
int bvh_cuda_model_material(const bvh_cuda_model* model, size_t i, BvhCudaMaterialView* out);
/* glTF: get_scene_instances(Mat4::IDENTITY); OBJ: 0 instances. */
size_t bvh_cuda_model_instance_count(const bvh_cuda_model* model);
int bvh_cuda_model_instance(const bvh_cuda_model* model, size_t i, BvhCudaInstanceView* out);

#ifdef __cplusplus
}
#endif
#endif /* BVH_CUDA_MODELS_H */
