"""POD layouts shared with the reference (numpy structured dtypes, byte-identical to the Rust `repr(C)` structs).

BvhNode   crates/bvh/src/blas.rs:10-17      (WGSL mirror shaders/utils/bvh.wgsl:11-16)
TlasNode  crates/bvh/src/tlas.rs:7-14       (WGSL mirror shaders/utils/bvh.wgsl:4-9)
Instance  crates/components/src/shared.rs:67-75  (column-major Mat4 x2, mesh, material, junk[2])
MeshInfo  crates/components/src/shared.rs:29-39
"""
import numpy as np

BVH_NODE = np.dtype(
    [("min", "<f4", (3,)), ("left_first", "<u4"), ("max", "<f4", (3,)), ("count", "<u4")], align=False
)
TLAS_NODE = np.dtype(
    [("min", "<f4", (3,)), ("left_right", "<u4"), ("max", "<f4", (3,)), ("instance_idx", "<u4")], align=False
)
INSTANCE = np.dtype(
    [
        ("transform", "<f4", (16,)),
        ("inv_transform", "<f4", (16,)),
        ("mesh", "<u4"),
        ("material", "<u4"),
        ("junk", "<u4", (2,)),
    ],
    align=False,
)
MESH_INFO = np.dtype(
    [
        ("min", "<f4", (3,)),
        ("index_count", "<u4"),
        ("max", "<f4", (3,)),
        ("base_index", "<u4"),
        ("vertex_offset", "<i4"),
        ("bvh_index", "<u4"),
        ("junk", "<u4", (2,)),
    ],
    align=False,
)

assert BVH_NODE.itemsize == 32 and TLAS_NODE.itemsize == 32
assert INSTANCE.itemsize == 144 and MESH_INFO.itemsize == 48

MAX_DIST = np.float32(1e30)  # crates/bvh/src/intersection.rs:3
NO_HIT = np.uint32(0xFFFFFFFF)

# status codes of the C ABI (include/bvh_cuda.h)
OK, EINVAL, EDEGENERATE, ECUDA, ENOMEM = 0, -1, -2, -3, -4
