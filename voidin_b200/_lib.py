"""Loader for libbvh_cuda.so (the C ABI of include/bvh_cuda.h).  There is no CPU fallback: a missing library or a
missing CUDA device is an error."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BVH_CUDA_LIB selects another build of the same library (scripts/variants.py compares compile-time variants)
LIB_PATH = os.environ.get("BVH_CUDA_LIB") or os.path.join(_HERE, "libbvh_cuda.so")

# every symbol include/bvh_cuda.h declares
SYMBOLS = [
    "bvh_cuda_abi_version",
    "bvh_cuda_create",
    "bvh_cuda_destroy",
    "bvh_cuda_last_error",
    "bvh_cuda_launch_count",
    "bvh_cuda_set_profiling",
    "bvh_cuda_blas_build",
    "bvh_cuda_blas_build_dev",
    "bvh_cuda_blas_build_batch_dev",
    "bvh_cuda_blas_build_batch_async_dev",
    "bvh_cuda_blas_build_finish",
    "bvh_cuda_blas_last_order",
    "bvh_cuda_blas_last_stats",
    "bvh_cuda_tlas_build",
    "bvh_cuda_tlas_build_dev",
    "bvh_cuda_scene_upload",
    "bvh_cuda_scene_wrap_dev",
    "bvh_cuda_scene_refresh_dev",
    "bvh_cuda_scene_instance_boxes_dev",
    "bvh_cuda_scene_free",
    "bvh_cuda_trace_blas",
    "bvh_cuda_trace_blas_dev",
    "bvh_cuda_trace_blas_recursive",
    "bvh_cuda_trace_blas_recursive_dev",
    "bvh_cuda_trace_closest",
    "bvh_cuda_trace_closest_dev",
    "bvh_cuda_trace_any",
    "bvh_cuda_trace_any_dev",
    "bvh_cuda_gen_primary_rays_dev",
    "bvh_cuda_gen_shadow_rays_dev",
    "bvh_cuda_gen_area_shadow_rays_dev",
    "bvh_cuda_instances_rotate_z_dev",
]


class BuildStats(C.Structure):
    _fields_ = [
        ("sum_interior_prims", C.c_uint64),
        ("n_nodes", C.c_uint32),
        ("interior_nodes", C.c_uint32),
        ("grid_levels", C.c_uint32),
        ("big_block_tasks", C.c_uint32),
        ("block_tasks", C.c_uint32),
        ("warp_node_tasks", C.c_uint32),
        ("warp_tasks", C.c_uint32),
        ("kernel_launches", C.c_uint32),
        ("ms_setup", C.c_float),
        ("ms_grid", C.c_float),
        ("ms_big_block", C.c_float),
        ("ms_block", C.c_float),
        ("ms_warp_node", C.c_float),
        ("ms_warp", C.c_float),
        ("ms_emit", C.c_float),
        ("ms_total", C.c_float),
        ("ms_thread", C.c_float),
        ("thread_tasks", C.c_uint32),
        ("grid_nodes", C.c_uint32),
        ("cluster_tasks", C.c_uint32),
        ("grid_interior_prims", C.c_uint64),
        ("ms_cluster", C.c_float),
        ("reserved0", C.c_uint32),
    ]

    def as_dict(self):
        return {k: (float(getattr(self, k)) if k.startswith("ms_") else int(getattr(self, k))) for k, _ in self._fields_}


class SceneDesc(C.Structure):
    _fields_ = [
        ("tlas_nodes", C.c_void_p), ("n_tlas_nodes", C.c_size_t),
        ("tlas_children", C.c_void_p),
        ("instances", C.c_void_p), ("n_instances", C.c_size_t),
        ("meshes", C.c_void_p), ("n_meshes", C.c_size_t),
        ("bvh_nodes", C.c_void_p), ("n_bvh_nodes", C.c_size_t),
        ("vertices", C.c_void_p), ("n_vertices", C.c_size_t),
        ("indices", C.c_void_p), ("n_indices", C.c_size_t),
    ]


class BvhCudaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"bvh_cuda error {code}: {message}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """dlopen the library and set prototypes.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  voidin_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    vp, sz, u32p = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32)
    lib.bvh_cuda_abi_version.restype = C.c_int
    lib.bvh_cuda_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.bvh_cuda_destroy.argtypes = [vp]
    lib.bvh_cuda_destroy.restype = None
    lib.bvh_cuda_last_error.argtypes = [vp]
    lib.bvh_cuda_last_error.restype = C.c_char_p
    lib.bvh_cuda_launch_count.argtypes = [vp]
    lib.bvh_cuda_launch_count.restype = C.c_uint64
    lib.bvh_cuda_set_profiling.argtypes = [vp, C.c_int]
    lib.bvh_cuda_blas_build.argtypes = [vp, vp, sz, vp, sz, vp, sz, u32p]
    lib.bvh_cuda_blas_build_dev.argtypes = [vp, vp, sz, vp, sz, vp, sz, u32p, vp]
    lib.bvh_cuda_blas_build_batch_dev.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, sz, u32p, vp]
    lib.bvh_cuda_blas_build_batch_async_dev.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, sz, vp, vp]
    lib.bvh_cuda_blas_build_finish.argtypes = [vp, u32p]
    lib.bvh_cuda_blas_last_order.argtypes = [vp, vp, sz]
    lib.bvh_cuda_blas_last_stats.argtypes = [vp, C.POINTER(BuildStats)]
    lib.bvh_cuda_tlas_build.argtypes = [vp, vp, sz, vp, sz, vp, vp]
    lib.bvh_cuda_tlas_build_dev.argtypes = [vp, vp, sz, vp, sz, vp, vp, vp]
    lib.bvh_cuda_scene_upload.argtypes = [vp, C.POINTER(SceneDesc), C.POINTER(vp)]
    lib.bvh_cuda_scene_wrap_dev.argtypes = [vp, C.POINTER(SceneDesc), vp, C.POINTER(vp)]
    lib.bvh_cuda_scene_refresh_dev.argtypes = [vp, vp, C.POINTER(SceneDesc), vp]
    lib.bvh_cuda_scene_instance_boxes_dev.argtypes = [vp, vp, C.c_int, vp]
    lib.bvh_cuda_scene_free.argtypes = [vp, vp]
    lib.bvh_cuda_scene_free.restype = None
    lib.bvh_cuda_trace_blas.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, vp, sz, vp, vp]
    lib.bvh_cuda_trace_blas_dev.argtypes = [vp, vp, vp, vp, vp, vp, sz, vp, vp, vp]
    lib.bvh_cuda_trace_blas_recursive.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, vp, sz, C.c_uint32, C.c_float, vp, vp]
    lib.bvh_cuda_trace_blas_recursive_dev.argtypes = [vp, vp, vp, vp, vp, vp, sz, C.c_uint32, C.c_float, vp, vp, vp]
    lib.bvh_cuda_trace_closest.argtypes = [vp, vp, vp, vp, sz, C.c_float, vp, vp, vp]
    lib.bvh_cuda_trace_closest_dev.argtypes = [vp, vp, vp, vp, sz, C.c_float, vp, vp, vp, vp]
    lib.bvh_cuda_trace_any.argtypes = [vp, vp, vp, vp, sz, C.c_float, vp]
    lib.bvh_cuda_trace_any_dev.argtypes = [vp, vp, vp, vp, sz, C.c_float, vp, vp]
    lib.bvh_cuda_gen_primary_rays_dev.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp]
    lib.bvh_cuda_gen_shadow_rays_dev.argtypes = [vp, vp, vp, sz, vp, vp, vp, vp]
    lib.bvh_cuda_gen_area_shadow_rays_dev.argtypes = [vp, vp, vp, vp, sz, vp, vp, vp, vp]
    lib.bvh_cuda_instances_rotate_z_dev.argtypes = [vp, vp, vp, sz, C.c_float, C.c_float, C.c_int, vp]
    for name in SYMBOLS:
        getattr(lib, name)  # AttributeError here means the header and the library disagree
    _lib = lib
    return lib
