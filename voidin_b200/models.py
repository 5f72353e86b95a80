"""Host-side mirror of voidin's asset importers, on top of the C ABI in include/bvh_cuda_models.h (the loaders live in
libbvh_cuda.so, `csrc/models.cpp`):

    ObjModel.import_(path)      <- ObjModel::import      crates/app/src/models/mod.rs:20-57   (tobj GPU_LOAD_OPTIONS)
    GltfDocument.import_(path)  <- GltfDocument::import  crates/app/src/models/gltf_model/mod.rs:103-207

Both return the per-mesh arrays that voidin passes to `App::add_mesh` -> `MeshPool::add`
(crates/pools/src/mesh/mod.rs:309-351), i.e. the input of `BvhBuilder::new(vertices, indices).build()`.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _lib
from ._lib import BvhCudaError


class _MeshView(C.Structure):
    _fields_ = [
        ("positions", C.POINTER(C.c_float)), ("normals", C.POINTER(C.c_float)), ("texcoords", C.POINTER(C.c_float)),
        ("tangents", C.POINTER(C.c_float)), ("indices", C.POINTER(C.c_uint32)),
        ("n_vertices", C.c_size_t), ("n_normals", C.c_size_t), ("n_texcoords", C.c_size_t), ("n_indices", C.c_size_t),
        ("material", C.c_int32), ("gltf_mesh", C.c_int32), ("gltf_primitive", C.c_int32), ("reserved", C.c_int32),
        ("name", C.c_char_p),
    ]


class _MaterialView(C.Structure):
    _fields_ = [("base_color", C.c_float * 4), ("name", C.c_char_p)]


class _InstanceView(C.Structure):
    _fields_ = [("transform", C.c_float * 16), ("mesh", C.c_uint32), ("material", C.c_int32)]


@dataclass
class MeshRef:
    """`MeshRef` of crates/pools/src/mesh/mod.rs:24-31 plus where it came from."""
    vertices: np.ndarray      # [V, 3] f32
    normals: np.ndarray       # [V, 3] f32 (may be shorter for an OBJ that mixes corners with / without vn)
    tangents: np.ndarray      # [V, 4] f32
    tex_coords: np.ndarray    # [V, 2] f32
    indices: np.ndarray       # [I] u32, mesh-local
    material: int = -1
    name: str = ""
    gltf_mesh: int = -1
    gltf_primitive: int = -1


@dataclass
class ModelInstance:
    transform: np.ndarray     # [4, 4] f32, math convention (column-major storage transposed)
    mesh: int
    material: int


@dataclass
class Model:
    meshes: List[MeshRef] = field(default_factory=list)
    materials: List[dict] = field(default_factory=list)
    instances: List[ModelInstance] = field(default_factory=list)


_proto_done = False


def _lib_models():
    global _proto_done
    lib = _lib.load()
    if not _proto_done:
        vp = C.c_void_p
        lib.bvh_cuda_model_load_obj.argtypes = [C.c_char_p, C.POINTER(vp)]
        lib.bvh_cuda_model_load_gltf.argtypes = [C.c_char_p, C.POINTER(vp)]
        lib.bvh_cuda_model_free.argtypes = [vp]
        lib.bvh_cuda_model_free.restype = None
        lib.bvh_cuda_model_last_error.restype = C.c_char_p
        for n in ("mesh", "material", "instance"):
            getattr(lib, f"bvh_cuda_model_{n}_count").argtypes = [vp]
            getattr(lib, f"bvh_cuda_model_{n}_count").restype = C.c_size_t
        lib.bvh_cuda_model_mesh.argtypes = [vp, C.c_size_t, C.POINTER(_MeshView)]
        lib.bvh_cuda_model_material.argtypes = [vp, C.c_size_t, C.POINTER(_MaterialView)]
        lib.bvh_cuda_model_instance.argtypes = [vp, C.c_size_t, C.POINTER(_InstanceView)]
        _proto_done = True
    return lib


def _arr(ptr, n, dtype, cols):
    if not ptr or n == 0:
        return np.zeros((0, cols) if cols > 1 else (0,), dtype=dtype)
    a = np.ctypeslib.as_array(ptr, shape=(n * cols,)).astype(dtype, copy=True)
    return a.reshape(n, cols) if cols > 1 else a


def _load(path: str, fn_name: str) -> Model:
    lib = _lib_models()
    h = C.c_void_p()
    rc = getattr(lib, fn_name)(os.fsencode(path), C.byref(h))
    if rc != 0:
        raise BvhCudaError(rc, (lib.bvh_cuda_model_last_error() or b"").decode("utf-8", "replace"))
    try:
        out = Model()
        mv, tv, iv = _MeshView(), _MaterialView(), _InstanceView()
        for i in range(lib.bvh_cuda_model_mesh_count(h)):
            lib.bvh_cuda_model_mesh(h, i, C.byref(mv))
            out.meshes.append(MeshRef(
                vertices=_arr(mv.positions, mv.n_vertices, np.float32, 3),
                normals=_arr(mv.normals, mv.n_normals, np.float32, 3),
                tangents=_arr(mv.tangents, mv.n_vertices, np.float32, 4),
                tex_coords=_arr(mv.texcoords, mv.n_texcoords, np.float32, 2),
                indices=_arr(mv.indices, mv.n_indices, np.uint32, 1),
                material=mv.material, name=(mv.name or b"").decode("utf-8", "replace"),
                gltf_mesh=mv.gltf_mesh, gltf_primitive=mv.gltf_primitive))
        for i in range(lib.bvh_cuda_model_material_count(h)):
            lib.bvh_cuda_model_material(h, i, C.byref(tv))
            out.materials.append({"name": (tv.name or b"").decode("utf-8", "replace"), "base_color": np.array(tv.base_color[:], dtype=np.float32)})
        for i in range(lib.bvh_cuda_model_instance_count(h)):
            lib.bvh_cuda_model_instance(h, i, C.byref(iv))
            out.instances.append(ModelInstance(np.array(iv.transform[:], dtype=np.float32).reshape(4, 4).T.copy(), iv.mesh, iv.material))
        return out
    finally:
        lib.bvh_cuda_model_free(h)


class ObjModel:
    """ObjModel::import (crates/app/src/models/mod.rs:20-57)."""

    @staticmethod
    def import_(path: str) -> Model:
        return _load(path, "bvh_cuda_model_load_obj")


class GltfDocument:
    """GltfDocument::import + get_scene_instances(Mat4::IDENTITY) (crates/app/src/models/gltf_model/mod.rs)."""

    @staticmethod
    def import_(path: str) -> Model:
        return _load(path, "bvh_cuda_model_load_gltf")


def find_asset(name: str) -> Optional[str]:
    """Where a voidin asset would be if the user dropped it in: $VOIDIN_ASSETS/<name>, <repo>/assets/<name>."""
    roots = [os.environ.get("VOIDIN_ASSETS"), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")]
    for r in roots:
        if r and os.path.exists(os.path.join(r, name)):
            return os.path.join(r, name)
    return None


def load_single_mesh(path: str):
    """All meshes of an .obj / .gltf / .glb file pooled into ONE (vertices [V,3] f32, indices [3N] u32) pair, the way
    bench.py wants its `dragon.obj` / `bunny.obj`: indices rebased by the running vertex count."""
    model = (ObjModel if path.lower().endswith(".obj") else GltfDocument).import_(path)
    vs, is_, base = [], [], 0
    for m in model.meshes:
        vs.append(m.vertices)
        is_.append(m.indices[: m.indices.size // 3 * 3] + np.uint32(base))
        base += m.vertices.shape[0]
    if not vs:
        raise BvhCudaError(-1, f"{path}: no meshes")
    return np.ascontiguousarray(np.concatenate(vs), dtype=np.float32), np.ascontiguousarray(np.concatenate(is_), dtype=np.uint32)
