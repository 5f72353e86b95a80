"""ctypes loader for the CPU oracle (oracle/bvh_oracle.cpp).  TEST INFRASTRUCTURE — only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product package (voidin_b200) never does."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbvh_oracle.so")

BVH_NODE = np.dtype([("min", "<f4", (3,)), ("left_first", "<u4"), ("max", "<f4", (3,)), ("count", "<u4")])
TLAS_NODE = np.dtype([("min", "<f4", (3,)), ("left_right", "<u4"), ("max", "<f4", (3,)), ("instance_idx", "<u4")])

OK, EINVAL, EDEGENERATE = 0, -1, -2


class BuildStats(C.Structure):
    _fields_ = [
        ("sum_interior_prims", C.c_uint64),
        ("interior_nodes", C.c_uint32),
        ("max_depth", C.c_uint32),
        ("candidates", C.c_uint64),
        ("unexamined_was_left", C.c_uint64),
        ("nan_candidates", C.c_uint64),
        ("final_pivot_differs", C.c_uint32),
        ("reserved", C.c_uint32),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_ if k != "reserved"}


class RayStats(C.Structure):
    _fields_ = [
        ("pops", C.c_uint64),
        ("interior_visits", C.c_uint64),
        ("triangle_tests", C.c_uint64),
        ("instance_visits", C.c_uint64),
        ("max_stack", C.c_uint64),
        ("hits", C.c_uint64),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def build_lib(force: bool = False) -> str:
    src = os.path.join(_HERE, "bvh_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libbvh_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_lib())
        _lib.oracle_shuffle_seq.restype = C.c_uint32
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def blas_build(vertices, indices, model: bool = False):
    """BvhBuilder::new(vertices, indices).build() (blas.rs:51-103).
    Returns (status, nodes, permuted_indices, prim_order, stats_dict)."""
    v = _f32(vertices).reshape(-1, 3)
    idx = np.array(indices, dtype=np.uint32, copy=True).reshape(-1)
    n = idx.size // 3
    nodes = np.zeros(max(2 * n, 2), dtype=BVH_NODE)
    order = np.zeros(max(n, 1), dtype=np.uint32)
    m = C.c_uint32(0)
    st = BuildStats()
    if model:
        rc = lib().oracle_blas_build_model(_p(v), C.c_size_t(v.shape[0]), _p(idx), C.c_size_t(n), _p(nodes),
                                           C.byref(m), _p(order))
    else:
        rc = lib().oracle_blas_build(_p(v), C.c_size_t(v.shape[0]), _p(idx), C.c_size_t(n), _p(nodes),
                                     C.byref(m), _p(order), C.byref(st))
    return rc, nodes[: m.value].copy(), idx, order[:n], st.as_dict()


def shuffle_seq(ids, flags_by_id):
    ids = np.array(ids, dtype=np.uint32, copy=True)
    fl = np.ascontiguousarray(flags_by_id, dtype=np.uint8)
    piv = lib().oracle_shuffle_seq(_p(ids), _p(fl), C.c_uint32(ids.size))
    return int(piv), ids


def tlas_build(instances, meshes):
    """Tlas::build (tlas.rs:31-85).  Returns (status, nodes[2I+1], children[2I+1,2], calls, pairs)."""
    inst = np.ascontiguousarray(instances)
    mesh = np.ascontiguousarray(meshes)
    n = inst.shape[0]
    nodes = np.zeros(2 * n + 1, dtype=TLAS_NODE)
    kids = np.zeros((2 * n + 1, 2), dtype=np.uint32)
    calls, pairs = C.c_uint64(0), C.c_uint64(0)
    rc = lib().oracle_tlas_build(_p(inst), C.c_size_t(n), _p(mesh), C.c_size_t(mesh.shape[0]), _p(nodes),
                                 _p(kids), C.byref(calls), C.byref(pairs))
    return rc, nodes, kids, calls.value, pairs.value


def trace_blas(nodes, vertices, perm_indices, ray_o, ray_d, threads: int = 1):
    """Bvh::traverse_iter per ray (blas.rs:247-295).  Returns (t, tri, stats)."""
    o, d = _f32(ray_o).reshape(-1, 3), _f32(ray_d).reshape(-1, 3)
    r = o.shape[0]
    t = np.empty(r, dtype=np.float32)
    tri = np.empty(r, dtype=np.uint32)
    st = RayStats()
    v = _f32(vertices)
    idx = np.ascontiguousarray(perm_indices, dtype=np.uint32)
    nd = np.ascontiguousarray(nodes)
    lib().oracle_trace_blas(_p(nd), _p(v), _p(idx), _p(o), _p(d), C.c_size_t(r), _p(t), _p(tri), C.byref(st),
                            C.c_int(threads))
    return t, tri, st.as_dict()


def trace_blas_recursive(nodes, vertices, perm_indices, ray_o, ray_d, node_idx: int = 0, t0: float = 1e30):
    """Bvh::traverse per ray (blas.rs:211-245).  Returns (hit[bool], t)."""
    o, d = _f32(ray_o).reshape(-1, 3), _f32(ray_d).reshape(-1, 3)
    r = o.shape[0]
    t = np.empty(r, dtype=np.float32)
    hit = np.empty(r, dtype=np.uint8)
    v = _f32(vertices)
    idx = np.ascontiguousarray(perm_indices, dtype=np.uint32)
    nd = np.ascontiguousarray(nodes)
    lib().oracle_trace_blas_recursive(_p(nd), _p(v), _p(idx), _p(o), _p(d), C.c_size_t(r), C.c_uint32(node_idx), C.c_float(t0),
                                      _p(t), _p(hit))
    return hit.astype(bool), t


def trace_scene(tlas, children, instances, meshes, bvh_nodes, vertices, indices, ray_o, ray_d, tmax=1e30,
                any_hit: bool = False, threads: int = 1):
    """traverse_tlas per ray (bvh.wgsl:89-123).  Returns (t, tri, inst, occluded, stats)."""
    o, d = _f32(ray_o).reshape(-1, 3), _f32(ray_d).reshape(-1, 3)
    r = o.shape[0]
    t = np.empty(r, dtype=np.float32)
    tri = np.empty(r, dtype=np.uint32)
    ins = np.empty(r, dtype=np.uint32)
    occ = np.empty(r, dtype=np.uint8)
    st = RayStats()
    keep = [np.ascontiguousarray(x) for x in (tlas, instances, meshes, bvh_nodes)]
    v = _f32(vertices)
    idx = np.ascontiguousarray(indices, dtype=np.uint32)
    kids = None if children is None else np.ascontiguousarray(children, dtype=np.uint32)
    lib().oracle_trace_scene(_p(keep[0]), _p(kids), _p(keep[1]), _p(keep[2]), _p(keep[3]), _p(v), _p(idx), _p(o),
                             _p(d), C.c_size_t(r), C.c_float(tmax), C.c_int(1 if any_hit else 0), _p(t), _p(tri),
                             _p(ins), _p(occ), C.byref(st), C.c_int(threads))
    return t, tri, ins, occ, st.as_dict()


def brute_force(vertices, indices, ray_o, ray_d, mode: int):
    o, d = _f32(ray_o).reshape(-1, 3), _f32(ray_d).reshape(-1, 3)
    t = np.empty(o.shape[0], dtype=np.float32)
    v = _f32(vertices)
    idx = np.ascontiguousarray(indices, dtype=np.uint32)
    lib().oracle_brute_force(_p(v), _p(idx), C.c_size_t(idx.size // 3), _p(o), _p(d), C.c_size_t(o.shape[0]),
                             C.c_int(mode), _p(t))
    return t


def instances_rotate_z(instances, ids, sin_a: float, cos_a: float, update_inverse: bool = False):
    """compute_update.wgsl:12-27 on a copy of `instances`; ids=None rotates all of them."""
    inst = np.array(instances, copy=True)
    if ids is None:
        lib().oracle_instances_rotate_z(_p(inst), None, C.c_size_t(inst.shape[0]), C.c_float(sin_a), C.c_float(cos_a),
                                        C.c_int(1 if update_inverse else 0))
    else:
        i = np.ascontiguousarray(ids, dtype=np.uint32)
        lib().oracle_instances_rotate_z(_p(inst), _p(i), C.c_size_t(i.size), C.c_float(sin_a), C.c_float(cos_a),
                                        C.c_int(1 if update_inverse else 0))
    return inst


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def gen_primary_rays(clip_to_world, width: int, height: int):
    m = _f32(clip_to_world).reshape(16)
    ro = np.empty((width * height, 3), dtype=np.float32)
    rd = np.empty((width * height, 3), dtype=np.float32)
    lib().oracle_gen_primary_rays(_p(m), C.c_uint32(width), C.c_uint32(height), _p(ro), _p(rd))
    return ro, rd


def gen_shadow_rays(pos, nor, light):
    p, n = _f32(pos).reshape(-1, 3), _f32(nor).reshape(-1, 3)
    l = _f32(light).reshape(3)
    ro, rd = np.empty_like(p), np.empty_like(p)
    lib().oracle_gen_shadow_rays(_p(p), _p(n), C.c_size_t(p.shape[0]), _p(l), _p(ro), _p(rd))
    return ro, rd


def gen_area_shadow_rays(pos, nor, uv, corners):
    """CPU twin of bvh_cuda_gen_area_shadow_rays_dev (rect light, corner order of crates/pools/src/light.rs:28-52)."""
    p, n, w = _f32(pos).reshape(-1, 3), _f32(nor).reshape(-1, 3), _f32(uv).reshape(-1, 2)
    c = _f32(corners).reshape(4, 3)
    ro, rd = np.empty_like(p), np.empty_like(p)
    lib().oracle_gen_area_shadow_rays(_p(p), _p(n), _p(w), C.c_size_t(p.shape[0]), _p(c), _p(ro), _p(rd))
    return ro, rd
