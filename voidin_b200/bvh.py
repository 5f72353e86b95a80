"""Host-side mirror of voidin's `crates/bvh` public interface (crates/bvh/src/lib.rs:5-7) on top of the C ABI:

    BvhBuilder(vertices, indices).set_bin_number(n).build() -> Bvh      blas.rs:51-103
    Bvh.traverse_iter(vertices, indices, ray) -> Dist                  blas.rs:247-295
    Tlas.empty(); Tlas.build(instances, meshes)                        tlas.rs:27-85
    Ray(orig, dir); Dist = ("Hit", t) | "Miss"                         intersection.rs:22-26,57-66

Same names, argument meaning and ownership as the Rust API: `indices` is permuted IN PLACE, the builder is
consumed by build(), Tlas.nodes is replaced wholesale.  Where Rust panics / never terminates this raises
BvhCudaError.  Batched (array-of-rays) and device-pointer variants are additions for the benchmark.
All computation happens in libbvh_cuda.so on the GPU; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .types import BVH_NODE, INSTANCE, MESH_INFO, TLAS_NODE, MAX_DIST


def _vp(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Context:
    """One bvh_cuda_ctx (one per host thread and device)."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.bvh_cuda_create(device, C.byref(h))
        if rc != 0:
            raise _lib.BvhCudaError(rc, "bvh_cuda_create failed (no CUDA device? voidin_b200 has no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.bvh_cuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != 0:
            raise _lib.BvhCudaError(rc, self.lib.bvh_cuda_last_error(self.h).decode())

    def set_profiling(self, enable: bool):
        self.check(self.lib.bvh_cuda_set_profiling(self.h, 1 if enable else 0))

    @property
    def launch_count(self) -> int:
        return int(self.lib.bvh_cuda_launch_count(self.h))

    def last_build_stats(self) -> dict:
        st = _lib.BuildStats()
        self.check(self.lib.bvh_cuda_blas_last_stats(self.h, C.byref(st)))
        return st.as_dict()

    def last_order(self, n_tris: int) -> np.ndarray:
        out = np.empty(n_tris, dtype=np.uint32)
        self.check(self.lib.bvh_cuda_blas_last_order(self.h, _vp(out), n_tris))
        return out

    # ---- device-pointer entry points (ints are raw device addresses, e.g. torch.Tensor.data_ptr()) ----
    def blas_build_dev(self, d_vertices: int, n_vertices: int, d_indices: int, n_tris: int, d_nodes: int,
                       nodes_cap: int, stream: int = 0) -> int:
        m = C.c_uint32(0)
        self.check(self.lib.bvh_cuda_blas_build_dev(self.h, d_vertices, n_vertices, d_indices, n_tris, d_nodes,
                                                    nodes_cap, C.byref(m), stream))
        return int(m.value)

    def blas_build_batch_dev(self, d_vertices: int, n_vertices: int, d_indices: int, n_indices: int, d_mesh_info: int,
                             n_meshes: int, d_nodes: int, nodes_cap: int, stream: int = 0) -> int:
        """Forest build of a pooled scene (MeshPool::add x n_meshes); fills MeshInfo.bvh_index on the device."""
        m = C.c_uint32(0)
        self.check(self.lib.bvh_cuda_blas_build_batch_dev(self.h, d_vertices, n_vertices, d_indices, n_indices, d_mesh_info,
                                                          n_meshes, d_nodes, nodes_cap, C.byref(m), stream))
        return int(m.value)

    def blas_build_batch_async_dev(self, d_vertices: int, n_vertices: int, d_indices: int, n_indices: int, d_mesh_info: int,
                                   n_meshes: int, d_nodes: int, nodes_cap: int, d_result: int = 0, stream: int = 0) -> None:
        """Stream-ordered forest build: enqueues the whole build on `stream` and returns; `blas_build_finish` collects it.
        d_result (device, 4 x u32, optional): {nodes, status bits, interior nodes, 0} written when the build completes."""
        self.check(self.lib.bvh_cuda_blas_build_batch_async_dev(self.h, d_vertices, n_vertices, d_indices, n_indices, d_mesh_info or None,
                                                                n_meshes, d_nodes, nodes_cap, d_result or None, stream))

    def blas_build_finish(self) -> int:
        """Waits for the pending asynchronous build of this context; returns its total node count (raises on a failed build)."""
        m = C.c_uint32(0)
        self.check(self.lib.bvh_cuda_blas_build_finish(self.h, C.byref(m)))
        return int(m.value)

    def tlas_build_dev(self, d_instances: int, n_inst: int, d_meshes: int, n_mesh: int, d_nodes: int,
                       d_children: int, stream: int = 0):
        self.check(self.lib.bvh_cuda_tlas_build_dev(self.h, d_instances, n_inst, d_meshes, n_mesh, d_nodes,
                                                    d_children or None, stream))

    def trace_blas_dev(self, d_nodes, d_vertices, d_indices, d_ro, d_rd, n_rays, d_t, d_tri, stream: int = 0):
        self.check(self.lib.bvh_cuda_trace_blas_dev(self.h, d_nodes, d_vertices, d_indices, d_ro, d_rd, n_rays, d_t,
                                                    d_tri, stream))


def gen_primary_rays_dev(ctx: Context, clip_to_world, width: int, height: int, d_ro: int, d_rd: int, stream: int = 0):
    """Primary rays per pixel (src/bin/bvh_cpu.rs:72-84).  clip_to_world: 16 floats, column-major (glam Mat4)."""
    m = np.ascontiguousarray(clip_to_world, dtype=np.float32).reshape(16)
    ctx.check(ctx.lib.bvh_cuda_gen_primary_rays_dev(ctx.h, _vp(m), width, height, d_ro, d_rd, stream))


def gen_shadow_rays_dev(ctx: Context, d_pos: int, d_nor: int, n: int, light_pos, d_ro: int, d_rd: int, stream: int = 0):
    """ray_new(pos + nor*1e-4, light.position - pos) (src/bin/raytraced_shadows.wgsl:93-99)."""
    l = np.ascontiguousarray(light_pos, dtype=np.float32).reshape(3)
    ctx.check(ctx.lib.bvh_cuda_gen_shadow_rays_dev(ctx.h, d_pos, d_nor, n, _vp(l), d_ro, d_rd, stream))


def gen_area_shadow_rays_dev(ctx: Context, d_pos: int, d_nor: int, d_uv: int, n: int, corners, d_ro: int, d_rd: int, stream: int = 0):
    c = np.ascontiguousarray(corners, dtype=np.float32).reshape(12)
    ctx.check(ctx.lib.bvh_cuda_gen_area_shadow_rays_dev(ctx.h, d_pos, d_nor, d_uv, n, _vp(c), d_ro, d_rd, stream))


def instances_rotate_z_dev(ctx: Context, d_instances: int, d_ids: int | None, n: int, sin_a: float, cos_a: float,
                           update_inverse: bool = False, stream: int = 0):
    """shaders/compute_update.wgsl:12-27 on device-resident instances (d_ids: device u32 ids, None/0 = all n):
    transform = from_rotation_z(+-angle) * transform.  Rebuild the TLAS afterwards (Tlas.build / tlas_build_dev)."""
    ctx.check(ctx.lib.bvh_cuda_instances_rotate_z_dev(ctx.h, d_instances, d_ids or None, n, float(sin_a), float(cos_a),
                                                      1 if update_inverse else 0, stream))


_default_ctx: Context | None = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


@dataclass
class Ray:  # intersection.rs:57-66
    orig: np.ndarray
    dir: np.ndarray

    @staticmethod
    def new(orig, dir):
        return Ray(np.asarray(orig, dtype=np.float32), np.asarray(dir, dtype=np.float32))


MISS = "Miss"


def Hit(t: float):
    return ("Hit", float(t))


class Bvh:  # blas.rs:206-208
    def __init__(self, nodes: np.ndarray, ctx: Context):
        self.nodes = nodes
        self._ctx = ctx

    def traverse_iter_batch(self, vertices, indices, ray_o, ray_d):
        """Bvh::traverse_iter for an array of rays.  Returns (t, tri): t = 1e30 / tri = 0xFFFFFFFF on a miss."""
        ctx = self._ctx
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        o = np.ascontiguousarray(ray_o, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(ray_d, dtype=np.float32).reshape(-1, 3)
        t = np.empty(o.shape[0], dtype=np.float32)
        tri = np.empty(o.shape[0], dtype=np.uint32)
        nodes = np.ascontiguousarray(self.nodes)
        ctx.check(ctx.lib.bvh_cuda_trace_blas(ctx.h, _vp(nodes), nodes.shape[0], _vp(v), v.shape[0], _vp(idx),
                                              idx.size // 3, _vp(o), _vp(d), o.shape[0], _vp(t), _vp(tri)))
        return t, tri

    def traverse_batch(self, vertices, indices, ray_o, ray_d, node_idx: int = 0, t: float = 1e30):
        """Bvh::traverse (recursive variant, blas.rs:211-245) for an array of rays.  Returns (hit[bool], t)."""
        ctx = self._ctx
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        o = np.ascontiguousarray(ray_o, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(ray_d, dtype=np.float32).reshape(-1, 3)
        tt = np.empty(o.shape[0], dtype=np.float32)
        hit = np.empty(o.shape[0], dtype=np.uint8)
        nodes = np.ascontiguousarray(self.nodes)
        ctx.check(ctx.lib.bvh_cuda_trace_blas_recursive(ctx.h, _vp(nodes), nodes.shape[0], _vp(v), v.shape[0], _vp(idx), idx.size // 3,
                                                        _vp(o), _vp(d), o.shape[0], node_idx, t, _vp(tt), _vp(hit)))
        return hit.astype(bool), tt

    def traverse(self, vertices, indices, ray: Ray, node_idx: int = 0, t: float = 1e30):
        hit, tt = self.traverse_batch(vertices, indices, ray.orig[None, :], ray.dir[None, :], node_idx, t)
        return Hit(tt[0]) if hit[0] else MISS

    def traverse_iter(self, vertices, indices, ray: Ray):
        t, _ = self.traverse_iter_batch(vertices, indices, ray.orig[None, :], ray.dir[None, :])
        return MISS if t[0] >= MAX_DIST else Hit(t[0])


class BvhBuilder:  # blas.rs:41-67
    def __init__(self, vertices: np.ndarray, indices: np.ndarray, ctx: Context | None = None):
        """`vertices`: [V,3] float32.  `indices`: C-contiguous uint32 array of 3*N entries ([N,3] or flat) that is
        permuted in place by build(), as `&mut [UVec3]` is in the reference."""
        self._ctx = ctx or default_context()
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        if not (isinstance(indices, np.ndarray) and indices.dtype == np.uint32 and indices.flags.c_contiguous
                and indices.flags.writeable):
            raise TypeError("indices must be a writable C-contiguous uint32 ndarray (it is permuted in place)")
        if indices.size % 3 != 0:
            raise _lib.BvhCudaError(-1, "indices.len() % 3 != 0 (bytemuck::cast_slice_mut panics, mesh/mod.rs:321)")
        self.indices = indices
        self.num_bins = 8

    def set_bin_number(self, num_bins: int) -> "BvhBuilder":
        self.num_bins = num_bins  # stored and never read, exactly like blas.rs:64-67 / :136
        return self

    def build(self) -> Bvh:
        ctx = self._ctx
        n = self.indices.size // 3
        nodes = np.zeros(max(2 * n, 2), dtype=BVH_NODE)
        m = C.c_uint32(0)
        ctx.check(ctx.lib.bvh_cuda_blas_build(ctx.h, _vp(self.vertices), self.vertices.shape[0], _vp(self.indices), n,
                                              _vp(nodes), nodes.shape[0], C.byref(m)))
        return Bvh(nodes[: m.value].copy(), ctx)


class Tlas:  # tlas.rs:22-29
    def __init__(self, ctx: Context | None = None):
        self.nodes = np.zeros(0, dtype=TLAS_NODE)
        self.children = np.zeros((0, 2), dtype=np.uint32)  # side buffer (see include/bvh_cuda.h)
        self._ctx = ctx or default_context()

    @staticmethod
    def empty(ctx: Context | None = None) -> "Tlas":
        return Tlas(ctx)

    def build(self, instances: np.ndarray, meshes: np.ndarray):
        ctx = self._ctx
        inst = np.ascontiguousarray(instances, dtype=INSTANCE)
        mesh = np.ascontiguousarray(meshes, dtype=MESH_INFO)
        n = inst.shape[0]
        nodes = np.zeros(2 * n + 1, dtype=TLAS_NODE)
        kids = np.zeros((2 * n + 1, 2), dtype=np.uint32)
        ctx.check(ctx.lib.bvh_cuda_tlas_build(ctx.h, _vp(inst), n, _vp(mesh), mesh.shape[0], _vp(nodes), _vp(kids)))
        self.nodes, self.children = nodes, kids


class Scene:
    """The trace bind group (crates/pools/src/mesh/mod.rs:136-238): tlas_nodes, instances, meshes, bvh_nodes,
    vertices, indices — uploaded once, then traversed with traverse_tlas semantics (shaders/utils/bvh.wgsl:89)."""

    def __init__(self, tlas_nodes, tlas_children, instances, meshes, bvh_nodes, vertices, indices,
                 ctx: Context | None = None, device_ptrs: bool = False, counts: dict | None = None, stream: int = 0):
        self._ctx = ctx or default_context()
        d = _lib.SceneDesc()
        if device_ptrs:
            d.tlas_nodes, d.tlas_children = tlas_nodes, tlas_children or None
            d.instances, d.meshes, d.bvh_nodes, d.vertices, d.indices = instances, meshes, bvh_nodes, vertices, indices
            d.n_tlas_nodes, d.n_instances, d.n_meshes = counts["tlas_nodes"], counts["instances"], counts["meshes"]
            d.n_bvh_nodes, d.n_vertices, d.n_indices = counts["bvh_nodes"], counts["vertices"], counts["indices"]
            h = C.c_void_p()
            self._ctx.check(self._ctx.lib.bvh_cuda_scene_wrap_dev(self._ctx.h, C.byref(d), stream, C.byref(h)))
            self.h = h
            self._desc = d
            return
        else:
            keep = [np.ascontiguousarray(tlas_nodes, dtype=TLAS_NODE),
                    None if tlas_children is None else np.ascontiguousarray(tlas_children, dtype=np.uint32),
                    np.ascontiguousarray(instances, dtype=INSTANCE), np.ascontiguousarray(meshes, dtype=MESH_INFO),
                    np.ascontiguousarray(bvh_nodes, dtype=BVH_NODE),
                    np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3),
                    np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)]
            d.tlas_nodes, d.n_tlas_nodes = keep[0].ctypes.data, keep[0].shape[0]
            d.tlas_children = None if keep[1] is None else keep[1].ctypes.data
            d.instances, d.n_instances = keep[2].ctypes.data, keep[2].shape[0]
            d.meshes, d.n_meshes = keep[3].ctypes.data, keep[3].shape[0]
            d.bvh_nodes, d.n_bvh_nodes = keep[4].ctypes.data, keep[4].shape[0]
            d.vertices, d.n_vertices = keep[5].ctypes.data, keep[5].shape[0]
            d.indices, d.n_indices = keep[6].ctypes.data, keep[6].size
            fn = self._ctx.lib.bvh_cuda_scene_upload
        h = C.c_void_p()
        self._ctx.check(fn(self._ctx.h, C.byref(d), C.byref(h)))
        self.h = h

    def refresh_dev(self, n_bvh_nodes: int | None = None, stream: int = 0):
        """Re-bake a device-wrapped scene after its buffers were rewritten in place."""
        d = self._desc
        if n_bvh_nodes is not None:
            d.n_bvh_nodes = n_bvh_nodes
        self._ctx.check(self._ctx.lib.bvh_cuda_scene_refresh_dev(self._ctx.h, self.h, C.byref(d), stream))

    def close(self):
        if getattr(self, "h", None) and self._ctx.h:
            self._ctx.lib.bvh_cuda_scene_free(self._ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _rays(ray_o, ray_d):
        o = np.ascontiguousarray(ray_o, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(ray_d, dtype=np.float32).reshape(-1, 3)
        return o, d

    def instance_boxes(self, enable: bool = True, stream: int = 0):
        """(Re)compute the tight per-instance world boxes the exact-order kernels cull with (wrapped scenes: call again after
        every change of the instance buffer; uploaded scenes have them from the start)."""
        ctx = self._ctx
        ctx.check(ctx.lib.bvh_cuda_scene_instance_boxes_dev(ctx.h, self.h, 1 if enable else 0, stream))

    def traverse_tlas(self, ray_o, ray_d, tmax: float = 1e30):
        """Closest hit.  Returns (t, tri, inst)."""
        ctx = self._ctx
        o, d = self._rays(ray_o, ray_d)
        r = o.shape[0]
        t = np.empty(r, dtype=np.float32)
        tri = np.empty(r, dtype=np.uint32)
        inst = np.empty(r, dtype=np.uint32)
        ctx.check(ctx.lib.bvh_cuda_trace_closest(ctx.h, self.h, _vp(o), _vp(d), r, tmax, _vp(t), _vp(tri), _vp(inst)))
        return t, tri, inst

    def occluded(self, ray_o, ray_d, tmax: float = 1e30):
        """Shadow rays: traverse_tlas(ray).hit with early exit (raytraced_shadows.wgsl:98-102)."""
        ctx = self._ctx
        o, d = self._rays(ray_o, ray_d)
        occ = np.empty(o.shape[0], dtype=np.uint8)
        ctx.check(ctx.lib.bvh_cuda_trace_any(ctx.h, self.h, _vp(o), _vp(d), o.shape[0], tmax, _vp(occ)))
        return occ

    def traverse_tlas_dev(self, d_ro: int, d_rd: int, n_rays: int, d_t: int, d_tri: int, d_inst: int,
                          tmax: float = 1e30, stream: int = 0):
        ctx = self._ctx
        ctx.check(ctx.lib.bvh_cuda_trace_closest_dev(ctx.h, self.h, d_ro, d_rd, n_rays, tmax, d_t, d_tri, d_inst, stream))

    def occluded_dev(self, d_ro: int, d_rd: int, n_rays: int, d_occ: int, tmax: float = 1e30, stream: int = 0):
        ctx = self._ctx
        ctx.check(ctx.lib.bvh_cuda_trace_any_dev(ctx.h, self.h, d_ro, d_rd, n_rays, tmax, d_occ, stream))
