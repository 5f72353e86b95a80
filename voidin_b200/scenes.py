"""Deterministic synthetic inputs for the BASELINE.json configs (meshes, instances, rays).

Everything here is plain numpy on the host: the arrays are *inputs* handed unchanged to both the CUDA
path and the CPU oracle, so nothing in this file has to be bit-reproduced on the device.

Reference shapes mirrored (paths relative to the voidin checkout):
  make_plane_mesh   crates/pools/src/mesh/plane.rs:5-38
  make_uv_sphere    crates/pools/src/mesh/sphere.rs:6-67   (same vertex/triangle order and counts)
  soup              src/bin/bvh_cpu.rs:42-50               (random unshared triangles)
  MeshPool.add      crates/pools/src/mesh/mod.rs:309-351   (pooled offsets -> MeshInfo)
  Instance.new      crates/components/src/shared.rs:90-98
  AreaLight         crates/pools/src/light.rs:28-52
"""
from __future__ import annotations

import numpy as np

from .types import INSTANCE, MESH_INFO

F32 = np.float32


# ----------------------------------------------------------------------------------------------
# meshes
# ----------------------------------------------------------------------------------------------
def make_plane_mesh(width: float = 1.0, height: float = 1.0):
    w, h = F32(width) / F32(2), F32(height) / F32(2)
    v = np.array([[-w, 0, -h], [-w, 0, h], [w, 0, h], [w, 0, -h]], dtype=F32)
    idx = np.array([0, 1, 2, 0, 2, 3], dtype=np.uint32)
    return v, idx


def _uv_sphere_indices(vside: int, uside: int) -> np.ndarray:
    i = np.arange(vside, dtype=np.int64)[:, None]
    j = np.arange(uside, dtype=np.int64)[None, :]
    k1 = i * (uside + 1) + j
    k2 = k1 + uside + 1
    t1 = np.stack([k1, k2, k1 + 1], axis=-1)  # emitted when i != 0
    t2 = np.stack([k1 + 1, k2, k2 + 1], axis=-1)  # always emitted (i != stack_count is always true)
    both = np.stack([t1, t2], axis=2)  # [vside, uside, 2, 3]
    rows = [both[0, :, 1, :].reshape(-1, 3)]
    if vside > 1:
        rows.append(both[1:].reshape(-1, 3))
    return np.concatenate(rows, axis=0).astype(np.uint32).reshape(-1)


def make_uv_sphere(radius: float = 1.0, resolution: int = 1):
    vside = 4 * resolution
    uside = 2 * vside
    v = (np.arange(vside + 1, dtype=F32) / F32(vside))[:, None]
    u = (np.arange(uside + 1, dtype=F32) / F32(uside))[None, :]
    pi = F32(np.pi)
    theta = F32(2) * pi * u + pi
    phi = pi * v
    r = F32(radius)
    x = np.cos(theta, dtype=F32) * np.sin(phi, dtype=F32) * r
    y = np.broadcast_to(-np.cos(phi, dtype=F32) * r, x.shape)
    z = np.sin(theta, dtype=F32) * np.sin(phi, dtype=F32) * r
    verts = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(F32)
    return np.ascontiguousarray(verts), _uv_sphere_indices(vside, uside)


def displaced_sphere(vside: int, uside: int, seed: int, amplitude: float = 0.25, radius: float = 1.0):
    """UV sphere whose radius is modulated by a fixed set of random low/medium-frequency lobes plus a little
    per-vertex jitter: a closed, bumpy, non-symmetric surface used as the stand-in for the Stanford meshes
    that are absent from the reference checkout (.MISSING_LARGE_BLOBS).  2*uside*vside - uside triangles."""
    rng = np.random.default_rng(seed)
    v = (np.arange(vside + 1, dtype=np.float64) / vside)[:, None]
    u = (np.arange(uside + 1, dtype=np.float64) / uside)[None, :]
    theta = 2 * np.pi * u + np.pi
    phi = np.pi * v
    d = np.stack([np.cos(theta) * np.sin(phi), np.broadcast_to(-np.cos(phi), (vside + 1, uside + 1)),
                  np.sin(theta) * np.sin(phi)], axis=-1)
    rad = np.ones((vside + 1, uside + 1))
    for k in range(12):
        w = rng.normal(size=3)
        w /= np.linalg.norm(w)
        freq = rng.uniform(1.5, 9.0)
        ph = rng.uniform(0, 2 * np.pi)
        rad += (amplitude / (1.5 + k * 0.5)) * np.sin(freq * (d @ w) * np.pi + ph)
    rad += rng.uniform(-1.0, 1.0, size=rad.shape) * (0.15 * np.pi / vside)
    # anisotropic stretch so the three axes have different extents
    stretch = np.array([1.0, 0.7, 1.35])
    verts = (d * rad[..., None] * radius * stretch).reshape(-1, 3).astype(F32)
    return np.ascontiguousarray(verts), _uv_sphere_indices(vside, uside)


def bunny_class(seed: int = 1):
    """~69 K triangles (Stanford bunny: 69 451)."""
    return displaced_sphere(132, 264, seed)


def dragon_class(seed: int = 2):
    """~871 K triangles (Stanford dragon: 871 414)."""
    return displaced_sphere(467, 934, seed)


def soup(n_tris: int, seed: int = 4, edge: float = 0.005):
    """n unshared random triangles: v0 uniform in [0,1)^3, v1/v2 = v0 + uniform[-edge,edge)^3."""
    rng = np.random.default_rng(seed)
    v0 = rng.random((n_tris, 1, 3), dtype=F32)
    e = (rng.random((n_tris, 2, 3), dtype=F32) * F32(2) - F32(1)) * F32(edge)
    verts = np.concatenate([v0, v0 + e], axis=1).reshape(-1, 3).astype(F32)
    idx = np.arange(3 * n_tris, dtype=np.uint32)
    return np.ascontiguousarray(verts), idx


def grid_mesh(nx: int, ny: int):
    """Flat regular grid in the xz plane: many exactly equal coordinates (exercises the strict-< tie rules)
    and a zero-extent axis."""
    xs, zs = np.meshgrid(np.arange(nx + 1, dtype=F32), np.arange(ny + 1, dtype=F32), indexing="ij")
    verts = np.stack([xs, np.zeros_like(xs), zs], axis=-1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    a = (i * (ny + 1) + j).reshape(-1)
    b = a + 1
    c = a + (ny + 1)
    d = c + 1
    idx = np.stack([a, b, c, b, d, c], axis=-1).reshape(-1).astype(np.uint32)
    return np.ascontiguousarray(verts.astype(F32)), idx


def load_obj_positions(path: str):
    """Minimal OBJ reader (positions only, fan triangulation) for assets/cube/cube.obj-style files."""
    pos, tris = [], []
    with open(path) as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                pos.append([float(x) for x in p[1:4]])
            elif p[0] == "f":
                ids = [int(tok.split("/")[0]) for tok in p[1:]]
                ids = [i - 1 if i > 0 else len(pos) + i for i in ids]
                for k in range(1, len(ids) - 1):
                    tris.append([ids[0], ids[k], ids[k + 1]])
    return np.asarray(pos, dtype=F32), np.asarray(tris, dtype=np.uint32).reshape(-1)


# ----------------------------------------------------------------------------------------------
# transforms / instances
# ----------------------------------------------------------------------------------------------
def mat_translation(t):
    m = np.eye(4)
    m[:3, 3] = t
    return m


def mat_scale(s):
    s = np.broadcast_to(np.asarray(s, dtype=np.float64), (3,))
    return np.diag([s[0], s[1], s[2], 1.0])


def mat_rotation_x(a):
    c, s = np.cos(a), np.sin(a)
    m = np.eye(4)
    m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, -s, s, c
    return m


def mat_from_quat(q):
    x, y, z, w = q
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 0],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w), 0],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y), 0],
            [0, 0, 0, 1],
        ]
    )


def make_instances(transforms, mesh_ids, material: int = 1) -> np.ndarray:
    """Instance::new (shared.rs:90-98).  `transforms`: [I,4,4] row-major math matrices (float64).  Stored
    column-major as f32; inv_transform is the float64 inverse rounded to f32 and is an *input* to both
    implementations (the reference computes it once with glam's Mat4::inverse at construction)."""
    t = np.asarray(transforms, dtype=np.float64).reshape(-1, 4, 4)
    t32 = t.astype(F32)
    inv = np.linalg.inv(t32.astype(np.float64)).astype(F32)
    out = np.zeros(t.shape[0], dtype=INSTANCE)
    out["transform"] = t32.transpose(0, 2, 1).reshape(-1, 16)
    out["inv_transform"] = inv.transpose(0, 2, 1).reshape(-1, 16)
    out["mesh"] = np.asarray(mesh_ids, dtype=np.uint32)
    out["material"] = material
    return out


def random_instances(n: int, n_meshes: int, seed: int = 3, extent: float = 500.0):
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    pos = rng.uniform(-extent, extent, size=(n, 3))
    sc = rng.uniform(0.5, 2.0, size=n)
    mesh = rng.integers(0, n_meshes, size=n)
    mats = np.empty((n, 4, 4))
    for i in range(n):
        mats[i] = mat_translation(pos[i]) @ mat_from_quat(q[i]) @ mat_scale(sc[i])
    return make_instances(mats, mesh)


class MeshPool:
    """Host mirror of the pooling bookkeeping in MeshPool::add (crates/pools/src/mesh/mod.rs:309-351): pooled
    vertices / permuted indices / BVH nodes plus one MeshInfo per mesh.  `builder(vertices, indices)` must
    return (nodes, permuted_indices) — the CUDA builder in production, the oracle in tests."""

    def __init__(self, builder):
        self._builder = builder
        self.vertices, self.indices, self.bvh_nodes, self.mesh_info = [], [], [], []
        self.vertex_offset = self.base_index = self.bvh_index = 0

    def add(self, vertices: np.ndarray, indices: np.ndarray) -> int:
        vertices = np.ascontiguousarray(vertices, dtype=F32).reshape(-1, 3)
        nodes, perm = self._builder(vertices, np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1))
        return self.add_built(vertices, perm, nodes)

    def add_built(self, vertices, perm_indices, nodes) -> int:
        info = np.zeros((), dtype=MESH_INFO)
        info["min"] = vertices.min(axis=0)  # over all positions, referenced or not (mesh/mod.rs:22-27,333)
        info["max"] = vertices.max(axis=0)
        info["vertex_offset"] = self.vertex_offset
        info["base_index"] = self.base_index
        info["index_count"] = perm_indices.size
        info["bvh_index"] = self.bvh_index
        self.vertices.append(vertices)
        self.indices.append(perm_indices)
        self.bvh_nodes.append(nodes)
        self.mesh_info.append(info)
        self.vertex_offset += vertices.shape[0]
        self.base_index += perm_indices.size
        self.bvh_index += nodes.shape[0]
        return len(self.mesh_info) - 1

    def pooled(self):
        return (
            np.ascontiguousarray(np.concatenate(self.vertices, axis=0)),
            np.ascontiguousarray(np.concatenate(self.indices)),
            np.ascontiguousarray(np.concatenate(self.bvh_nodes)),
            np.ascontiguousarray(np.stack(self.mesh_info)),
        )


# ----------------------------------------------------------------------------------------------
# lights and rays
# ----------------------------------------------------------------------------------------------
def rect_light_corners(wh=(5.0, 8.0), translation=(0.0, 10.0, 15.0), rot_x=-np.pi / 4):
    """AreaLight::from_transform(wh, T(translation) * Rx(rot_x)) -> 4 corners (light.rs:28-52; the light of
    src/bin/model.rs:72-77)."""
    rot = mat_rotation_x(rot_x)[:3, :3]
    d = rot @ np.array([0.0, 0.0, 1.0])
    d /= np.linalg.norm(d)
    up = np.array([0.0, 1.0, 0.0])
    dirx = np.cross(up, d)
    diry = np.cross(d, dirx)
    dx = dirx * wh[0] / 2
    dy = diry * wh[1] / 2
    t = np.asarray(translation, dtype=np.float64)
    return np.stack([t - dx - dy, t + dx - dy, t + dx + dy, t - dx + dy])


def rays_toward_box(n: int, bmin, bmax, seed: int = 11, radius_scale: float = 2.0):
    """Config 1: origin uniform on a sphere of radius 2*|diag| about the box centre, unit direction toward a
    uniform point inside the box."""
    rng = np.random.default_rng(seed)
    bmin, bmax = np.asarray(bmin, np.float64), np.asarray(bmax, np.float64)
    c = 0.5 * (bmin + bmax)
    r = radius_scale * np.linalg.norm(bmax - bmin)
    o = rng.normal(size=(n, 3))
    o = c + r * o / np.linalg.norm(o, axis=1, keepdims=True)
    tgt = bmin + rng.random((n, 3)) * (bmax - bmin)
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.ascontiguousarray(o.astype(F32)), np.ascontiguousarray(d.astype(F32))


def rays_sphere_to_cube(n: int, radius: float = 1200.0, extent: float = 500.0, seed: int = 13):
    """Config 3: primary-ray-like; origin uniform on a sphere, unit direction toward a uniform point in
    [-extent,extent]^3."""
    rng = np.random.default_rng(seed)
    o = rng.normal(size=(n, 3))
    o = radius * o / np.linalg.norm(o, axis=1, keepdims=True)
    tgt = rng.uniform(-extent, extent, size=(n, 3))
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.ascontiguousarray(o.astype(F32)), np.ascontiguousarray(d.astype(F32))


def shadow_rays(n: int, world_tris: np.ndarray, light_corners: np.ndarray, seed: int = 12, chunk: int = 1 << 22):
    """Config 2 / 5: origin = uniform surface point (triangle picked by area, uniform barycentrics) pushed
    1e-4 along the geometric normal (raytraced_shadows.wgsl:98); direction = (uniform point on the rect
    light) - origin, NOT normalised, no t-max.  `world_tris`: [T,3,3] float world-space triangles."""
    rng = np.random.default_rng(seed)
    tris = np.asarray(world_tris, dtype=np.float64)
    e1 = tris[:, 1] - tris[:, 0]
    e2 = tris[:, 2] - tris[:, 0]
    nrm = np.cross(e1, e2)
    area = 0.5 * np.linalg.norm(nrm, axis=1)
    cdf = np.cumsum(area)
    cdf /= cdf[-1]
    unit = nrm / np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-300)
    lc = np.asarray(light_corners, dtype=np.float64)
    o_out = np.empty((n, 3), dtype=F32)
    d_out = np.empty((n, 3), dtype=F32)
    for b in range(0, n, chunk):
        m = min(chunk, n - b)
        t = np.minimum(np.searchsorted(cdf, rng.random(m)), len(cdf) - 1)
        r1 = np.sqrt(rng.random(m))
        r2 = rng.random(m)
        p = tris[t, 0] + e1[t] * (r1 * (1 - r2))[:, None] + e2[t] * (r1 * r2)[:, None]
        o = p + 1e-4 * unit[t]
        a, bb = rng.random(m), rng.random(m)
        tgt = lc[0] + (lc[1] - lc[0]) * a[:, None] + (lc[3] - lc[0]) * bb[:, None]
        o32 = o.astype(F32)
        o_out[b : b + m] = o32
        d_out[b : b + m] = (tgt - o32.astype(np.float64)).astype(F32)
    return o_out, d_out


def world_triangles(vertices, indices, transform=None) -> np.ndarray:
    v = np.asarray(vertices, dtype=np.float64)
    if transform is not None:
        t = np.asarray(transform, dtype=np.float64)
        v = v @ t[:3, :3].T + t[:3, 3]
    return v[np.asarray(indices, dtype=np.int64).reshape(-1, 3)]


def dragon_scene_instances(lift: float = 0.9):
    """Config 2 scene: mesh 0 = ground plane scaled x200 (raytraced_shadows.rs:36-40), mesh 1 = the dragon-class
    mesh lifted so that it rests on the ground.  Returns row-major transforms and mesh ids for make_instances."""
    return np.stack([mat_scale(200.0), mat_translation([0.0, lift, 0.0])]), np.array([0, 1])


def gbuffer_shadow_rays(n: int, dragon_world_tris: np.ndarray, light_corners: np.ndarray, seed: int = 12,
                        ground_radius: float = 6.0, ground_frac: float = 0.5, coherent: bool = True):
    """Config 2 rays as a G-buffer would produce them (raytraced_shadows.wgsl:73-99 runs once per pixel and shades
    what the camera sees: the model and the ground around it).  `1-ground_frac` of the origins are area-uniform
    points on the model surface emitted in surface-raster order (stratified along the mesh's own triangle order,
    which follows its (u,v) parametrisation, so consecutive rays start on neighbouring triangles like neighbouring
    pixels do); the rest are the pixels of a square raster of the ground plane (y = 0, normal +y) of half-width
    `ground_radius` around the model, row-major.  Origin is pushed 1e-4 along the geometric normal; direction =
    (uniform random point on the rect light) - origin, NOT normalised, no t-max.
    coherent=False applies one fixed random permutation (worst case for SIMT traversal; reported separately).
    Area-weighted sampling over the whole x200 ground quad (40 000 area units against ~12 for the model) would send
    99.97 % of the rays from empty ground and make the traversal trivial."""
    rng = np.random.default_rng(seed)
    n_ground = int(n * ground_frac)
    n_model = n - n_ground
    tris = np.asarray(dragon_world_tris, dtype=np.float64)
    e1 = tris[:, 1] - tris[:, 0]
    e2 = tris[:, 2] - tris[:, 0]
    nrm = np.cross(e1, e2)
    area = 0.5 * np.linalg.norm(nrm, axis=1)
    cdf = np.cumsum(area)
    cdf /= cdf[-1]
    unit = nrm / np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-300)
    lc = np.asarray(light_corners, dtype=np.float64)
    o = np.empty((n, 3), dtype=F32)
    d = np.empty((n, 3), dtype=F32)
    chunk = 1 << 22
    for b0 in range(0, n_model, chunk):
        m = min(chunk, n_model - b0)
        strat = (np.arange(b0, b0 + m) + rng.random(m)) / n_model  # stratified: monotone in ray index
        t = np.minimum(np.searchsorted(cdf, strat), len(cdf) - 1)
        r1 = np.sqrt(rng.random(m))
        r2 = rng.random(m)
        p = tris[t, 0] + e1[t] * (r1 * (1 - r2))[:, None] + e2[t] * (r1 * r2)[:, None]
        o32 = (p + 1e-4 * unit[t]).astype(F32)
        tgt = lc[0] + (lc[1] - lc[0]) * rng.random(m)[:, None] + (lc[3] - lc[0]) * rng.random(m)[:, None]
        o[b0:b0 + m] = o32
        d[b0:b0 + m] = (tgt - o32.astype(np.float64)).astype(F32)
    side = int(np.ceil(np.sqrt(max(n_ground, 1))))
    for b0 in range(0, n_ground, chunk):
        m = min(chunk, n_ground - b0)
        k = np.arange(b0, b0 + m)
        px = (k % side + rng.random(m)) / side
        pz = (k // side + rng.random(m)) / side
        p = np.stack([(2 * px - 1) * ground_radius, np.full(m, 1e-4), (2 * pz - 1) * ground_radius], axis=1)
        o32 = p.astype(F32)
        tgt = lc[0] + (lc[1] - lc[0]) * rng.random(m)[:, None] + (lc[3] - lc[0]) * rng.random(m)[:, None]
        o[n_model + b0:n_model + b0 + m] = o32
        d[n_model + b0:n_model + b0 + m] = (tgt - o32.astype(np.float64)).astype(F32)
    if not coherent:
        perm = np.random.default_rng(seed + 2000).permutation(n)
        o, d = o[perm], d[perm]
    return np.ascontiguousarray(o), np.ascontiguousarray(d)


# ----------------------------------------------------------------------------------------------
# minimal glTF 2.0 reader (positions + indices of every mesh primitive; no node transforms), the subset of
# crates/app/src/models/gltf_model/mod.rs:103-155 that feeds MeshPool::add
# ----------------------------------------------------------------------------------------------
def load_gltf_primitives(path: str):
    """Returns [(vertices [V,3] f32, indices [3N] u32), ...] for a .glb or a .gltf with external .bin buffers."""
    import json
    import os
    import struct

    with open(path, "rb") as f:
        raw = f.read()
    if raw[:4] == b"glTF":
        length = struct.unpack_from("<I", raw, 8)[0]
        off, doc, bins = 12, None, []
        while off < length:
            clen, ctype = struct.unpack_from("<II", raw, off)
            chunk = raw[off + 8: off + 8 + clen]
            if ctype == 0x4E4F534A:
                doc = json.loads(chunk.decode("utf-8"))
            elif ctype == 0x004E4942:
                bins.append(chunk)
            off += 8 + clen
        buffers = bins
    else:
        doc = json.loads(raw.decode("utf-8"))
        base = os.path.dirname(path)
        buffers = [open(os.path.join(base, b["uri"]), "rb").read() for b in doc["buffers"]]
    ctype_np = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
    ncomp = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4}

    def accessor(i):
        a = doc["accessors"][i]
        bv = doc["bufferViews"][a["bufferView"]]
        dt = np.dtype(ctype_np[a["componentType"]])
        n = ncomp[a["type"]]
        start = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
        stride = bv.get("byteStride", 0) or dt.itemsize * n
        buf = buffers[bv["buffer"]]
        if stride == dt.itemsize * n:
            return np.frombuffer(buf, dtype=dt, count=a["count"] * n, offset=start).reshape(a["count"], n)
        rows = np.lib.stride_tricks.as_strided(np.frombuffer(buf, dtype=np.uint8, offset=start),
                                               shape=(a["count"], dt.itemsize * n), strides=(stride, 1))
        return np.ascontiguousarray(rows).view(dt).reshape(a["count"], n)

    out = []
    for mesh in doc.get("meshes", []):
        for prim in mesh["primitives"]:
            if prim.get("mode", 4) != 4 or "indices" not in prim:
                continue
            v = np.ascontiguousarray(accessor(prim["attributes"]["POSITION"]), dtype=F32)
            idx = np.ascontiguousarray(accessor(prim["indices"]).reshape(-1).astype(np.uint32))
            out.append((v, idx[: idx.size // 3 * 3]))
    return out


# ----------------------------------------------------------------------------------------------
# BASELINE config 3 inputs (shared by bench.py, the GPU parity tests and tests/golden/make_golden_large.py)
# ----------------------------------------------------------------------------------------------
def config3_meshes():
    """bunny-class, dragon-class and a DamagedHelmet-sized stand-in for ferris3d (the real assets are absent from the
    reference checkout, .MISSING_LARGE_BLOBS)."""
    return [bunny_class(), dragon_class(), displaced_sphere(62, 124, 3)]


def pooled_mesh_info(meshes) -> np.ndarray:
    """MeshInfo of `meshes` pooled in order, as MeshPool::add lays them out (crates/pools/src/mesh/mod.rs:322-347);
    bvh_index is left 0 (the forest build fills it in)."""
    info = np.zeros(len(meshes), dtype=MESH_INFO)
    info["index_count"] = [i.size for _, i in meshes]
    info["vertex_offset"] = np.concatenate([[0], np.cumsum([v.shape[0] for v, _ in meshes])[:-1]])
    info["base_index"] = np.concatenate([[0], np.cumsum(info["index_count"])[:-1]])
    info["min"] = [v.min(0) for v, _ in meshes]
    info["max"] = [v.max(0) for v, _ in meshes]
    return info


def config3_scene_inputs(n_inst: int, meshes=None):
    """(instances, mesh_info) of config 3: `n_inst` instances, mesh uniform, T(uniform [-500,500]^3) R(uniform
    quaternion) S(uniform [0.5,2]), seed 3."""
    meshes = config3_meshes() if meshes is None else meshes
    return random_instances(n_inst, len(meshes), seed=3, extent=500.0), pooled_mesh_info(meshes)
