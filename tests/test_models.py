"""Asset loaders (include/bvh_cuda_models.h, csrc/models.cpp) against independent Python readers of the same files.

Reference behaviour restated: ObjModel::import = tobj::load_obj(path, GPU_LOAD_OPTIONS) (crates/app/src/models/mod.rs:20-57),
GltfDocument::make_meshes / get_scene_instances (crates/app/src/models/gltf_model/mod.rs:103-207).
No GPU needed.  The files of the reference checkout are used when /root/reference exists (this container); everywhere else
the synthetic files written by the tests themselves and the committed fixture tests/golden/real_meshes.npz +
tests/golden/models_golden.json (made by tests/golden/make_models_golden.py) are what is compared.
"""
import base64
import hashlib
import json
import os
import struct
import urllib.parse

import numpy as np
import pytest

from voidin_b200 import models as M
from voidin_b200 import scenes as S
from voidin_b200._lib import BvhCudaError

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ASSETS = "/root/reference/assets"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF_ASSETS), reason="reference checkout not present")


# ---- independent restatement of tobj's single-index export (python dicts, line by line) ------------------------------
def py_tobj(path):
    pos, tex, nrm = [], [], []
    models, faces, name, mat = [], [], "unnamed_object", None
    mats = {}

    def export():
        nonlocal faces
        vmap, P, T, N, I = {}, [], [], [], []
        for f in faces:
            if len(f) < 3:
                continue
            for k in range(2, len(f)):
                for c in (f[0], f[k - 1], f[k]):
                    if c not in vmap:
                        vmap[c] = len(vmap)
                        P.append(pos[c[0]])
                        if tex and c[1] is not None:
                            T.append(tex[c[1]][:2])
                        if nrm and c[2] is not None:
                            N.append(nrm[c[2]])
                    I.append(vmap[c])
        models.append(dict(name=name, material=mat, positions=np.array(P, np.float32).reshape(-1, 3),
                           texcoords=np.array(T, np.float32).reshape(-1, 2), normals=np.array(N, np.float32).reshape(-1, 3),
                           indices=np.array(I, np.uint32)))
        faces = []

    for line in open(path):
        p = line.split()
        if not p or p[0].startswith("#"):
            continue
        if p[0] == "v":
            pos.append([np.float32(x) for x in p[1:4]])
        elif p[0] == "vt":
            tex.append([np.float32(x) for x in p[1:3]])
        elif p[0] == "vn":
            nrm.append([np.float32(x) for x in p[1:4]])
        elif p[0] in ("f", "l", "p"):
            face = []
            for tok in p[1:]:
                parts = tok.split("/") + ["", ""]
                sizes = (len(pos), len(tex), len(nrm))
                idx = []
                for s, n in zip(parts[:3], sizes):
                    idx.append(None if s == "" else (n + int(s) if int(s) < 0 else int(s) - 1))
                face.append(tuple(idx))
            faces.append(face)
        elif p[0] in ("o", "g"):
            if faces:
                export()
            name = line.strip()[1:].strip() or "unnamed_object"
        elif p[0] == "mtllib":
            mp = os.path.join(os.path.dirname(path), p[1])
            if os.path.exists(mp):
                for ml in open(mp):
                    q = ml.split()
                    if q and q[0] == "newmtl":
                        mats[ml.strip()[6:].strip()] = len(mats)
        elif p[0] == "usemtl":
            new = mats.get(line.strip()[6:].strip())
            if new != mat and faces:
                export()
            mat = new
    if faces:
        export()
    return models


def check_obj(path):
    got = M.ObjModel.import_(path)
    want = py_tobj(path)
    assert len(got.meshes) == len(want)
    for g, w in zip(got.meshes, want):
        assert g.name == w["name"]
        assert g.material == (-1 if w["material"] is None else w["material"])
        assert g.vertices.tobytes() == w["positions"].tobytes()
        assert g.normals.tobytes() == w["normals"].tobytes()
        assert g.tex_coords.tobytes() == w["texcoords"].tobytes()
        assert (g.indices == w["indices"]).all()
        assert g.tangents.shape == (g.vertices.shape[0], 4) and not g.tangents.any()
    return got


def test_obj_single_index_triangulation_negative_indices_groups(tmp_path):
    (tmp_path / "m.mtl").write_text("newmtl red\nKd 1 0 0.25\nnewmtl blue stone\nKd 0 0 1\n")
    (tmp_path / "a.obj").write_text("""# quad + pentagon + shared corners with different normals, relative indices, points and lines
mtllib m.mtl
o first
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
vt 0 0
vt 1 0
vt 1 1
vn 0 0 1
vn 0 1 0
usemtl red
f 1/1/1 2/2/1 3/3/1 4/1/1
f 1/1/2 2/2/1 3/3/1
p 1
l 1 2
v 0.5 1.5 1e-3
f -5//1 -4//1 -3//1 -2//1 -1//1
usemtl blue stone
f 1 2 3
g second group
v 2 0 0
v 3 0 0
v 3 1 0
f 6 7 8
f 8/3 7/2 6/1
""")
    got = check_obj(str(tmp_path / "a.obj"))
    assert [m.name for m in got.meshes] == ["first", "first", "second group"]
    assert [m.material for m in got.meshes] == [0, 1, 1]
    # quad -> 2 triangles, triangle -> 1, pentagon -> 3 (fan 0,i-1,i)
    assert got.meshes[0].indices.size == 3 * 6
    assert got.materials[0]["base_color"].tolist() == [1.0, 0.0, 0.25, 0.5]
    assert got.materials[1]["name"] == "blue stone"


def test_obj_errors(tmp_path):
    with pytest.raises(BvhCudaError):
        M.ObjModel.import_(str(tmp_path / "missing.obj"))
    (tmp_path / "bad.obj").write_text("v 0 0 0\nv 1 0 0\nf 1 2 9\n")
    with pytest.raises(BvhCudaError):
        M.ObjModel.import_(str(tmp_path / "bad.obj"))


def test_obj_random_soup_roundtrip(tmp_path):
    """A larger file: written with repr-precision floats, read back bit-exactly; positions-only faces keep the
    vertex numbering of first use."""
    rng = np.random.default_rng(3)
    v = rng.standard_normal((5000, 3)).astype(np.float32)
    f = rng.integers(0, 5000, size=(12000, 3))
    with open(tmp_path / "soup.obj", "w") as fh:
        for p in v:
            fh.write("v %s %s %s\n" % tuple(repr(float(x)) for x in p))
        for t in f:
            fh.write("f %d %d %d\n" % tuple(int(x) + 1 for x in t))
    got = check_obj(str(tmp_path / "soup.obj"))
    m = got.meshes[0]
    assert m.vertices[m.indices].tobytes() == v[f.reshape(-1)].tobytes()


# ---- glTF --------------------------------------------------------------------------------------------------------------
def _quat_matrix(t, r, s):
    """gltf crate Transform::Decomposed -> matrix(), in float32 with the same operation order (numpy scalars)."""
    f = np.float32
    x, y, z, w = (f(a) for a in r)
    x2, y2, z2 = x + x, y + y, z + z
    xx2, xy2, xz2, yy2, yz2, zz2 = x2 * x, x2 * y, x2 * z, y2 * y, y2 * z, z2 * z
    sy2, sz2, sx2 = y2 * w, z2 * w, x2 * w
    one = f(1)
    R = np.array([[one - yy2 - zz2, xy2 + sz2, xz2 - sy2, 0], [xy2 - sz2, one - xx2 - zz2, yz2 + sx2, 0],
                  [xz2 + sy2, yz2 - sx2, one - xx2 - yy2, 0], [0, 0, 0, 1]], dtype=f)  # rows here = columns
    T = np.eye(4, dtype=f)
    T[3, :3] = [f(a) for a in t]
    Sm = np.diag([f(s[0]), f(s[1]), f(s[2]), f(1)]).astype(f)
    return _mul_cols(_mul_cols(T, R), Sm)


def _mul_cols(A, B):
    """A, B given as arrays of columns; returns columns of A*B with ((a*x + b*y) + c*z) + d*w in float32."""
    out = np.zeros((4, 4), np.float32)
    for c in range(4):
        acc = A[0] * B[c, 0]
        acc = acc + A[1] * B[c, 1]
        acc = acc + A[2] * B[c, 2]
        acc = acc + A[3] * B[c, 3]
        out[c] = acc
    return out


def py_gltf(path):
    """Independent reader: make_meshes + get_scene_instances semantics (POSITION/NORMAL contiguous, children first)."""
    raw = open(path, "rb").read()
    glb_bin = None
    if raw[:4] == b"glTF":
        off, doc = 12, None
        while off < len(raw):
            clen, ctype = struct.unpack_from("<II", raw, off)
            chunk = raw[off + 8: off + 8 + clen]
            if ctype == 0x4E4F534A:
                doc = json.loads(chunk.decode())
            elif ctype == 0x004E4942 and glb_bin is None:
                glb_bin = chunk
            off += 8 + clen
    else:
        doc = json.loads(raw.decode())
    bufs = []
    for b in doc.get("buffers", []):
        if "uri" not in b:
            bufs.append(glb_bin)
        elif b["uri"].startswith("data:"):
            bufs.append(base64.b64decode(b["uri"].split(",", 1)[1]))
        else:
            bufs.append(open(os.path.join(os.path.dirname(path), urllib.parse.unquote(b["uri"])), "rb").read())
    npdt = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
    ncomp = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4}

    def start(a):
        bv = doc["bufferViews"][a["bufferView"]]
        return bufs[bv["buffer"]], bv.get("byteOffset", 0) + a.get("byteOffset", 0), bv.get("byteStride", 0)

    def contiguous(i):
        a = doc["accessors"][i]
        buf, st, _ = start(a)
        n = a["count"] * ncomp[a["type"]]
        return np.frombuffer(buf, npdt[a["componentType"]], n, st).reshape(a["count"], -1)

    def strided(i):
        a = doc["accessors"][i]
        buf, st, stride = start(a)
        dt = np.dtype(npdt[a["componentType"]])
        w = dt.itemsize * ncomp[a["type"]]
        stride = stride or w
        rows = [np.frombuffer(buf, dt, ncomp[a["type"]], st + k * stride) for k in range(a["count"])]
        return np.stack(rows) if rows else np.zeros((0, ncomp[a["type"]]), dt)

    meshes, key = [], {}
    for mi, mesh in enumerate(doc.get("meshes", [])):
        for pi, prim in enumerate(mesh["primitives"]):
            at = prim["attributes"]
            if "POSITION" not in at or "NORMAL" not in at:
                continue
            v = contiguous(at["POSITION"]).astype(np.float32)
            n = contiguous(at["NORMAL"]).astype(np.float32)
            nv = v.shape[0]
            uv = np.zeros((nv, 2), np.float32)
            if "TEXCOORD_0" in at:
                t = strided(at["TEXCOORD_0"])
                t = t.astype(np.float32) / np.float32({np.dtype(np.uint8): 255, np.dtype(np.uint16): 65535}.get(t.dtype, 1))
                uv[: min(nv, len(t))] = t[:nv]
            tan = np.tile(np.array([0, 1, 0, 1], np.float32), (nv, 1))
            if "TANGENT" in at:
                t = strided(at["TANGENT"]).astype(np.float32)
                tan[: min(nv, len(t))] = t[:nv]
            idx = strided(prim["indices"]).reshape(-1).astype(np.uint32) if "indices" in prim else np.arange(nv, dtype=np.uint32)
            key[(mi, pi)] = len(meshes)
            meshes.append(dict(v=v, n=n, uv=uv, tan=tan, idx=idx, material=prim.get("material", -1), name=mesh.get("name", "")))
    inst = []

    def walk(ni, parent):
        node = doc["nodes"][ni]
        if "matrix" in node:
            local = np.array(node["matrix"], dtype=np.float64).astype(np.float32).reshape(4, 4)  # rows = columns
        else:
            local = _quat_matrix(np.array(node.get("translation", [0, 0, 0]), np.float64).astype(np.float32),
                                 np.array(node.get("rotation", [0, 0, 0, 1]), np.float64).astype(np.float32),
                                 np.array(node.get("scale", [1, 1, 1]), np.float64).astype(np.float32))
        world = _mul_cols(parent, local)
        for c in node.get("children", []):
            walk(c, world)
        if "mesh" in node:
            for pi, prim in enumerate(doc["meshes"][node["mesh"]]["primitives"]):
                if (node["mesh"], pi) in key:
                    inst.append((world.copy(), key[(node["mesh"], pi)], prim.get("material", -1)))

    for sc in doc.get("scenes", []):
        for r in sc.get("nodes", []):
            walk(r, np.eye(4, dtype=np.float32))
    return meshes, inst


def check_gltf(path):
    got = M.GltfDocument.import_(path)
    meshes, inst = py_gltf(path)
    assert len(got.meshes) == len(meshes)
    for g, w in zip(got.meshes, meshes):
        assert g.vertices.tobytes() == w["v"].tobytes()
        assert g.normals.tobytes() == w["n"].tobytes()
        assert g.tex_coords.tobytes() == w["uv"].tobytes()
        assert g.tangents.tobytes() == w["tan"].tobytes()
        assert (g.indices == w["idx"]).all()
        assert g.material == w["material"] and g.name == w["name"]
    assert len(got.instances) == len(inst)
    for gi, (world, mesh, mat) in zip(got.instances, inst):
        assert gi.mesh == mesh and gi.material == mat
        assert gi.transform.T.tobytes() == world.tobytes()  # gi.transform is math convention, world is columns
    return got


def _write_synthetic_gltf(tmp_path, glb: bool, embed: bool):
    """Two meshes (u16 indices + u8 normalised texcoords; no indices + strided f32 texcoords), a primitive without
    NORMAL (skipped), a node hierarchy with matrix and TRS nodes, a mesh used by two nodes."""
    rng = np.random.default_rng(11)
    blob = bytearray()
    views, accs = [], []

    def add(arr, stride=0, target=None, normalized=False, typ=None):
        nonlocal blob
        while len(blob) % 4:
            blob += b"\0"
        off = len(blob)
        a = np.ascontiguousarray(arr)
        if stride:
            w = a.dtype.itemsize * a.shape[1]
            rows = bytearray()
            for r in a:
                rows += r.tobytes() + b"\xAB" * (stride - w)
            blob += rows
        else:
            blob += a.tobytes()
        v = {"buffer": 0, "byteOffset": off, "byteLength": len(blob) - off}
        if stride:
            v["byteStride"] = stride
        views.append(v)
        ct = {np.dtype(np.float32): 5126, np.dtype(np.uint16): 5123, np.dtype(np.uint8): 5121, np.dtype(np.uint32): 5125}[a.dtype]
        t = typ or {1: "SCALAR", 2: "VEC2", 3: "VEC3", 4: "VEC4"}[1 if a.ndim == 1 else a.shape[1]]
        acc = {"bufferView": len(views) - 1, "componentType": ct, "count": int(a.shape[0]), "type": t}
        if normalized:
            acc["normalized"] = True
        accs.append(acc)
        return len(accs) - 1

    v0 = rng.standard_normal((40, 3)).astype(np.float32)
    n0 = rng.standard_normal((40, 3)).astype(np.float32)
    i0 = rng.integers(0, 40, 60).astype(np.uint16)
    uv0 = rng.integers(0, 256, (40, 2)).astype(np.uint8)
    v1 = rng.standard_normal((30, 3)).astype(np.float32)
    n1 = rng.standard_normal((30, 3)).astype(np.float32)
    uv1 = rng.random((30, 2)).astype(np.float32)
    tan1 = rng.standard_normal((30, 4)).astype(np.float32)
    v2 = rng.standard_normal((3, 3)).astype(np.float32)
    prims0 = [{"attributes": {"POSITION": add(v0), "NORMAL": add(n0), "TEXCOORD_0": add(uv0, stride=4, normalized=True)},
               "indices": add(i0), "material": 1},
              {"attributes": {"POSITION": add(v2)}}]  # no NORMAL: skipped by make_meshes
    prims1 = [{"attributes": {"POSITION": add(v1), "NORMAL": add(n1), "TEXCOORD_0": add(uv1, stride=16), "TANGENT": add(tan1)}}]
    doc = {
        "asset": {"version": "2.0"},
        "scenes": [{"nodes": [0, 3]}], "scene": 0,
        "nodes": [
            {"children": [1, 2], "translation": [1.5, -2.0, 0.25], "rotation": [0.1825742, 0.3651484, 0.5477226, 0.7302967],
             "scale": [2.0, 0.5, 1.25], "mesh": 1},
            {"mesh": 0, "matrix": [1, 0, 0, 0, 0, 0.5, 0, 0, 0, 0, 2, 0, 3, 4, 5, 1]},
            {"mesh": 0, "rotation": [0.0, 0.7071068, 0.0, 0.7071068], "name": 'child \u00e9 "q" \U0001F600'},
            {"mesh": 1, "scale": [0.1, 0.1, 0.1]},
        ],
        "meshes": [{"primitives": prims0, "name": "m0"}, {"primitives": prims1}],
        "materials": [{"name": "a"}, {"name": "b", "pbrMetallicRoughness": {"baseColorFactor": [0.5, 0.25, 1, 1]}}],
        "accessors": accs, "bufferViews": views,
    }
    if glb:
        doc["buffers"] = [{"byteLength": len(blob)}]
        text = json.dumps(doc).encode()
        text += b" " * (-len(text) % 4)
        b = bytes(blob) + b"\0" * (-len(blob) % 4)
        out = b"glTF" + struct.pack("<II", 2, 12 + 8 + len(text) + 8 + len(b)) + struct.pack("<II", len(text), 0x4E4F534A) + text
        out += struct.pack("<II", len(b), 0x004E4942) + b
        p = tmp_path / "s.glb"
        p.write_bytes(out)
        return str(p)
    if embed:
        doc["buffers"] = [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(bytes(blob)).decode()}]
    else:
        (tmp_path / "s data.bin").write_bytes(bytes(blob))
        doc["buffers"] = [{"byteLength": len(blob), "uri": "s%20data.bin"}]
    p = tmp_path / "s.gltf"
    p.write_text(json.dumps(doc, indent=1))
    return str(p)


@pytest.mark.parametrize("glb,embed", [(True, False), (False, True), (False, False)])
def test_gltf_synthetic(tmp_path, glb, embed):
    got = check_gltf(_write_synthetic_gltf(tmp_path, glb, embed))
    assert len(got.meshes) == 2 and len(got.instances) == 4
    # children before the node's own primitives (gltf_model/mod.rs:183-205): node 0 = [child 1, child 2, own], then node 3
    assert [i.mesh for i in got.instances] == [0, 0, 1, 1]
    assert got.materials[1]["base_color"].tolist() == [0.5, 0.25, 1.0, 1.0]


def test_gltf_errors(tmp_path):
    (tmp_path / "x.gltf").write_text("{ not json")
    with pytest.raises(BvhCudaError):
        M.GltfDocument.import_(str(tmp_path / "x.gltf"))
    (tmp_path / "y.gltf").write_text(json.dumps({"asset": {"version": "2.0"}, "buffers": [{"byteLength": 4, "uri": "nope.bin"}]}))
    with pytest.raises(BvhCudaError):
        M.GltfDocument.import_(str(tmp_path / "y.gltf"))


# ---- the reference checkout's own assets --------------------------------------------------------------------------------
def _digest(model):
    h = hashlib.sha256()
    for m in model.meshes:
        for a in (m.vertices, m.normals, m.tex_coords, m.tangents, m.indices):
            h.update(np.ascontiguousarray(a).tobytes())
    for i in model.instances:
        h.update(np.ascontiguousarray(i.transform).tobytes() + struct.pack("<Ii", i.mesh, i.material))
    return h.hexdigest()


REAL = [("cube/cube.obj", "cube"), ("glTF-Sample-Models/2.0/DamagedHelmet/glTF-Binary/DamagedHelmet.glb", "helmet"),
        ("glTF-Sample-Models/2.0/AntiqueCamera/glTF/AntiqueCamera.gltf", "camera")]


@needs_ref
@pytest.mark.parametrize("rel,name", REAL)
def test_reference_assets(rel, name):
    path = os.path.join(REF_ASSETS, rel)
    got = check_obj(path) if rel.endswith(".obj") else check_gltf(path)
    gold = json.load(open(os.path.join(HERE, "golden", "models_golden.json")))[name]
    assert _digest(got) == gold["sha256"]
    assert [int(m.vertices.shape[0]) for m in got.meshes] == gold["vertices"]
    assert [int(m.indices.size) for m in got.meshes] == gold["indices"] and len(got.instances) == gold["instances"]
    # same triangles as the committed positions+indices fixture (which the BLAS parity tests build from)
    fx = np.load(os.path.join(HERE, "golden", "real_meshes.npz"))
    for k, m in enumerate(got.meshes):
        fv, fi = (fx["cube_v"], fx["cube_i"]) if name == "cube" else (fx[f"{name}{k}_v"], fx[f"{name}{k}_i"])
        assert m.vertices[m.indices].tobytes() == fv[fi].tobytes()


def test_load_single_mesh_pools_and_rebases(tmp_path):
    (tmp_path / "two.obj").write_text("o a\nv 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\no b\nv 0 0 1\nv 1 0 1\nv 0 1 1\nf 4 5 6\n")
    v, i = M.load_single_mesh(str(tmp_path / "two.obj"))
    assert v.shape == (6, 3) and i.tolist() == [0, 1, 2, 3, 4, 5]
    assert M.find_asset("definitely-not-there.obj") is None
    assert S is not None


def test_models_c_abi_argument_checks(tmp_path):
    """Raw C ABI: null arguments and out-of-range indices come back as BVH_CUDA_EINVAL (-1), nothing crashes; an OBJ has no
    instances; the error text of a failing load is kept per thread."""
    import ctypes as C

    lib = M._lib_models()
    h = C.c_void_p()
    assert lib.bvh_cuda_model_load_obj(None, C.byref(h)) == -1
    assert lib.bvh_cuda_model_load_gltf(b"/nonexistent/x.glb", C.byref(h)) == -1 and not h.value
    assert b"cannot read" in lib.bvh_cuda_model_last_error()
    (tmp_path / "t.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    assert lib.bvh_cuda_model_load_obj(str(tmp_path / "t.obj").encode(), C.byref(h)) == 0 and h.value
    assert lib.bvh_cuda_model_mesh_count(h) == 1 and lib.bvh_cuda_model_instance_count(h) == 0
    mv = M._MeshView()
    assert lib.bvh_cuda_model_mesh(h, 1, C.byref(mv)) == -1 and lib.bvh_cuda_model_mesh(h, 0, None) == -1
    assert lib.bvh_cuda_model_mesh(h, 0, C.byref(mv)) == 0 and mv.n_vertices == 3 and mv.n_indices == 3 and mv.name == b"unnamed_object"
    assert lib.bvh_cuda_model_instance(h, 0, C.byref(M._InstanceView())) == -1
    assert lib.bvh_cuda_model_material(h, 0, C.byref(M._MaterialView())) == -1
    lib.bvh_cuda_model_free(h)
    lib.bvh_cuda_model_free(None)
    assert lib.bvh_cuda_model_mesh_count(None) == 0
