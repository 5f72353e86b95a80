"""Shared test inputs and structural checks (no reference / oracle knowledge in the checks themselves)."""
import hashlib
import os

import numpy as np

from voidin_b200 import scenes as S

REF_CUBE = "/root/reference/assets/cube/cube.obj"


def small_meshes():
    """(name, vertices, indices) — the meshes every voidin run builds (mesh/mod.rs:267-274) plus adversarial ones."""
    out = [
        ("plane", *S.make_plane_mesh()),
        ("uv_sphere_1", *S.make_uv_sphere(1.0, 1)),
        ("uv_sphere_10", *S.make_uv_sphere(1.0, 10)),
        ("soup_1", *S.soup(1, 101, 0.05)),
        ("soup_2", *S.soup(2, 102, 0.05)),
        ("soup_3", *S.soup(3, 103, 0.05)),
        ("soup_4", *S.soup(4, 104, 0.05)),
        ("soup_5", *S.soup(5, 105, 0.05)),
        ("soup_33", *S.soup(33, 133, 0.05)),
        ("soup_1000", *S.soup(1000, 7, 0.05)),
        ("soup_2049", *S.soup(2049, 2149, 0.02)),
        ("soup_20000", *S.soup(20000, 20100, 0.02)),
        ("grid_20x20_zero_extent_axis", *S.grid_mesh(20, 20)),
        ("grid_64x3", *S.grid_mesh(64, 3)),
        ("displaced_sphere_5k", *S.displaced_sphere(36, 72, 9)),
        ("dup_centroids_3", *dup_centroids()),
    ]
    out += real_meshes()
    return out


def real_meshes():
    """Real geometry of the reference checkout (cube.obj, DamagedHelmet.glb, AntiqueCamera.gltf), extracted once by
    tests/golden/make_real_meshes.py into a committed .npz so it travels to the GPU box."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "real_meshes.npz")
    if not os.path.exists(path):
        return []
    d = np.load(path)
    names = sorted({k[:-2] for k in d.files})
    return [("real_" + n, np.ascontiguousarray(d[n + "_v"]), np.ascontiguousarray(d[n + "_i"])) for n in names]


def dup_centroids():
    """Groups of <=3 triangles sharing one centroid (the largest degenerate group the reference survives)."""
    rng = np.random.default_rng(5)
    verts, idx = [], []
    for g in range(40):
        c = rng.random(3)
        for k in range(3):
            d = rng.normal(size=(2, 3)) * 0.05
            tri = np.stack([c + d[0], c + d[1], c - d[0] - d[1]])
            idx.extend(range(len(verts), len(verts) + 3))
            verts.extend(tri)
    return np.asarray(verts, dtype=np.float32), np.asarray(idx, dtype=np.uint32)


def sha(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def check_bvh_structure(vertices, orig_indices, nodes, perm_indices, order=None):
    """Checks that do not depend on the oracle being right by construction."""
    v = np.asarray(vertices, dtype=np.float32).reshape(-1, 3)
    n = orig_indices.size // 3
    tri_o = np.asarray(orig_indices).reshape(-1, 3)
    tri_p = np.asarray(perm_indices).reshape(-1, 3)
    # (1) the permuted index buffer is a permutation of the original triangles
    if order is not None:
        assert (np.sort(order) == np.arange(n)).all()
        assert (tri_p == tri_o[order]).all()
    else:
        key = lambda t: np.sort(np.ascontiguousarray(t).view([("a", "<u4"), ("b", "<u4"), ("c", "<u4")]).reshape(-1), order=("a", "b", "c"))
        assert (key(tri_p) == key(tri_o)).all()
    count = nodes["count"]
    lf = nodes["left_first"]
    interior = np.zeros(len(nodes), dtype=bool)
    # (3) node 1 all-zero, nodes.len() == 2 + 2*interior, leaf <=> count in {1,2,3}
    assert nodes[1].tobytes() == b"\0" * 32
    reach = np.zeros(len(nodes), dtype=bool)
    covered = np.zeros(n, dtype=np.int32)
    tmin = v[tri_p].min(axis=1)
    tmax = v[tri_p].max(axis=1)
    stack = [(0, 0, n)]
    while stack:
        i, s, c = stack.pop()
        reach[i] = True
        # (2) every node box equals the exact union of its range's triangle vertices
        assert (nodes["min"][i] == tmin[s:s + c].min(axis=0)).all(), i
        assert (nodes["max"][i] == tmax[s:s + c].max(axis=0)).all(), i
        if count[i] > 0:
            assert c == count[i] and lf[i] == s and 1 <= c <= 3
            covered[s:s + c] += 1
        else:
            assert c > 3
            interior[i] = True
            l, r = int(lf[i]), int(lf[i]) + 1
            # children counts are not stored for interior children; recover the split from the left subtree's extent
            lc = _subtree_count(nodes, l)
            stack.append((r, s + lc, c - lc))
            stack.append((l, s, lc))
    assert (covered == 1).all()
    assert len(nodes) == 2 + 2 * int(interior.sum())
    assert reach[0] and not reach[1] and reach[2:].all()


def _subtree_count(nodes, i):
    tot, st = 0, [i]
    while st:
        k = st.pop()
        if nodes["count"][k] > 0:
            tot += int(nodes["count"][k])
        else:
            st.append(int(nodes["left_first"][k]))
            st.append(int(nodes["left_first"][k]) + 1)
    return tot


def make_scene(builder, n_inst=40, seed=3, extent=20.0):
    """Three pooled meshes + random instances; `builder(v, idx) -> (nodes, permuted_indices)`."""
    pool = S.MeshPool(builder)
    pool.add(*S.make_plane_mesh())
    pool.add(*S.make_uv_sphere(1.0, 10))
    pool.add(*S.soup(3000, 9, 0.05))
    verts, inds, nodes, infos = pool.pooled()
    inst = S.random_instances(n_inst, 3, seed=seed, extent=extent)
    return verts, inds, nodes, infos, inst
