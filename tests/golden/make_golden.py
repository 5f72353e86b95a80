"""Writes tests/golden/oracle_hashes.json: SHA-256 of the oracle's outputs on fixed seeded inputs, plus a few
literal values.  Run from the repo root:  python tests/golden/make_golden.py
The reference itself cannot be executed here (Rust, no toolchain) and ships no vectors, so these pin the oracle
against regressions, not against the reference."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

from oracle import oracle as O  # noqa: E402
from voidin_b200 import scenes as S  # noqa: E402
from helpers import sha  # noqa: E402


def compute():
    out = {}
    for name, (v, idx) in {
        "plane": S.make_plane_mesh(),
        "uv_sphere_1": S.make_uv_sphere(1.0, 1),
        "uv_sphere_10": S.make_uv_sphere(1.0, 10),
        "soup_1000_seed7": S.soup(1000, 7, 0.05),
        "soup_100000_seed0_edge0.01": S.soup(100000, 0, 0.01),
        "grid_20x20": S.grid_mesh(20, 20),
    }.items():
        rc, nodes, perm, order, st = O.blas_build(v, idx)
        out[name] = {"rc": rc, "n_nodes": int(len(nodes)), "interior": st["interior_nodes"],
                     "S": st["sum_interior_prims"], "max_depth": st["max_depth"],
                     "input": sha(v, idx), "nodes": sha(nodes), "indices": sha(perm), "order": sha(order)}
    # scene: TLAS + traversal
    def builder(v, i):
        rc, nodes, perm, _, _ = O.blas_build(v, i)
        return nodes, perm
    pool = S.MeshPool(builder)
    pool.add(*S.make_plane_mesh()); pool.add(*S.make_uv_sphere(1.0, 10)); pool.add(*S.soup(3000, 9, 0.05))
    verts, inds, nodes, infos = pool.pooled()
    inst = S.random_instances(200, 3, seed=3, extent=20.0)
    rc, tl, kids, calls, pairs = O.tlas_build(inst, infos)
    ro, rd = S.rays_toward_box(5000, [-20, -20, -20], [20, 20, 20], seed=77)
    t, tri, ins, occ, st = O.trace_scene(tl, kids, inst, infos, nodes, verts, inds, ro, rd)
    out["scene_200"] = {"tlas": sha(tl), "children": sha(kids), "calls": int(calls), "pairs": int(pairs),
                        "t": sha(t), "tri": sha(tri), "inst": sha(ins), "occ": sha(occ), "hits": st["hits"],
                        "pops": st["pops"]}
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_hashes.json")
    with open(path, "w") as f:
        json.dump(compute(), f, indent=1, sort_keys=True)
    print("wrote", path)
