"""Extracts positions + triangle indices of the real meshes that ship with the reference checkout (the ones
SURVEY.md §4 lists: assets/cube/cube.obj, DamagedHelmet.glb, AntiqueCamera.gltf) into tests/golden/real_meshes.npz,
so the GPU box (which has no /root/reference) can run parity on real geometry.  Only vertex positions and indices are
kept (asset data, not reference source).  Run from the repo root:  python tests/golden/make_real_meshes.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from voidin_b200 import scenes as S  # noqa: E402

REF = "/root/reference/assets"
out = {}
v, i = S.load_obj_positions(os.path.join(REF, "cube/cube.obj"))
out["cube_v"], out["cube_i"] = v, i
for name, rel in [("helmet", "glTF-Sample-Models/2.0/DamagedHelmet/glTF-Binary/DamagedHelmet.glb"),
                  ("camera", "glTF-Sample-Models/2.0/AntiqueCamera/glTF/AntiqueCamera.gltf")]:
    for k, (v, i) in enumerate(S.load_gltf_primitives(os.path.join(REF, rel))):
        out[f"{name}{k}_v"], out[f"{name}{k}_i"] = v, i
        print(name, k, v.shape, i.size // 3)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "real_meshes.npz"), **out)
print("wrote real_meshes.npz", {k: a.shape for k, a in out.items()})
