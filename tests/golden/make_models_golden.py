"""Digests of what the asset loaders (csrc/models.cpp) return for the meshes that ship with the reference checkout
(cube.obj, DamagedHelmet.glb, AntiqueCamera.gltf) -> tests/golden/models_golden.json.  tests/test_models.py first
checks the loaders against independent Python readers of the same files, then pins them to these digests.
Run from the repo root (needs /root/reference):  python tests/golden/make_models_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_models import REAL, REF_ASSETS, _digest  # noqa: E402
from voidin_b200 import models as M  # noqa: E402

out = {}
for rel, name in REAL:
    path = os.path.join(REF_ASSETS, rel)
    m = (M.ObjModel if rel.endswith(".obj") else M.GltfDocument).import_(path)
    out[name] = {"sha256": _digest(m), "vertices": [int(x.vertices.shape[0]) for x in m.meshes],
                 "indices": [int(x.indices.size) for x in m.meshes], "instances": len(m.instances)}
    print(name, out[name])
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "models_golden.json"), "w"), indent=1)
