"""Child process of the GPU tests that trace under a different environment (the library reads its A/B switches once per
process): loads a scene + rays from an .npz, traces closest-hit and any-hit through the C ABI, writes the results.
    python tests/_child_trace.py in.npz out.npz"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voidin_b200 as vb  # noqa: E402
from voidin_b200.types import INSTANCE, MESH_INFO  # noqa: E402


def main(src, dst):
    d = np.load(src)
    ctx = vb.Context(0)
    kids = d["kids"] if d["kids"].size else None
    scene = vb.Scene(d["tlas"], kids, d["inst"].view(INSTANCE).reshape(-1), d["infos"].view(MESH_INFO).reshape(-1), d["nodes"],
                     d["verts"], d["inds"], ctx)
    t, tri, ins = scene.traverse_tlas(d["ro"], d["rd"])
    occ = scene.occluded(d["ro"], d["rd"])
    np.savez(dst, t=t, tri=tri, ins=ins, occ=occ)


if __name__ == "__main__":
    main(*sys.argv[1:3])
