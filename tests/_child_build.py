"""Child process of the GPU tests that need a different environment (the library reads its A/B switches once per
process): builds every mesh of an .npz (v0, i0, v1, i1, ...) either one by one or as one forest build and writes the
nodes / permuted indices / build statistics to another .npz.   python tests/_child_build.py in.npz out.npz single|forest"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voidin_b200 as vb  # noqa: E402


def main(src, dst, mode):
    d = np.load(src)
    k = len(d.files) // 2
    meshes = [(np.ascontiguousarray(d[f"v{i}"]), np.ascontiguousarray(d[f"i{i}"])) for i in range(k)]
    ctx = vb.Context(0)
    out = {}
    if mode == "forest":
        import torch
        from voidin_b200 import multi_gpu as MG

        dev = torch.device("cuda", 0)
        tm = [(torch.from_numpy(v.reshape(-1)).to(dev), torch.from_numpy(i.view(np.int32)).to(dev)) for v, i in meshes]
        outs = MG.cuda_build_batch_fn(ctx)(tm)
        for i, (nodes, perm) in enumerate(outs):
            out[f"n{i}"] = nodes.cpu().numpy()
            out[f"g{i}"] = perm.cpu().numpy().view(np.uint32)
    else:
        for i, (v, idx) in enumerate(meshes):
            gi = idx.copy()
            out[f"n{i}"] = vb.BvhBuilder(v, gi, ctx).build().nodes.view(np.int32).reshape(-1)
            out[f"g{i}"] = gi
    out["stats"] = np.frombuffer(json.dumps(ctx.last_build_stats()).encode(), dtype=np.uint8)
    np.savez(dst, **out)


if __name__ == "__main__":
    main(*sys.argv[1:4])
