"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): every tier of the BLAS build, TLAS
build and both traversals on small inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import voidin_b200 as vb
from voidin_b200 import scenes as S

ctx = vb.Context(0)
def builder(v, i):
    i2 = i.copy(); b = vb.BvhBuilder(v, i2, ctx).build(); return b.nodes, i2
pool = S.MeshPool(builder)
big = int(sys.argv[1]) if len(sys.argv) > 1 else 40000  # 130000: grid tier on 512-slot tiles at the root, 256 below
for n in (3, 20, 200, 3000, big):
    pool.add(*S.soup(n, 70 + n, 0.05))
verts, inds, nodes, infos = pool.pooled()
inst = S.random_instances(50, len(infos), seed=4, extent=5.0)
tl = vb.Tlas.empty(ctx); tl.build(inst, infos)
scene = vb.Scene(tl.nodes, tl.children, inst, infos, nodes, verts, inds, ctx)
ro, rd = S.rays_toward_box(20000, [-5, -5, -5], [5, 5, 5], seed=3)
t, tri, ins = scene.traverse_tlas(ro, rd); occ = scene.occluded(ro, rd)
v, i = S.soup(3000, 3070, 0.05); i2 = i.copy(); b = vb.BvhBuilder(v, i2, ctx).build()
t2, tri2 = b.traverse_iter_batch(v, i2, ro[:5000] * 0.1 + 0.5, rd[:5000])
print("ok", len(nodes), int((tri != 0xFFFFFFFF).sum()), int(occ.sum()), int((tri2 != 0xFFFFFFFF).sum()))
