#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "trace or any_hit or smoke" > gpurun_out/r03k_pytest.log 2>&1 ) 2>&1 | grep real
tail -3 gpurun_out/r03k_pytest.log
for st in 1 2; do for lg in 21 24; do
BVH_CUDA_TRACE_STREAMS=$st timeout 300 python - <<PY
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import voidin_b200 as vb
from voidin_b200 import scenes as S
ctx = vb.Context(0)
def gb(v, i):
    gi = np.array(i, dtype=np.uint32, copy=True); b = vb.BvhBuilder(v, gi, ctx).build(); return b.nodes, gi
dv, di = S.dragon_class(); pl_v, pl_i = S.make_plane_mesh()
pool = S.MeshPool(gb); pool.add(pl_v, pl_i); pool.add(dv, di)
pv, pi, pn, pinf = pool.pooled()
mats, mids = S.dragon_scene_instances(); inst = S.make_instances(mats, mids)
tl = vb.Tlas.empty(ctx); tl.build(inst, pinf)
scene = vb.Scene(tl.nodes, tl.children, inst, pinf, pn, pv, pi, ctx)
n = 1 << $lg
ro, rd = S.gbuffer_shadow_rays(n, S.world_triangles(dv, di, mats[1]), S.rect_light_corners(), seed=12)
# block-cyclic 64 Ki chunks of a 16 Mi batch, like a rank of an 8-GPU run, when n is small
h_ro = torch.from_numpy(ro.reshape(-1)).pin_memory(); h_rd = torch.from_numpy(rd.reshape(-1)).pin_memory()
h_occ = torch.empty(n, dtype=torch.uint8).pin_memory()
import ctypes as C
lib = ctx.lib
ts = []
for k in range(6):
    t0 = time.perf_counter()
    ctx.check(lib.bvh_cuda_trace_any(ctx.h, scene.h, h_ro.data_ptr(), h_rd.data_ptr(), n, 1e30, h_occ.data_ptr()))
    ts.append(time.perf_counter() - t0)
print(f"streams=$st rays=2^$lg e2e {1e3*np.mean(ts[1:]):.3f} ms  {n/np.mean(ts[1:])/1e6:.0f} Mrays/s  occ {h_occ.float().mean():.4f}")
PY
done; done
