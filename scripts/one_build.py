"""Profiling driver: a few dragon-class BLAS builds (device-resident) and one 4 Mi-ray any-hit trace, nothing else.
Used under ncu:  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python scripts/one_build.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import voidin_b200 as vb
from voidin_b200 import scenes as S

n_builds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 22)
dev = torch.device("cuda", 0)
ctx = vb.Context(0); ctx.set_profiling(True)
dv, di = S.dragon_class(); pv, pi = S.make_plane_mesh()
mats, mids = S.dragon_scene_instances(); inst = S.make_instances(mats, mids)
d_v = torch.from_numpy(dv.reshape(-1)).to(dev); d_i0 = torch.from_numpy(di.view(np.int32)).to(dev)
n = di.size // 3
d_nodes = torch.zeros(2 * n * 8, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for k in range(n_builds):
    d_i = d_i0.clone(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    m = ctx.blas_build_dev(d_v.data_ptr(), dv.shape[0], d_i.data_ptr(), n, d_nodes.data_ptr(), 2 * n, stream)
    torch.cuda.synchronize()
    print(f"build {k}: {1e3*(time.perf_counter()-t0):.3f} ms wall", ctx.last_build_stats(), flush=True)
if n_rays:
    def builder(v, i):
        i2 = i.copy(); b = vb.BvhBuilder(v, i2, ctx).build(); return b.nodes, i2
    pool = S.MeshPool(builder); pool.add(pv, pi); pool.add(dv, di)
    verts, inds, nodes, infos = pool.pooled()
    tl = vb.Tlas.empty(ctx); tl.build(inst, infos)
    scene = vb.Scene(tl.nodes, tl.children, inst, infos, nodes, verts, inds, ctx)
    ro, rd = S.gbuffer_shadow_rays(n_rays, S.world_triangles(dv, di, mats[1]), S.rect_light_corners(), seed=12)
    d_ro = torch.from_numpy(ro.reshape(-1)).to(dev); d_rd = torch.from_numpy(rd.reshape(-1)).to(dev)
    d_occ = torch.empty(n_rays, dtype=torch.uint8, device=dev)
    for k in range(4):
        if k == 2:
            perm = torch.randperm(n_rays).to(dev)
            d_ro = d_ro.view(-1, 3)[perm].contiguous().view(-1); d_rd = d_rd.view(-1, 3)[perm].contiguous().view(-1)
            print("-- incoherent (random permutation) --")
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); scene.occluded_dev(d_ro.data_ptr(), d_rd.data_ptr(), n_rays, d_occ.data_ptr(), 1e30, stream); e1.record()
        torch.cuda.synchronize()
        print(f"trace any {k}: {e0.elapsed_time(e1):.3f} ms  {n_rays/e0.elapsed_time(e1)/1e3:.1f} Mrays/s occ={d_occ.float().mean().item():.4f}", flush=True)
