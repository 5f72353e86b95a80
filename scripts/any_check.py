"""Any-hit (order-free kernel) vs oracle on several scenes, plus timing.  Usage: python scripts/any_check.py [n_rays_log2]"""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import voidin_b200 as vb
from voidin_b200 import scenes as S
from oracle import oracle
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import make_scene

ctx = vb.Context(0)

def gpu_build(v, i):
    gi = np.array(i, dtype=np.uint32, copy=True)
    b = vb.BvhBuilder(v, gi, ctx).build()
    return b.nodes, gi

def check(name, scene_args, ro, rd, tmax=1e30):
    tl_nodes, kids, inst, infos, nodes, verts, inds = scene_args
    scene = vb.Scene(tl_nodes, kids, inst, infos, nodes, verts, inds, ctx)
    occ = scene.occluded(ro, rd, tmax=tmax)
    _, _, _, oocc, _ = oracle.trace_scene(tl_nodes, kids, inst, infos, nodes, verts, inds, ro, rd, tmax=tmax, any_hit=True,
                                          threads=oracle.max_threads())
    bad = int((occ != oocc).sum())
    print(f"[{name}] rays={len(ro)} tmax={tmax} occluded={int(oocc.sum())} mismatches={bad}", flush=True)
    return bad

def parity():
    bad = 0
    verts, inds, nodes, infos, inst = make_scene(gpu_build, n_inst=300)
    tl = vb.Tlas.empty(ctx); tl.build(inst, infos)
    args = (tl.nodes, tl.children, inst, infos, nodes, verts, inds)
    ro, rd = S.rays_toward_box(400_000, [-20, -20, -20], [20, 20, 20], seed=77)
    bad += check("scene300", args, ro, rd)
    bad += check("scene300 tmax=0.8", args, ro, rd, tmax=0.8)
    bad += check("scene300 tmax=3e38", args, ro[:64], rd[:64], tmax=3e38)
    # axis-aligned and zero directions -> deferral path
    ro2 = ro[:60000].copy(); rd2 = rd[:60000].copy()
    rd2[::3, 0] = 0.0; rd2[1::3, 1] = 0.0; rd2[2::7] = 0.0
    bad += check("scene300 zero-comps", args, ro2, rd2)
    # grazing: rays aimed exactly at vertices of a mesh (identity instance)
    v, idx = S.bunny_class()
    bn, gi = gpu_build(v, idx)
    pool = S.MeshPool(gpu_build); pool.add(v, idx)
    pv, pi, pn, pinf = pool.pooled()
    inst1 = S.make_instances(np.eye(4)[None], [0])
    tl1 = vb.Tlas.empty(ctx); tl1.build(inst1, pinf)
    args1 = (tl1.nodes, tl1.children, inst1, pinf, pn, pv, pi)
    rng = np.random.default_rng(5)
    tgt = v[rng.integers(0, len(v), 300_000)]
    org = (rng.normal(size=tgt.shape) * 3).astype(np.float32)
    bad += check("bunny vertex-grazing", args1, org, (tgt - org).astype(np.float32))
    tri3 = np.asarray(idx).reshape(-1, 3)
    tri = tri3[rng.integers(0, len(tri3), 300_000)]
    mid = ((v[tri[:, 0]] + v[tri[:, 1]]) * np.float32(0.5)).astype(np.float32)
    bad += check("bunny edge-grazing", args1, org, (mid - org).astype(np.float32))
    print("ANY-HIT OK" if bad == 0 else f"ANY-HIT FAIL {bad}")



if os.environ.get("PARITY", "1") == "1":
    parity()

# timing on the dragon scene
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
dv, di = S.dragon_class()
pl_v, pl_i = S.make_plane_mesh()
pool = S.MeshPool(gpu_build); pool.add(pl_v, pl_i); pool.add(dv, di)
pv, pi, pn, pinf = pool.pooled()
mats, mesh_ids = S.dragon_scene_instances()
inst = S.make_instances(mats, mesh_ids)
tl = vb.Tlas.empty(ctx); tl.build(inst, pinf)
scene = vb.Scene(tl.nodes, tl.children, inst, pinf, pn, pv, pi, ctx)
n = 1 << lg
ro, rd = S.gbuffer_shadow_rays(n, S.world_triangles(dv, di, mats[1]), S.rect_light_corners(), seed=12,
                               coherent=os.environ.get("RAYS", "coherent") == "coherent")
dev = torch.device("cuda", 0)
d_ro = torch.from_numpy(ro).to(dev); d_rd = torch.from_numpy(rd).to(dev)
d_occ = torch.empty(n, dtype=torch.uint8, device=dev)
from voidin_b200 import _lib
L = _lib.load()
ms = None
for votes in [int(x) for x in os.environ.get("VOTES", "1").split(",")]:
    L.bvh_cuda_debug_set_votes(votes, 1)
    for _ in range(3):
        scene.occluded_dev(d_ro.data_ptr(), d_rd.data_ptr(), n, d_occ.data_ptr(), 1e30, 0)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        scene.occluded_dev(d_ro.data_ptr(), d_rd.data_ptr(), n, d_occ.data_ptr(), 1e30, 0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"  votes={votes}: {ms:.3f} ms  {n/ms/1e3:.1f} Mrays/s", flush=True)
occ = d_occ.cpu().numpy()
k = min(n, 1 << 19)
_, _, _, oocc, _ = oracle.trace_scene(tl.nodes, tl.children, inst, pinf, pn, pv, pi, ro[:k], rd[:k], any_hit=True, threads=oracle.max_threads())
print(f"[dragon any-hit] mode={os.environ.get('BVH_CUDA_ANYHIT','fast')} rays={n} {ms:.3f} ms  {n/ms/1e3:.1f} Mrays/s  occluded={occ.mean():.4f} mismatches(first {k})={int((occ[:k]!=oocc).sum())}")

if os.environ.get("SIZES"):
    L.bvh_cuda_debug_set_votes(1, 1)
    def timeit(off, cnt, reps=5):
        for _ in range(2):
            scene.occluded_dev(d_ro.data_ptr() + 12 * off, d_rd.data_ptr() + 12 * off, cnt, d_occ.data_ptr(), 1e30, 0)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            scene.occluded_dev(d_ro.data_ptr() + 12 * off, d_rd.data_ptr() + 12 * off, cnt, d_occ.data_ptr(), 1e30, 0)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    for cnt in [1, 32, 1 << 10, 1 << 14, 1 << 17, 1 << 20, 1 << 21, 1 << 22, n // 2, n]:
        if cnt <= n:
            print(f"  first {cnt:9d} rays: {timeit(0, cnt):.3f} ms", flush=True)
    print(f"  model half : {timeit(0, n // 2):.3f} ms   ground half: {timeit(n // 2, n // 2):.3f} ms", flush=True)
