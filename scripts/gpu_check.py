"""First-contact GPU diagnostics: BLAS parity against the oracle over a ladder of sizes that exercises each
tier (warp sub-tree, block task queue, grid-wide levels), then TLAS and traversal.  Prints the first
mismatch in detail.  Run under gpurun; output goes to stdout."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import voidin_b200 as vb
from voidin_b200 import scenes as S
from oracle import oracle as O

ctx = vb.Context(0)
ok_all = True

def check_blas(name, v, idx):
    global ok_all
    gi = idx.copy()
    t0 = time.time()
    try:
        bvh = vb.BvhBuilder(v, gi, ctx).build()
    except vb.BvhCudaError as e:
        rc, *_ = O.blas_build(v, idx)
        print(f"[{name}] GPU error: {e}; oracle rc={rc}")
        ok_all = ok_all and (rc == e.code)
        return
    dt = time.time() - t0
    rc, onodes, oidx, oorder, st = O.blas_build(v, idx)
    n = idx.size // 3
    same_n = len(bvh.nodes) == len(onodes)
    same_nodes = same_n and bvh.nodes.tobytes() == onodes.tobytes()
    same_idx = (gi == oidx).all()
    stats = ctx.last_build_stats()
    print(f"[{name}] n={n} M={len(bvh.nodes)}/{len(onodes)} nodes_ok={same_nodes} idx_ok={same_idx} "
          f"S={stats['sum_interior_prims']}/{st['sum_interior_prims']} levels={stats['grid_levels']} t2b={stats['big_block_tasks']} t2={stats['block_tasks']} t2w={stats['warp_node_tasks']} "
          f"t3={stats['warp_tasks']} launches={stats['kernel_launches']} e2e={dt*1e3:.2f}ms", flush=True)
    if not (same_nodes and same_idx):
        ok_all = False
        order = ctx.last_order(n)
        bad = np.nonzero(order != oorder)[0]
        print("   order mismatches:", len(bad), "first at", bad[:8])
        if same_n:
            a = bvh.nodes.view(np.uint8).reshape(-1, 32); b = onodes.view(np.uint8).reshape(-1, 32)
            badn = np.nonzero((a != b).any(axis=1))[0]
            print("   node mismatches:", len(badn), "first", badn[:8])
            for k in badn[:3]:
                print("     gpu", bvh.nodes[k], "\n     ora", onodes[k])
        is_perm = (np.sort(order) == np.arange(n)).all()
        print("   gpu order is a permutation:", is_perm)

for n in [1, 2, 3, 4, 5, 7, 8, 16, 31, 32]:
    check_blas(f"soup{n}", *S.soup(n, 100 + n, 0.05))
check_blas("plane", *S.make_plane_mesh())
check_blas("sphere1", *S.make_uv_sphere(1, 1))
for n in [33, 64, 100, 500, 2000, 2048]:
    check_blas(f"soup{n}", *S.soup(n, 100 + n, 0.05))
check_blas("grid20", *S.grid_mesh(20, 20))
check_blas("sphere10", *S.make_uv_sphere(1, 10))
for n in [2049, 5000, 20000, 100000]:
    check_blas(f"soup{n}", *S.soup(n, 100 + n, 0.02))
check_blas("bunny", *S.bunny_class())
# degenerate: identical triangles
v = np.tile(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), (8, 1)); idx = np.arange(24, dtype=np.uint32)
check_blas("degenerate8", v, idx)

# TLAS + traversal
def builder(vv, ii):
    i2 = ii.copy(); b = vb.BvhBuilder(vv, i2, ctx).build(); return b.nodes, i2
pool = S.MeshPool(builder)
pool.add(*S.make_plane_mesh()); pool.add(*S.make_uv_sphere(1, 10)); pool.add(*S.soup(3000, 9, 0.05))
verts, inds, nodes, infos = pool.pooled()
for I in [1, 2, 3, 10, 100, 1000]:
    inst = S.random_instances(I, 3, seed=I, extent=20.0)
    tl = vb.Tlas.empty(ctx); t0 = time.time(); tl.build(inst, infos); dt = time.time() - t0
    rc, otl, okids, calls, pairs = O.tlas_build(inst, infos)
    okk = tl.nodes.tobytes() == otl.tobytes() and (tl.children == okids).all()
    ok_all = ok_all and okk
    print(f"[tlas I={I}] ok={okk} e2e={dt*1e3:.2f}ms calls={calls}", flush=True)
    if not okk:
        a = tl.nodes.view(np.uint8).reshape(-1, 32); b = otl.view(np.uint8).reshape(-1, 32)
        badn = np.nonzero((a != b).any(axis=1))[0]; print("   bad nodes", len(badn), badn[:8])
    scene = vb.Scene(tl.nodes, tl.children, inst, infos, nodes, verts, inds, ctx)
    ro, rd = S.rays_toward_box(20000, [-20, -20, -20], [20, 20, 20], seed=77)
    t, tri, ins = scene.traverse_tlas(ro, rd); occ = scene.occluded(ro, rd)
    ot, otri, oins, _, st = O.trace_scene(otl, okids, inst, infos, nodes, verts, inds, ro, rd, threads=8)
    _, _, _, oocc, _ = O.trace_scene(otl, okids, inst, infos, nodes, verts, inds, ro, rd, any_hit=True, threads=8)
    okt = (tri == otri).all() and (ins == oins).all() and (t == ot).all() and (occ == oocc).all()
    ok_all = ok_all and okt
    print(f"   trace ok={okt} hits={int((otri != 0xFFFFFFFF).sum())} t_bitexact={(t == ot).all()} maxstack={st['max_stack']}", flush=True)
# BLAS-only Rust mode
v, idx = S.bunny_class(); gi = idx.copy(); bvh = vb.BvhBuilder(v, gi, ctx).build()
ro, rd = S.rays_toward_box(50000, v.min(0), v.max(0), seed=11)
t, tri = bvh.traverse_iter_batch(v, gi, ro, rd)
ot, otri, st = O.trace_blas(bvh.nodes, v, gi, ro, rd, threads=8)
okb = (tri == otri).all() and (t == ot).all()
ok_all = ok_all and okb
print(f"[trace_blas bunny] ok={okb} hits={int((otri != 0xFFFFFFFF).sum())} maxstack={st['max_stack']}")
print("ALL OK" if ok_all else "SOME FAILED")
