"""GPU parity tests proper: the CUDA path, called through the C ABI (via the host mirror), against the oracle on
the same seeded inputs — bit-exact nodes, topology, primitive order and hit ids; t within 1e-6 relative (in
practice bit-exact, since the kernels are compiled without FMA contraction)."""
import json
import os

import numpy as np
import pytest

import voidin_b200 as vb
from voidin_b200 import scenes as S

from helpers import check_bvh_structure, make_scene, sha, small_meshes

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_hashes.json")
T_RTOL = 1e-6  # north_star: "hit distance t agrees within 1e-6 relative"


def gpu_build(ctx, v, idx):
    gi = np.array(idx, dtype=np.uint32, copy=True)
    bvh = vb.BvhBuilder(v, gi, ctx).build()
    return bvh, gi


@pytest.mark.parametrize("name,v,idx", small_meshes(), ids=lambda x: x if isinstance(x, str) else None)
def test_blas_bit_exact(ctx, oracle, name, v, idx):
    bvh, gi = gpu_build(ctx, v, idx)
    rc, onodes, oidx, oorder, st = oracle.blas_build(v, idx)
    assert rc == 0
    assert len(bvh.nodes) == len(onodes)
    assert bvh.nodes.tobytes() == onodes.tobytes()          # node AABBs + topology
    assert (gi == oidx).all()                               # permuted index buffer
    assert (ctx.last_order(idx.size // 3) == oorder).all()  # primitive order
    gst = ctx.last_build_stats()
    assert gst["sum_interior_prims"] == st["sum_interior_prims"] and gst["interior_nodes"] == st["interior_nodes"]


# sizes chosen so that the root level runs on every tile size the grid tier picks per level on a 148-SM part
# (256-slot tiles up to ~113 K triangles, then 512, 1024, 2048), plus one just above the grid-tier threshold
@pytest.mark.parametrize("n,seed,edge", [(25_000, 5, 0.02), (100_000, 0, 0.01), (150_000, 7, 0.01), (300_000, 21, 0.01),
                                         (600_000, 9, 0.005)])
def test_blas_bit_exact_grid_tier(ctx, oracle, n, seed, edge):
    v, idx = S.soup(n, seed, edge)
    bvh, gi = gpu_build(ctx, v, idx)
    rc, onodes, oidx, _, _ = oracle.blas_build(v, idx)
    assert rc == 0 and bvh.nodes.tobytes() == onodes.tobytes() and (gi == oidx).all()
    st = ctx.last_build_stats()
    assert st["grid_levels"] >= (2 if n >= 100_000 else 1) and st["cluster_tasks"] == 0  # (the cluster tier is opt-in)
    assert st["big_block_tasks"] > 0 and st["block_tasks"] > 0
    assert st["warp_node_tasks"] > 0 and st["warp_tasks"] > 0 and st["thread_tasks"] > 0


def _build_in_child(meshes, mode, env):
    """Builds `meshes` in a child process with extra environment (tests/_child_build.py); returns ([(nodes, perm)], stats)."""
    import subprocess
    import sys
    import tempfile

    with tempfile.TemporaryDirectory() as td:
        src, dst = os.path.join(td, "in.npz"), os.path.join(td, "out.npz")
        np.savez(src, **{f"v{i}": m[0] for i, m in enumerate(meshes)}, **{f"i{i}": m[1] for i, m in enumerate(meshes)})
        subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "_child_build.py"), src, dst, mode],
                       check=True, env=dict(os.environ, **env), timeout=600)
        d = np.load(dst)
        outs = [(d[f"n{i}"].copy(), d[f"g{i}"].copy()) for i in range(len(meshes))]
        return outs, json.loads(bytes(d["stats"]).decode())


@pytest.mark.parametrize("mode", ["smem", "global"])
def test_blas_cluster_tier_variants_bit_exact(oracle, mode):
    """BVH_CUDA_TC=smem|global (read once per process, hence the child process) routes nodes of 16 385..262 144 triangles
    to the opt-in cluster tier: one node per 16-CTA thread-block cluster, order resident in distributed shared memory
    (k_tcs) or in global memory (k_tc).  Same bytes as the oracle for single meshes around the tier's boundaries (one
    with -0.0 coordinates) and for a forest with five times more cluster tasks than co-resident clusters."""
    nz = S.soup(40_000, 12, 0.02)
    q = np.float32(1 / 64)
    nzv = (np.round(nz[0] / q) * q).astype(np.float32) - np.float32(0.5)
    nzv[nzv == 0] = np.float32(-0.0)
    singles = [S.soup(25_000, 5, 0.02), S.soup(150_000, 7, 0.01), S.soup(300_000, 21, 0.01), (nzv, nz[1])]
    outs, st = _build_in_child(singles, "single", {"BVH_CUDA_TC": mode})
    assert st["cluster_tasks"] >= 1, st
    for (v, idx), (nodes, perm) in zip(singles, outs):
        rc, onodes, oidx, _, _ = oracle.blas_build(v, idx)
        assert rc == 0 and nodes.tobytes() == onodes.tobytes() and (perm == oidx).all()
    meshes = [S.soup(30_000 + 997 * k, 400 + k, 0.02) for k in range(36)]
    outs, st = _build_in_child(meshes, "forest", {"BVH_CUDA_TC": mode})
    assert st["cluster_tasks"] >= len(meshes) and st["grid_levels"] == 0, st
    for (v, idx), (nodes, perm) in zip(meshes, outs):
        rc, onodes, oidx, _, _ = oracle.blas_build(v, idx)
        assert rc == 0 and nodes.tobytes() == onodes.tobytes() and (perm == oidx).all()


def test_blas_context_reuse_across_sizes(ctx, oracle):
    """One context, alternating large and small meshes: the workspace is re-carved per build, so stale bytes of an
    earlier build must never be mistaken for published queue slots (regression: epoch/ready collision)."""
    big = S.soup(6000, 31, 0.05)
    for k in range(6):
        bvh, gi = gpu_build(ctx, *big)
        for n in (4, 5, 9, 40, 300):
            v, idx = S.soup(n, 1000 + 10 * k + n, 0.05)
            bvh, gi = gpu_build(ctx, v, idx)
            rc, onodes, oidx, _, _ = oracle.blas_build(v, idx)
            assert bvh.nodes.tobytes() == onodes.tobytes() and (gi == oidx).all()


def test_blas_bit_exact_tile_scan_path(ctx, oracle):
    """> 512 tiles in one node (N > 1 048 576) switches the grid tier to the explicit per-node tile scan."""
    v, idx = S.soup(1_300_000, 41, 0.004)
    bvh, gi = gpu_build(ctx, v, idx)
    rc, onodes, oidx, _, _ = oracle.blas_build(v, idx)
    assert rc == 0 and bvh.nodes.tobytes() == onodes.tobytes() and (gi == oidx).all()


def test_blas_matches_committed_golden_hashes(ctx):
    gold = json.load(open(GOLDEN))
    for name, (v, idx) in {"uv_sphere_10": S.make_uv_sphere(1.0, 10), "soup_1000_seed7": S.soup(1000, 7, 0.05),
                           "soup_100000_seed0_edge0.01": S.soup(100000, 0, 0.01), "grid_20x20": S.grid_mesh(20, 20)}.items():
        assert sha(v, idx) == gold[name]["input"]
        bvh, gi = gpu_build(ctx, v, idx)
        assert sha(bvh.nodes) == gold[name]["nodes"] and sha(gi) == gold[name]["indices"]
        assert sha(ctx.last_order(idx.size // 3)) == gold[name]["order"]


def test_blas_full_size_dragon_class_properties(ctx, oracle):
    """BASELINE config 2 size (871 K triangles): size-independent properties + idempotence; the bit-exact
    comparison with the (slow, single-threaded) oracle at this size lives in test_blas_dragon_class_vs_oracle."""
    v, idx = S.dragon_class()
    bvh, gi = gpu_build(ctx, v, idx)
    n = idx.size // 3
    order = ctx.last_order(n)
    check_bvh_structure(v, idx, bvh.nodes, gi, order)
    bvh2, gi2 = gpu_build(ctx, v, idx)  # deterministic
    assert bvh2.nodes.tobytes() == bvh.nodes.tobytes() and (gi2 == gi).all()


def test_blas_dragon_class_vs_oracle(ctx, oracle):
    v, idx = S.dragon_class()
    bvh, gi = gpu_build(ctx, v, idx)
    rc, onodes, oidx, _, _ = oracle.blas_build(v, idx)
    assert rc == 0 and bvh.nodes.tobytes() == onodes.tobytes() and (gi == oidx).all()


def test_blas_errors(ctx):
    v, idx = S.soup(4, 1, 0.05)
    with pytest.raises(vb.BvhCudaError) as e:
        vb.BvhBuilder(v, np.zeros(0, np.uint32), ctx).build()  # empty mesh: blas.rs:84 panics
    assert e.value.code == vb.types.EINVAL
    bad = idx.copy(); bad[5] = 10_000
    with pytest.raises(vb.BvhCudaError) as e:
        vb.BvhBuilder(v, bad, ctx).build()
    assert e.value.code == vb.types.EINVAL
    dv = np.tile(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), (8, 1))
    with pytest.raises(vb.BvhCudaError) as e:
        vb.BvhBuilder(dv, np.arange(24, dtype=np.uint32), ctx).build()
    assert e.value.code == vb.types.EDEGENERATE
    # the context stays usable after an error
    bvh, _ = gpu_build(ctx, v, idx)
    assert len(bvh.nodes) == 4


def test_set_bin_number_is_inert_like_the_reference(ctx):
    v, idx = S.soup(500, 3, 0.05)
    a = vb.BvhBuilder(v, idx.copy(), ctx).build()
    b = vb.BvhBuilder(v, idx.copy(), ctx).set_bin_number(32).build()  # blas.rs:64-67,136
    assert a.nodes.tobytes() == b.nodes.tobytes()


@pytest.mark.parametrize("n_inst", [1, 2, 3, 10, 100, 1000, 3000, 12288, 12289, 14001])
def test_tlas_bit_exact(ctx, oracle, n_inst):
    def builder(v, i):
        b, gi = gpu_build(ctx, v, i)
        return b.nodes, gi

    verts, inds, nodes, infos, _ = make_scene(builder)
    inst = S.random_instances(n_inst, 3, seed=n_inst, extent=20.0)
    tl = vb.Tlas.empty(ctx)
    tl.build(inst, infos)
    rc, otl, okids, _, _ = oracle.tlas_build(inst, infos)
    assert rc == 0 and tl.nodes.tobytes() == otl.tobytes() and (tl.children == okids).all()


def test_tlas_ties_on_a_regular_grid(ctx, oracle):
    """Identical meshes on a lattice: many exactly equal union areas, so first-index tie-breaking decides."""
    v, idx = S.make_uv_sphere(1.0, 1)
    b, gi = gpu_build(ctx, v, idx)
    pool = S.MeshPool(lambda vv, ii: (b.nodes, gi))
    pool.add(v, idx)
    verts, inds, nodes, infos = pool.pooled()
    mats = np.stack([S.mat_translation([3.0 * (k % 16), 0.0, 3.0 * (k // 16)]) for k in range(256)])
    inst = S.make_instances(mats, np.zeros(256, dtype=np.int64))
    tl = vb.Tlas.empty(ctx)
    tl.build(inst, infos)
    rc, otl, okids, _, _ = oracle.tlas_build(inst, infos)
    assert tl.nodes.tobytes() == otl.tobytes() and (tl.children == okids).all()


def test_tlas_ties_on_a_large_lattice_cluster_kernel(ctx, oracle):
    """128 x 100 lattice of identical boxes: exercises the cluster (distributed shared memory) chain kernel with exact ties."""
    v, idx = S.make_uv_sphere(1.0, 1)
    b, gi = gpu_build(ctx, v, idx)
    pool = S.MeshPool(lambda vv, ii: (b.nodes, gi))
    pool.add(v, idx)
    verts, inds, nodes, infos = pool.pooled()
    mats = np.stack([S.mat_translation([3.0 * (k % 128), 0.0, 3.0 * (k // 128)]) for k in range(128 * 100)])
    inst = S.make_instances(mats, np.zeros(128 * 100, dtype=np.int64))
    tl = vb.Tlas.empty(ctx)
    tl.build(inst, infos)
    rc, otl, okids, _, _ = oracle.tlas_build(inst, infos)
    assert tl.nodes.tobytes() == otl.tobytes() and (tl.children == okids).all()


def test_tlas_signed_zero_boxes_bit_exact(ctx, oracle):
    """Instance boxes whose faces sit exactly at +-0: the folds keep the first zero they meet (tlas.rs:43,69-70)."""
    v = np.array([[-0.0, 0.0, -0.0], [1, 0.0, 0.0], [0.0, 1, -0.0], [0.0, -0.0, 1]], dtype=np.float32)
    idx = np.array([0, 1, 2, 0, 2, 3, 0, 3, 1, 1, 3, 2], dtype=np.uint32)
    b, gi = gpu_build(ctx, v, idx)
    pool = S.MeshPool(lambda vv, ii: (b.nodes, gi))
    pool.add(v, idx)
    verts, inds, nodes, infos = pool.pooled()
    mats = []
    for k in range(60):
        m = np.eye(4)
        m[0, 0] = -1.0 if k % 2 else 1.0  # mirrored instances turn +0 into -0
        m[1, 1] = -1.0 if k % 3 == 0 else 1.0
        m[:3, 3] = [0.0 if k % 4 else 2.0 * (k // 4), 0.0, 0.0 if k % 5 else 1.0 * k]
        mats.append(m)
    inst = S.make_instances(np.stack(mats), np.zeros(60, dtype=np.int64))
    tl = vb.Tlas.empty(ctx)
    tl.build(inst, infos)
    rc, otl, okids, _, _ = oracle.tlas_build(inst, infos)
    assert tl.nodes.tobytes() == otl.tobytes() and (tl.children == okids).all()


def test_trace_two_level_ids_exact(ctx, oracle):
    def builder(v, i):
        b, gi = gpu_build(ctx, v, i)
        return b.nodes, gi

    verts, inds, nodes, infos, inst = make_scene(builder, n_inst=300)
    tl = vb.Tlas.empty(ctx)
    tl.build(inst, infos)
    scene = vb.Scene(tl.nodes, tl.children, inst, infos, nodes, verts, inds, ctx)
    ro, rd = S.rays_toward_box(200_000, [-20, -20, -20], [20, 20, 20], seed=77)
    t, tri, ins = scene.traverse_tlas(ro, rd)
    occ = scene.occluded(ro, rd)
    ot, otri, oins, _, _ = oracle.trace_scene(tl.nodes, tl.children, inst, infos, nodes, verts, inds, ro, rd,
                                               threads=oracle.max_threads())
    _, _, _, oocc, _ = oracle.trace_scene(tl.nodes, tl.children, inst, infos, nodes, verts, inds, ro, rd, any_hit=True,
                                          threads=oracle.max_threads())
    assert (tri == otri).all() and (ins == oins).all()
    assert np.allclose(t, ot, rtol=T_RTOL, atol=0.0)
    assert (occ == oocc).all() and (occ == (ot < np.float32(1e30))).all()
    assert (otri != 0xFFFFFFFF).sum() > 10_000
    # unnormalised directions (shadow rays, raytraced_shadows.wgsl:98-99) and a finite tmax
    rd2 = (rd * np.float32(7.5)).astype(np.float32)
    t2, tri2, ins2 = scene.traverse_tlas(ro, rd2, tmax=2.0)
    ot2, otri2, oins2, _, _ = oracle.trace_scene(tl.nodes, tl.children, inst, infos, nodes, verts, inds, ro, rd2, tmax=2.0,
                                                 threads=oracle.max_threads())
    assert (tri2 == otri2).all() and (ins2 == oins2).all() and np.allclose(t2, ot2, rtol=T_RTOL, atol=0.0)
    # packed 16+16 children (no side buffer) give the same answer while I <= 32767
    scene_packed = vb.Scene(tl.nodes, None, inst, infos, nodes, verts, inds, ctx)
    t3, tri3, ins3 = scene_packed.traverse_tlas(ro[:20000], rd[:20000])
    assert (tri3 == otri[:20000]).all() and (ins3 == oins[:20000]).all()


def test_async_builds_on_two_contexts_overlap_and_match(oracle):
    """bvh_cuda_blas_build_batch_async_dev: two builds enqueued on two contexts / two streams before either is collected;
    d_result is valid on the stream; nodes and permuted indices equal the oracle's; a second enqueue on a context whose
    build is still pending is refused; a degenerate mesh reports EDEGENERATE from finish()."""
    import torch

    dev = torch.device("cuda", 0)
    meshes = [S.soup(60_000, 31, 0.01), S.displaced_sphere(70, 140, 9)]
    ctxs = [vb.Context(0), vb.Context(0)]
    streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    bufs = []
    torch.cuda.synchronize()
    for (v, idx), c, st in zip(meshes, ctxs, streams):
        n = idx.size // 3
        with torch.cuda.stream(st):
            d_v = torch.from_numpy(v.reshape(-1)).to(dev, non_blocking=False)
            d_i = torch.from_numpy(idx.view(np.int32).copy()).to(dev)
            d_n = torch.zeros(2 * n * 8, dtype=torch.int32, device=dev)
            d_r = torch.full((4,), -1, dtype=torch.int32, device=dev)
        st.synchronize()
        c.blas_build_batch_async_dev(d_v.data_ptr(), v.shape[0], d_i.data_ptr(), 3 * n, 0, 1, d_n.data_ptr(), 2 * n, d_r.data_ptr(), st.cuda_stream)
        bufs.append((d_v, d_i, d_n, d_r, n))
    with pytest.raises(vb.BvhCudaError):  # one build in flight per context
        ctxs[0].blas_build_batch_async_dev(bufs[0][0].data_ptr(), meshes[0][0].shape[0], bufs[0][1].data_ptr(), 3 * bufs[0][4], 0, 1,
                                           bufs[0][2].data_ptr(), 2 * bufs[0][4], 0, streams[0].cuda_stream)
    for (v, idx), c, st, (d_v, d_i, d_n, d_r, n) in zip(meshes, ctxs, streams, bufs):
        # the result words are stream-ordered: readable on the build's stream without calling finish first
        with torch.cuda.stream(st):
            res = d_r.clone()
        st.synchronize()
        m = c.blas_build_finish()
        rc, onodes, oidx, _, _ = oracle.blas_build(v, idx)
        assert rc == 0 and int(res[0]) == m == len(onodes) and int(res[1]) == 0
        assert d_n[: 8 * m].cpu().numpy().tobytes() == onodes.tobytes()
        assert (d_i.cpu().numpy().view(np.uint32) == oidx).all()
    # degenerate input: eight coincident triangles (the reference would not terminate, blas.rs:115,139)
    v = np.zeros((3, 3), dtype=np.float32)
    idx = np.tile(np.arange(3, dtype=np.uint32), 8)
    d_v, d_i = torch.from_numpy(v.reshape(-1)).to(dev), torch.from_numpy(idx.view(np.int32)).to(dev)
    d_n = torch.zeros(16 * 8, dtype=torch.int32, device=dev)
    ctxs[0].blas_build_batch_async_dev(d_v.data_ptr(), 3, d_i.data_ptr(), 24, 0, 1, d_n.data_ptr(), 16, 0, 0)
    with pytest.raises(vb.BvhCudaError) as e:
        ctxs[0].blas_build_finish()
    assert e.value.code == -2


def test_instance_culling_keeps_results_on_a_config3_like_scene(ctx, oracle):
    """Uploaded scenes cull instance visits with tight world boxes (csrc/trace.cu k_instance_wbox).  BASELINE config 3's shape:
    thousands of instances far from the origin, whose TLAS leaf boxes are stretched back to it (tlas.rs:39), so nearly every
    visit is one the culling drops.  Closest-hit ids / t and the exact-order any-hit flags must equal the oracle's, for primary-like
    rays and for rays aimed exactly at mesh vertices of far instances (grazing the instance's own box), incl. a sheared and a
    mirrored instance, a singular one (never culled) and non-uniform scales."""
    def builder(v, i):
        b, gi = gpu_build(ctx, v, i)
        return b.nodes, gi

    verts, inds, nodes, infos, _ = make_scene(builder, n_inst=4)
    inst = S.random_instances(3000, infos.shape[0], seed=33, extent=400.0)
    # odd ones: non-uniform scale, shear, mirror, singular (inverse of a singular matrix is whatever the caller supplies)
    def set_tf(k, m):
        inst["transform"][k] = m.T.reshape(-1).astype(np.float32)
        inst["inv_transform"][k] = np.linalg.inv(m).T.reshape(-1).astype(np.float32)
    base = np.eye(4); base[:3, 3] = [120.0, -80.0, 60.0]
    set_tf(0, base @ np.diag([3.0, 0.4, 1.5, 1.0]))
    sh = np.eye(4); sh[0, 1] = 0.7; sh[2, 0] = -0.3
    set_tf(1, base @ sh)
    set_tf(2, base @ np.diag([-1.0, 1.0, 1.0, 1.0]))
    inst["inv_transform"][3] = 0.0
    inst["inv_transform"][3][15] = 1.0
    tl = vb.Tlas.empty(ctx)
    tl.build(inst, infos)
    scene = vb.Scene(tl.nodes, tl.children, inst, infos, nodes, verts, inds, ctx)
    ro, rd = S.rays_sphere_to_cube(60_000, 900.0, 400.0, seed=13)
    # rays toward vertices of the meshes of far instances, through the instance's forward transform
    rng = np.random.default_rng(8)
    pick = rng.integers(0, len(inst), 40_000)
    vo = infos["vertex_offset"][inst["mesh"][pick]].astype(np.int64)
    nv = np.append(infos["vertex_offset"][1:], verts.shape[0]).astype(np.int64)[inst["mesh"][pick]] - vo
    local = verts[vo + (rng.integers(0, 1 << 30, pick.size) % nv)]
    M = inst["transform"][pick].reshape(-1, 4, 4).transpose(0, 2, 1).astype(np.float64)
    world = np.einsum("nij,nj->ni", M[:, :3, :3], local.astype(np.float64)) + M[:, :3, 3]
    org = (rng.normal(size=world.shape) * 300.0)
    ro = np.concatenate([ro, org.astype(np.float32)])
    rd = np.concatenate([rd, (world - org).astype(np.float32)])
    t, tri, ins = scene.traverse_tlas(ro, rd)
    ot, otri, oins, _, st = oracle.trace_scene(tl.nodes, tl.children, inst, infos, nodes, verts, inds, ro, rd, threads=oracle.max_threads())
    assert (tri == otri).all() and (ins == oins).all() and np.allclose(t, ot, rtol=T_RTOL, atol=0.0)
    assert (otri != 0xFFFFFFFF).sum() > 5_000 and st["instance_visits"] > 50 * len(ro)  # the reference really does enter them all
    # exact-order any-hit (tmax above 1e30 routes every ray through k_trace_scene<true>; the reference then visits EVERY node of
    # every instance it enters, so only a handful of rays: 48 rays cost the oracle ~3e8 node visits)
    sel = np.r_[0:24, len(ro) - 24:len(ro)]
    occ = scene.occluded(ro[sel], rd[sel], tmax=3e38)
    _, _, _, oocc, _ = oracle.trace_scene(tl.nodes, tl.children, inst, infos, nodes, verts, inds, ro[sel], rd[sel], tmax=3e38,
                                          any_hit=True, threads=oracle.max_threads())
    assert (occ == oocc).all()


def test_host_pointer_trace_pipeline_equals_device_calls(ctx):
    """bvh_cuda_trace_any / _closest with host pointers cut the batch into chunks that alternate between two compute streams
    and the scene's two control slots (csrc/api.cu).  1.3 Mi rays = 3 chunks: the results must equal one device-pointer
    launch over the same rays, bit for bit, on both slots, pinned or pageable."""
    import torch

    def builder(v, i):
        b, gi = gpu_build(ctx, v, i)
        return b.nodes, gi

    verts, inds, nodes, infos, inst = make_scene(builder, n_inst=200)
    tl = vb.Tlas.empty(ctx)
    tl.build(inst, infos)
    scene = vb.Scene(tl.nodes, tl.children, inst, infos, nodes, verts, inds, ctx)
    n = (1 << 20) + (1 << 18) + 777
    ro, rd = S.rays_toward_box(n, [-20, -20, -20], [20, 20, 20], seed=91)
    dev = torch.device("cuda", 0)
    d_ro, d_rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    d_occ = torch.empty(n, dtype=torch.uint8, device=dev)
    d_t = torch.empty(n, dtype=torch.float32, device=dev)
    d_tri = torch.empty(n, dtype=torch.int32, device=dev)
    d_ins = torch.empty(n, dtype=torch.int32, device=dev)
    scene.occluded_dev(d_ro.data_ptr(), d_rd.data_ptr(), n, d_occ.data_ptr())
    scene.traverse_tlas_dev(d_ro.data_ptr(), d_rd.data_ptr(), n, d_t.data_ptr(), d_tri.data_ptr(), d_ins.data_ptr())
    torch.cuda.synchronize()
    for pinned in (False, True):
        h_ro = torch.from_numpy(ro).pin_memory().numpy() if pinned else ro
        h_rd = torch.from_numpy(rd).pin_memory().numpy() if pinned else rd
        for _ in range(2):  # the second call reuses both deferral lists and control slots
            occ = scene.occluded(h_ro, h_rd)
            t, tri, ins = scene.traverse_tlas(h_ro, h_rd)
            assert (occ == d_occ.cpu().numpy()).all()
            assert t.tobytes() == d_t.cpu().numpy().tobytes()
            assert (tri == d_tri.cpu().numpy().view(np.uint32)).all() and (ins == d_ins.cpu().numpy().view(np.uint32)).all()
    assert 0.05 < occ.mean() < 0.95


def test_trace_with_staged_tlas_top_equals_default(ctx, oracle):
    """BVH_CUDA_TLAS_TOP=1 (read once per process, hence the child) serves the 255 TLAS nodes nearest the root, node 0 and
    their child pairs from shared memory (csrc/trace.cu TlasTop; off by default because it measured slower,
    profiles/r02_tlas_top_ab.txt).  Same hit ids, same bits of t, same occlusion flags as the default kernels and the
    oracle, with a TLAS larger than the staged part (1 201 nodes) and with the side `children` buffer in use."""
    import subprocess
    import sys
    import tempfile

    def builder(v, i):
        b, gi = gpu_build(ctx, v, i)
        return b.nodes, gi

    verts, inds, nodes, infos, inst = make_scene(builder, n_inst=600)
    tl = vb.Tlas.empty(ctx)
    tl.build(inst, infos)
    scene = vb.Scene(tl.nodes, tl.children, inst, infos, nodes, verts, inds, ctx)
    ro, rd = S.rays_toward_box(100_000, [-20, -20, -20], [20, 20, 20], seed=78)
    t, tri, ins = scene.traverse_tlas(ro, rd)
    occ = scene.occluded(ro, rd)
    _, otri, oins, _, _ = oracle.trace_scene(tl.nodes, tl.children, inst, infos, nodes, verts, inds, ro, rd, threads=oracle.max_threads())
    assert (tri == otri).all() and (ins == oins).all()
    with tempfile.TemporaryDirectory() as td:
        src, dst = os.path.join(td, "in.npz"), os.path.join(td, "out.npz")
        np.savez(src, tlas=tl.nodes, kids=tl.children, inst=inst.view(np.uint8), infos=infos.view(np.uint8), nodes=nodes, verts=verts,
                 inds=inds, ro=ro, rd=rd)
        subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "_child_trace.py"), src, dst],
                       check=True, env=dict(os.environ, BVH_CUDA_TLAS_TOP="1"), timeout=600)
        d = np.load(dst)
        assert (d["tri"] == tri).all() and (d["ins"] == ins).all() and d["t"].tobytes() == t.tobytes() and (d["occ"] == occ).all()


def test_any_hit_order_free_kernel_equals_reference_order(ctx, oracle):
    """trace_any runs an order-free kernel (separate interior / leaf stacks, missed interior children skipped) and hands
    rays with unusable reciprocals to the exact-order kernel.  Its answer must equal traverse_tlas(ray).hit of the
    reference order bit for bit, also for grazing rays (aimed exactly at vertices and edge midpoints, where a
    triangle test can pass although the leaf's own box test fails), finite / huge tmax and zero direction components."""
    def builder(v, i):
        b, gi = gpu_build(ctx, v, i)
        return b.nodes, gi

    verts, inds, nodes, infos, inst = make_scene(builder, n_inst=300)
    tl = vb.Tlas.empty(ctx)
    tl.build(inst, infos)
    scene = vb.Scene(tl.nodes, tl.children, inst, infos, nodes, verts, inds, ctx)
    args = (tl.nodes, tl.children, inst, infos, nodes, verts, inds)
    ro, rd = S.rays_toward_box(200_000, [-20, -20, -20], [20, 20, 20], seed=78)

    def same(sc, a, o, d, tmax=1e30):
        occ = sc.occluded(o, d, tmax=tmax)
        _, _, _, oocc, _ = oracle.trace_scene(*a, o, d, tmax=tmax, any_hit=True, threads=oracle.max_threads())
        assert (occ == oocc).all(), f"{int((occ != oocc).sum())} of {len(o)} rays differ (tmax={tmax})"
        return int(oocc.sum())

    assert same(scene, args, ro, rd) > 10_000
    t_closest, _, _ = scene.traverse_tlas(ro, rd)
    t_med = float(np.median(t_closest[t_closest < np.float32(1e30)]))
    n_all = int((t_closest < np.float32(1e30)).sum())
    n_fin = same(scene, args, ro, rd, tmax=t_med)         # finite tmax: a missed far child is not pushed
    assert 0 < n_fin < n_all
    same(scene, args, ro[:64], rd[:64], tmax=3e38)        # above the miss value (every node is visited): exact kernel
    rz = rd[:60_000].copy()
    rz[::3, 0] = 0.0
    rz[1::3, 1] = 0.0
    rz[2::7] = 0.0
    same(scene, args, ro[:60_000], rz)                    # infinite / NaN reciprocals: deferred to the exact kernel
    # grazing rays on one mesh under an identity instance
    v, idx = S.displaced_sphere(96, 192, 4)
    pool = S.MeshPool(builder)
    pool.add(v, idx)
    pv, pi, pn, pinf = pool.pooled()
    inst1 = S.make_instances(np.eye(4)[None], [0])
    tl1 = vb.Tlas.empty(ctx)
    tl1.build(inst1, pinf)
    scene1 = vb.Scene(tl1.nodes, tl1.children, inst1, pinf, pn, pv, pi, ctx)
    args1 = (tl1.nodes, tl1.children, inst1, pinf, pn, pv, pi)
    rng = np.random.default_rng(5)
    tri3 = np.asarray(idx).reshape(-1, 3)
    org = (rng.normal(size=(150_000, 3)) * 3).astype(np.float32)
    tgt = v[rng.integers(0, len(v), len(org))]
    same(scene1, args1, org, (tgt - org).astype(np.float32))
    t = tri3[rng.integers(0, len(tri3), len(org))]
    mid = ((v[t[:, 0]] + v[t[:, 1]]) * np.float32(0.5)).astype(np.float32)
    same(scene1, args1, org, (mid - org).astype(np.float32))
    # rays that start on a vertex (origin exactly on box planes)
    same(scene1, args1, tgt, (org - tgt).astype(np.float32))
    # an instance rotated by 45 degrees about z: world directions (1, 1, z) have all components non-zero, but the
    # object-space direction gets an exact zero (c*1 - s*1 with c == s in f32) -> deferred at instance entry
    c45 = float(np.float32(np.sqrt(0.5)))
    rot = np.array([[c45, -c45, 0, 0], [c45, c45, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float64)
    inst2 = S.make_instances(np.stack([rot, np.eye(4)]), [0, 0])
    inst2["inv_transform"][0] = np.array([[c45, c45, 0, 0], [-c45, c45, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32).T.reshape(-1)
    tl2 = vb.Tlas.empty(ctx)
    tl2.build(inst2, pinf)
    scene2 = vb.Scene(tl2.nodes, tl2.children, inst2, pinf, pn, pv, pi, ctx)
    args2 = (tl2.nodes, tl2.children, inst2, pinf, pn, pv, pi)
    o2 = (rng.normal(size=(50_000, 3)) * 0.3 + np.array([-3.0, -3.0, 0.0])).astype(np.float32)
    d2 = np.stack([np.ones(len(o2)), np.ones(len(o2)), rng.normal(size=len(o2)) * 0.2], axis=1).astype(np.float32)
    assert same(scene2, args2, o2, d2) > 1000


def test_trace_blas_rust_mode_ids_exact(ctx, oracle):
    v, idx = S.bunny_class()
    bvh, gi = gpu_build(ctx, v, idx)
    ro, rd = S.rays_toward_box(200_000, v.min(0), v.max(0), seed=11)
    t, tri = bvh.traverse_iter_batch(v, gi, ro, rd)
    ot, otri, _ = oracle.trace_blas(bvh.nodes, v, gi, ro, rd, threads=oracle.max_threads())
    assert (tri == otri).all() and np.allclose(t, ot, rtol=T_RTOL, atol=0.0)
    one = bvh.traverse_iter(v, gi, vb.Ray.new(ro[0], rd[0]))
    assert one == (vb.MISS if ot[0] >= np.float32(1e30) else vb.Hit(ot[0]))


def test_trace_empty_and_degenerate_rays(ctx, oracle):
    def builder(v, i):
        b, gi = gpu_build(ctx, v, i)
        return b.nodes, gi

    verts, inds, nodes, infos, inst = make_scene(builder, n_inst=5)
    tl = vb.Tlas.empty(ctx)
    tl.build(inst, infos)
    scene = vb.Scene(tl.nodes, tl.children, inst, infos, nodes, verts, inds, ctx)
    t, tri, ins = scene.traverse_tlas(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert t.size == 0
    # axis-aligned directions (zero components -> infinite reciprocals) and zero-length directions
    ro = np.array([[0, 50, 0], [0, 0, 0], [5, 5, 5], [1, 2, 3]], np.float32)
    rd = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 0], [0, 0, 1]], np.float32)
    t, tri, ins = scene.traverse_tlas(ro, rd)
    ot, otri, oins, _, _ = oracle.trace_scene(tl.nodes, tl.children, inst, infos, nodes, verts, inds, ro, rd)
    assert (tri == otri).all() and (ins == oins).all() and (t == ot).all()


def test_blas_batch_forest_build_equals_per_mesh_builds(ctx, oracle):
    """bvh_cuda_blas_build_batch_dev (MeshPool::add x n in one forest build) must give, mesh by mesh, exactly the nodes
    and permuted indices of separate builds, pooled with the reference's offsets (mesh/mod.rs:310-331)."""
    import torch
    from voidin_b200 import multi_gpu as MG

    meshes = [S.make_plane_mesh(), S.soup(3, 1, 0.05), S.make_uv_sphere(1.0, 1), S.soup(40_000, 2, 0.02), S.soup(300, 3, 0.05),
              S.make_uv_sphere(1.0, 10), S.soup(2500, 4, 0.05), S.soup(1, 5, 0.05), S.displaced_sphere(36, 72, 9), S.soup(33, 6, 0.05)]
    dev = torch.device("cuda", 0)
    tm = [(torch.from_numpy(v.reshape(-1)).to(dev), torch.from_numpy(i.view(np.int32)).to(dev)) for v, i in meshes]
    outs = MG.cuda_build_batch_fn(ctx)(tm)
    st = ctx.last_build_stats()
    total = 0
    for (v, idx), (nodes, perm) in zip(meshes, outs):
        rc, onodes, oidx, _, _ = oracle.blas_build(v, idx)
        assert rc == 0
        assert nodes.cpu().numpy().tobytes() == onodes.tobytes()
        assert (perm.cpu().numpy().view(np.uint32) == oidx).all()
        total += len(onodes)
    assert st["n_nodes"] == total
    # and through the sharded-scene assembly on one rank
    tri = [i.size // 3 for _, i in meshes]
    vc = [v.shape[0] for v, _ in meshes]
    bounds = np.stack([np.stack([v.min(0), v.max(0)]) for v, _ in meshes])
    sc = MG.build_sharded({k: tm[k] for k in range(len(tm))}, len(tm), vc, tri, bounds, None, 0, 1,
                          build_batch_fn=MG.cuda_build_batch_fn(ctx))
    pool = S.MeshPool(lambda v, i: (lambda r: (r[1], r[2]))(oracle.blas_build(v, i)))
    for v, idx in meshes:
        pool.add(v, idx)
    verts, inds, nodes, infos = pool.pooled()
    assert sc.bvh_nodes.cpu().numpy().tobytes() == nodes.tobytes() and sc.mesh_info.tobytes() == infos.tobytes()
    assert (sc.indices.cpu().numpy().view(np.uint32) == inds).all()


def test_blas_forest_with_more_grid_tiles_than_blocks(ctx, oracle):
    """A forest whose grid-tier levels have more tiles than the cooperative kernel has blocks while no single node is large
    enough for the tile-scan path: every block walks several tiles per phase (PA leaves each tile's ballots for PB)."""
    import torch
    from voidin_b200 import multi_gpu as MG

    meshes = [S.soup(30_000 + 997 * k, 400 + k, 0.02) for k in range(36)]  # 36 x ~15 tiles of 2048 > 444 blocks
    dev = torch.device("cuda", 0)
    tm = [(torch.from_numpy(v.reshape(-1)).to(dev), torch.from_numpy(i.view(np.int32)).to(dev)) for v, i in meshes]
    outs = MG.cuda_build_batch_fn(ctx)(tm)
    st = ctx.last_build_stats()
    assert st["grid_levels"] >= 1 and st["grid_nodes"] >= len(meshes)
    for (v, idx), (nodes, perm) in zip(meshes, outs):
        rc, onodes, oidx, _, _ = oracle.blas_build(v, idx)
        assert rc == 0
        assert nodes.cpu().numpy().tobytes() == onodes.tobytes()
        assert (perm.cpu().numpy().view(np.uint32) == oidx).all()


def test_blas_batch_rejects_bad_mesh_table(ctx):
    import torch
    from voidin_b200.types import MESH_INFO

    dev = torch.device("cuda", 0)
    v, idx = S.soup(100, 1, 0.05)
    info = np.zeros(2, dtype=MESH_INFO)
    info["index_count"] = [150, 150]
    info["base_index"] = [0, 151]  # not back to back
    info["vertex_offset"] = [0, 0]
    d_info = torch.from_numpy(info.view(np.uint8).reshape(-1)).to(dev)
    d_v = torch.from_numpy(v.reshape(-1)).to(dev)
    d_i = torch.from_numpy(idx.view(np.int32)).to(dev)
    nodes = torch.empty(2 * 100 * 8, dtype=torch.int32, device=dev)
    with pytest.raises(vb.BvhCudaError) as e:
        ctx.blas_build_batch_dev(d_v.data_ptr(), 300, d_i.data_ptr(), 300, d_info.data_ptr(), 2, nodes.data_ptr(), 200)
    assert e.value.code == vb.types.EINVAL


def test_trace_blas_recursive_variant(ctx, oracle):
    """Bvh::traverse (blas.rs:211-245): Hit(t) once the start node's box is hit (t = t0 when no triangle is closer)."""
    v, idx = S.displaced_sphere(36, 72, 9)
    bvh, gi = gpu_build(ctx, v, idx)
    ro, rd = S.rays_toward_box(50_000, v.min(0) * 1.5, v.max(0) * 1.5, seed=21)
    for node_idx, t0 in [(0, 1e30), (0, 3.0), (2, 1e30)]:
        hit, t = bvh.traverse_batch(v, gi, ro, rd, node_idx, t0)
        ohit, ot = oracle.trace_blas_recursive(bvh.nodes, v, gi, ro, rd, node_idx, t0)
        assert (hit == ohit).all() and np.allclose(t, ot, rtol=T_RTOL, atol=0.0) and (t == ot).all()
    # closest distance agrees with traverse_iter wherever a triangle is hit
    hit, t = bvh.traverse_batch(v, gi, ro, rd)
    ti, tri = bvh.traverse_iter_batch(v, gi, ro, rd)
    has = tri != 0xFFFFFFFF
    assert (t[has] == ti[has]).all() and hit[has].all()
    assert bvh.traverse(v, gi, vb.Ray.new(ro[0], rd[0])) == (vb.Hit(t[0]) if hit[0] else vb.MISS)


def test_ray_generation_bit_exact(ctx, oracle):
    """Primary rays from clip_to_world (bvh_cpu.rs:72-84) and G-buffer shadow rays (raytraced_shadows.wgsl:93-99)."""
    import torch

    dev = torch.device("cuda", 0)
    # a perspective * look-at camera, inverted in float64 and rounded (the reference's proj_view.inverse() is an input here)
    f, aspect, zn, zf = 1.0 / np.tan(0.4), 1.0, 0.1, 100.0
    proj = np.array([[f / aspect, 0, 0, 0], [0, f, 0, 0], [0, 0, zf / (zn - zf), zn * zf / (zn - zf)], [0, 0, -1, 0]])
    eye, tgt, up = np.array([3.0, 2.0, 5.0]), np.array([0.0, 0.5, 0.0]), np.array([0.0, 1.0, 0.0])
    fw = (tgt - eye) / np.linalg.norm(tgt - eye); rt = np.cross(fw, up); rt /= np.linalg.norm(rt); u2 = np.cross(rt, fw)
    view = np.eye(4); view[0, :3], view[1, :3], view[2, :3] = rt, u2, -fw; view[:3, 3] = -view[:3, :3] @ eye
    c2w = np.linalg.inv(proj @ view).astype(np.float32).T.reshape(-1)  # column-major
    for w, h in [(640, 640), (320, 200)]:
        d_ro = torch.empty(w * h * 3, dtype=torch.float32, device=dev)
        d_rd = torch.empty(w * h * 3, dtype=torch.float32, device=dev)
        vb.gen_primary_rays_dev(ctx, c2w, w, h, d_ro.data_ptr(), d_rd.data_ptr())
        oro, ord_ = oracle.gen_primary_rays(c2w, w, h)
        assert (d_ro.cpu().numpy().reshape(-1, 3) == oro).all() and (d_rd.cpu().numpy().reshape(-1, 3) == ord_).all()
    rng = np.random.default_rng(5)
    pos = rng.normal(size=(100_000, 3)).astype(np.float32) * 5
    nor = rng.normal(size=(100_000, 3)).astype(np.float32)
    nor /= np.linalg.norm(nor, axis=1, keepdims=True)
    light = np.array([-3.0, 8.5, 10.0], dtype=np.float32)  # raytraced_shadows.rs:33
    d_p, d_n = torch.from_numpy(pos.reshape(-1)).to(dev), torch.from_numpy(nor.reshape(-1)).to(dev)
    d_ro, d_rd = torch.empty_like(d_p), torch.empty_like(d_p)
    vb.gen_shadow_rays_dev(ctx, d_p.data_ptr(), d_n.data_ptr(), pos.shape[0], light, d_ro.data_ptr(), d_rd.data_ptr())
    oro, ord_ = oracle.gen_shadow_rays(pos, nor, light)
    assert (d_ro.cpu().numpy().reshape(-1, 3) == oro).all() and (d_rd.cpu().numpy().reshape(-1, 3) == ord_).all()
    # area light: every direction ends on the rect
    corners = S.rect_light_corners().astype(np.float32)
    uv = rng.random((pos.shape[0], 2)).astype(np.float32)
    d_uv = torch.from_numpy(uv.reshape(-1)).to(dev)
    vb.gen_area_shadow_rays_dev(ctx, d_p.data_ptr(), d_n.data_ptr(), d_uv.data_ptr(), pos.shape[0], corners, d_ro.data_ptr(), d_rd.data_ptr())
    oro, ord_ = oracle.gen_area_shadow_rays(pos, nor, uv, corners)
    assert (d_ro.cpu().numpy().reshape(-1, 3) == oro).all() and (d_rd.cpu().numpy().reshape(-1, 3) == ord_).all()
    end = oro.astype(np.float64) + ord_.astype(np.float64)
    expect = corners[0] + (corners[1] - corners[0]) * uv[:, :1] + (corners[3] - corners[0]) * uv[:, 1:]
    assert np.abs(end - expect).max() < 1e-4


@pytest.mark.parametrize("n,q", [(20, 1 / 4), (300, 1 / 16), (3000, 1 / 16), (40_000, 1 / 64)])
def test_blas_negative_zero_inputs_bit_exact(ctx, oracle, n, q):
    """Inputs that mix -0.0 and +0.0 on a box face: the reference keeps the first zero its sequential fold meets
    (f32::min/max keep the accumulator, blas.rs:190-198).  The kernels reproduce that on a rare-path (first slot with
    a zero on that face, in the order the node sees when it computes its box), so even the sign bits match."""
    v, idx = S.soup(n, 77, 0.2)
    v = (np.round(v / np.float32(q)) * np.float32(q)).astype(np.float32)  # produces both -0.0 and +0.0
    assert (np.signbit(v) & (v == 0)).any()
    rc, onodes, oidx, oorder, _ = oracle.blas_build(v, idx)
    if rc != 0:
        pytest.skip("degenerate for the reference")
    bvh, gi = gpu_build(ctx, v, idx)
    assert bvh.nodes.tobytes() == onodes.tobytes() and (gi == oidx).all()


def test_animated_instances_frame_loop(ctx, oracle):
    """the per-frame loop compute_update.wgsl implies — rotate the moving instances on the device
    (bit-exact against the oracle twin, given sin/cos), rebuild the TLAS in place, trace through the wrapped scene."""
    import torch
    from voidin_b200.types import INSTANCE, TLAS_NODE

    def builder(v, i):
        b, gi = gpu_build(ctx, v, i)
        return b.nodes, gi

    dev = torch.device("cuda", 0)
    verts, inds, nodes, infos, inst = make_scene(builder, n_inst=200)
    inst = inst.copy()
    inst["transform"][::5, 14] = -20.0  # some instances beyond z = -15 spin the other way (compute_update.wgsl:20-25)
    n = inst.shape[0]
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
    d_inst, d_info, d_nodes = to_dev(inst), to_dev(infos), to_dev(nodes)
    d_verts, d_inds = to_dev(verts), to_dev(inds)
    d_tlas = torch.empty((2 * n + 1) * 32, dtype=torch.uint8, device=dev)
    d_kids = torch.empty((2 * n + 1) * 8, dtype=torch.uint8, device=dev)
    moving = np.arange(0, n, 2, dtype=np.uint32)
    d_ids = torch.from_numpy(moving.view(np.int32)).to(dev)
    ctx.tlas_build_dev(d_inst.data_ptr(), n, d_info.data_ptr(), infos.shape[0], d_tlas.data_ptr(), d_kids.data_ptr())
    scene = vb.Scene(d_tlas.data_ptr(), d_kids.data_ptr(), d_inst.data_ptr(), d_info.data_ptr(), d_nodes.data_ptr(),
                     d_verts.data_ptr(), d_inds.data_ptr(), ctx, device_ptrs=True,
                     counts={"tlas_nodes": 2 * n + 1, "instances": n, "meshes": infos.shape[0], "bvh_nodes": nodes.shape[0],
                             "vertices": verts.shape[0], "indices": inds.size})
    ro, rd = S.rays_toward_box(50_000, [-20, -20, -20], [20, 20, 20], seed=91)
    d_ro, d_rd = torch.from_numpy(ro.reshape(-1)).to(dev), torch.from_numpy(rd.reshape(-1)).to(dev)
    d_t = torch.empty(len(ro), dtype=torch.float32, device=dev)
    d_tri = torch.empty(len(ro), dtype=torch.int32, device=dev)
    d_ins = torch.empty(len(ro), dtype=torch.int32, device=dev)
    ref = inst
    for frame in range(3):
        time_s, dt = 0.7 + 0.016 * frame, 0.016
        ang = np.float32(2.0 * np.sin(time_s * 0.5)) * np.float32(dt)
        s_a, c_a = float(np.sin(np.float32(ang))), float(np.cos(np.float32(ang)))
        vb.instances_rotate_z_dev(ctx, d_inst.data_ptr(), d_ids.data_ptr(), moving.size, s_a, c_a, update_inverse=True)
        ref = oracle.instances_rotate_z(ref, moving, s_a, c_a, update_inverse=True)
        got = d_inst.cpu().numpy().view(INSTANCE)
        assert got.tobytes() == ref.tobytes()
        ctx.tlas_build_dev(d_inst.data_ptr(), n, d_info.data_ptr(), infos.shape[0], d_tlas.data_ptr(), d_kids.data_ptr())
        rc, otl, okids, _, _ = oracle.tlas_build(ref, infos)
        assert d_tlas.cpu().numpy().view(TLAS_NODE).tobytes() == otl.tobytes()
        if frame >= 1:
            scene.instance_boxes(True)  # wrapped scene: the culling boxes follow the rewritten instance buffer (frame 0: off)
        scene.traverse_tlas_dev(d_ro.data_ptr(), d_rd.data_ptr(), len(ro), d_t.data_ptr(), d_tri.data_ptr(), d_ins.data_ptr())
        ot, otri, oins, _, _ = oracle.trace_scene(otl, okids, ref, infos, nodes, verts, inds, ro, rd, threads=oracle.max_threads())
        assert (d_tri.cpu().numpy().view(np.uint32) == otri).all() and (d_ins.cpu().numpy().view(np.uint32) == oins).all()
        assert np.allclose(d_t.cpu().numpy(), ot, rtol=T_RTOL, atol=0.0)
    # the stale-inverse variant is exactly the shader: inv_transform untouched
    before = d_inst.cpu().numpy().view(INSTANCE).copy()
    vb.instances_rotate_z_dev(ctx, d_inst.data_ptr(), None, n, 0.01, float(np.cos(np.float32(0.01))))
    after = d_inst.cpu().numpy().view(INSTANCE)
    assert (after["inv_transform"] == before["inv_transform"]).all() and not (after["transform"] == before["transform"]).all()
    assert after.tobytes() == oracle.instances_rotate_z(before, None, 0.01, float(np.cos(np.float32(0.01)))).tobytes()
