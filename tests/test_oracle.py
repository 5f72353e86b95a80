"""CPU tests of the oracle: the sequential restatement against the independently written scan/bin/rank model,
structural self-checks, the closed-form shuffle against the sequential loop, TLAS invariants, brute-force
traversal, and the committed golden hashes."""
import json
import os

import numpy as np
import pytest

from voidin_b200 import scenes as S

from helpers import check_bvh_structure, make_scene, sha, small_meshes

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_hashes.json")


@pytest.mark.parametrize("name,v,idx", small_meshes(), ids=lambda x: x if isinstance(x, str) else None)
def test_sequential_equals_model_and_structure(oracle, name, v, idx):
    rc, nodes, perm, order, st = oracle.blas_build(v, idx)
    rc2, nodes2, perm2, order2, _ = oracle.blas_build(v, idx, model=True)
    assert rc == 0 and rc2 == 0
    assert nodes.tobytes() == nodes2.tobytes()
    assert (perm == perm2).all() and (order == order2).all()
    check_bvh_structure(v, idx, nodes, perm, order)
    assert len(nodes) == 2 + 2 * st["interior_nodes"]


def _shuffle_closed_form(flags):
    """SURVEY.md Appendix B as used by the kernels, on a bare L/R flag vector; returns (dest[], pivot)."""
    n = len(flags)
    L = np.asarray(flags, dtype=bool)
    nL = int(L.sum())
    RF = np.concatenate([[0], np.cumsum(~L)[:-1]])
    LF = np.arange(n) - RF
    tab = np.zeros(n, dtype=np.int64)
    for j in range(n):
        if L[j]:
            tab[n - 1 - (nL - LF[j] - 1)] = j
        else:
            tab[RF[j]] = j
    f = 0
    for j in range(n):
        lnext = int(L[j + 1]) if j + 1 < n else 0
        lbb = nL - LF[j] - int(L[j]) - lnext
        if j + 2 <= n and lbb >= RF[j]:
            f += 1
    pivot = nL - int(L[f])
    dest = np.zeros(n, dtype=np.int64)
    for j in range(n):
        if j < f:
            dest[j] = j if L[j] else (n - 1 if RF[j] == 0 else tab[n - RF[j]] - 1)
        elif j == f:
            dest[j] = pivot
        else:
            dest[j] = tab[nL - LF[j] - 1] if L[j] else j - 1
    return dest, pivot


def test_closed_form_shuffle_equals_sequential(oracle):
    rng = np.random.default_rng(0)
    for trial in range(3000):
        n = int(rng.integers(1, 40))
        p = rng.random()
        flags = (rng.random(n) < p).astype(np.uint8)
        piv, ids = oracle.shuffle_seq(np.arange(n, dtype=np.uint32), flags)
        dest, pivot = _shuffle_closed_form(flags)
        out = np.empty(n, dtype=np.int64)
        out[dest] = np.arange(n)
        assert pivot == piv, (flags, pivot, piv)
        assert (out == ids).all(), (flags, out, ids)


def _shuffle_kernel_form(flags):
    """The shuffle as the CUDA tiers compute it since the end of round 1 (p_t1_table in blas_grid.cuh, k_t2 / k_t2w in blas_block.cuh):
    the boundary element f from nL and the flags at nL-1, nL, nL+1 instead of a predicate per slot, and a rank->position
    table that only holds the entries that can be looked up (R's at j <= nL, L's at j >= nL); every other entry is
    poisoned here so that a look-up outside the filter fails the test."""
    n = len(flags)
    L = np.asarray(flags, dtype=bool)
    nL = int(L.sum())
    RF = np.concatenate([[0], np.cumsum(~L)[:-1]])
    LF = np.arange(n) - RF
    at = lambda j: int(L[j]) if 0 <= j < n else 0
    l0, l1, l2 = (at(nL - 1) if nL else 0), at(nL), at(nL + 1)
    if nL >= 1 and not (nL + 1 <= n and l0 + l1 <= 1):
        f, lf = nL - 1, l0
    elif not (nL + 2 <= n and l1 + l2 == 0):
        f, lf = nL, l1
    else:
        f, lf = nL + 1, l2
    pivot = nL - lf
    tab = np.full(n, -10**9, dtype=np.int64)
    for j in range(n):
        if L[j]:
            if j >= nL:
                tab[n - 1 - (nL - LF[j] - 1)] = j
        elif j <= nL:
            tab[RF[j]] = j
    dest = np.zeros(n, dtype=np.int64)
    for j in range(n):
        if j < f:
            dest[j] = j if L[j] else (n - 1 if RF[j] == 0 else tab[n - RF[j]] - 1)
        elif j == f:
            dest[j] = pivot
        else:
            dest[j] = tab[nL - LF[j] - 1] if L[j] else j - 1
    return dest, pivot, f


def test_kernel_form_of_the_shuffle_equals_sequential(oracle):
    rng = np.random.default_rng(5)
    cases = [np.zeros(n, np.uint8) for n in (1, 2, 3, 7)] + [np.ones(n, np.uint8) for n in (1, 2, 3, 7)]
    for trial in range(4000):
        n = int(rng.integers(1, 70))
        cases.append((rng.random(n) < rng.random()).astype(np.uint8))
    for flags in cases:
        n = len(flags)
        piv, ids = oracle.shuffle_seq(np.arange(n, dtype=np.uint32), flags)
        dest, pivot, f = _shuffle_kernel_form(flags)
        ref_dest, ref_pivot = _shuffle_closed_form(flags)
        assert pivot == piv == ref_pivot, (flags, pivot, piv)
        assert (dest >= 0).all() and (dest == ref_dest).all(), (flags, dest, ref_dest)
        out = np.empty(n, dtype=np.int64)
        out[dest] = np.arange(n)
        assert (out == ids).all(), (flags, out, ids)


def test_same_axis_planes_only_move_the_suffix(oracle):
    """Groundwork for DESIGN.md section 9 item 1: the 7 planes of an axis are visited in increasing order, the left set of
    plane b contains that of plane b-1, and partition_shuffle leaves [0, pivot) all-left -- so from the second plane of an
    axis on the front cursor walks over the previous pivot's prefix without a swap, and the shuffle is the identity there
    and equal to the shuffle of the suffix alone.  On uniformly spread centroids only ~63 % of an axis' element-shuffles
    touch anything."""
    rng = np.random.default_rng(21)
    active = total = 0
    for trial in range(1500):
        n = int(rng.integers(1, 300))
        k = rng.integers(0, 8, size=n)  # plane counts on one axis, by primitive
        ids = np.arange(n, dtype=np.uint32)
        prev = 0
        for b in range(1, 8):
            fl = (k < b).astype(np.uint8)
            piv, out = oracle.shuffle_seq(ids, fl)
            piv2, out2 = oracle.shuffle_seq(ids[prev:], fl) if n > prev else (0, ids[prev:])
            assert (out[:prev] == ids[:prev]).all() and (out[prev:] == out2).all() and piv == prev + piv2
            total += n
            active += n - prev
            ids, prev = out, piv
    assert 0.55 < active / total < 0.70


def test_frozen_prefix_buffer_plan_equals_full_shuffles(oracle):
    """The ping-pong plan for the suffix-only shuffles (shuffle_models.axis_with_frozen_prefix): pivots and the merged
    order after an axis must be those of the seven full shuffles; the merge pass copies well under one array."""
    import shuffle_models as M
    rng = np.random.default_rng(33)
    copied = total = 0
    for trial in range(800):
        n = int(rng.integers(1, 300))
        k = rng.integers(0, 8, size=n) if trial % 4 else np.full(n, rng.integers(0, 8))
        ids = rng.permutation(n).astype(np.uint32)
        want, pivs = ids, []
        for b in range(1, 8):
            piv, want = oracle.shuffle_seq(want, (k < b).astype(np.uint8))
            pivs.append(piv)
        got_p, got, moved = M.axis_with_frozen_prefix(ids, k, oracle.shuffle_seq)
        assert got_p == pivs and (got == want).all(), (n, k, got_p, pivs)
        copied += moved
        total += n
    assert copied < 0.6 * total


def test_tiled_shuffle_models_equal_sequential(oracle):
    """The tiled two-phase form the grid tier runs, and the one-phase form planned for it (no global table, every move
    executed by the lighter of the two tiles it involves): both must reproduce partition_shuffle on any flag vector and
    tile size; the one-phase form must also keep the number of partner tiles a tile loads small whatever the left
    fraction of the plane is."""
    import shuffle_models as M
    rng = np.random.default_rng(11)
    worst = 0
    cases = []
    for trial in range(600):
        n = int(rng.integers(1, 400))
        tile = int(rng.choice([4, 8, 16, 32]))
        kind = trial % 3
        if kind == 0:    # random order, any left fraction (extreme ones included)
            flags = (rng.random(n) < rng.choice([0.01, 0.05, 0.125, 0.5, 0.875, 0.95, 0.99, rng.random()])).astype(np.uint8)
        elif kind == 1:  # the order the previous plane of the same axis leaves behind: an all-L front, then a mix
            cut = int(rng.integers(0, n + 1))
            flags = np.concatenate([np.ones(cut, np.uint8), (rng.random(n - cut) < rng.random()).astype(np.uint8)])
        else:            # all L / all R / a single odd one out
            flags = np.full(n, trial % 2, np.uint8)
            if n > 2 and trial % 4 < 2:
                flags[int(rng.integers(0, n))] ^= 1
        cases.append((flags, tile))
    for flags, tile in cases:
        n = len(flags)
        piv, ids = oracle.shuffle_seq(np.arange(n, dtype=np.uint32), flags)
        want = np.empty(n, dtype=np.int64)
        want[ids.astype(np.int64)] = np.arange(n)  # want[j] = destination of the element that starts at j
        d2, p2 = M.two_phase(flags, tile)
        assert p2 == piv and (d2 == want).all(), (flags, tile)
        d1, p1, loads, moved = M.one_phase(flags, tile)
        assert p1 == piv and (moved == 1).all() and (d1 == want).all(), (flags, tile, d1, want, moved)
        worst = max(worst, int(loads.max()) if len(loads) else 0)
    assert worst <= 6, worst


def test_pull_form_equals_sequential(oracle):
    """The one-barrier form the grid tier runs when every tile has a block of its own (csrc/blas_grid_pull.cuh): every
    OUTPUT slot finds its source from per-tile counts, the boundary flags and a rank-select in the partner tile's ballots.
    Exhaustive over all flag vectors up to 11 elements, then random ones at any left fraction."""
    import itertools
    import shuffle_models as M
    for n in range(1, 12):
        for bits in itertools.product([0, 1], repeat=n):
            flags = np.array(bits, np.uint8)
            piv, ids = oracle.shuffle_seq(np.arange(n, dtype=np.uint32), flags)
            src, p, _ = M.pull_form(flags, 4, warp_slots=2)
            assert p == piv and (src == ids).all(), (bits, src, ids)
    rng = np.random.default_rng(5)
    for trial in range(300):
        n = int(rng.integers(1, 600))
        tile = int(rng.choice([8, 16, 32, 64]))
        if trial % 2:
            flags = (rng.random(n) < rng.choice([0.01, 0.125, 0.5, 0.875, 0.99, rng.random()])).astype(np.uint8)
        else:
            cut = int(rng.integers(0, n + 1))
            flags = np.concatenate([np.ones(cut, np.uint8), (rng.random(n - cut) < rng.random()).astype(np.uint8)])
        piv, ids = oracle.shuffle_seq(np.arange(n, dtype=np.uint32), flags)
        src, p, _ = M.pull_form(flags, tile)
        assert p == piv and (src == ids).all(), (flags, tile)


def test_empty_and_invalid_inputs(oracle):
    v, idx = S.soup(4, 1, 0.05)
    rc, *_ = oracle.blas_build(v, np.zeros(0, dtype=np.uint32))
    assert rc == oracle.EINVAL
    bad = idx.copy()
    bad[5] = 10_000
    rc, *_ = oracle.blas_build(v, bad)
    assert rc == oracle.EINVAL


def test_degenerate_input_is_reported(oracle):
    # >= 4 triangles with identical centroids: every candidate has an empty left side (blas.rs:139,115)
    v = np.tile(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), (8, 1))
    idx = np.arange(24, dtype=np.uint32)
    rc, _, perm, _, _ = oracle.blas_build(v, idx)
    assert rc == oracle.EDEGENERATE
    assert (perm == idx).all()  # untouched
    rc2, *_ = oracle.blas_build(v, idx, model=True)
    assert rc2 == oracle.EDEGENERATE


def test_tlas_invariants(oracle):
    def builder(v, i):
        rc, nodes, perm, _, _ = oracle.blas_build(v, i)
        assert rc == 0
        return nodes, perm

    verts, inds, nodes, infos, _ = make_scene(builder)
    for n_inst in (1, 2, 3, 17, 300):
        inst = S.random_instances(n_inst, 3, seed=n_inst, extent=20.0)
        rc, tl, kids, calls, pairs = oracle.tlas_build(inst, infos)
        assert rc == 0 and len(tl) == 2 * n_inst + 1  # tlas.rs:32
        # self-merged root (tlas.rs:61): both children of node 0 are the same node
        assert kids[0][0] == kids[0][1]
        assert tl["left_right"][0] == np.uint32((int(kids[0][0]) + (int(kids[0][1]) << 16)) & 0xFFFFFFFF)
        leaves = tl[1:n_inst + 1]
        assert (leaves["left_right"] == 0).all() and (leaves["instance_idx"] == np.arange(n_inst)).all()
        # leaf boxes contain the untransformed local box (tlas.rs:39 quirk)
        mi = infos[inst["mesh"]]
        assert (leaves["min"] <= mi["min"]).all() and (leaves["max"] >= mi["max"]).all()
        # every interior box is the exact union of its children
        for k in range(n_inst + 1, 2 * n_inst + 1):
            a, b = kids[k]
            assert (tl["min"][k] == np.minimum(tl["min"][a], tl["min"][b])).all()
            assert (tl["max"][k] == np.maximum(tl["max"][a], tl["max"][b])).all()
            assert tl["instance_idx"][k] == 0xFFFFFFFF


def test_tlas_best_match_cache_model_is_exact_and_saves_little(oracle):
    """tests/tlas_cache_model.py: Tlas::build with every slot's find_best_match answer cached and repaired per merge.  The model
    reproduces the oracle's bytes (random instances, lattices full of exact ties, the stale slot `a`), which pins the cache
    rules — and it shows why the CUDA chain does not use them: the slots the walk visits next are the ones a merge has just
    invalidated, so 2.8-2.9 full scans per merge remain against the reference's ~3.1 calls."""
    import tlas_cache_model as TM
    from voidin_b200.types import MESH_INFO

    infos = np.zeros(3, dtype=MESH_INFO)
    infos["min"] = [[-1, -1, -1], [-0.5, -2, -0.5], [-3, -0.2, -1]]
    infos["max"] = [[1, 1, 1], [0.5, 2, 0.5], [3, 0.2, 1]]
    cases = [S.random_instances(n, 3, seed=n, extent=e) for n, e in ((1, 20.0), (2, 20.0), (5, 500.0), (17, 20.0), (300, 500.0), (700, 20.0))]
    mats = []
    for x in range(6):
        for y in range(5):
            for z in range(4):
                m = np.eye(4); m[:3, 3] = [3.0 * x, 3.0 * y, 3.0 * z]; mats.append(m)
    cases.append(S.make_instances(np.stack(mats), [0] * len(mats)))
    for inst in cases:
        n = len(inst)
        rc, tl, kids, calls, _ = oracle.tlas_build(inst, infos)
        lo, hi, lr, k2, st = TM.build(tl["min"][1:n + 1].copy(), tl["max"][1:n + 1].copy())
        assert rc == 0 and lo.tobytes() == np.ascontiguousarray(tl["min"]).tobytes() and hi.tobytes() == np.ascontiguousarray(tl["max"]).tobytes()
        assert (lr == tl["left_right"]).all() and (k2 == kids).all()
        assert st["merges"] == n and st["scans"] + st["lookups"] + 1 >= calls - n  # every reference call is a scan or a lookup
        if n >= 300:
            assert st["scans"] > 2.0 * n  # the cache removes well under half of the scans


def test_instance_cull_boxes_never_drop_a_hit(oracle):
    """CPU replica of csrc/trace.cu k_instance_wbox / wbox_miss (the tight per-instance world boxes the exact-order kernels cull
    with): on a config-3-like scene no ray is culled at the instance the oracle reports its hit in -- also for rays aimed exactly
    at mesh vertices -- while nearly every other (ray, instance) pair is.  The GPU suite checks the kernels themselves against the
    oracle's ids (test_instance_culling_keeps_results_on_a_config3_like_scene)."""
    F = np.float32

    def builder(v, i):
        rc, nodes, perm, _, _ = oracle.blas_build(v, i)
        return nodes, perm

    verts, inds, nodes, infos, _ = make_scene(builder, n_inst=4)
    inst = S.random_instances(1200, infos.shape[0], seed=33, extent=400.0)
    rc, tl, kids, _, _ = oracle.tlas_build(inst, infos)
    ro, rd = S.rays_sphere_to_cube(12_000, 900.0, 400.0, seed=13)
    rng = np.random.default_rng(8)
    pick = rng.integers(0, len(inst), 12_000)
    vo = infos["vertex_offset"][inst["mesh"][pick]].astype(np.int64)
    nv = np.append(infos["vertex_offset"][1:], verts.shape[0]).astype(np.int64)[inst["mesh"][pick]] - vo
    local = verts[vo + (rng.integers(0, 1 << 30, pick.size) % nv)]
    M = inst["transform"][pick].reshape(-1, 4, 4).transpose(0, 2, 1).astype(np.float64)
    world = np.einsum("nij,nj->ni", M[:, :3, :3], local.astype(np.float64)) + M[:, :3, 3]
    org = rng.normal(size=world.shape) * 300.0
    ro = np.concatenate([ro, org.astype(F)])
    rd = np.concatenate([rd, (world - org).astype(F)])
    _, tri, ins, _, st = oracle.trace_scene(tl, kids, inst, infos, nodes, verts, inds, ro, rd, threads=oracle.max_threads())
    # the boxes: BLAS root box through the inverse of the linear part of inv_transform, in double, grown by 1e-3 of its size
    A = inst["inv_transform"].reshape(-1, 4, 4).transpose(0, 2, 1).astype(np.float64)
    Linv = np.linalg.inv(A[:, :3, :3])
    root = nodes[infos["bvh_index"][inst["mesh"]]]
    bits = np.array([[(c >> k) & 1 for k in range(3)] for c in range(8)])
    corners = np.where(bits[None, :, :] == 1, root["max"][:, None, :].astype(np.float64), root["min"][:, None, :].astype(np.float64))
    w = np.einsum("nij,ncj->nci", Linv, corners - A[:, None, :3, 3])
    mn, mx = w.min(1), w.max(1)
    amax = np.maximum(np.abs(mn), np.abs(mx)).max(1)
    g = 1e-3 * (amax + (mx - mn).max(1))
    lo, hi, low = (mn - g[:, None]).astype(F), (mx + g[:, None]).astype(F), (amax + g).astype(F)

    def miss(e, d, j):
        inv = (F(1) / d).astype(F)
        m = (F(1e-4) * (np.abs(e).max(1) + low[j])).astype(F)
        a = ((lo[j] - m[:, None] - e) * inv).astype(F)
        b = ((hi[j] + m[:, None] - e) * inv).astype(F)
        tmx, tmn = np.maximum(a, b).min(1), np.minimum(a, b).max(1)
        return (tmx < tmn) | (tmx < 0)

    h = tri != 0xFFFFFFFF
    assert h.sum() > 3_000 and st["instance_visits"] > 100 * len(ro)
    assert not miss(ro[h], rd[h], ins[h]).any()
    ri, jj = rng.integers(0, len(ro), 20_000), rng.integers(0, len(inst), 20_000)
    assert miss(ro[ri], rd[ri], jj).mean() > 0.99


def test_traversal_equals_brute_force(oracle):
    v, idx = S.displaced_sphere(36, 72, 9)
    rc, nodes, perm, _, _ = oracle.blas_build(v, idx)
    ro, rd = S.rays_toward_box(3000, v.min(0), v.max(0), seed=11)
    t, tri, st = oracle.trace_blas(nodes, v, perm, ro, rd)
    bf = oracle.brute_force(v, perm, ro, rd, mode=0)
    assert (t == bf).all()
    hit = tri != 0xFFFFFFFF
    assert hit.sum() > 500 and (t[~hit] == np.float32(1e30)).all()
    # two-level, WGSL semantics, single identity instance == brute force with culling
    pool = S.MeshPool(lambda vv, ii: (nodes, perm))
    pool.add(v, idx)
    verts, inds, bnodes, infos = pool.pooled()
    inst = S.make_instances(np.eye(4)[None], [0])
    rc, tl, kids, _, _ = oracle.tlas_build(inst, infos)
    t2, tri2, ins2, occ, _ = oracle.trace_scene(tl, kids, inst, infos, bnodes, verts, inds, ro, rd)
    bf2 = oracle.brute_force(v, perm, ro, rd, mode=1)
    assert (t2 == bf2).all()
    _, _, _, occ_any, _ = oracle.trace_scene(tl, kids, inst, infos, bnodes, verts, inds, ro, rd, any_hit=True)
    assert (occ_any == (t2 < np.float32(1e30))).all() and (occ == occ_any).all()
    # packed 16+16 children decode to the same traversal when I <= 32767
    t3, tri3, ins3, _, _ = oracle.trace_scene(tl, None, inst, infos, bnodes, verts, inds, ro, rd)
    assert (t3 == t2).all() and (tri3 == tri2).all() and (ins3 == ins2).all()


def test_recursive_traverse_matches_iterative(oracle):
    v, idx = S.displaced_sphere(36, 72, 9)
    rc, nodes, perm, _, _ = oracle.blas_build(v, idx)
    ro, rd = S.rays_toward_box(3000, v.min(0) * 1.5, v.max(0) * 1.5, seed=21)
    hit, t = oracle.trace_blas_recursive(nodes, v, perm, ro, rd)
    ti, tri, _ = oracle.trace_blas(nodes, v, perm, ro, rd)
    has = tri != 0xFFFFFFFF
    assert has.sum() > 300 and (t[has] == ti[has]).all() and hit[has].all()
    # box hit without a triangle hit: Hit(t0), the quirk of blas.rs:235,244
    only_box = hit & ~has
    assert only_box.sum() > 0 and (t[only_box] == np.float32(1e30)).all()


def test_real_meshes_match_the_survey_statistics(oracle):
    """SURVEY.md Appendix F (an independent scratch restatement made during the survey): cube.obj 420 nodes / 209
    interior / depth 11; DamagedHelmet 13 916 nodes / 6 957 interior / depth 19 / 2 474 final-pivot mismatches."""
    from helpers import real_meshes

    got = {n: oracle.blas_build(v, i) for n, v, i in real_meshes()}
    if not got:
        pytest.skip("tests/golden/real_meshes.npz missing")
    rc, nodes, _, _, st = got["real_cube"]
    assert rc == 0 and len(nodes) == 420 and st["interior_nodes"] == 209 and st["max_depth"] == 11 and st["final_pivot_differs"] == 61 and st["nan_candidates"] == 38
    rc, nodes, _, _, st = got["real_helmet0"]
    assert rc == 0 and len(nodes) == 13916 and st["interior_nodes"] == 6957 and st["max_depth"] == 19
    assert st["final_pivot_differs"] == 2474 and st["nan_candidates"] == 1920 and st["candidates"] == 21 * 6957  # the survey counts 22 shuffles per node: 153 054


def test_golden_hashes(oracle):
    """Regression pin of the oracle's own outputs (tests/golden/make_golden.py wrote them; the reference has no
    golden vectors of its own — 'parity unpinned', see oracle/bvh_oracle.cpp header)."""
    with open(GOLDEN) as f:
        gold = json.load(f)
    from golden.make_golden import compute

    now = compute()
    assert now == gold


def test_instances_rotate_z_properties(oracle):
    """compute_update.wgsl:12-27 twin: rotation about z, opposite sense beyond z = -15, inv_transform untouched unless
    asked; with update_inverse the pair stays inverse to each other."""
    inst = S.random_instances(64, 3, seed=4, extent=30.0)
    ang = 0.05
    s, c = float(np.sin(np.float32(ang))), float(np.cos(np.float32(ang)))
    out = oracle.instances_rotate_z(inst, None, s, c)
    assert (out["inv_transform"] == inst["inv_transform"]).all()
    T0 = inst["transform"].reshape(-1, 4, 4).transpose(0, 2, 1).astype(np.float64)
    T1 = out["transform"].reshape(-1, 4, 4).transpose(0, 2, 1).astype(np.float64)
    for k in range(len(inst)):
        a = ang if T0[k][2, 3] > -15.0 else -ang
        R = np.array([[np.cos(a), -np.sin(a), 0, 0], [np.sin(a), np.cos(a), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
        assert np.allclose(T1[k], R @ T0[k], rtol=1e-5, atol=1e-5)
    assert (T0[:, 2, 3] <= -15.0).any() and (T0[:, 2, 3] > -15.0).any()
    ids = np.array([3, 9, 11], dtype=np.uint32)
    out2 = oracle.instances_rotate_z(inst, ids, s, c, update_inverse=True)
    untouched = np.setdiff1d(np.arange(len(inst)), ids)
    assert out2[untouched].tobytes() == inst[untouched].tobytes()
    for k in ids:
        T = out2["transform"][k].reshape(4, 4).T.astype(np.float64)
        I = out2["inv_transform"][k].reshape(4, 4).T.astype(np.float64)
        assert np.allclose(T @ I, np.eye(4), atol=1e-4)
