"""Host-side logic of the N>1 path on CPU: world_size-2 gloo processes build disjoint mesh shards with the oracle as
the build callback, all-gather them, and must assemble exactly the pooled scene a single process produces."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from voidin_b200 import multi_gpu as MG  # noqa: E402
from voidin_b200 import scenes as S  # noqa: E402


def _meshes():
    out = []
    for i, n in enumerate([300, 40, 1200, 7, 650, 90, 2, 500]):
        v, idx = S.soup(n, 500 + i, 0.05) if i % 2 else S.displaced_sphere(6 + i, 12 + 2 * i, 5000 + i)
        out.append((v, idx))
    return out


def _oracle_build_fn():
    from oracle import oracle as O

    def fn(v, idx):
        rc, nodes, perm, _, _ = O.blas_build(v.numpy().reshape(-1, 3), idx.numpy().view(np.uint32))
        assert rc == 0
        return torch.from_numpy(nodes.view(np.int32).reshape(-1).copy()), torch.from_numpy(perm.view(np.int32).copy())

    return fn


def _run(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    meshes = _meshes()
    tri = [m[1].size // 3 for m in meshes]
    vc = [m[0].shape[0] for m in meshes]
    bounds = np.stack([np.stack([m[0].min(0), m[0].max(0)]) for m in meshes])
    plan = MG.lpt_assignment(tri, world)
    mine = {i: (torch.from_numpy(meshes[i][0].reshape(-1)), torch.from_numpy(meshes[i][1].view(np.int32))) for i in plan[rank]}
    sc = MG.build_sharded(mine, len(meshes), vc, tri, bounds, _oracle_build_fn(), rank, world)
    ret[rank] = (sc.vertices.numpy().tobytes(), sc.indices.numpy().tobytes(), sc.bvh_nodes.numpy().tobytes(),
                 sc.mesh_info.tobytes(), sc.n_nodes)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_lpt_assignment_is_balanced_and_deterministic():
    tri = [100, 5, 70, 70, 3, 40, 40, 1]
    plan = MG.lpt_assignment(tri, 3)
    assert sorted(sum(plan, [])) == list(range(8))
    loads = [sum(tri[i] for i in p) for p in plan]
    assert max(loads) - min(loads) <= max(tri)
    assert plan == MG.lpt_assignment(tri, 3)
    assert MG.lpt_assignment(tri, 1) == [list(range(8))]


def test_ray_ranges_partition_the_batch():
    for n, w in [(10, 3), (16, 8), (5, 8), (0, 2), (1 << 20, 7)]:
        r = [MG.ray_range(k, w, n) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1


def test_ray_chunks_are_a_block_cyclic_partition():
    for n, w, c in [(10, 3, 4), (1 << 20, 8, 1 << 16), (100, 2, 7), (5, 8, 64), (0, 2, 16)]:
        seen = np.zeros(n, dtype=np.int32)
        sizes = []
        for r in range(w):
            ch = MG.ray_chunks(r, w, n, c)
            assert all(b % c == 0 and (b // c) % w == r and b < e <= min(n, b + c) for b, e in ch)
            for b, e in ch:
                seen[b:e] += 1
            sizes.append(sum(e - b for b, e in ch))
        assert (seen == 1).all() and max(sizes) - min(sizes) <= c


@pytest.mark.timeout(180)
def test_sharded_build_world2_equals_single_process():
    from oracle import oracle as O

    meshes = _meshes()
    pool = S.MeshPool(lambda v, i: (lambda r: (r[1], r[2]))(O.blas_build(v, i)))
    for v, idx in meshes:
        pool.add(v, idx)
    verts, inds, nodes, infos = pool.pooled()
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_run, args=(2, port, ret), nprocs=2, join=True)
    for rank in (0, 1):
        v, i, n, info, n_nodes = ret[rank]
        assert v == verts.tobytes() and i == inds.tobytes() and n == nodes.tobytes()
        assert info == infos.tobytes()
        assert sum(n_nodes) == nodes.shape[0]
    # world_size 1 degenerates to the plain pooled scene
    tri = [m[1].size // 3 for m in meshes]
    vc = [m[0].shape[0] for m in meshes]
    bounds = np.stack([np.stack([m[0].min(0), m[0].max(0)]) for m in meshes])
    mine = {i: (torch.from_numpy(meshes[i][0].reshape(-1)), torch.from_numpy(meshes[i][1].view(np.int32))) for i in range(len(meshes))}
    sc = MG.build_sharded(mine, len(meshes), vc, tri, bounds, _oracle_build_fn(), 0, 1)
    assert sc.bvh_nodes.numpy().tobytes() == nodes.tobytes() and sc.mesh_info.tobytes() == infos.tobytes()
