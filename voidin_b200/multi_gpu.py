"""Multi-GPU use of the hot path: only where it shards naturally.

* BLAS builds of a multi-mesh scene are independent (`MeshPool::add` is per mesh, crates/pools/src/mesh/mod.rs:309):
  meshes are assigned to ranks (LPT by triangle count), every rank builds its own, then ONE all-gather of
  padded per-rank slabs {vertices | permuted indices | BVH nodes} gives every rank the whole pooled scene; the
  pooled offsets (vertex_offset / base_index / bvh_index) are prefix sums in mesh-id order, computed identically on
  every rank exactly as MeshPool::add does (mesh/mod.rs:310-331).
* The TLAS build is a sequential chain: built redundantly on every rank (deterministic, identical bytes).
* Rays are independent: contiguous ranges per rank, scene replicated, no collective.
* A single BLAS does not shard ("replicas only").

The functions are backend-agnostic (torch tensors on any device, any torch.distributed backend), so the host logic
is exercised on CPU with gloo in tests/; on the B200 box the build callback is the CUDA builder and the backend
NCCL over NVLink."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Sequence

import numpy as np
import torch
import torch.distributed as dist

from .types import MESH_INFO


def lpt_assignment(tri_counts: Sequence[int], world: int) -> list[list[int]]:
    """Longest-processing-time-first: meshes sorted by (triangles desc, id asc) go to the least loaded rank (ties:
    lowest rank).  Deterministic, so every rank derives the same plan without communication."""
    order = sorted(range(len(tri_counts)), key=lambda i: (-int(tri_counts[i]), i))
    load = [0] * world
    out: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(tri_counts[i])
    return [sorted(x) for x in out]


def ray_range(rank: int, world: int, n_rays: int) -> tuple[int, int]:
    """Contiguous shard [begin, end) of rank; sizes differ by at most one."""
    q, rem = divmod(n_rays, world)
    b = rank * q + min(rank, rem)
    return b, b + q + (1 if rank < rem else 0)


def ray_chunks(rank: int, world: int, n_rays: int, chunk: int = 1 << 16) -> list[tuple[int, int]]:
    """Block-cyclic shard of a ray batch: chunk k (rays [k*chunk, (k+1)*chunk)) goes to rank k % world.  A frame's rays are
    not uniform in cost (pixels on the model against pixels on empty ground), so contiguous halves leave one rank with
    all the expensive rays; chunks of 64 Ki rays keep the coherence inside a warp / block and balance the ranks."""
    out = []
    for k in range(rank, (n_rays + chunk - 1) // chunk, world):
        out.append((k * chunk, min(n_rays, (k + 1) * chunk)))
    return out


@dataclass
class PooledScene:
    vertices: torch.Tensor   # float32 [V*3]
    indices: torch.Tensor    # int32   [sum index_count]   (bit pattern of u32)
    bvh_nodes: torch.Tensor  # int32   [M*8]               (32-byte BvhNode records)
    mesh_info: np.ndarray    # MESH_INFO[n_meshes] (host)
    n_nodes: list[int]


def build_sharded(meshes_of_rank: dict[int, tuple[torch.Tensor, torch.Tensor]], n_meshes: int,
                  vert_counts: Sequence[int], tri_counts: Sequence[int], mesh_bounds: np.ndarray,
                  build_fn: Callable[[torch.Tensor, torch.Tensor], tuple[torch.Tensor, torch.Tensor]],
                  rank: int, world: int, group=None, timings: dict | None = None, build_batch_fn=None) -> PooledScene:
    """meshes_of_rank: mesh id -> (vertices float32 [V*3], indices int32 [3N]) for the ids lpt_assignment gives this
    rank.  vert_counts / tri_counts / mesh_bounds ([n_meshes,2,3] min/max over all positions, mesh/mod.rs:22-27) are
    known to every rank.  build_fn(vertices, indices) -> (nodes int32 [M*8], permuted indices int32 [3N])."""
    plan = lpt_assignment(tri_counts, world)
    mine = plan[rank]
    assert sorted(meshes_of_rank) == mine, "rank holds meshes that the LPT plan did not assign to it"
    dev = next(iter(meshes_of_rank.values()))[0].device if meshes_of_rank else torch.device("cpu")

    # ---- local builds ----
    built = {}
    if build_batch_fn is not None and mine:
        # one forest build for all local meshes (bvh_cuda_blas_build_batch_dev)
        outs = build_batch_fn([meshes_of_rank[mid] for mid in mine])
        for mid, (nodes, perm) in zip(mine, outs):
            built[mid] = (meshes_of_rank[mid][0].reshape(-1), perm.reshape(-1), nodes.reshape(-1))
    else:
        for mid in mine:
            v, idx = meshes_of_rank[mid]
            nodes, perm = build_fn(v, idx)
            built[mid] = (v.reshape(-1), perm.reshape(-1), nodes.reshape(-1))
    counts = torch.zeros(n_meshes, dtype=torch.int64, device=dev)
    for mid in mine:
        counts[mid] = built[mid][2].numel() // 8
    if timings is not None and dev.type == "cuda":
        timings["after_build"] = torch.cuda.Event(enable_timing=True)
        timings["after_build"].record()

    # ---- node counts of every mesh (tiny all-reduce), then one all-gather of padded slabs ----
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    n_nodes = [int(x) for x in counts.tolist()]

    def slab_words(r):  # int32 words of rank r's slab
        return sum(3 * vert_counts[m] + 3 * tri_counts[m] + 8 * n_nodes[m] for m in plan[r])

    pad = max(slab_words(r) for r in range(world))
    parts = []
    for mid in mine:
        v, perm, nodes = built[mid]
        parts += [v.view(torch.int32), perm, nodes]
    slab = torch.cat(parts) if parts else torch.zeros(0, dtype=torch.int32, device=dev)
    send = torch.zeros(pad, dtype=torch.int32, device=dev)
    send[: slab.numel()] = slab
    if world > 1:
        recv = torch.empty(world * pad, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(recv, send, group=group)
    else:
        recv = send
    if timings is not None and dev.type == "cuda":
        timings["after_gather"] = torch.cuda.Event(enable_timing=True)
        timings["after_gather"].record()

    # ---- assemble the pooled buffers in mesh-id order (the order MeshPool::add would have seen) ----
    where = {}
    for r in range(world):
        off = r * pad
        for m in plan[r]:
            nv, ni, nn = 3 * vert_counts[m], 3 * tri_counts[m], 8 * n_nodes[m]
            where[m] = (off, off + nv, off + nv + ni, off + nv + ni + nn)
            off += nv + ni + nn
    vs = [recv[where[m][0]:where[m][1]] for m in range(n_meshes)]
    is_ = [recv[where[m][1]:where[m][2]] for m in range(n_meshes)]
    ns = [recv[where[m][2]:where[m][3]] for m in range(n_meshes)]
    vertices = torch.cat(vs).view(torch.float32)
    indices = torch.cat(is_)
    bvh_nodes = torch.cat(ns)
    if timings is not None:
        timings["gather_bytes_received"] = 4 * (world - 1) * pad  # what the collective delivered to this rank
        if dev.type == "cuda":
            timings["after_assemble"] = torch.cuda.Event(enable_timing=True)
            timings["after_assemble"].record()
    info = np.zeros(n_meshes, dtype=MESH_INFO)
    info["min"] = mesh_bounds[:, 0]
    info["max"] = mesh_bounds[:, 1]
    info["index_count"] = 3 * np.asarray(tri_counts, dtype=np.int64)
    info["vertex_offset"] = np.concatenate([[0], np.cumsum(vert_counts)[:-1]])
    info["base_index"] = np.concatenate([[0], np.cumsum(3 * np.asarray(tri_counts, dtype=np.int64))[:-1]])
    info["bvh_index"] = np.concatenate([[0], np.cumsum(n_nodes)[:-1]])
    return PooledScene(vertices, indices, bvh_nodes, info, n_nodes)


def cuda_build_fn(ctx, stream: int = 0):
    """build_fn for CUDA tensors: calls bvh_cuda_blas_build_dev on device pointers (indices are copied first, because
    the builder permutes them in place)."""

    def fn(v: torch.Tensor, idx: torch.Tensor):
        n = idx.numel() // 3
        work = idx.clone()
        nodes = torch.empty(2 * n * 8, dtype=torch.int32, device=v.device)
        m = ctx.blas_build_dev(v.data_ptr(), v.numel() // 3, work.data_ptr(), n, nodes.data_ptr(), 2 * n, stream)
        return nodes[: m * 8], work

    return fn


def cuda_build_batch_fn(ctx, stream: int = 0):
    """build_batch_fn for CUDA tensors: pools the given meshes (vertex_offset / base_index prefix sums, as
    MeshPool::add does), runs ONE forest build and splits the pooled outputs back per mesh."""

    def fn(meshes):
        dev = meshes[0][0].device
        vcount = [m[0].numel() // 3 for m in meshes]
        icount = [m[1].numel() for m in meshes]
        info = np.zeros(len(meshes), dtype=MESH_INFO)
        info["index_count"] = icount
        info["vertex_offset"] = np.concatenate([[0], np.cumsum(vcount)[:-1]])
        info["base_index"] = np.concatenate([[0], np.cumsum(icount)[:-1]])
        d_info = torch.from_numpy(info.view(np.uint8).reshape(-1)).to(dev)
        verts = torch.cat([m[0].reshape(-1) for m in meshes])
        inds = torch.cat([m[1].reshape(-1) for m in meshes])  # fresh buffer: permuted in place
        n_tris = inds.numel() // 3
        nodes = torch.empty(2 * n_tris * 8, dtype=torch.int32, device=dev)
        ctx.blas_build_batch_dev(verts.data_ptr(), verts.numel() // 3, inds.data_ptr(), inds.numel(), d_info.data_ptr(),
                                 len(meshes), nodes.data_ptr(), 2 * n_tris, stream)
        bvh_index = d_info.cpu().numpy().view(MESH_INFO)["bvh_index"].astype(np.int64)
        total = int(ctx.last_build_stats()["n_nodes"])
        ends = np.concatenate([bvh_index[1:], [total]])
        out = []
        for k in range(len(meshes)):
            b0 = int(info["base_index"][k])
            out.append((nodes[8 * int(bvh_index[k]): 8 * int(ends[k])], inds[b0: b0 + icount[k]]))
        return out

    return fn
