#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its config 2: dragon-class SAH BLAS build (Mtris/s) followed by 16 M any-hit
shadow rays toward a rect area light (Mrays/s), on N B200s of one node; plus a config-5 leg (1 024 meshes, builds
sharded over the ranks + NCCL all-gather, 1 Gi rays ray-sharded) in the same JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one batch of synthetic input, inputs resident in HBM:
    BLAS builds of the ground plane and the dragon-class mesh (871 422 triangles), one forest build -> `value` (Mtris/s)
    TLAS build over the 2 instances
    16 M any-hit shadow rays through the two-level BVH                                         -> `rays.value`
Top-level `metric` is the first half of BASELINE.json's metric (build Mtris/s on dragon); the second half (shadow-ray
Mrays/s) is the `rays` object with the same sub-keys.  N > 1: a single BLAS does not shard, so the job is N DISTINCT
dragon-class meshes, one built per rank in place inside its slot of the pooled scene arrays, followed by NCCL
all-gathers (nodes, permuted indices, vertices) so that every GPU holds every mesh: `value` = N x tris / (build +
all-gather), "scaling": "weak" (never a replicated build multiplied by N).  Rays: the FIXED 16 Mi batch is split over
the ranks (strong scaling, no collective) -> `rays.value`; `rays.weak` repeats round 1's 16 Mi rays per GPU.

`e2e` repeats the measurement through the host-pointer C ABI (bvh_cuda_blas_build / bvh_cuda_trace_any) with
pinned HOST buffers, H2D/D2H copies inside the timed region.  `--impl reference` times the CPU oracle port
(oracle/, the restatement of the reference's Rust; the reference itself cannot be compiled: no cargo/rustc).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAYS = 16 * 1024 * 1024
def workload_label():
    from voidin_b200 import models as M
    mesh = "assets/dragon.obj (read by the product's OBJ loader)" if M.find_asset("dragon.obj") else "dragon_class (871422 tris, stand-in for assets/dragon.obj)"
    return f"config2: {mesh} BLAS build + 16Mi any-hit shadow rays toward a rect area light"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def asset_or_standin(name: str, seed=None):
    """(vertices, indices, label): a dropped-in voidin asset ($VOIDIN_ASSETS/<name> or <repo>/assets/<name>) read by the
    product's own loader (ObjModel.import_ = tobj GPU_LOAD_OPTIONS, crates/app/src/models/mod.rs:24), else the synthetic
    stand-in of the same size class (the reference checkout lists bunny.obj / dragon.obj in .MISSING_LARGE_BLOBS).
    seed: a distinct mesh of the same size (weak-scaled builds): the asset uniformly rescaled, or the stand-in reseeded."""
    from voidin_b200 import models as M
    from voidin_b200 import scenes as S

    path = M.find_asset(name)
    if path:
        v, i = M.load_single_mesh(path)
        if seed is not None:
            v = (v * np.float32(1.0 + 1e-3 * seed)).astype(np.float32)
        return v, i, f"assets/{name} ({i.size // 3} tris, loaded by bvh_cuda_model_load_obj)"
    if name == "dragon.obj":
        v, i = S.dragon_class() if seed is None else S.dragon_class(seed=seed)
        return v, i, f"dragon_class ({i.size // 3} tris, stand-in for assets/dragon.obj)"
    v, i = S.bunny_class()
    return v, i, f"bunny_class ({i.size // 3} tris, stand-in for assets/bunny.obj)"


def build_inputs(rank: int, n_rays: int):
    from voidin_b200 import scenes as S

    dv, di, _ = asset_or_standin("dragon.obj")
    pv, pi = S.make_plane_mesh()
    mats, mesh_ids = S.dragon_scene_instances()
    inst = S.make_instances(mats, mesh_ids)
    ro, rd = S.gbuffer_shadow_rays(n_rays, S.world_triangles(dv, di, mats[1]), S.rect_light_corners(), seed=12 + rank)
    return dv, di, pv, pi, inst, ro, rd


def measured_traffic():
    """Per-launch ncu figures (DRAM bytes, L2 peak, issue-slot use) written from the committed summaries under profiles/
    (r02_traffic.json, else round 1's); None when neither exists."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            return json.load(open(p))
    return None


def build_bytes(n_tris, n_verts, S_sum, M):
    """12N + 12V + 36N + 44*S + 32*M + 24N."""
    return 12 * n_tris + 12 * n_verts + 36 * n_tris + 44 * S_sum + 32 * M + 24 * n_tris


def ray_bytes(counters_per_ray, out_bytes):
    """24 + out + 32*pops + 64*interior + 48*tri tests + 192*instance visits."""
    c = counters_per_ray
    return 24 + out_bytes + 32 * c["pops"] + 64 * c["interior_visits"] + 48 * c["triangle_tests"] + 192 * c["instance_visits"]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms from before the warm-up until after the timed region;
    the summary uses the samples whose arrival time falls inside the timed window (falls back to all samples taken
    while the GPU was busy when the window is shorter than the sampling period)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        self.t0 = self.t1 = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        parsed = []
        for ts, r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                parsed.append((ts, float(f[0]), float(f[1]), float(f[2]), [nme for k, nme in enumerate(names) if f[3 + k].lower().startswith("active")]))
            except ValueError:
                continue
        if not parsed:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        inwin = [p for p in parsed if self.t0 is not None and self.t0 - 0.05 <= p[0] <= self.t1 + 0.1]
        src = "timed window"
        if not inwin:
            pmax = max(p[3] for p in parsed)
            inwin = [p for p in parsed if p[3] > 0.6 * pmax] or parsed
            src = "whole run, busy samples (timed window shorter than the sampling period)"
        reasons = sorted({r for p in inwin for r in p[4]})
        return {"sm_mhz": float(np.median([p[1] for p in inwin])), "sm_max_mhz": float(max(p[2] for p in parsed)), "reasons": reasons,
                "samples": len(inwin), "samples_total": len(parsed), "power_w_max": max(p[3] for p in inwin), "source": src}


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference's Rust on the host cores, on the SAME workload and `config` as the GPU arm:
    every step builds the full dragon-class mesh (single-threaded, like the reference) and traces a bounded 1 Mi-ray
    sample of the 16 Mi batch on all host threads.  Rank 0 only."""
    if rank != 0:
        return
    from oracle import oracle as O
    from voidin_b200 import scenes as S

    threads = O.max_threads()
    total = args.steps + args.warmup
    dv, di, _ = asset_or_standin("dragon.obj")
    pv, pi = S.make_plane_mesh()
    mats, mesh_ids = S.dragon_scene_instances()
    inst = S.make_instances(mats, mesh_ids)
    n_rays = 1 << 20
    ro, rd = S.gbuffer_shadow_rays(n_rays, S.world_triangles(dv, di, mats[1]), S.rect_light_corners(), seed=12)
    sample = (f"every step: one full dragon_class + plane build ({di.size // 3 + 2} tris, 1 thread, as the reference builds) and "
              f"{n_rays} of the {args.rays} any-hit rays on {threads} threads (std::thread over rays; the reference itself is single-threaded)")
    bt, rt = [], []
    for step in range(total):
        t0 = time.perf_counter()
        rc, pn, pperm, _, _ = O.blas_build(pv, pi)
        rc, dn, dperm, _, st = O.blas_build(dv, di)
        t1 = time.perf_counter()
        pool = S.MeshPool(None)
        pool.add_built(pv, pperm, pn); pool.add_built(dv, dperm, dn)
        verts, inds, nodes, infos = pool.pooled()
        rc, tl, kids, _, _ = O.tlas_build(inst, infos)
        t2 = time.perf_counter()
        O.trace_scene(tl, kids, inst, infos, nodes, verts, inds, ro, rd, any_hit=True, threads=threads)
        t3 = time.perf_counter()
        if step >= args.warmup:
            bt.append(t1 - t0); rt.append(t3 - t2)
    n_tris = di.size // 3 + 2
    b_ms, r_ms = 1e3 * float(np.mean(bt)), 1e3 * float(np.mean(rt))
    val = n_tris / (b_ms * 1e-3) / 1e6
    rval = n_rays / (r_ms * 1e-3) / 1e6
    line = {
        "impl": "reference", "metric": "dragon_blas_build_Mtris_per_s", "value": val, "unit": "Mtris/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": b_ms + r_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config2_dict(n_tris, args.rays, world),
        "cpu_baseline": {"value": val, "unit": "Mtris/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Mtris/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rays": {"metric": "shadow_ray_Mrays_per_s", "value": rval, "unit": "Mrays/s", "ms": r_ms,
                 "cpu_baseline": {"value": rval, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample},
                 "e2e": {"value": rval, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
        "phase_ms": {"build": b_ms, "trace": r_ms},
        "note": "reference = C++ oracle port of crates/bvh (Rust toolchain absent, so oracle/_ref cannot exist); build is single-threaded like the reference",
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(dv, di, pv, pi, inst):
    """Bounded CPU sample on rank 0 at N=1: one full single-threaded dragon-class build (the reference path is
    single-threaded) and 2^19 any-hit rays on 1 thread; also returns the oracle's per-ray visit counters, which
    feed the traversal byte model."""
    from oracle import oracle as O
    from voidin_b200 import scenes as S

    t0 = time.perf_counter()
    rc, pn, pperm, _, _ = O.blas_build(pv, pi)
    rc, dn, dperm, _, st = O.blas_build(dv, di)
    t1 = time.perf_counter()
    pool = S.MeshPool(None)
    pool.add_built(pv, pperm, pn); pool.add_built(dv, dperm, dn)
    verts, inds, nodes, infos = pool.pooled()
    rc, tl, kids, _, _ = O.tlas_build(inst, infos)
    mats, _ = S.dragon_scene_instances()
    n_s = 1 << 19
    ro, rd = S.gbuffer_shadow_rays(n_s, S.world_triangles(dv, di, mats[1]), S.rect_light_corners(), seed=12)
    t2 = time.perf_counter()
    *_, occ, rs = O.trace_scene(tl, kids, inst, infos, nodes, verts, inds, ro, rd, any_hit=True, threads=1)
    t3 = time.perf_counter()
    n_tris = di.size // 3 + 2
    per_ray = {k: rs[k] / n_s for k in ("pops", "interior_visits", "triangle_tests", "instance_visits")}
    build = {"value": n_tris / (t1 - t0) / 1e6, "unit": "Mtris/s", "cores": 1, "kind": "port",
             "sample": f"one full dragon_class + plane build ({n_tris} tris), single thread, {t1 - t0:.2f} s"}
    rays = {"value": n_s / (t3 - t2) / 1e6, "unit": "Mrays/s", "cores": 1, "kind": "port",
            "sample": f"{n_s} any-hit rays of the same distribution, single thread, {t3 - t2:.2f} s", "occluded_frac": float(occ.mean())}
    return build, rays, per_ray, {"S": st["sum_interior_prims"], "M": int(len(dn))}


def run_small_configs(args, local_rank):
    """BASELINE config 1 (bunny-class BLAS + 1 Mi closest-hit rays, Rust semantics = Bvh::traverse_iter) and config 3
    (TLAS over many randomly transformed instances of three meshes + two-level closest-hit).  Single GPU, one JSON line.
    Secondary workloads: not the line the driver reads."""
    import torch

    import voidin_b200 as vb
    from voidin_b200 import scenes as S
    from voidin_b200.types import MESH_INFO

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = vb.Context(local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    steps = max(1, args.steps)

    def timed(fn, reps):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.mean(ts))

    def up(a, dt):
        return torch.from_numpy(np.ascontiguousarray(a).view(dt).reshape(-1)).to(dev)

    out = {"n_gpus": 1, "data": "synthetic", "dtype": "f32", "steps": steps}
    if args.workload == "soup":
        # BASELINE config 4: 64 Mi-triangle random soup, one BLAS (HBM-bound top levels, grid-wide passes)
        n = (1 << 26) if args.meshes == 1024 else args.meshes  # --meshes doubles as the triangle count here
        g = torch.Generator(device=dev); g.manual_seed(4)
        v0 = torch.rand((n, 1, 3), generator=g, device=dev, dtype=torch.float32)
        e = (torch.rand((n, 2, 3), generator=g, device=dev, dtype=torch.float32) * 2 - 1) * 0.005
        d_v = torch.cat([v0, v0 + e], dim=1).reshape(-1).contiguous()
        del v0, e
        d_i0 = torch.arange(3 * n, device=dev, dtype=torch.int32)
        d_i = d_i0.clone()
        d_nodes = torch.empty(2 * n * 8, dtype=torch.int32, device=dev)
        ctx.set_profiling(True)
        mm = [0]

        def build():
            d_i.copy_(d_i0)
            mm[0] = ctx.blas_build_dev(d_v.data_ptr(), 3 * n, d_i.data_ptr(), n, d_nodes.data_ptr(), 2 * n, stream)

        b_ms = timed(build, max(1, min(steps, 3)))
        st = ctx.last_build_stats()
        bb = build_bytes(n, 3 * n, st["sum_interior_prims"], st["n_nodes"])
        peak, src = peaks()
        # order must be a permutation and every leaf range covered exactly once: cheap device-side property checks
        order = torch.sort(d_i.view(-1, 3)[:, 0] // 3)[0]
        ok_perm = bool((order == torch.arange(n, device=dev, dtype=torch.int32)).all().item())
        tr = measured_traffic() or {}
        soup_tr = tr.get("k_t1_coop_soup8Mi") if n == (1 << 23) else None
        grid_bytes = 44.0 * st["grid_interior_prims"] + 48.0 * st["grid_nodes"]
        grid_gbps = grid_bytes / (max(st["ms_grid"], 1e-9) * 1e-3) / 1e9
        cpu = None
        if not args.no_cpu_baseline:
            # bounded sample: the oracle on a 2^20-triangle soup of the same distribution (the 2^26 one takes 83 minutes,
            # tests/golden/oracle_hashes_large.json)
            from oracle import oracle as O
            ns = 1 << 20
            gs = torch.Generator(device=dev); gs.manual_seed(4)
            sv0 = torch.rand((ns, 1, 3), generator=gs, device=dev, dtype=torch.float32)
            se = (torch.rand((ns, 2, 3), generator=gs, device=dev, dtype=torch.float32) * 2 - 1) * 0.005
            sv = torch.cat([sv0, sv0 + se], dim=1).reshape(-1, 3).cpu().numpy()
            t0 = time.perf_counter()
            rc, _, _, _, _ = O.blas_build(sv, np.arange(3 * ns, dtype=np.uint32))
            dt = time.perf_counter() - t0
            cpu = {"value": ns / dt / 1e6, "unit": "Mtris/s", "cores": 1, "kind": "port",
                   "sample": f"one {ns}-triangle soup of the same distribution, single thread, {dt:.2f} s (rc {rc})"}
        out.update({"metric": "soup_blas_build_Mtris_per_s", "value": n / (b_ms * 1e-3) / 1e6, "unit": "Mtris/s", "ms_per_step": b_ms,
                    "higher_is_better": True,
                    "config": {"workload": f"config4: {n}-triangle random soup, single BLAS",
                               "parity": "nodes / primitive order / permuted indices of the 2^25 and 2^26 soups are pinned to oracle SHA-256 digests "
                                         "(tests/test_gpu_fullsize.py::test_soup_above_2_pow_24_matches_oracle_hash)"},
                    "stats": st, "permutation_ok": ok_perm,
                    "roofline": {"bound": "hbm", "kernel": "k_t1_coop (grid tier)", "achieved": grid_gbps, "peak": peak, "unit": "GB/s",
                                 "frac": grid_gbps / peak, "peak_source": src, "algorithmic_bytes": grid_bytes, "launch_ms": st["ms_grid"],
                                 "share_of_build": st["ms_grid"] / st["ms_total"],
                                 "traffic": (soup_tr["dram_bytes"] if soup_tr else None),
                                 "traffic_over_algorithmic": (soup_tr["dram_bytes"] / grid_bytes if soup_tr else None),
                                 "traffic_note": (soup_tr["reading"] if soup_tr else "ncu DRAM bytes are kept for the 2^23-triangle soup (profiles/r02o_k_t1_coop_soup8M_ncu_summary.txt)"),
                                 "whole_build": {"achieved": bb / (b_ms * 1e-3) / 1e9, "frac": bb / (b_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": bb}},
                    "cpu_baseline": cpu})
    elif args.workload == "bunny":
        v, idx, bunny_label = asset_or_standin("bunny.obj")
        n = idx.size // 3
        d_v, d_i0 = up(v, np.float32), up(idx, np.int32)
        d_i = d_i0.clone()
        d_nodes = torch.zeros(2 * n * 8, dtype=torch.int32, device=dev)
        m = [0]

        def build():
            d_i.copy_(d_i0)
            m[0] = ctx.blas_build_dev(d_v.data_ptr(), v.shape[0], d_i.data_ptr(), n, d_nodes.data_ptr(), 2 * n, stream)

        b_ms = timed(build, steps)
        n_rays = 1 << 20 if args.rays == N_RAYS else args.rays
        ro, rd = S.rays_toward_box(n_rays, v.min(0), v.max(0), seed=11)
        d_ro, d_rd = up(ro, np.float32), up(rd, np.float32)
        d_t = torch.empty(n_rays, dtype=torch.float32, device=dev)
        d_tri = torch.empty(n_rays, dtype=torch.int32, device=dev)
        r_ms = timed(lambda: ctx.trace_blas_dev(d_nodes.data_ptr(), d_v.data_ptr(), d_i.data_ptr(), d_ro.data_ptr(), d_rd.data_ptr(),
                                                n_rays, d_t.data_ptr(), d_tri.data_ptr(), stream), steps)
        st = ctx.last_build_stats()
        peak, src = peaks()
        bb = build_bytes(n, v.shape[0], st["sum_interior_prims"], st["n_nodes"])
        cpu = cpu_rays = None
        per_ray = None
        if not args.no_cpu_baseline:
            from oracle import oracle as O
            t0 = time.perf_counter()
            rc, on, oi, _, _ = O.blas_build(v, idx)
            t1 = time.perf_counter()
            nsr = min(n_rays, 1 << 18)
            _, otri, rst = O.trace_blas(on, v, oi, ro[:nsr], rd[:nsr])
            t2 = time.perf_counter()
            cpu = {"value": n / (t1 - t0) / 1e6, "unit": "Mtris/s", "cores": 1, "kind": "port",
                   "sample": f"one full build of the same mesh ({n} tris), single thread, {t1 - t0:.2f} s (rc {rc})"}
            cpu_rays = {"value": nsr / (t2 - t1) / 1e6, "unit": "Mrays/s", "cores": 1, "kind": "port",
                        "sample": f"{nsr} of the {n_rays} closest-hit rays, Bvh::traverse_iter, single thread, {t2 - t1:.2f} s",
                        "ids_equal_gpu": bool((d_tri[:nsr].cpu().numpy().view(np.uint32) == otri).all())}
            per_ray = {k: rst[k] / nsr for k in rst if isinstance(rst[k], (int, float))}
        out.update({"metric": "bunny_blas_build_Mtris_per_s", "value": n / (b_ms * 1e-3) / 1e6, "unit": "Mtris/s", "ms_per_step": b_ms + r_ms,
                    "higher_is_better": True,
                    "config": {"workload": f"config1: {bunny_label} BLAS + {n_rays} closest-hit rays, Bvh::traverse_iter semantics"},
                    "phase_ms": {"build": b_ms, "trace": r_ms}, "stats": st,
                    "roofline": {"bound": "hbm", "kernel": "whole build (all tiers; a 69 K-triangle mesh spends 4 of its ~20 levels in the grid tier)",
                                 "achieved": bb / (b_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": bb / (b_ms * 1e-3) / 1e9 / peak,
                                 "peak_source": src, "algorithmic_bytes": bb, "traffic": None},
                    "cpu_baseline": cpu,
                    "rays": {"metric": "closest_hit_Mrays_per_s", "value": n_rays / (r_ms * 1e-3) / 1e6, "unit": "Mrays/s",
                             "hit_frac": float((d_tri != -1).float().mean().item()), "cpu_baseline": cpu_rays,
                             "oracle_counters_per_ray": per_ray}})
    else:
        meshes = [asset_or_standin("bunny.obj")[:2], asset_or_standin("dragon.obj")[:2], S.displaced_sphere(62, 124, 3)]  # third = DamagedHelmet-sized stand-in for ferris3d
        n_inst = args.instances
        tm = [(up(v, np.float32), up(i, np.int32)) for v, i in meshes]
        info = np.zeros(len(meshes), dtype=MESH_INFO)
        info["index_count"] = [i.size for _, i in meshes]
        info["vertex_offset"] = np.concatenate([[0], np.cumsum([v.shape[0] for v, _ in meshes])[:-1]])
        info["base_index"] = np.concatenate([[0], np.cumsum(info["index_count"])[:-1]])
        info["min"] = [v.min(0) for v, _ in meshes]
        info["max"] = [v.max(0) for v, _ in meshes]
        d_info = up(info, np.uint8)
        d_verts = torch.cat([a for a, _ in tm])
        d_inds0 = torch.cat([b for _, b in tm])
        d_inds = d_inds0.clone()
        n_tris = d_inds.numel() // 3
        d_nodes = torch.zeros(2 * n_tris * 8, dtype=torch.int32, device=dev)
        mm = [0]

        def build():
            d_inds.copy_(d_inds0)
            mm[0] = ctx.blas_build_batch_dev(d_verts.data_ptr(), d_verts.numel() // 3, d_inds.data_ptr(), d_inds.numel(), d_info.data_ptr(),
                                             len(meshes), d_nodes.data_ptr(), 2 * n_tris, stream)

        b_ms = timed(build, steps)
        inst = S.random_instances(n_inst, len(meshes), seed=3, extent=500.0)
        d_inst = up(inst, np.uint8)
        d_tlas = torch.zeros((2 * n_inst + 1) * 8, dtype=torch.int32, device=dev)
        d_kids = torch.zeros((2 * n_inst + 1) * 2, dtype=torch.int32, device=dev)
        t_ms = timed(lambda: ctx.tlas_build_dev(d_inst.data_ptr(), n_inst, d_info.data_ptr(), len(meshes), d_tlas.data_ptr(), d_kids.data_ptr(), stream), 1)
        scene = vb.Scene(d_tlas.data_ptr(), d_kids.data_ptr(), d_inst.data_ptr(), d_info.data_ptr(), d_nodes.data_ptr(), d_verts.data_ptr(),
                         d_inds.data_ptr(), ctx, device_ptrs=True,
                         counts={"tlas_nodes": 2 * n_inst + 1, "instances": n_inst, "meshes": len(meshes), "bvh_nodes": mm[0],
                                 "vertices": d_verts.numel() // 3, "indices": d_inds.numel()}, stream=stream)
        # tight per-instance world boxes for the exact-order kernels (a wrapped scene: recomputed whenever the instance buffer changes)
        scene.instance_boxes(True, stream)
        n_rays = 1 << 20 if args.rays == N_RAYS else args.rays
        ro, rd = S.rays_sphere_to_cube(n_rays, 1200.0, 500.0, seed=13)
        d_ro, d_rd = up(ro, np.float32), up(rd, np.float32)
        d_t = torch.empty(n_rays, dtype=torch.float32, device=dev)
        d_tri = torch.empty(n_rays, dtype=torch.int32, device=dev)
        d_ins = torch.empty(n_rays, dtype=torch.int32, device=dev)
        r_ms = timed(lambda: scene.traverse_tlas_dev(d_ro.data_ptr(), d_rd.data_ptr(), n_rays, d_t.data_ptr(), d_tri.data_ptr(), d_ins.data_ptr(),
                                                     1e30, stream), max(1, min(steps, 3)))
        cpu = cpu_rays = None
        pairs = None
        if not args.no_cpu_baseline:
            from oracle import oracle as O
            ns = min(n_inst, 32767)  # the oracle's chain is O(I^2): 12 s at 32 767, 116 s at 100 000
            t0 = time.perf_counter()
            rc, otl, okids, calls, opairs = O.tlas_build(inst[:ns], info)
            dt = time.perf_counter() - t0
            cpu = {"value": dt * 1e3, "unit": "ms", "cores": 1, "kind": "port",
                   "sample": f"Tlas::build over the first {ns} of the {n_inst} instances, single thread (rc {rc}), {opairs / dt / 1e9:.2f} G pair evaluations/s"}
            if ns == n_inst:
                pairs = int(opairs)
                cpu["tlas_bytes_equal_gpu"] = bool(d_tlas.cpu().numpy().view(np.uint8).tobytes() == otl.tobytes())
                # two-level closest hit on a bounded sample of the rays (every ray enters a large share of the stretched boxes)
                nsr = min(n_rays, 1 << 13)
                hn = d_nodes[: 8 * mm[0]].cpu().numpy().view(vb.BVH_NODE)
                t1 = time.perf_counter()
                _, otri, oins, _, _ = O.trace_scene(otl, okids, inst, d_info.cpu().numpy().view(MESH_INFO).reshape(-1), hn, d_verts.cpu().numpy().reshape(-1, 3), d_inds.cpu().numpy().view(np.uint32),
                                                   ro[:nsr], rd[:nsr])
                dtr = time.perf_counter() - t1
                cpu_rays = {"value": nsr / dtr / 1e6, "unit": "Mrays/s", "cores": 1, "kind": "port",
                            "sample": f"{nsr} of the {n_rays} two-level closest-hit rays, single thread, {dtr:.2f} s",
                            "ids_equal_gpu": bool((d_tri[:nsr].cpu().numpy().view(np.uint32) == otri).all() and (d_ins[:nsr].cpu().numpy().view(np.uint32) == oins).all())}
        # per-frame loop of the `model` demo (compute_update.wgsl + the TLAS rebuild the reference lacks)
        frame = [0]

        def animated_frame():
            tm_s, dt = 0.7 + 0.016 * frame[0], 0.016
            frame[0] += 1
            ang = np.float32(2.0 * np.sin(tm_s * 0.5)) * np.float32(dt)
            vb.instances_rotate_z_dev(ctx, d_inst.data_ptr(), None, n_inst, float(np.sin(ang)), float(np.cos(ang)), True, stream)
            ctx.tlas_build_dev(d_inst.data_ptr(), n_inst, d_info.data_ptr(), len(meshes), d_tlas.data_ptr(), d_kids.data_ptr(), stream)
            scene.instance_boxes(True, stream)
            scene.traverse_tlas_dev(d_ro.data_ptr(), d_rd.data_ptr(), n_rays, d_t.data_ptr(), d_tri.data_ptr(), d_ins.data_ptr(), 1e30, stream)

        f_ms = timed(animated_frame, max(1, min(steps, 3)))
        out["animated_frame"] = {"ms": f_ms, "what": "rotate all instances (compute_update.wgsl) + TLAS rebuild + the same closest-hit rays"}
        peak, src = peaks()
        tlas_bytes = 144 * n_inst + 48 * len(meshes) + 32 * (2 * n_inst + 1)
        out["roofline"] = {"bound": "hbm", "kernel": "k_tlas_chain / k_tlas_chain_cluster (Tlas::build: ~3.1 I dependent arg-mins over the live slots)",
                           "achieved": tlas_bytes / (t_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": tlas_bytes / (t_ms * 1e-3) / 1e9 / peak,
                           "peak_source": src, "algorithmic_bytes": tlas_bytes, "traffic": None,
                           "pair_evaluations": pairs, "pair_evaluations_per_s": (pairs / (t_ms * 1e-3) if pairs else None),
                           "note": "SURVEY 8(d): B_tlas = 144 I + 48 n_mesh + 32 (2I+1); a latency-bound dependent chain, so the HBM fraction says nothing "
                                   "and pair evaluations/s is the throughput figure"}
        out["cpu_baseline"] = cpu
        out.update({"metric": "tlas_build_ms", "value": t_ms, "unit": "ms", "higher_is_better": False, "ms_per_step": b_ms + t_ms + r_ms,
                    "config": {"workload": f"config3: TLAS over {n_inst} random instances of 3 meshes ({n_tris} tris, forest BLAS build) + {n_rays} two-level closest-hit rays",
                               "note": "Tlas::build seeds every leaf box with the untransformed local mesh box (tlas.rs:39), so instances far from the origin get boxes stretched to the origin and most rays enter a large share of them: reference behaviour, reproduced bit-exactly"},
                    "phase_ms": {"blas_forest_build": b_ms, "tlas_build": t_ms, "trace": r_ms},
                    "blas": {"value": n_tris / (b_ms * 1e-3) / 1e6, "unit": "Mtris/s"},
                    "rays": {"metric": "closest_hit_Mrays_per_s", "value": n_rays / (r_ms * 1e-3) / 1e6, "unit": "Mrays/s",
                             "hit_frac": float((d_tri != -1).float().mean().item()), "cpu_baseline": cpu_rays}})
    print(json.dumps(out), flush=True)


def device_spheres(mids, vside, dev, side):
    """Config-5 meshes generated ON THE DEVICE (1 024 numpy meshes would cost a minute of host time per run): displaced UV
    spheres like scenes.displaced_sphere — same topology and triangle order (scenes._uv_sphere_indices), radius modulated by
    12 random lobes whose parameters come from a CPU generator seeded with 5000 + mesh id — each authored at its lattice
    cell (spacing 3), see the note on Tlas::build's local-box seed in run_scene1024.  Returns ({id: (verts, idx)}, bounds)."""
    import torch

    from voidin_b200 import scenes as S

    uside = 2 * vside
    v = (torch.arange(vside + 1, dtype=torch.float64, device=dev) / vside)[:, None]
    u = (torch.arange(uside + 1, dtype=torch.float64, device=dev) / uside)[None, :]
    theta, phi = 2 * np.pi * u + np.pi, np.pi * v
    d = torch.stack([torch.cos(theta) * torch.sin(phi), (-torch.cos(phi)).expand(vside + 1, uside + 1),
                     torch.sin(theta) * torch.sin(phi)], dim=-1).reshape(-1, 3)          # [V,3]
    idx = torch.from_numpy(S._uv_sphere_indices(vside, uside).view(np.int32)).to(dev)
    stretch = torch.tensor([1.0, 0.7, 1.35], dtype=torch.float64, device=dev)
    out, bounds = {}, {}
    for mid in mids:
        g = np.random.default_rng(5000 + mid)
        w = g.normal(size=(12, 3)); w /= np.linalg.norm(w, axis=1, keepdims=True)
        freq, ph = g.uniform(1.5, 9.0, size=12), g.uniform(0, 2 * np.pi, size=12)
        amp = 0.25 / (1.5 + 0.5 * np.arange(12))
        wt = torch.from_numpy(w).to(dev)
        rad = 1.0 + (torch.from_numpy(amp).to(dev)[None, :] * torch.sin(torch.from_numpy(freq).to(dev)[None, :] * (d @ wt.T) * np.pi
                                                                          + torch.from_numpy(ph).to(dev)[None, :])).sum(dim=1)
        cell = torch.tensor([3.0 * (mid % side), 0.0, 3.0 * (mid // side)], dtype=torch.float64, device=dev)
        verts = (d * rad[:, None] * stretch + cell).to(torch.float32).contiguous()
        bounds[mid] = (verts.min(dim=0)[0].cpu().numpy(), verts.max(dim=0)[0].cpu().numpy())
        out[mid] = (verts.view(-1), idx)
    return out, bounds


def config5_leg(args, ctx, rank, world, dev, stream, steps, warm):
    """BASELINE config 5: `--meshes` distinct synthetic meshes; BLAS builds sharded over the ranks (LPT), ONE all-gather
    of {vertices | permuted indices | nodes} slabs over NCCL so every GPU holds the pooled scene, TLAS per rank, then
    `--c5-rays` any-hit shadow rays (TOTAL, split over the ranks in contiguous ranges: strong scaling, no collective).
    Returns the leg's dict on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist

    import voidin_b200 as vb
    from voidin_b200 import multi_gpu as MG
    from voidin_b200 import scenes as S

    n_meshes, vside = args.meshes, args.mesh_res
    uside = 2 * vside
    tri_counts = [2 * uside * vside - uside] * n_meshes
    vert_counts = [(uside + 1) * (vside + 1)] * n_meshes
    plan = MG.lpt_assignment(tri_counts, world)
    side = int(np.ceil(np.sqrt(n_meshes)))
    # Each mesh is authored at its lattice cell (spacing 3) and instanced with the identity.  Tlas::build seeds every
    # leaf box with the UNTRANSFORMED local mesh box (tlas.rs:39), so translating centred meshes by their instance
    # transform instead would stretch every leaf box from the origin to the instance and make the TLAS useless
    # (measured with the oracle: ~200 instance entries per ray on this scene).
    mine, bnd = device_spheres(plan[rank], vside, dev, side)
    bounds_local = np.zeros((n_meshes, 2, 3), dtype=np.float32)
    for mid, (lo, hi) in bnd.items():
        bounds_local[mid, 0], bounds_local[mid, 1] = lo, hi
    b_t = torch.from_numpy(bounds_local).to(dev)
    if world > 1:
        dist.all_reduce(b_t, op=dist.ReduceOp.SUM)  # every mesh is owned by exactly one rank
    bounds = b_t.cpu().numpy()
    inst = S.make_instances(np.stack([np.eye(4)] * n_meshes), np.arange(n_meshes))
    d_inst = torch.from_numpy(inst.view(np.uint8).reshape(-1)).to(dev)
    d_tlas = torch.zeros((2 * n_meshes + 1) * 8, dtype=torch.int32, device=dev)
    d_kids = torch.zeros((2 * n_meshes + 1) * 2, dtype=torch.int32, device=dev)
    # rays of this rank's contiguous shard of the fixed total, generated on the device from the global ray index
    n_total = args.c5_rays
    rb, re_ = MG.ray_range(rank, world, n_total)
    n_rays = re_ - rb
    d_ro = torch.empty(3 * n_rays, dtype=torch.float32, device=dev)
    d_rd = torch.empty(3 * n_rays, dtype=torch.float32, device=dev)
    ext = 3.0 * side
    for c0 in range(0, n_rays, 1 << 24):  # in chunks: the int64 hashing temporaries are 8x the ray bytes
        c1 = min(n_rays, c0 + (1 << 24))
        gi = torch.arange(rb + c0, rb + c1, device=dev, dtype=torch.int64)

        def u01(salt):
            x = (gi * 6364136223846793005 + (salt * 1442695040888963407) % (1 << 62)) & 0x7FFFFFFFFFFFFFFF
            x = ((x >> 29) ^ x) * 0x2545F4914F6CDD1D & 0x7FFFFFFFFFFFFFFF
            return ((x >> 11) & 0xFFFFFF).to(torch.float32) / 16777216.0

        o = torch.stack([u01(15) * (ext + 6) - 4.5, torch.full((c1 - c0,), -1.6, device=dev), u01(16) * (ext + 6) - 4.5], dim=1)
        tgt = torch.stack([ext / 2 - 1.5 + (u01(17) - 0.5) * 20, torch.full((c1 - c0,), 30.0, device=dev),
                           ext / 2 - 1.5 + (u01(18) - 0.5) * 20], dim=1)
        d_ro[3 * c0:3 * c1] = o.reshape(-1)
        d_rd[3 * c0:3 * c1] = (tgt - o).reshape(-1)
        del gi, o, tgt
    d_occ = torch.empty(n_rays, dtype=torch.uint8, device=dev)
    build_fn = MG.cuda_build_fn(ctx, stream)
    build_batch_fn = MG.cuda_build_batch_fn(ctx, stream)
    state = {}

    def step(ev=None):
        if ev: ev[0].record()
        tm = {}
        sc = MG.build_sharded(mine, n_meshes, vert_counts, tri_counts, bounds, build_fn, rank, world, timings=tm, build_batch_fn=build_batch_fn)
        if ev: ev[1].record()
        d_infos = torch.from_numpy(sc.mesh_info.view(np.uint8).reshape(-1)).to(dev)
        ctx.tlas_build_dev(d_inst.data_ptr(), n_meshes, d_infos.data_ptr(), n_meshes, d_tlas.data_ptr(), d_kids.data_ptr(), stream)
        if ev: ev[2].record()
        scene = vb.Scene(d_tlas.data_ptr(), d_kids.data_ptr(), d_inst.data_ptr(), d_infos.data_ptr(), sc.bvh_nodes.data_ptr(),
                         sc.vertices.data_ptr(), sc.indices.data_ptr(), ctx, device_ptrs=True,
                         counts={"tlas_nodes": 2 * n_meshes + 1, "instances": n_meshes, "meshes": n_meshes,
                                 "bvh_nodes": sum(sc.n_nodes), "vertices": sum(vert_counts), "indices": 3 * sum(tri_counts)}, stream=stream)
        if ev: ev[3].record()
        scene.occluded_dev(d_ro.data_ptr(), d_rd.data_ptr(), n_rays, d_occ.data_ptr(), 1e30, stream)
        if ev: ev[4].record()
        torch.cuda.synchronize()
        state["gather_ms"] = tm["after_build"].elapsed_time(tm["after_gather"])
        state["assemble_ms"] = tm["after_gather"].elapsed_time(tm["after_assemble"])
        state["gather_bytes"] = tm["gather_bytes_received"]
        scene.close()
        del sc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        step()
    barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(steps)]
    gms, ams = [], []
    l0 = ctx.launch_count
    for k in range(steps):
        barrier()
        step(evs[k])
        gms.append(state["gather_ms"]); ams.append(state["assemble_ms"])
    barrier()
    launches = ctx.launch_count - l0
    tot = torch.tensor([sum(e[0].elapsed_time(e[1]) for e in evs), sum(e[1].elapsed_time(e[2]) for e in evs),
                        sum(e[2].elapsed_time(e[3]) for e in evs), sum(e[3].elapsed_time(e[4]) for e in evs), sum(gms), sum(ams)],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    b_ms, t_ms, k_ms, r_ms, g_ms, a_ms = [float(x) / steps for x in tot.tolist()]
    occ = float(d_occ.float().mean().item())
    del d_ro, d_rd, d_occ, mine
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    total_tris = sum(tri_counts)
    gbytes = state["gather_bytes"]
    return {
        "workload": f"config5: {n_meshes} meshes x {tri_counts[0]} tris, BLAS builds sharded (LPT) + one NCCL all-gather, {n_total} any-hit rays in total, ray-sharded",
        "scaling": "strong (fixed scene and fixed ray batch split over the ranks)",
        "build": {"metric": "multi_mesh_blas_build_Mtris_per_s", "value": total_tris / (b_ms * 1e-3) / 1e6, "unit": "Mtris/s",
                  "ms_build_plus_gather": b_ms, "ms_local_forest_build": b_ms - g_ms - a_ms, "ms_all_gather": g_ms, "ms_assemble": a_ms,
                  "all_gather_bytes_received_per_gpu": int(gbytes),
                  "all_gather_GBps_per_gpu": (gbytes / (g_ms * 1e-3) / 1e9) if (world > 1 and g_ms > 0) else None,
                  "collective": "torch.distributed all_gather_into_tensor (NCCL over NVLink), zero-padded per-rank slabs" if world > 1 else "none (one rank)"},
        "tlas_ms": t_ms, "scene_bake_ms": k_ms,
        "rays": {"metric": "shadow_ray_Mrays_per_s", "value": n_total / (r_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms": r_ms,
                 "rays_total": n_total, "occluded_frac_rank0": occ},
        "steps": steps, "warmup": warm, "gpu_launches": int(launches), "tris": total_tris}


def run_scene1024(args, rank, local_rank, world):
    """--workload scene1024: the config-5 leg alone, one JSON line."""
    import torch
    import torch.distributed as dist

    import voidin_b200 as vb

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = vb.Context(local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    leg = config5_leg(args, ctx, rank, world, dev, stream, max(1, min(args.steps, 5)), max(1, min(args.warmup, 2)))
    if rank == 0:
        st = ctx.last_build_stats()  # this rank's shard: one forest build over its meshes
        peak, src = peaks()
        vside = args.mesh_res
        n_shard = leg["tris"] // world
        v_shard = (2 * vside + 1) * (vside + 1) * (args.meshes // world)
        bb = build_bytes(n_shard, v_shard, st["sum_interior_prims"], st["n_nodes"])
        lb_ms = leg["build"]["ms_local_forest_build"]
        leg["roofline"] = {"bound": "hbm", "kernel": "whole forest build of this rank's shard (all tiers)", "achieved": bb / (lb_ms * 1e-3) / 1e9,
                           "peak": peak, "unit": "GB/s", "frac": bb / (lb_ms * 1e-3) / 1e9 / peak, "peak_source": src, "algorithmic_bytes": bb,
                           "traffic": None, "S": st["sum_interior_prims"], "M": st["n_nodes"]}
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import oracle as O
            from voidin_b200 import scenes as S
            sv, si = S.displaced_sphere(vside, 2 * vside, 5000)
            t0 = time.perf_counter()
            rc, _, _, _, _ = O.blas_build(sv, si)
            dt = time.perf_counter() - t0
            cpu = {"value": si.size // 3 / dt / 1e6, "unit": "Mtris/s", "cores": 1, "kind": "port",
                   "sample": f"one displaced sphere of the scene's mesh size ({si.size // 3} tris) of {args.meshes}, single thread, {dt:.2f} s (rc {rc})"}
        leg["cpu_baseline"] = cpu
        print(json.dumps({"metric": leg["build"]["metric"], "value": leg["build"]["value"], "unit": "Mtris/s", "n_gpus": world,
                          "steps": leg["steps"], "warmup": leg["warmup"], "ms_per_step": leg["build"]["ms_build_plus_gather"] + leg["tlas_ms"] + leg["scene_bake_ms"] + leg["rays"]["ms"],
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": leg["workload"]}, "roofline": leg["roofline"], "cpu_baseline": leg["cpu_baseline"],
                          "config5": leg}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def config2_dict(n_tris, n_rays, world):
    """`config` of the headline line; the reference arm prints the same dict (it times the same workload on the CPU)."""
    return {"workload": workload_label(), "tris": n_tris, "rays": n_rays, "l2": "flushed between timed iterations (256 MiB fill)",
            "multi_gpu": ("one rank" if world == 1 else
                          f"build: {world} distinct dragon-class meshes, one per rank (a single BLAS does not shard), then NCCL all-gather of "
                          "vertices / permuted indices / nodes so that every GPU holds all of them (weak scaling); rays: the fixed "
                          f"{n_rays}-ray batch split over the ranks, scene replicated, no collective (strong scaling)")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=N_RAYS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="dragon", choices=["dragon", "scene1024", "bunny", "instances", "soup"],
                    help="dragon = BASELINE config 2 (default, the bench line the driver reads); scene1024 = config 5")
    ap.add_argument("--meshes", type=int, default=1024)
    ap.add_argument("--instances", type=int, default=32767, help="instances workload (config 3): instance count")
    ap.add_argument("--mesh-res", type=int, default=181, help="scene1024: vside of each displaced sphere (181 -> 130682 tris)")
    ap.add_argument("--c5-rays", type=int, default=1 << 30, help="config-5 leg: any-hit rays in TOTAL (split over the ranks)")
    ap.add_argument("--no-config5", action="store_true", help="skip the config-5 leg of the default workload")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "scene1024":
        run_scene1024(args, rank, local_rank, world)
        return
    if args.workload in ("bunny", "instances", "soup"):
        run_small_configs(args, local_rank)
        return
    run_dragon(args, rank, local_rank, world)


def run_dragon(args, rank, local_rank, world):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import voidin_b200 as vb
    from voidin_b200 import multi_gpu as MG
    from voidin_b200 import scenes as S
    from voidin_b200.types import MESH_INFO

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- inputs.  Mesh 2+rank is built by this rank; mesh 2 (rank 0's) is the one every rank traces. ----
    dv0, di0, pv, pi, inst, ro, rd = build_inputs(0, args.rays)
    dv, di = (dv0, di0) if rank == 0 else asset_or_standin("dragon.obj", seed=2 + rank)[:2]
    assert dv.shape == dv0.shape and di.shape == di0.shape
    n_dtris, n_ptris = di.size // 3, pi.size // 3
    n_tris = n_dtris + n_ptris
    n_rays = ro.shape[0]
    # this rank's shard of the fixed batch: block-cyclic 64 Ki-ray chunks (contiguous halves would give one rank all the
    # rays that start on the model and the other the cheap ground rays), packed once, untimed, as a renderer's ray
    # generation would emit them
    shard_idx = np.concatenate([np.arange(b0, e0) for b0, e0 in MG.ray_chunks(rank, world, n_rays)]) if world > 1 else None
    ro_s, rd_s = (ro, rd) if world == 1 else (np.ascontiguousarray(ro[shard_idx]), np.ascontiguousarray(rd[shard_idx]))
    n_shard = ro_s.shape[0]
    ctx = vb.Context(local_rank)
    ctx.set_profiling(True)
    stream = torch.cuda.current_stream().cuda_stream

    def dev_t(a, dtype):
        return torch.from_numpy(np.ascontiguousarray(a).view(dtype).reshape(-1)).to(dev)

    n_verts = pv.shape[0] + dv.shape[0]
    nodes_cap = 2 * n_tris
    d_pi_src, d_di_src = dev_t(pi, np.int32), dev_t(di, np.int32)
    # The builder wants room for 2N nodes (blas.rs:52) and uses ~0.9N: nodes are built into a private buffer, and only the
    # used part travels.  Node counts differ between the ranks' meshes; the slot is sized for the largest (one throw-away
    # build + a scalar all-reduce, before anything is timed; builds are deterministic).
    d_nodes = torch.zeros(nodes_cap * 8, dtype=torch.int32, device=dev)
    d_infos = torch.zeros(2 * 48, dtype=torch.uint8, device=dev)
    infos = np.zeros(2, dtype=MESH_INFO)
    infos["min"][0], infos["max"][0] = pv.min(0), pv.max(0)
    infos["min"][1], infos["max"][1] = dv.min(0), dv.max(0)
    infos["index_count"] = [pi.size, di.size]
    infos["base_index"] = [0, pi.size]
    infos["vertex_offset"] = [0, pv.shape[0]]
    infos_dev_src = torch.from_numpy(infos.view(np.uint8).reshape(-1).copy()).to(dev)
    d_infos.copy_(infos_dev_src)
    tmp_v = torch.cat([dev_t(pv, np.float32), dev_t(dv, np.float32)])
    tmp_i = torch.cat([d_pi_src, d_di_src])
    m_own = ctx.blas_build_batch_dev(tmp_v.data_ptr(), n_verts, tmp_i.data_ptr(), 3 * n_tris, d_infos.data_ptr(), 2, d_nodes.data_ptr(), nodes_cap, stream)
    m_t = torch.tensor([m_own], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(m_t, op=dist.ReduceOp.MAX)
    m_max = int(m_t.item())
    # Pooled arrays of the whole job: one slot per rank, [vertices | indices | nodes] (MeshPool::add order inside a slot:
    # plane, dragon).  Vertices and indices are built on IN PLACE inside the rank's slot and the used nodes are copied
    # next to them, so ONE all-gather sends straight from the slot and delivers straight into the arrays the traversal
    # reads: no packing or re-assembly passes on either side.
    off_i = 3 * n_verts
    off_n = (off_i + 3 * n_tris + 7) // 8 * 8          # nodes 32-byte aligned inside the slot
    slot_w = (off_n + 8 * m_max + 23) // 24 * 24        # slots start 32-byte aligned and on a whole vertex
    pool = torch.zeros(world * slot_w, dtype=torch.int32, device=dev)
    slot = pool[rank * slot_w:(rank + 1) * slot_w]
    d_verts = slot[:off_i].view(torch.float32)
    d_inds = slot[off_i:off_i + 3 * n_tris]
    d_slot_nodes = slot[off_n:off_n + 8 * m_max]
    d_verts.copy_(tmp_v)
    del tmp_v, tmp_i
    d_ro, d_rd = dev_t(ro, np.float32), dev_t(rd, np.float32)
    d_ro_s, d_rd_s = (d_ro, d_rd) if world == 1 else (dev_t(ro_s, np.float32), dev_t(rd_s, np.float32))
    d_inst = torch.from_numpy(inst.view(np.uint8).reshape(-1)).to(dev)
    d_tlas = torch.zeros((2 * 2 + 1) * 8, dtype=torch.int32, device=dev)
    d_kids = torch.zeros((2 * 2 + 1) * 2, dtype=torch.int32, device=dev)
    d_occ = torch.empty(n_rays, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # MeshInfo of rank 0's two meshes inside slot 0 of the pooled array, for the traversal (offsets relative to the slot's
    # vertex / index / node regions; the dragon's nodes follow the plane's two, blas.rs:93): the scene of BASELINE
    # config 2, identical on every rank
    tinfos = np.zeros(2, dtype=MESH_INFO)
    tinfos["min"][0], tinfos["max"][0] = pv.min(0), pv.max(0)
    tinfos["min"][1], tinfos["max"][1] = dv0.min(0), dv0.max(0)
    tinfos["index_count"] = [pi.size, di0.size]
    tinfos["base_index"] = [0, pi.size]
    tinfos["vertex_offset"] = [0, pv.shape[0]]
    tinfos["bvh_index"] = [0, 2]
    d_tinfos = torch.from_numpy(tinfos.view(np.uint8).reshape(-1).copy()).to(dev)

    state = {}

    def step_dev(ev=None):
        """One step on device-resident inputs.  ev: optional list of 5 torch events."""
        d_inds[: 3 * n_ptris].copy_(d_pi_src)      # builds permute indices in place: restore the unpermuted input
        d_inds[3 * n_ptris:].copy_(d_di_src)
        if ev: ev[0].record()
        # MeshPool::add x 2 as ONE forest build over the rank's slot (fills MeshInfo.bvh_index on the device)
        d_infos.copy_(infos_dev_src)
        m_total = ctx.blas_build_batch_dev(d_verts.data_ptr(), n_verts, d_inds.data_ptr(), 3 * n_tris, d_infos.data_ptr(), 2,
                                           d_nodes.data_ptr(), nodes_cap, stream)
        state["stats"] = ctx.last_build_stats()
        if ev: ev[1].record()
        d_slot_nodes.copy_(d_nodes[: 8 * m_max])
        if world > 1:  # every GPU gets every rank's BLAS (and its geometry): one in-place all-gather, slot to slot
            dist.all_gather_into_tensor(pool, slot)
        if ev: ev[2].record()
        ctx.tlas_build_dev(d_inst.data_ptr(), 2, d_tinfos.data_ptr(), 2, d_tlas.data_ptr(), d_kids.data_ptr(), stream)
        if ev: ev[3].record()
        if "scene" not in state:
            state["scene"] = vb.Scene(d_tlas.data_ptr(), d_kids.data_ptr(), d_inst.data_ptr(), d_tinfos.data_ptr(), pool.data_ptr() + 4 * off_n,
                                      pool.data_ptr(), pool.data_ptr() + 4 * off_i, ctx, device_ptrs=True,
                                      counts={"tlas_nodes": 5, "instances": 2, "meshes": 2, "bvh_nodes": m_max, "vertices": n_verts,
                                              "indices": 3 * n_tris}, stream=stream)
        else:
            state["scene"].refresh_dev(m_max, stream)  # re-bake the traversal copy of the freshly permuted triangles
        state["scene"].occluded_dev(d_ro_s.data_ptr(), d_rd_s.data_ptr(), n_shard, d_occ.data_ptr(), 1e30, stream)
        if ev: ev[4].record()
        state["M"] = m_total

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- CPU baseline (rank 0, N=1 only) ----
    cpu_build = cpu_rays = None
    per_ray = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_build, cpu_rays, per_ray, _ = cpu_baseline_leg(dv, di, pv, pi, inst)

    sampler = ClockSampler(local_rank)
    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier()

    # ---- device-resident timing ----
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    launches0 = ctx.launch_count
    barrier()
    t_wall0 = time.perf_counter()
    phase_lib = []
    for k in range(args.steps):
        flush.fill_(k & 0xFF)  # L2 flush between timed iterations (untimed: outside the event brackets)
        step_dev(evs[k])
        phase_lib.append(state["stats"])
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count - launches0
    sampler.window(t_wall0, t_wall0 + t_wall)
    el = lambda a, b: float(np.sum([e[a].elapsed_time(e[b]) for e in evs]))
    tot = torch.tensor([el(0, 1), el(1, 2), el(2, 3), el(3, 4), el(0, 4), el(0, 2)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    b_tot, g_tot, t_tot, r_tot, s_tot, bg_tot = [float(x) for x in tot.tolist()]
    occ_frac = float(d_occ[:n_shard].float().mean().item())
    host_scene = state["scene"]

    def timed_trace(d_o, d_d, n, reps=4):
        ts = []
        for k in range(reps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); host_scene.occluded_dev(d_o, d_d, n, d_occ.data_ptr(), 1e30, stream); e1.record()
            torch.cuda.synchronize()
            if k > 0: ts.append(e0.elapsed_time(e1))
        t = torch.tensor([float(np.mean(ts))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # the collective alone, back to back (no build in front of it): what NCCL needs for this message size
    bare_ms = None
    if world > 1:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dist.all_gather_into_tensor(pool, slot)
        e1.record(); torch.cuda.synchronize()
        bt_ = torch.tensor([e0.elapsed_time(e1) / 5], dtype=torch.float64, device=dev)
        dist.all_reduce(bt_, op=dist.ReduceOp.MAX)
        bare_ms = float(bt_.item())
    # weak-scaled rays (N > 1 only): every rank traces the whole batch
    weak_ms = timed_trace(d_ro.data_ptr(), d_rd.data_ptr(), n_rays) if world > 1 else None
    # incoherent variant: the rank's shard under one fixed random permutation (device-side gather, untimed)
    gperm = torch.Generator(device="cpu"); gperm.manual_seed(2012 + rank)
    perm = torch.randperm(n_shard, generator=gperm).to(dev)
    d_ro_i = d_ro_s.view(-1, 3)[perm].contiguous().view(-1); d_rd_i = d_rd_s.view(-1, 3)[perm].contiguous().view(-1)
    del perm
    inc_ms = timed_trace(d_ro_i.data_ptr(), d_rd_i.data_ptr(), n_shard)
    del d_ro_i, d_rd_i

    # ---- end-to-end through the host-pointer C ABI, pinned host buffers ----
    h_dv = torch.from_numpy(dv.reshape(-1)).pin_memory()
    h_di = torch.from_numpy(di.view(np.int32)).pin_memory()
    h_di_work = torch.empty_like(h_di).pin_memory()
    h_nodes = torch.empty(2 * n_dtris * 8, dtype=torch.int32).pin_memory()
    h_ro, h_rd = torch.from_numpy(ro.reshape(-1)).pin_memory(), torch.from_numpy(rd.reshape(-1)).pin_memory()
    h_ro_s, h_rd_s = (h_ro, h_rd) if world == 1 else (torch.from_numpy(ro_s.reshape(-1)).pin_memory(), torch.from_numpy(rd_s.reshape(-1)).pin_memory())
    h_occ = torch.empty(n_rays, dtype=torch.uint8).pin_memory()
    lib = ctx.lib
    e2e_steps = max(3, min(args.steps, 5))
    eb, er, ew = [], [], []
    for k in range(e2e_steps + 1):
        h_di_work.copy_(h_di)
        m = C.c_uint32(0)
        barrier()
        t0 = time.perf_counter()
        ctx.check(lib.bvh_cuda_blas_build(ctx.h, h_dv.data_ptr(), dv.shape[0], h_di_work.data_ptr(), n_dtris, h_nodes.data_ptr(),
                                          2 * n_dtris, C.byref(m)))
        t1 = time.perf_counter()
        ctx.check(lib.bvh_cuda_trace_any(ctx.h, host_scene.h, h_ro_s.data_ptr(), h_rd_s.data_ptr(), n_shard, 1e30, h_occ.data_ptr()))
        t2 = time.perf_counter()
        if world > 1:
            barrier()
            t3 = time.perf_counter()
            ctx.check(lib.bvh_cuda_trace_any(ctx.h, host_scene.h, h_ro.data_ptr(), h_rd.data_ptr(), n_rays, 1e30, h_occ.data_ptr()))
            t4 = time.perf_counter()
        else:
            t3 = t4 = 0.0
        if k > 0:
            eb.append(t1 - t0); er.append(t2 - t1); ew.append(t4 - t3)
    e2e_t = torch.tensor([float(np.mean(eb)), float(np.mean(er)), float(np.mean(ew))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_b, e2e_r, e2e_w = [float(x) for x in e2e_t.tolist()]
    m_dragon = int(m.value)
    clocks = sampler.stop()

    # ---- config-5 leg (sharded multi-mesh build + all-gather, strong-scaled rays) ----
    host_scene.close()
    del d_ro, d_rd, d_ro_s, d_rd_s, flush
    torch.cuda.empty_cache()
    c5 = None if args.no_config5 else config5_leg(args, ctx, rank, world, dev, stream, steps=2, warm=1)

    if rank == 0:
        peak, peak_src = peaks()
        st = phase_lib[-1]  # stats of the step's forest build (ground plane + dragon-class mesh)
        bb = build_bytes(n_tris, n_verts, st["sum_interior_prims"], st["n_nodes"])
        b_ms_step, g_ms_step = b_tot / args.steps, g_tot / args.steps
        bg_ms_step = bg_tot / args.steps
        r_ms_step = r_tot / args.steps
        value = world * n_tris / (bg_ms_step * 1e-3) / 1e6
        rvalue = n_rays / (r_ms_step * 1e-3) / 1e6
        lib_ms = {k: float(np.mean([p[k] for p in phase_lib])) for k in ("ms_setup", "ms_grid", "ms_cluster", "ms_big_block", "ms_block", "ms_warp_node", "ms_warp", "ms_thread", "ms_emit", "ms_total")}
        tr = measured_traffic()
        # Dominant kernel = k_t1_coop (grid tier, one launch per build).  Its algorithmic bytes: per level and per primitive
        # of the nodes it splits, id 4 + centroid 12 + AABB 24 read, id 4 written (44 B), plus one 48-byte record
        # per node; S_grid and the node count are counted on the device (BvhCudaBuildStats.grid_*).
        s_grid = float(np.mean([p["grid_interior_prims"] for p in phase_lib]))
        n_grid = float(np.mean([p["grid_nodes"] for p in phase_lib]))
        grid_bytes = 44.0 * s_grid + 48.0 * n_grid
        grid_ms = lib_ms["ms_grid"] + lib_ms["ms_cluster"]
        grid_gbps = grid_bytes / (grid_ms * 1e-3) / 1e9
        build_roof = {"bound": "hbm", "achieved": grid_gbps, "peak": peak, "unit": "GB/s", "frac": grid_gbps / peak,
                      "traffic": (tr["k_t1_coop"] if tr else None), "peak_source": peak_src,
                      "kernel": "k_t1_coop (grid tier: every node above 16384 triangles, one cooperative launch per build)",
                      "launch_ms": grid_ms, "share_of_build": grid_ms / lib_ms["ms_total"],
                      "algorithmic_bytes": grid_bytes, "S_grid": s_grid, "nodes_grid": n_grid,
                      "traffic_note": "ncu DRAM bytes of that launch; far below the algorithmic bytes because the working set stays in L2",
                      "issue": (tr.get("k_t1_coop_issue") if tr else None),
                      "whole_build": {"achieved": bb / (lib_ms["ms_total"] * 1e-3) / 1e9, "frac": bb / (lib_ms["ms_total"] * 1e-3) / 1e9 / peak,
                                      "algorithmic_bytes": bb, "S": st["sum_interior_prims"], "M": st["n_nodes"],
                                      "note": "plane + dragon forest build, all tiers (latency-bound chain of dependent passes); per-phase device ms in phase_ms_dragon"}}
        if per_ray is None:
            per_ray = {"pops": 0.0, "interior_visits": 0.0, "triangle_tests": 0.0, "instance_visits": 0.0}
            rb_note = "visit counters unavailable (CPU leg skipped): compulsory bytes only"
        else:
            rb_note = "per-ray visit counters from the oracle on a 2^19-ray sample of the same distribution"
        rbytes = ray_bytes(per_ray, 1)
        scene_bytes = 32 * state["M"] + 12 * n_verts + 12 * n_tris + 5 * 32 + 2 * 192
        ray_ms_1gpu = r_ms_step  # the shard's launch
        ray_roof = ray_roofline(rbytes, per_ray, n_shard, ray_ms_1gpu, scene_bytes, tr, peak, peak_src, rb_note)
        gather_bytes = 4 * (world - 1) * slot_w
        line = {
            "metric": "dragon_blas_build_Mtris_per_s", "value": value, "unit": "Mtris/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": s_tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config2_dict(n_tris, n_rays, world),
            "phase_ms": {"build": b_ms_step, "all_gather": g_ms_step, "tlas": t_tot / args.steps, "trace_shard": r_ms_step},
            "phase_ms_dragon": lib_ms,
            "roofline": build_roof,
            "cpu_baseline": cpu_build,
            "e2e": {"value": world * n_dtris / e2e_b / 1e6, "unit": "Mtris/s", "ms": e2e_b * 1e3,
                    "h2d_bytes_per_step": int(dv.nbytes + di.nbytes), "d2h_bytes_per_step": int(32 * m_dragon + di.nbytes),
                    "api": "bvh_cuda_blas_build (host pointers, pinned), one distinct mesh per rank"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "multi_gpu": None if world == 1 else {
                "what": "per step every rank builds its own dragon-class mesh in place inside its slot of the pooled array, then ONE in-place NCCL all-gather (vertices | permuted indices | used nodes) gives every GPU all meshes; `value` = world x tris / (build + all-gather)",
                "bare_all_gather_ms": bare_ms,
                "all_gather_ms": g_ms_step, "all_gather_bytes_received_per_gpu": int(gather_bytes),
                "all_gather_GBps_per_gpu": gather_bytes / (g_ms_step * 1e-3) / 1e9 if g_ms_step > 0 else None,
                "build_only_Mtris_per_s_per_gpu": n_tris / (b_ms_step * 1e-3) / 1e6},
            "rays": {"metric": "shadow_ray_Mrays_per_s", "value": rvalue, "unit": "Mrays/s", "ms": r_ms_step, "occluded_frac": occ_frac,
                     "scaling": "strong: the fixed batch of %d rays split over %d rank(s) in block-cyclic 64 Ki-ray chunks, max over ranks" % (n_rays, world),
                     "order": "surface-raster (G-buffer-like) ray order; incoherent = same rays, one fixed random permutation",
                     "incoherent": {"value": n_rays / (inc_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms": inc_ms},
                     "weak": None if world == 1 else {"value": world * n_rays / (weak_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms": weak_ms,
                                                      "what": "every rank traces the whole batch (work per GPU fixed)",
                                                      "e2e": {"value": world * n_rays / e2e_w / 1e6, "unit": "Mrays/s", "ms": e2e_w * 1e3,
                                                              "h2d_bytes_per_step": int(ro.nbytes + rd.nbytes), "d2h_bytes_per_step": int(n_rays)}},
                     "roofline": ray_roof, "cpu_baseline": cpu_rays,
                     "e2e": {"value": n_rays / e2e_r / 1e6, "unit": "Mrays/s", "ms": e2e_r * 1e3,
                             "h2d_bytes_per_step": int(24 * n_shard), "d2h_bytes_per_step": int(n_shard),
                             "api": "bvh_cuda_trace_any (host pointers, pinned), the rank's shard of the batch"}},
            "config5": c5,
            "wall_s_timed_region_incl_flush": t_wall,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def ray_roofline(rbytes, per_ray, n_rays, ms, scene_bytes, tr, peak, peak_src, note):
    """Traversal is served from L2 (the 60 MB scene is resident), so the byte model of SURVEY 8(d) is compared with the
    L2 bandwidth, and the DRAM side is reported separately as measured bytes against the HBM peak; the issue-slot
    fraction (ncu) says how close the kernel is to its real bound, instruction issue."""
    l2 = tr.get("l2_peak_GBps") if tr else None
    issue = tr.get("k_trace_any_issue") if tr else None
    model_gbps = rbytes * n_rays / (ms * 1e-3) / 1e9
    dram = tr.get("k_trace_any_16Mi") if (tr and n_rays == N_RAYS) else None
    out = {"bound": "hbm", "achieved": ((dram / (ms * 1e-3) / 1e9) if dram else (25 * n_rays + scene_bytes) / (ms * 1e-3) / 1e9),
           "peak": peak, "unit": "GB/s", "peak_source": peak_src, "traffic": dram,
           "achieved_is": "measured DRAM bytes of one launch (ncu) / launch time" if dram else "compulsory bytes (24 B in + 1 B out per ray + the scene once) / launch time",
           "kernel": "k_trace_any (one launch per step; the exact-order kernel that takes deferred rays runs empty)",
           "l2_model": {"bytes_per_ray": rbytes, "GBps": model_gbps, "l2_peak_GBps": l2, "frac_of_l2": (model_gbps / l2) if l2 else None,
                        "note": "SURVEY 8(d) byte model x rays / time: what the kernel asks its L1 for; most of it is served by L1 (hit rate 73 %)"},
           "l2_measured": (None if not (issue and l2 and n_rays == N_RAYS) else
                           {"bytes_per_launch": issue["l2_bytes_per_launch"], "GBps": issue["l2_bytes_per_launch"] / (ms * 1e-3) / 1e9,
                            "frac_of_l2_peak": issue["l2_bytes_per_launch"] / (ms * 1e-3) / 1e9 / l2,
                            "note": "lts__t_sectors.sum x 32 B of one launch (ncu) / this run's launch time, against the measured L2 read peak"}),
           "issue": (tr.get("k_trace_any_issue") if tr else None),
           "counters_per_ray": per_ray, "note": note}
    out["frac"] = out["achieved"] / peak
    return out


if __name__ == "__main__":
    main()
