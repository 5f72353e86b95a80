"""A/B timing of compile-time variants of libbvh_cuda.so on the GPU box.  For every voidin_b200/variants/libbvh_cuda_<name>.so
(built with `make -C voidin_b200/csrc variant NAME=<name> EXTRA=-D...`) a child process builds the dragon-class mesh a few
times with per-phase timing, checks nodes and primitive order byte-for-byte against the CPU oracle (computed once), and
runs the small-mesh parity ladder.  Usage: python scripts/variants.py [name[:ENV=VAL[,ENV=VAL]] ...]
(a name may carry extra environment for that run, e.g. base:BVH_CUDA_T1_GRID=296).  The round-1 results are tabulated in
profiles/r01_build_variants_ab.txt."""
import glob, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

CACHE = "/tmp/dragon_oracle.npz"


def child(name):
    import torch
    import voidin_b200 as vb
    from voidin_b200 import scenes as S
    from oracle import oracle as O
    import helpers
    ctx = vb.Context(0); ctx.set_profiling(True)
    dev = torch.device("cuda", 0)
    ok = True
    # small ladder through the host API
    for nm, v, idx in helpers.small_meshes() + [(f"soup{n}", *S.soup(n, 300 + n, 0.05)) for n in (6, 12, 17, 24, 40, 300, 3000, 40000)]:
        gi = idx.copy()
        try:
            bvh = vb.BvhBuilder(v, gi, ctx).build()
        except vb.BvhCudaError as e:
            rc, *_ = O.blas_build(v, idx)
            if rc != e.code: ok = False; print(f"[{name}] {nm}: GPU error {e} oracle rc {rc}")
            continue
        rc, onodes, oidx, _, _ = O.blas_build(v, idx)
        good = rc == 0 and bvh.nodes.tobytes() == onodes.tobytes() and (gi == oidx).all()
        if not good: ok = False; print(f"[{name}] MISMATCH on {nm} (n={idx.size // 3})", flush=True)
    # forest build: 40 small meshes of assorted sizes in one batch
    # dragon-class
    d = np.load(CACHE)
    dv, di = S.dragon_class()
    n = di.size // 3
    d_v = torch.from_numpy(dv.reshape(-1)).to(dev); d_i0 = torch.from_numpy(di.view(np.int32)).to(dev)
    d_nodes = torch.zeros(2 * n * 8, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    rows = []
    for k in range(6):
        d_i = d_i0.clone(); torch.cuda.synchronize()
        m = ctx.blas_build_dev(d_v.data_ptr(), dv.shape[0], d_i.data_ptr(), n, d_nodes.data_ptr(), 2 * n, stream)
        torch.cuda.synchronize()
        st = ctx.last_build_stats()
        if k: rows.append(st)
    nodes = d_nodes.cpu().numpy().view(np.uint8)[: 32 * m]
    same = m == len(d["nodes"]) // 32 and nodes.tobytes() == d["nodes"].tobytes() and (d_i.cpu().numpy().view(np.uint32) == d["idx"]).all()
    if not same: ok = False; print(f"[{name}] MISMATCH on the dragon-class mesh", flush=True)
    keys = [k for k in rows[0] if k.startswith("ms_")]
    avg = {k: round(float(np.mean([r[k] for r in rows])), 4) for k in keys}
    mn = {k: round(float(np.min([r[k] for r in rows])), 4) for k in keys}
    cnt = {k: rows[0][k] for k in rows[0] if not k.startswith("ms_")}
    print(json.dumps({"variant": name, "parity": ok, "ms_avg": avg, "ms_min": mn, "counts": cnt}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2]); sys.exit(0)
    if not os.path.exists(CACHE):
        from voidin_b200 import scenes as S
        from oracle import oracle as O
        dv, di = S.dragon_class()
        t0 = time.time()
        rc, onodes, oidx, _, _ = O.blas_build(dv, di)
        assert rc == 0
        np.savez(CACHE, nodes=onodes.view(np.uint8).reshape(-1), idx=oidx)
        print(f"oracle: dragon-class build {time.time() - t0:.1f} s on the host", flush=True)
    names = sys.argv[1:] or sorted(os.path.basename(p)[len("libbvh_cuda_"):-3] for p in glob.glob(os.path.join(ROOT, "voidin_b200", "variants", "libbvh_cuda_*.so")))
    for spec in names:  # "<name>" or "<name>:KEY=VAL[,KEY=VAL]" (extra environment for that run)
        name, _, extra = spec.partition(":")
        env = dict(os.environ, BVH_CUDA_LIB=os.path.join(ROOT, "voidin_b200", "variants", f"libbvh_cuda_{name}.so"))
        for kv in filter(None, extra.split(",")):
            k, _, v = kv.partition("=")
            env[k] = v
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", spec], env=env, timeout=600)
        if r.returncode: print(f"[{name}] child exited with {r.returncode}", flush=True)
