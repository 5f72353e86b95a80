import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import voidin_b200 as vb
from voidin_b200 import scenes as S
dev = torch.device("cuda", 0)
ctx = vb.Context(0); ctx.set_profiling(True)
dv, di = S.dragon_class()
d_v = torch.from_numpy(dv.reshape(-1)).to(dev); d_i0 = torch.from_numpy(di.view(np.int32)).to(dev)
n = di.size // 3
d_nodes = torch.zeros(2 * n * 8, dtype=torch.int32, device=dev)
buf = (C.c_ulonglong * 32)()
names = ["init", "bounds", "flags", "bins", "select", "count_final", "tilescan", "table", "scatter", "children", "nextlevel", "pull"]
for k in range(3):
    d_i = d_i0.clone(); torch.cuda.synchronize()
    ctx.blas_build_dev(d_v.data_ptr(), dv.shape[0], d_i.data_ptr(), n, d_nodes.data_ptr(), 2 * n, 0)
    st = ctx.last_build_stats()
    ctx.lib.bvh_cuda_debug_t1_timing(buf)
    print(f"build {k}: grid {st['ms_grid']:.3f} ms, levels {st['grid_levels']}")
    tw = tb = 0
    for i, nm in enumerate(names):
        print(f"   {nm:12s} work {buf[2*i]/1e3:9.1f} us   barrier {buf[2*i+1]/1e3:9.1f} us")
        tw += buf[2*i]; tb += buf[2*i+1]
    print(f"   total work {tw/1e3:.1f} us, barrier {tb/1e3:.1f} us")
    pl = (C.c_ulonglong * 4096)()
    if hasattr(ctx.lib, "bvh_cuda_debug_t1_pull") and ctx.lib.bvh_cuda_debug_t1_pull(pl):
        a = np.array(pl[:], dtype=np.float64).reshape(8, 512)[:, :444] / 22e3  # us per shuffle, level 0, thread 0 of every block
        for k, nm in enumerate(["prefix", "partner cache", "pass 1 (select)", "pass 2 (gather+store)"]):
            print(f"   level-0 pull {nm:22s}: per-block us/shuffle med {np.median(a[k]):.2f} p90 {np.percentile(a[k], 90):.2f} max {a[k].max():.2f}  by decile "
                  f"{[round(float(a[k][i*42:(i+1)*42].mean()), 2) for i in range(10)]}")
    blk = (C.c_ulonglong * 2048)()
    if ctx.lib.bvh_cuda_debug_t1_blocks(blk):
        for kind, nm in enumerate(["table", "scatter"]):
            a = np.array(blk[kind * 1024: kind * 1024 + 444], dtype=np.float64) / 22e3  # us per shuffle, level 0
            order = np.argsort(a)
            print(f"   level-0 {nm}: per-block work us/shuffle min {a.min():.2f} med {np.median(a):.2f} p90 {np.percentile(a,90):.2f} max {a.max():.2f}; slowest blocks {order[-6:].tolist()} fastest {order[:4].tolist()}")
            print("     by tile decile:", [round(float(a[i*42:(i+1)*42].mean()),2) for i in range(10)])
        ts = np.array(blk[900:1024], dtype=np.float64)
        meta = np.array(blk[1024 + 900:2048], dtype=np.uint64)
        print("   per level (us, tiles, nodes):", [(round(float(ts[i + 1] - ts[i]) / 1e3, 1), int(meta[i] & 0xFFFFFFFF), int(meta[i] >> 32))
                                                   for i in range(len(ts) - 1) if ts[i] > 0 and ts[i + 1] > 0])
