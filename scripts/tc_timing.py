"""Per-node timeline of the cluster tier (library built with -DBVH_TC_TIMING: make -C voidin_b200/csrc variant NAME=tctime
EXTRA=-DBVH_TC_TIMING, run with BVH_CUDA_LIB=voidin_b200/variants/libbvh_cuda_tctime.so).  Builds the dragon-class mesh a few
times and prints, for the last build, one row per cluster-tier node: size, cluster, wait before the pop, and the time spent in
bounds / flags / 21 candidate shuffles / bins + select / final shuffle + children, all in microseconds."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import voidin_b200 as vb
from voidin_b200 import scenes as S

dev = torch.device("cuda", 0)
ctx = vb.Context(0); ctx.set_profiling(True)
which = sys.argv[1] if len(sys.argv) > 1 else "dragon"
dv, di = S.dragon_class() if which == "dragon" else S.soup(int(which), 4, 0.005)
n = di.size // 3
d_v = torch.from_numpy(dv.reshape(-1)).to(dev); d_i0 = torch.from_numpy(di.view(np.int32)).to(dev)
d_nodes = torch.zeros(2 * n * 8, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
buf = np.zeros((4100, 8), dtype=np.uint64)
for k in range(4):
    d_i = d_i0.clone(); torch.cuda.synchronize()
    ctx.blas_build_dev(d_v.data_ptr(), dv.shape[0], d_i.data_ptr(), n, d_nodes.data_ptr(), 2 * n, stream)
    torch.cuda.synchronize()
    rows = ctx.lib.bvh_cuda_debug_tc_log(buf.ctypes.data_as(C.c_void_p), 4100)
st = ctx.last_build_stats()
print({k: round(v, 3) if isinstance(v, float) else v for k, v in st.items()})
if rows <= 0:
    print("library was not built with -DBVH_TC_TIMING"); sys.exit(0)
ph = buf[rows].astype(np.float64)
if ph[6] > 0:
    print("per shuffle (cycles, rank-0 thread 0, mean over %d shuffles): A %.0f  sync1 %.0f  B %.0f  sync2 %.0f  C %.0f  sync3 %.0f" % ((ph[6],) + tuple(ph[:6] / ph[6])))
b = buf[:rows].astype(np.int64)
t0 = b[:, 1].min()
order = np.argsort(b[:, 2])
print(f"{rows} cluster-tier nodes; tier span {(b[:, 7].max() - t0) / 1e3:.1f} us")
print("      n  cl   pop@us  wait  bounds  flags   cand21  bins+sel  final  total")
for r in order:
    nn, cl = int(b[r, 0] & 0xFFFFFFFF), int(b[r, 0] >> 32)
    t = b[r, 1:8]
    d = np.diff(t) / 1e3
    print(f"{nn:7d} {cl:3d} {(t[1] - t0) / 1e3:8.1f} {d[0]:5.1f} {d[1]:7.1f} {d[2]:6.1f} {d[3]:8.1f} {d[4]:9.1f} {d[5]:6.1f} {(t[6] - t[1]) / 1e3:7.1f}")
