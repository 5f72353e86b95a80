"""Summarise ncu artefacts brought back in gpurun_out/ into small text files under profiles/ (tracked).
usage: python scripts/ncu_summary.py <round-tag>   e.g. r01"""
import collections, csv, glob, os, re, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg"]

for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{tag}_*.ncu-rep"))):
    name = os.path.basename(rep)[:-8]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    with open(os.path.join(OUT, name + "_ncu_summary.txt"), "w") as f:
        for vals in rows[2:]:
            kn = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"kernel: {kn}\n(ncu --set full --clock-control none; one launch, values are per launch)\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"  {k:78s} {vals[i]:>16s} {units[i]}\n")
            f.write("  -- warp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active.ratio), top 8 --\n")
            st = []
            for i, h in enumerate(hdr):
                m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active\.ratio", h)
                if m:
                    try:
                        st.append((float(vals[i].replace(",", "")), m.group(1)))
                    except ValueError:
                        pass
            for v, n in sorted(st, reverse=True)[:8]:
                f.write(f"  {n:40s} {v:8.3f}\n")
    print("wrote", name)

for lst in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{tag}_launches*.csv"))):
    lines = [l for l in open(lst) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        v = float(r["Metric Value"].replace(",", ""))
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = os.path.join(OUT, os.path.basename(lst)[:-4] + "_summary.txt")
    with open(out, "w") as f:
        f.write(f"source: {os.path.basename(lst)} ({len(rows)} launches; ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare shares)\n")
        f.write(f"{'kernel':60s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{n:60s} {c:8d} {t/1e3:12.1f} {t/c/1e3:10.2f} {100*t/tot:6.1f}%\n")
    print("wrote", out)
