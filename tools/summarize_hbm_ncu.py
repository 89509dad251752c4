"""profiles/ helper: per-kernel HBM figures from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum,dram__throughput... --csv` log of tools/run_hbm_kernels.py (last = warm launch of each kernel)."""
import collections, csv, json, os, re, sys
path = sys.argv[1]
peak = 6487.4
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
lines = [l for l in open(path) if not l.startswith("==")]
d = collections.OrderedDict()
for r in csv.DictReader(lines):
    name = r["Kernel Name"]
    m = re.search(r"(\w+_kernel)", name)
    short = m.group(1) if m else name[:40]
    d.setdefault((r["ID"], short), {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
keep = ("rmsnorm", "layernorm", "swiglu", "rope", "gelu", "ce_rows", "pool_", "adam")
last = collections.OrderedDict()
for (id_, name), m in d.items():
    if not name.startswith(keep) or "gpu__time_duration.sum" not in m:
        continue
    t = m["gpu__time_duration.sum"]
    tus = t[0] / 1e3 if t[1] in ("ns", "nsecond") else (t[0] if t[1] in ("us", "usecond") else t[0] * 1e3)
    rd = m["dram__bytes_read.sum"][0] * unit.get(m["dram__bytes_read.sum"][1], 1)
    wr = m["dram__bytes_write.sum"][0] * unit.get(m["dram__bytes_write.sum"][1], 1)
    last[name] = (tus, rd / 1e6, wr / 1e6, (rd + wr) / tus / 1e3, m["dram__throughput.avg.pct_of_peak_sustained_elapsed"][0])
print(f"# HBM-bound training kernels at cfg-3 / cfg-2 shapes (ncu, --clock-control none; DRAM bytes = dram__bytes_read.sum + dram__bytes_write.sum);")
print(f"# fraction = measured DRAM GB/s / {peak} GB/s (MEASURED_PEAKS.json hbm_gbs, 'of measured')")
print(f"{'kernel':28s} {'us':>8s} {'read MB':>9s} {'write MB':>9s} {'GB/s':>8s} {'frac':>6s} {'ncu dram %':>10s}")
for k, o in last.items():
    print(f"{k:28s} {o[0]:8.1f} {o[1]:9.1f} {o[2]:9.1f} {o[3]:8.0f} {o[3] / peak:6.2f} {o[4]:10.1f}")
