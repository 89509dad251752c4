"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, share, average."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"]); u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)          # -> us
    name = row["Kernel Name"]
    m = re.search(r"(\w+_kernel(<[^>]*>)?)", name)
    short = m.group(1) if m else name[:50]
    agg[short][0] += 1; agg[short][1] += v; tot += v
print(f"total {tot/1e3:.2f} ms over {sum(n for n,_ in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:48s} n={n:5d} total={t/1e3:8.2f} ms share={t/tot*100:5.1f}% avg={t/n:8.1f} us")
