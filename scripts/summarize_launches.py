"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum` log into a per-kernel table (count, total, mean)."""
import csv, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])[:90]
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':90s} {'count':>6s} {'total_us':>11s} {'mean_us':>9s} {'share':>6s}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:90s} {n:6d} {us:11.1f} {us / n:9.1f} {100 * us / tot:5.1f}%")
print(f"TOTAL {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms")
