"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/summarize_ncu.py gpurun_out/launches.csv > profiles/xxx_summary.txt"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name", "gpu__time_duration.sum") != "gpu__time_duration.sum":
        continue
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (ValueError, KeyError):
        continue
    unit = row.get("Metric Unit", "ns")
    v = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    key = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    agg[key][0] += 1
    agg[key][1] += v
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, {tot:.3f} ms total "
      "(ncu per-launch times: cold cache, serialised - compare SHARES)")
print(f"{'kernel':60s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'avg_us':>9s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} {v[0]:8d} {v[1]:10.3f} {v[1] / tot:7.1%} {1e3 * v[1] / v[0]:9.1f}")
