"""Per-launch table (time, DRAM bytes, achieved GB/s) from an ncu CSV written with
--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv.
usage: python tools/launch_table.py launches.csv [first_kernel_substring] [max_rows]"""
import collections
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    d = collections.OrderedDict()
    for row in csv.DictReader(lines):
        key = (int(row["ID"]), row["Kernel Name"])
        val = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6,
                 "Gbyte": 1e9}.get(unit, 1.0)
        d.setdefault(key, {})[row["Metric Name"]] = val * scale
    return d


def main():
    d = load(sys.argv[1])
    first = sys.argv[2] if len(sys.argv) > 2 else None
    nmax = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
    items = list(d.items())
    if first:
        starts = [i for i, (k, v) in enumerate(items) if first in k[1]]
        items = items[starts[-1]:]
    items = items[:nmax]
    tot = 0.0
    print(f"{'kernel':58s} {'us':>9s} {'DRAM MB':>9s} {'GB/s':>8s}")
    for (i, name), v in items:
        t = v.get("gpu__time_duration.sum", 0.0)
        b = v.get("dram__bytes_read.sum", 0.0) + v.get("dram__bytes_write.sum", 0.0)
        tot += t
        short = name.split("(")[0].replace("pdr::", "").replace("void ", "")[:58]
        print(f"{short:58s} {t:9.1f} {b / 1e6:9.2f} {b / t / 1e3 if t else 0:8.1f}")
    print(f"{'total':58s} {tot:9.1f}")


if __name__ == "__main__":
    main()
