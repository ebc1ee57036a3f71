"""ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum CSV of the conv
launches of ONE forward -> profiles/<tag>_conv_traffic.json (bench.py's roofline.traffic)."""
import csv
import json
import sys

src, dst, note = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if len(r) > 10 and r[0].isdigit()]
per = {}
for r in rows:
    d = per.setdefault(r[0], {"name": r[4]})
    v = float(r[-1].replace(",", ""))
    unit = r[-2]
    if r[-3].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    elif r[-3].startswith("gpu__time"):
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}[unit]
    d[r[-3]] = v
launches = [d for d in per.values() if "conv_tc" in d["name"] or "conv_halo" in d["name"]]  # split-K finish folded in below
finish = [d for d in per.values() if "splitk_finish" in d["name"]]
rd = sum(d.get("dram__bytes_read.sum", 0) for d in launches + finish)
wr = sum(d.get("dram__bytes_write.sum", 0) for d in launches + finish)
ms = sum(d.get("gpu__time_duration.sum", 0) for d in launches + finish)
out = {"source": note, "launches": len(launches), "splitk_finish_launches": len(finish),
       "dram_read_bytes": rd, "dram_write_bytes": wr,
       "avg_traffic_bytes_per_launch": (rd + wr) / max(len(launches), 1),
       "sum_duration_ms_under_ncu": ms}
json.dump(out, open(dst, "w"), indent=1)
print(out)
