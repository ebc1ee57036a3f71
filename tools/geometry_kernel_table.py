"""Per-kernel evidence table for the HBM-bound (geometry) kernels from one `ncu --set full` capture:
   ncu -i capture.ncu-rep --page raw --csv > raw.csv ; python tools/geometry_kernel_table.py raw.csv
Prints markdown: launch time, DRAM bytes (read + written), achieved GB/s and its fraction of the
measured HBM peak (MEASURED_PEAKS.json), SM throughput %, achieved occupancy, registers."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, body = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def num(r, key, default=0.0):
    if key not in ix or r[ix[key]] in ("", "n/a"):
        return default
    v = float(r[ix[key]].replace(",", ""))
    u = units[ix[key]]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6,
             "Gbyte": 1e9}.get(u, 1.0)
    return v * scale


agg = {}
order = []
for r in body:
    name = r[ix["Kernel Name"]].split("(")[0].replace("pdr::", "")
    t = num(r, "gpu__time_duration.sum")
    b = num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")
    a = agg.setdefault(name, dict(n=0, t=0.0, b=0.0, sm=0.0, occ=0.0, regs=0, grid=0, block=0))
    if a["n"] == 0:
        order.append(name)
    a["n"] += 1
    a["t"] += t
    a["b"] += b
    a["sm"] += num(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed")
    a["occ"] += num(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    a["regs"] = int(num(r, "launch__registers_per_thread"))
    a["grid"] = int(num(r, "launch__grid_size"))
    a["block"] = int(num(r, "launch__block_size"))
print(f"| kernel | launches | us / launch | DRAM MB / launch | GB/s | of {peak:.0f} GB/s | SM thr % | occupancy % | regs | grid x block |")
print("|---|---|---|---|---|---|---|---|---|---|")
tt = tb = 0.0
for name in order:
    a = agg[name]
    n = a["n"]
    gbs = a["b"] / a["t"] / 1e3 if a["t"] else 0.0
    tt += a["t"]
    tb += a["b"]
    print(f"| `{name}` | {n} | {a['t'] / n:.1f} | {a['b'] / n / 1e6:.2f} | {gbs:.0f} | {gbs / peak:.3f} | "
          f"{a['sm'] / n:.0f} | {a['occ'] / n:.0f} | {a['regs']} | {a['grid']} x {a['block']} |")
print(f"| **all** | | {tt:.0f} (sum) | {tb / 1e6:.1f} (sum) | {tb / tt / 1e3:.0f} | {tb / tt / 1e3 / peak:.3f} | | | | |")
