"""Key metrics of an `ncu --set full` capture, one block per launch:
   ncu -i capture.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_full_summary.py raw.csv > profiles/xxx_ncu_full_summary.txt"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print(r[ix["Kernel Name"]].split("(")[0])
    for k in KEYS:
        if k in ix and r[ix[k]] not in ("", "n/a"):
            print(f"    {k:86s} {r[ix[k]]:>16s} {units[ix[k]]}")
