#!/bin/bash
# ncu source-level profile of the shipped HPR exact + filter kernels at 2 views (the configs[0] case)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'hpr_filter|hpr_exact' -s 4 -c 2 -o gpurun_out/r02w_hpr_full \
    python bench.py --config 0 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02w_ncu_full.log 2>&1
ncu -i gpurun_out/r02w_hpr_full.ncu-rep --page raw --csv > gpurun_out/r02w_hpr_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r02w_hpr_full.ncu-rep --page source --csv > gpurun_out/r02w_hpr_full_source.csv 2>/dev/null
rm -f gpurun_out/r02w_hpr_full.ncu-rep
ls -la gpurun_out/r02w_*
