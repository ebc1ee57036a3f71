#!/bin/bash
# HPR exact pass: points per warp A/B (16 / 8 / 4), parity per variant, one full ncu capture with source counters
# NOTE: the PDR_HPR_* environment switch used below existed only in the experimental build this job measured
# (results: profiles/r02u_filter_experiment.md); the committed kernels ignore it.
mkdir -p gpurun_out
for p in 16 8 4; do
PDR_HPR_POINTS=$p timeout 600 python -m pytest tests/test_hpr_gpu.py tests/test_production_goldens_gpu.py \
    -q -p no:cacheprovider > gpurun_out/r02q_pytest_p$p.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02q_pytest_p$p.log
PDR_HPR_POINTS=$p timeout 300 python bench.py --config 0 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02q_bench_config0_p$p.json 2>> gpurun_out/r02q_bench.err
PDR_HPR_POINTS=$p timeout 300 python bench.py --config 0 --views 8 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02q_bench_config0_8views_p$p.json 2>> gpurun_out/r02q_bench.err
PDR_HPR_POINTS=$p timeout 300 ncu -k regex:hpr_ --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02q_hpr_2views_p$p.csv \
    python bench.py --config 0 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02q_ncu.log 2>&1
PDR_HPR_POINTS=$p timeout 300 ncu -k regex:hpr_ --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02q_hpr_8views_p$p.csv \
    python bench.py --config 0 --views 8 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02q_ncu.log 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'hpr_exact|hpr_filter' -s 4 -c 2 -o gpurun_out/r02q_hpr_full \
    python bench.py --config 0 --views 8 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02q_ncu_full.log 2>&1
ncu -i gpurun_out/r02q_hpr_full.ncu-rep --page raw --csv > gpurun_out/r02q_hpr_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r02q_hpr_full.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/r02q_hpr_full_source.csv 2>/dev/null || \
ncu -i gpurun_out/r02q_hpr_full.ncu-rep --page source --csv > gpurun_out/r02q_hpr_full_source.csv 2>/dev/null
ls -la gpurun_out/r02q_hpr_full*
for p in 16 8 4; do tail -2 gpurun_out/r02q_pytest_p$p.log | head -1; head -c 200 gpurun_out/r02q_bench_config0_p$p.json | cut -c 30-75; head -c 200 gpurun_out/r02q_bench_config0_8views_p$p.json | cut -c 30-75; done
