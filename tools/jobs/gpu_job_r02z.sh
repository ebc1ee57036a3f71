#!/bin/bash
# HPR filter grid size 40 / 32 / 24: parity per variant, configs[0] at 2 and 8 views, kernel times, configs[1] project stage
# NOTE: the PDR_HPR_* environment switch used below existed only in the experimental build this job measured
# (results: profiles/r02u_filter_experiment.md); the committed kernels ignore it.
mkdir -p gpurun_out
for g in 40 32 24; do
PDR_HPR_GRID=$g timeout 600 python -m pytest tests/test_hpr_gpu.py tests/test_production_goldens_gpu.py \
    -q -p no:cacheprovider > gpurun_out/r02z_pytest_g$g.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02z_pytest_g$g.log
PDR_HPR_GRID=$g timeout 300 python bench.py --config 0 --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02z_bench_config0_g$g.json 2>> gpurun_out/r02z_bench.err
PDR_HPR_GRID=$g timeout 300 python bench.py --config 0 --views 8 --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02z_bench_config0_8views_g$g.json 2>> gpurun_out/r02z_bench.err
PDR_HPR_GRID=$g timeout 300 ncu -k regex:'hpr_filter|hpr_exact' --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02z_hpr_2views_g$g.csv \
    python bench.py --config 0 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02z_ncu.log 2>&1
PDR_HPR_GRID=$g timeout 300 ncu -k regex:'hpr_filter|hpr_exact' --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02z_hpr_8views_g$g.csv \
    python bench.py --config 0 --views 8 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02z_ncu.log 2>&1
PDR_HPR_GRID=$g timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-extras > gpurun_out/r02z_bench_1gpu_g$g.json 2>> gpurun_out/r02z_bench.err
done
for g in 40 32 24; do tail -2 gpurun_out/r02z_pytest_g$g.log | head -1; head -c 200 gpurun_out/r02z_bench_config0_g$g.json | cut -c 30-75; head -c 200 gpurun_out/r02z_bench_config0_8views_g$g.json | cut -c 30-75; grep -o '"project": [0-9.]*' gpurun_out/r02z_bench_1gpu_g$g.json; done
