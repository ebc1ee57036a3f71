#!/bin/bash
# HPR exact pass with E2 in shared memory: parity tests, A/B of 10 vs 8 warps per block, launch lists
# NOTE: the PDR_HPR_* environment switch used below existed only in the experimental build this job measured
# (results: profiles/r02u_filter_experiment.md); the committed kernels ignore it.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hpr_gpu.py tests/test_geometry_gpu.py tests/test_production_goldens_gpu.py \
    -q -p no:cacheprovider > gpurun_out/r02p_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02p_pytest.log
PDR_HPR_WARPS=8 timeout 600 python -m pytest tests/test_hpr_gpu.py tests/test_production_goldens_gpu.py \
    -q -p no:cacheprovider > gpurun_out/r02p_pytest_w8.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02p_pytest_w8.log
for w in 10 8; do
PDR_HPR_WARPS=$w timeout 300 python bench.py --config 0 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02p_bench_config0_w$w.json 2>> gpurun_out/r02p_bench.err
PDR_HPR_WARPS=$w timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02p_config0_launches_w$w.csv \
    python bench.py --config 0 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02p_ncu0.log 2>&1
PDR_HPR_WARPS=$w timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02p_config0_8views_launches_w$w.csv \
    python bench.py --config 0 --views 8 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02p_ncu8.log 2>&1
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-extras > gpurun_out/r02p_bench_1gpu.json 2>> gpurun_out/r02p_bench.err
tail -2 gpurun_out/r02p_pytest.log; tail -2 gpurun_out/r02p_pytest_w8.log; grep hpr_exact gpurun_out/r02p_*launches*.csv | awk -F'"' '{print $1, $(NF-1)}' | tail -8; head -c 250 gpurun_out/r02p_bench_config0_w10.json; echo; head -c 250 gpurun_out/r02p_bench_config0_w8.json; echo; grep -o '"stage_ms": {[^}]*}' gpurun_out/r02p_bench_1gpu.json
