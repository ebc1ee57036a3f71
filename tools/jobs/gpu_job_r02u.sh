#!/bin/bash
# pruned HPR filter: points per warp 8 / 16 / 32
# NOTE: the PDR_HPR_* environment switch used below existed only in the experimental build this job measured
# (results: profiles/r02u_filter_experiment.md); the committed kernels ignore it.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hpr_gpu.py tests/test_production_goldens_gpu.py tests/test_geometry_gpu.py \
    -q -p no:cacheprovider > gpurun_out/r02u_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02u_pytest.log
for p in 8 16 32; do
PDR_HPR_FILTER_POINTS=$p timeout 300 python bench.py --config 0 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02u_bench_config0_p$p.json 2>> gpurun_out/r02u_bench.err
PDR_HPR_FILTER_POINTS=$p timeout 300 python bench.py --config 0 --views 8 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02u_bench_config0_8views_p$p.json 2>> gpurun_out/r02u_bench.err
PDR_HPR_FILTER_POINTS=$p timeout 300 ncu -k regex:'hpr_filter|hpr_exact' --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r02u_hpr_2views_p$p.csv \
    python bench.py --config 0 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02u_ncu.log 2>&1
PDR_HPR_FILTER_POINTS=$p timeout 300 ncu -k regex:'hpr_filter|hpr_exact' --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r02u_hpr_8views_p$p.csv \
    python bench.py --config 0 --views 8 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02u_ncu.log 2>&1
done
tail -2 gpurun_out/r02u_pytest.log
for p in 8 16 32; do head -c 200 gpurun_out/r02u_bench_config0_p$p.json | cut -c 30-75; head -c 200 gpurun_out/r02u_bench_config0_8views_p$p.json | cut -c 30-75; done
grep -h "hpr_filter" gpurun_out/r02u_hpr_*views_p*.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-120 | tail -24
