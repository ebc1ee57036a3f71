#!/bin/bash
# pruned, spatially ordered HPR filter: parity, A/B of 2 vs 3 blocks per SM, launch lists
# NOTE: the PDR_HPR_* environment switch used below existed only in the experimental build this job measured
# (results: profiles/r02u_filter_experiment.md); the committed kernels ignore it.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hpr_gpu.py tests/test_geometry_gpu.py tests/test_production_goldens_gpu.py \
    tests/test_default_flow_gpu.py -q -p no:cacheprovider > gpurun_out/r02s_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02s_pytest.log
PDR_HPR_FILTER_BLOCKS=3 timeout 600 python -m pytest tests/test_hpr_gpu.py tests/test_production_goldens_gpu.py \
    -q -p no:cacheprovider > gpurun_out/r02s_pytest_b3.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02s_pytest_b3.log
for b in 2 3; do
PDR_HPR_FILTER_BLOCKS=$b timeout 300 python bench.py --config 0 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02s_bench_config0_b$b.json 2>> gpurun_out/r02s_bench.err
PDR_HPR_FILTER_BLOCKS=$b timeout 300 python bench.py --config 0 --views 8 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02s_bench_config0_8views_b$b.json 2>> gpurun_out/r02s_bench.err
PDR_HPR_FILTER_BLOCKS=$b timeout 300 ncu -k regex:hpr_ --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02s_hpr_2views_b$b.csv \
    python bench.py --config 0 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02s_ncu.log 2>&1
PDR_HPR_FILTER_BLOCKS=$b timeout 300 ncu -k regex:hpr_ --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02s_hpr_8views_b$b.csv \
    python bench.py --config 0 --views 8 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02s_ncu.log 2>&1
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-extras > gpurun_out/r02s_bench_1gpu.json 2>> gpurun_out/r02s_bench.err
tail -2 gpurun_out/r02s_pytest.log; tail -2 gpurun_out/r02s_pytest_b3.log
for b in 2 3; do head -c 200 gpurun_out/r02s_bench_config0_b$b.json | cut -c 30-75; head -c 200 gpurun_out/r02s_bench_config0_8views_b$b.json | cut -c 30-75; done
grep -o '"stage_ms": {[^}]*}' gpurun_out/r02s_bench_1gpu.json
