#!/bin/bash
# box-pruned HPR exact pass: parity tests, config-0 line, launch list at 2 and 8 views, memcheck of the HPR tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hpr_gpu.py tests/test_geometry_gpu.py tests/test_production_goldens_gpu.py \
    tests/test_default_flow_gpu.py -q -rA -p no:cacheprovider > gpurun_out/r02o_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02o_pytest.log
timeout 300 python bench.py --config 0 --steps 20 --warmup 5 > gpurun_out/r02o_bench_config0.json 2> gpurun_out/r02o_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02o_config0_launches.csv \
    python bench.py --config 0 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02o_ncu0.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02o_config0_8views_launches.csv \
    python bench.py --config 0 --views 8 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02o_ncu8.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-extras > gpurun_out/r02o_bench_1gpu.json 2>> gpurun_out/r02o_bench.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -p no:cacheprovider tests/test_hpr_gpu.py \
    > gpurun_out/r02o_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r02o_memcheck.log
tail -5 gpurun_out/r02o_pytest.log; head -c 400 gpurun_out/r02o_bench_config0.json; echo; grep -c hpr_ gpurun_out/r02o_config0_launches.csv; tail -2 gpurun_out/r02o_memcheck.log; head -c 300 gpurun_out/r02o_bench_1gpu.json
