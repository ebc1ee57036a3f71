#!/bin/bash
# final HPR (cheap scan pre-test, E blocks two per round): parity, config-0 lines, launch lists, memcheck, 1-GPU stage times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hpr_gpu.py tests/test_geometry_gpu.py tests/test_production_goldens_gpu.py \
    tests/test_default_flow_gpu.py -q -p no:cacheprovider > gpurun_out/r02r_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02r_pytest.log
timeout 300 python bench.py --config 0 --steps 20 --warmup 5 > gpurun_out/r02r_bench_config0.json 2>> gpurun_out/r02r_bench.err
timeout 300 python bench.py --config 0 --views 8 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02r_bench_config0_8views.json 2>> gpurun_out/r02r_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02r_config0_launches.csv \
    python bench.py --config 0 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02r_ncu.log 2>&1
timeout 300 ncu -k regex:hpr_ --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02r_hpr_8views.csv \
    python bench.py --config 0 --views 8 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02r_ncu.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-extras > gpurun_out/r02r_bench_1gpu.json 2>> gpurun_out/r02r_bench.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -p no:cacheprovider tests/test_hpr_gpu.py \
    > gpurun_out/r02r_memcheck_hpr.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r02r_memcheck_hpr.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest -q -p no:cacheprovider tests/test_hpr_gpu.py \
    > gpurun_out/r02r_racecheck_hpr.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r02r_racecheck_hpr.log
tail -2 gpurun_out/r02r_pytest.log; head -c 230 gpurun_out/r02r_bench_config0.json; echo; head -c 230 gpurun_out/r02r_bench_config0_8views.json; echo; grep -o '"stage_ms": {[^}]*}' gpurun_out/r02r_bench_1gpu.json; tail -2 gpurun_out/r02r_memcheck_hpr.log; tail -2 gpurun_out/r02r_racecheck_hpr.log
