#!/bin/bash
# round-2 second GPU call: new geometry kernels (raster / fill / HPR): parity, launch list, sanitizer
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/r02b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02b_pytest.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/r02b_geom_launches.csv \
    python bench.py --config 0 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02b_geom_ncu.log 2>&1
python bench.py --config 0 --steps 20 --warmup 5 > gpurun_out/r02b_bench_config0.json 2> gpurun_out/r02b_bench_config0.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -p no:cacheprovider \
    "tests/test_geometry_gpu.py::test_geometry_vs_reference_golden" tests/test_hpr_gpu.py \
    > gpurun_out/r02b_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r02b_memcheck.log
tail -5 gpurun_out/r02b_pytest.log; tail -3 gpurun_out/r02b_memcheck.log; cat gpurun_out/r02b_bench_config0.json | head -c 1500
