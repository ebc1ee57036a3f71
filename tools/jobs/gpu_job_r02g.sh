#!/bin/bash
# round-2 evidence run on the committed code: full GPU suite, smoke, bench line, geometry ncu, sanitizer
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/r02g_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02g_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02g_smoke.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r02g_bench_1gpu.json 2> gpurun_out/r02g_bench.err
timeout 200 python tools/exp_two_streams.py 4 > gpurun_out/r02g_four_engines.log 2>&1
echo "four engines exit $?" >> gpurun_out/r02g_four_engines.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/r02g_geom_launches_v2.csv \
    python bench.py --config 0 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02g_geom_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'raster_|fill_|hpr_|unproj_|splat_|compact_|scan_|point_vis|rescale|vertex_tr|mask_half|minmax|crop_params' \
    -s 44 -c 44 -o gpurun_out/r02g_geometry_full \
    python bench.py --config 0 --views 8 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02g_geom_full.log 2>&1
ncu -i gpurun_out/r02g_geometry_full.ncu-rep --page raw --csv > gpurun_out/r02g_geometry_full_raw.csv 2>/dev/null
rm -f gpurun_out/r02g_geometry_full.ncu-rep
python bench.py --config 0 --steps 20 --warmup 5 > gpurun_out/r02g_bench_config0.json 2>> gpurun_out/r02g_bench.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -p no:cacheprovider \
    "tests/test_geometry_gpu.py::test_geometry_vs_reference_golden" tests/test_hpr_gpu.py tests/test_conv_tc_gpu.py \
    > gpurun_out/r02g_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r02g_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest -q -p no:cacheprovider \
    "tests/test_geometry_gpu.py::test_geometry_vs_reference_golden" \
    > gpurun_out/r02g_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r02g_racecheck.log
tail -3 gpurun_out/r02g_pytest.log; tail -2 gpurun_out/r02g_smoke.log; tail -3 gpurun_out/r02g_four_engines.log; tail -2 gpurun_out/r02g_memcheck.log; tail -2 gpurun_out/r02g_racecheck.log; head -c 300 gpurun_out/r02g_bench_1gpu.json
