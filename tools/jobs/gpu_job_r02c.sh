#!/bin/bash
# round-2 third GPU call: raster fast path + HPR re-solve, geometry evidence (ncu --set full)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -k "geometry or hpr or production or default_flow or optimize or formats or config5 or neighbors" > gpurun_out/r02c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02c_pytest.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/r02c_geom_launches_v2.csv \
    python bench.py --config 0 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02c_geom_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/r02c_geom_launches_v8.csv \
    python bench.py --config 0 --views 8 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02c_geom_ncu8.log 2>&1
python bench.py --config 0 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02c_bench_config0.json 2> gpurun_out/r02c_bench_config0.err
python bench.py --config 0 --views 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02c_bench_config0_v8.json 2>> gpurun_out/r02c_bench_config0.err
# full-set capture of the geometry kernels of ONE step at 8 views (skip the warm-up step's launches)
ncu --set full --clock-control none --import-source on -k regex:'raster_|fill_|hpr_|unproj_|splat_|compact_|scan_|point_vis|rescale|vertex_tr|mask_half' \
    -s 45 -c 45 -o gpurun_out/r02c_geometry_full \
    python bench.py --config 0 --views 8 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02c_geom_full.log 2>&1
tail -3 gpurun_out/r02c_pytest.log; head -c 400 gpurun_out/r02c_bench_config0.json; echo; head -c 400 gpurun_out/r02c_bench_config0_v8.json; ls -la gpurun_out | grep r02c
