#!/bin/bash
# round-2 fourth GPU call: HPR v3 (lazy box, 8 points/warp), packed gn_apply: parity + timings
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/r02d_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02d_pytest.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/r02d_geom_launches_v8.csv \
    python bench.py --config 0 --views 8 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02d_geom_ncu8.log 2>&1
python bench.py --config 0 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02d_bench_config0.json 2> gpurun_out/r02d_bench.err
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02d_bench_config1.json 2>> gpurun_out/r02d_bench.err
tail -3 gpurun_out/r02d_pytest.log; grep "30k x 8" gpurun_out/r02d_pytest.log; head -c 300 gpurun_out/r02d_bench_config0.json; echo; head -c 300 gpurun_out/r02d_bench_config1.json
