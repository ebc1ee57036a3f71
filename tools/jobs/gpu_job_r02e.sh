#!/bin/bash
# round-2 fifth GPU call: split A/B producers in the halo conv (fused GroupNorm A/B), smem raster, HPR tweaks
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -k "geometry or hpr or production or unet_engine or conv_tc or optimize or formats" > gpurun_out/r02e_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02e_pytest.log
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02e_bench_default.json 2> gpurun_out/r02e_bench.err
PDR_FUSED_GN=1 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02e_bench_fused.json 2>> gpurun_out/r02e_bench.err
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02e_bench_default2.json 2>> gpurun_out/r02e_bench.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/r02e_geom_launches_v8.csv \
    python bench.py --config 0 --views 8 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02e_geom_ncu8.log 2>&1
python bench.py --config 0 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02e_bench_config0.json 2>> gpurun_out/r02e_bench.err
tail -3 gpurun_out/r02e_pytest.log
for f in default fused default2; do python -c "
import json,sys
j=json.load(open('gpurun_out/r02e_bench_$f.json')); print('$f', j['value'], j['roofline']['per_class_ms_per_forward'])"; done
