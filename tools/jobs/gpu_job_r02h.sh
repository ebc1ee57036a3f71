#!/bin/bash
# tcgen05 attention: parity + A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_unet_ops_gpu.py -q -rA -p no:cacheprovider -k "attention" > gpurun_out/r02h_pytest_attn.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02h_pytest_attn.log
tail -25 gpurun_out/r02h_pytest_attn.log
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "unet_engine or ddnm_gpu or geometry or production" > gpurun_out/r02h_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02h_pytest.log
tail -3 gpurun_out/r02h_pytest.log
PDR_NO_TC_ATTENTION=1 timeout 400 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02h_bench_old_attn.json 2> gpurun_out/r02h_bench.err
timeout 400 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02h_bench_tc_attn.json 2>> gpurun_out/r02h_bench.err
for f in old_attn tc_attn; do python -c "
import json
j=json.load(open('gpurun_out/r02h_bench_$f.json')); print('$f', j['value'], j['roofline']['per_class_ms_per_forward']['attention'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02h_geom_launches_v8.csv python bench.py --config 0 --views 8 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
