#!/bin/bash
# 8-softmax-warp tcgen05 attention: parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA -p no:cacheprovider -k "attention or unet_engine or ddnm" > gpurun_out/r02m_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02m_pytest.log
grep -E "passed|failed|attention \(pre|teacher" gpurun_out/r02m_pytest.log | tail -8
timeout 400 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err
python -c "
import json
j=json.load(open('gpurun_out/r02m_bench.json')); print(j['value'], j['ms_per_step'], j['roofline']['per_class_ms_per_forward'])"
PDR_QUICK=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attention_tc -c 10 --csv --log-file gpurun_out/r02m_attn_launches.csv python tools/bench_unet.py 8 > /dev/null 2>&1
grep attention_tc gpurun_out/r02m_attn_launches.csv | awk -F'","' '{print $NF}' | head -10
