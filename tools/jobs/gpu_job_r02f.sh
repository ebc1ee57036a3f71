#!/bin/bash
# round-2 sixth GPU call: TMA-store epilogue parity + A/B, power probe
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -k "conv_tc or unet_engine or unet_ops or ddnm_gpu" > gpurun_out/r02f_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02f_pytest.log
tail -3 gpurun_out/r02f_pytest.log
PDR_NO_TMA_STORE=1 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02f_bench_nostore.json 2> gpurun_out/r02f_bench.err
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02f_bench_tmastore.json 2>> gpurun_out/r02f_bench.err
PDR_NO_TMA_STORE=1 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02f_bench_nostore2.json 2>> gpurun_out/r02f_bench.err
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02f_bench_tmastore2.json 2>> gpurun_out/r02f_bench.err
python tools/power_probe.py > gpurun_out/r02f_power_probe.json 2> gpurun_out/r02f_power_probe.err
PDR_QUICK=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'conv_tc|conv_halo|splitk' -c 200 --csv --log-file gpurun_out/r02f_conv_launches.csv python tools/bench_unet.py 8 > gpurun_out/r02f_conv_ncu.log 2>&1
for f in nostore tmastore nostore2 tmastore2; do python -c "
import json
j=json.load(open('gpurun_out/r02f_bench_$f.json')); print('$f', j['value'], j['roofline']['per_class_ms_per_forward']['conv_tc'], j['roofline']['per_class_ms_per_forward']['stem'])"; done
cat gpurun_out/r02f_power_probe.json
