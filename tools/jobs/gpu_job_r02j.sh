#!/bin/bash
# selective GroupNorm fusion (PDR_FUSED_GN=2: only single-N-tile convs) A/B in one call
mkdir -p gpurun_out
python -m pytest tests/test_unet_engine_gpu.py -q -p no:cacheprovider > gpurun_out/r02j_pytest.log 2>&1
tail -2 gpurun_out/r02j_pytest.log
for mode in 0 2 1 0 2; do
  if [ "$mode" = "0" ]; then unset PDR_FUSED_GN; else export PDR_FUSED_GN=$mode; fi
  python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02j_bench_mode${mode}_$RANDOM.json 2>> gpurun_out/r02j_bench.err
done
unset PDR_FUSED_GN
for f in gpurun_out/r02j_bench_mode*.json; do python -c "
import json,sys
j=json.load(open('$f')); pc=j['roofline']['per_class_ms_per_forward']; print('$f', round(j['value'],5), round(j['ms_per_step'],1), round(pc['conv_tc'],2), round(pc['gn_apply'],2), j['clocks']['sm_mhz'])"; done
