#!/bin/bash
# device-side timestep-embedding cache: parity tests + same-call A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_engine_gpu.py tests/test_ddnm_gpu.py tests/test_default_flow_gpu.py tests/test_unet_ops_gpu.py \
    -q -p no:cacheprovider > gpurun_out/r02_embcache_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_embcache_pytest.log
tail -4 gpurun_out/r02_embcache_pytest.log
for i in 1 2; do
PDR_NO_EMB_CACHE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-extras > gpurun_out/r02_embcache_bench_off_$i.json 2>> gpurun_out/r02_embcache_bench.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-extras > gpurun_out/r02_embcache_bench_on_$i.json 2>> gpurun_out/r02_embcache_bench.err
done
for f in off_1 on_1 off_2 on_2; do python -c "
import json
j=json.loads([l for l in open('gpurun_out/r02_embcache_bench_$f.json') if l.startswith('{')][0]); print('$f', j['value'], j['ms_per_step'], j['roofline']['per_class_ms_per_forward']['linear'], j['clocks']['sm_mhz'], j['gpu_launches'])"; done
