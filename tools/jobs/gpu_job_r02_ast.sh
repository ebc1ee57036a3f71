#!/bin/bash
# fused GroupNorm convs with a fourth halo slot (AST 4, BST 8) vs the shipped 3 / 9: bitwise check + same-call A/B
mkdir -p gpurun_out
PDR_HALO_AST4=1 timeout 600 python -m pytest tests/test_unet_engine_gpu.py tests/test_conv_tc_gpu.py -q -p no:cacheprovider > gpurun_out/r02_ast_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_ast_pytest.log
for i in 1 2; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-extras > gpurun_out/r02_ast_bench_a3_$i.json 2>> gpurun_out/r02_ast_bench.err
PDR_HALO_AST4=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-extras > gpurun_out/r02_ast_bench_a4_$i.json 2>> gpurun_out/r02_ast_bench.err
done
tail -2 gpurun_out/r02_ast_pytest.log
for f in a3_1 a4_1 a3_2 a4_2; do python -c "
import json
j=json.loads([l for l in open('gpurun_out/r02_ast_bench_$f.json') if l.startswith('{')][0]); print('$f', j['value'], j['ms_per_step'], j['roofline']['per_class_ms_per_forward']['conv_tc'], j['clocks']['sm_mhz'])"; done
