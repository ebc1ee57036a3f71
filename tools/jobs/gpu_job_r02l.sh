#!/bin/bash
# check run: resample fusion (full GPU suite + bench line)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/r02l_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02l_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02l_smoke.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r02l_bench_1gpu.json 2> gpurun_out/r02l_bench.err
python bench.py --config 0 --steps 20 --warmup 5 > gpurun_out/r02l_bench_config0.json 2>> gpurun_out/r02l_bench.err
tail -3 gpurun_out/r02l_pytest.log; tail -2 gpurun_out/r02l_smoke.log; head -c 300 gpurun_out/r02l_bench_1gpu.json
