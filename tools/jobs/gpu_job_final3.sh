#!/bin/bash
# validation of the final tree of the round: full GPU suite, smoke, the default bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_final_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_final_bench_1gpu.json 2> gpurun_out/r02_final_bench.err
tail -3 gpurun_out/r02_final_pytest.log; tail -2 gpurun_out/r02_final_smoke.log; head -c 300 gpurun_out/r02_final_bench_1gpu.json
