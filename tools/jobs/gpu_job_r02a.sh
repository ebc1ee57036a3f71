#!/bin/bash
# round-2 first GPU call: full GPU test-suite, chain parity study, bench line, geometry launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/r02a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02a_pytest.log
python tools/chain_parity_report.py --out gpurun_out/r02a_chain_parity.json > gpurun_out/r02a_chain_parity.log 2>&1
echo "chain exit $?" >> gpurun_out/r02a_chain_parity.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench exit $?" >> gpurun_out/r02a_bench.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/r02a_geom_launches.csv \
    python bench.py --config 0 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02a_geom_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1
tail -3 gpurun_out/r02a_pytest.log; tail -2 gpurun_out/r02a_chain_parity.log; tail -c 600 gpurun_out/r02a_bench.err; tail -2 gpurun_out/r02a_smoke.log
