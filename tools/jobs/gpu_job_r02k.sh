#!/bin/bash
# final evidence run of round 2 on the committed code (fused GroupNorm default, tcgen05 attention)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/r02k_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02k_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02k_smoke.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r02k_bench_1gpu.json 2> gpurun_out/r02k_bench.err
PDR_QUICK=1 ncu --set full --clock-control none --import-source on -k regex:'attention' -c 16 -o gpurun_out/r02k_attention_full \
    python tools/bench_unet.py 8 > gpurun_out/r02k_attention_ncu.log 2>&1
ncu -i gpurun_out/r02k_attention_full.ncu-rep --page raw --csv > gpurun_out/r02k_attention_full_raw.csv 2>/dev/null
rm -f gpurun_out/r02k_attention_full.ncu-rep
PDR_QUICK=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02k_forward_launches.csv \
    python tools/bench_unet.py 8 > /dev/null 2>&1
python bench.py --config 0 --steps 20 --warmup 5 > gpurun_out/r02k_bench_config0.json 2>> gpurun_out/r02k_bench.err
tail -3 gpurun_out/r02k_pytest.log; tail -2 gpurun_out/r02k_smoke.log; head -c 300 gpurun_out/r02k_bench_1gpu.json
