#!/bin/bash
# view-sharded single-shape mode on N GPUs: bitwise check vs one GPU + latency per shape
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/verify_view_sharding.py 100 > gpurun_out/r02k_views_verify_${N}gpu.json 2> gpurun_out/r02k_views_verify_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --shard views --steps 3 --warmup 2 > gpurun_out/r02k_bench_views_${N}gpu.json 2> gpurun_out/r02k_bench_views_${N}gpu.err
if [ "$N" = "8" ] || [ "$N" = "4" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r02k_bench_${N}gpu.json 2> gpurun_out/r02k_bench_${N}gpu.err
fi
tail -2 gpurun_out/r02k_views_verify_${N}gpu.json; head -c 700 gpurun_out/r02k_bench_views_${N}gpu.json; echo; tail -3 gpurun_out/r02k_bench_views_${N}gpu.err
