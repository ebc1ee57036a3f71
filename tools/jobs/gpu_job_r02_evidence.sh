#!/bin/bash
# round-2 ncu evidence on the final kernels: launch list + DRAM traffic of one U-Net forward, full capture of the
# dominant (fused GroupNorm) conv launches, launch list of the default bench command
mkdir -p gpurun_out
PDR_QUICK=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 \
    --csv --log-file gpurun_out/r02_unet_forward_launches.csv python tools/bench_unet.py 8 > gpurun_out/r02_unet_forward_ncu.log 2>&1
PDR_QUICK=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -c 4 -o gpurun_out/r02_conv_halo_full \
    python tools/bench_unet.py 8 > gpurun_out/r02_conv_halo_full.log 2>&1
ncu -i gpurun_out/r02_conv_halo_full.ncu-rep --page raw --csv > gpurun_out/r02_conv_halo_full_raw.csv 2>/dev/null
rm -f gpurun_out/r02_conv_halo_full.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-extras > gpurun_out/r02_bench_under_ncu.log 2>&1
ls -la gpurun_out/r02_*; tail -2 gpurun_out/r02_unet_forward_ncu.log; tail -2 gpurun_out/r02_bench_under_ncu.log | cut -c1-200
