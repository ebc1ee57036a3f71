#!/bin/bash
# source-level ncu capture of the fused-GroupNorm halo conv (first two launches: in_layers.2 and out_layers.3 at 256^2)
mkdir -p gpurun_out
PDR_QUICK=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -c 2 -o gpurun_out/r02_transform_full \
    python tools/bench_unet.py 8 > gpurun_out/r02_transform_full.log 2>&1
ncu -i gpurun_out/r02_transform_full.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/r02_transform_source.csv 2>/dev/null || \
ncu -i gpurun_out/r02_transform_full.ncu-rep --page source --csv > gpurun_out/r02_transform_source.csv 2>/dev/null
ncu -i gpurun_out/r02_transform_full.ncu-rep --page source --csv > gpurun_out/r02_transform_source_sass.csv 2>/dev/null
rm -f gpurun_out/r02_transform_full.ncu-rep
ls -la gpurun_out/r02_transform*
