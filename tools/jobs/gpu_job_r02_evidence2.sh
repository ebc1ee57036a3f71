#!/bin/bash
# launch lists (repo kernels only: the first launches of these commands are torch's random weight initialisation)
mkdir -p gpurun_out
RX='^(conv_|gn_|sums8|attention|head_|linear_|stem_|resample|splitk|ddnm_|hpr_|raster_|splat_|fill_|unproj_|compact_|scan_|point_vis|mask_half|vertex_|rescale|minmax|crop_|atlas_|sparse_|paint_|dilate|count_|randn)'
PDR_QUICK=1 timeout 600 ncu -k regex:"$RX" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 360 \
    --csv --log-file gpurun_out/r02_unet_forward_launches.csv python tools/bench_unet.py 8 > gpurun_out/r02_unet_forward_ncu.log 2>&1
timeout 900 ncu -k regex:"$RX" --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-extras > gpurun_out/r02_bench_under_ncu.log 2>&1
grep -c "" gpurun_out/r02_unet_forward_launches.csv gpurun_out/r02_bench_launches.csv; tail -1 gpurun_out/r02_unet_forward_ncu.log
