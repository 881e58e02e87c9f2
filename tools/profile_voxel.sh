#!/bin/bash
# ncu launch list of one voxelised fusion at the BASELINE size (run under gpurun, 1 GPU)
tag=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'voxel|radix|seg_|gaussian_adapter' -s 60 -c 40 --csv \
    --log-file gpurun_out/launches_voxel_${tag}.csv python tools/kernel_bench.py --only voxel --no-torch --iters 2 > gpurun_out/launches_voxel_${tag}.log 2>&1
tail -45 gpurun_out/launches_voxel_${tag}.csv | cut -d, -f5,12-16 | tr -d '"'
