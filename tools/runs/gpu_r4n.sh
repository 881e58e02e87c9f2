#!/bin/bash
# launch list of one eager denoise step with the final build (attention: pair kernel for self and cross + merge kernel)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r2f.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-decoder --ncu-step > gpurun_out/launches_r2f.log 2>&1
echo "dit launches rc=$?"; wc -l gpurun_out/launches_r2f.csv
