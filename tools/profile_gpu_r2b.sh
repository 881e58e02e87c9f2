#!/bin/bash
# second half of tools/profile_gpu_r2.sh only (the --set full targets), for a re-run
tag=${1:-r2}
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --profile-from-start off -o /tmp/targets_${tag} -f \
    python tools/ncu_targets.py > gpurun_out/ncu_targets_${tag}.log 2>&1
echo "targets rc=$?"
ncu -i /tmp/targets_${tag}.ncu-rep --page raw --csv > gpurun_out/targets_${tag}_raw.csv 2> /dev/null
python tools/ncu_summary.py /tmp/targets_${tag}.ncu-rep > gpurun_out/targets_${tag}_summary.txt 2> /dev/null
ls -la /tmp/targets_${tag}.ncu-rep gpurun_out/
