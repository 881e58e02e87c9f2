#!/bin/bash
# ncu --set full of the reworked DiT attention kernels (self + cross) + the merge kernel; condensed on the box (the report exceeds the 64 MiB merge limit)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/targets_r4m -f \
    python tools/ncu_targets.py --only fmha_dit_self,fmha_dit_cross > gpurun_out/ncu_targets_r4m.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/targets_r4m.ncu-rep --page raw --csv > gpurun_out/targets_r4m_raw.csv 2> /dev/null
python tools/ncu_summary.py /tmp/targets_r4m.ncu-rep > gpurun_out/targets_r4m_summary.txt 2> /dev/null
head -60 gpurun_out/targets_r4m_summary.txt
