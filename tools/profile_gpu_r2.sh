#!/bin/bash
# Round-2 ncu evidence (run under gpurun, 1 GPU):  tools/profile_gpu_r2.sh <tag>   then here: python tools/make_profiles_r2.py <tag>
#   1. launch lists (device time of every kernel) of one eager denoise step of bench.py, one decoder forward, one Wan VAE decode
#   2. ONE --set full run over every hot kernel of the shipped library at its BASELINE shape (tools/ncu_targets.py)
tag=${1:-r2}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-decoder --ncu-step > gpurun_out/launches_${tag}.log 2>&1
echo "dit launches rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_decoder_${tag}.csv \
    python tools/decoder_profile.py --ncu > gpurun_out/launches_decoder_${tag}.log 2>&1
echo "decoder launches rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_vae_${tag}.csv \
    python tools/vae_bench.py --ncu > gpurun_out/launches_vae_${tag}.log 2>&1
echo "vae launches rc=$?"
timeout 1500 ncu --set full --clock-control none --profile-from-start off -o /tmp/targets_${tag} -f \
    python tools/ncu_targets.py > gpurun_out/ncu_targets_${tag}.log 2>&1
echo "targets rc=$?"
# the report itself is too large to travel back (64 MiB limit): condense it here
ncu -i /tmp/targets_${tag}.ncu-rep --page raw --csv > gpurun_out/targets_${tag}_raw.csv 2> /dev/null
python tools/ncu_summary.py /tmp/targets_${tag}.ncu-rep > gpurun_out/targets_${tag}_summary.txt 2> /dev/null
ls -la /tmp/targets_${tag}.ncu-rep gpurun_out/
