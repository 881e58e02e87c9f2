#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, 1 GPU):  tools/profile_gpu.sh <tag>
#   1. launch list (device time of every kernel) of one eager denoise step (cudaProfilerStart/Stop around it) of bench.py
#   2. --set full captures of the attention and GEMM kernels at the DiT shapes
tag=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --ncu-step > gpurun_out/launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fmha_fwd -s 3 -c 2 -o gpurun_out/fmha_${tag} -f \
    python tools/kernel_bench.py --only dit_self --iters 2 --no-torch --no-flush > gpurun_out/ncu_fmha_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fmha_fwd -s 3 -c 1 -o gpurun_out/fmha64_${tag} -f \
    python tools/kernel_bench.py --only dec_global --iters 2 --no-torch --no-flush > gpurun_out/ncu_fmha64_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 6 -c 1 -o gpurun_out/gemm_${tag} -f \
    python tools/kernel_bench.py --only dit_ffn1 --iters 2 --no-torch --no-flush > gpurun_out/ncu_gemm_${tag}.log 2>&1
ls -la gpurun_out/*.ncu-rep
