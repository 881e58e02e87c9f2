#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, 1 GPU):  tools/profile_gpu.sh <tag>   then here: python tools/make_profiles.py <tag>
#   1. launch lists (device time of every kernel) of one eager denoise step of bench.py and of one decoder forward
#   2. --set full captures of the attention kernels (d=128 DiT self-attention, d=64 decoder global attention), the bf16 GEMM,
#      a TF32 implicit-GEMM convolution and the Gaussian epilogue
tag=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-decoder --ncu-step > gpurun_out/launches_${tag}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_decoder_${tag}.csv \
    python tools/decoder_profile.py --ncu > gpurun_out/launches_decoder_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fmha_fwd -s 3 -c 1 -o gpurun_out/fmha_${tag} -f \
    python tools/kernel_bench.py --only dit_self --iters 2 --no-torch --no-flush > gpurun_out/ncu_fmha_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fmha_fwd -s 3 -c 1 -o gpurun_out/fmha64_${tag} -f \
    python tools/kernel_bench.py --only dec_global --iters 2 --no-torch --no-flush > gpurun_out/ncu_fmha64_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 6 -c 1 -o gpurun_out/gemm_${tag} -f \
    python tools/kernel_bench.py --only dit_ffn1 --iters 2 --no-torch --no-flush > gpurun_out/ncu_gemm_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gaussian_epilogue -c 1 -o gpurun_out/gauss_${tag} -f \
    python tools/decoder_profile.py --ncu-all > gpurun_out/ncu_gauss_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:voxel_reduce -s 2 -c 1 -o gpurun_out/voxel_${tag} -f \
    python tools/kernel_bench.py --only voxel --iters 2 --no-torch --no-flush > gpurun_out/ncu_voxel_${tag}.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'voxel|radix|seg_' -s 64 -c 32 --csv \
    --log-file gpurun_out/launches_voxel_${tag}.csv python tools/kernel_bench.py --only voxel --no-torch --iters 2 > gpurun_out/launches_voxel_${tag}.log 2>&1
ls -la gpurun_out/*_${tag}.ncu-rep gpurun_out/launches*_${tag}.csv
