#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/fmha_pair_check.py --iters 10 > gpurun_out/fmha_pair_check.log 2>&1
echo "pair check rc=$?"; tail -45 gpurun_out/fmha_pair_check.log
tools/gpu_ci.sh tests/test_parity_full_gpu.py tests/test_decoder_gpu.py tests/test_kernels_gpu.py tests/test_decoder_kernels_gpu.py > gpurun_out/ci_r2b.log 2>&1
grep -h "passed\|failed\|rc=\|rel-L2" gpurun_out/ci_r2b.log | tail -20
