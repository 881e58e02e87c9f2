#!/bin/bash
# pair attention kernel with its own output staging buffer: next item's Q load + first QK^T under the epilogue
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k fmha -x > gpurun_out/ci_r4e.log 2>&1
tail -5 gpurun_out/ci_r4e.log
timeout 200 python tools/fmha_pair_overhead.py 0 2>&1 | tee gpurun_out/fmha_pair_overhead_balanced.txt
timeout 400 python tools/fmha_split_check.py 2>&1 | tee gpurun_out/fmha_split_check_r4e.txt | tail -6
