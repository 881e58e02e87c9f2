#!/bin/bash
# key-split last wave of the pair attention kernel: correctness, then A/B timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k fmha -x > gpurun_out/ci_r3s.log 2>&1
tail -5 gpurun_out/ci_r3s.log
timeout 300 python tools/fmha_split_check.py 2>&1 | tee gpurun_out/fmha_split_check.txt | tail -12
