#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k fmha -x 2>&1 | tail -3
timeout 400 python tools/fmha_split_check.py 2>&1 | tee gpurun_out/fmha_split_check_r4i.txt | tail -6
