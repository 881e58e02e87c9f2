#!/bin/bash
# CUDA-core tail rows in the single-CTA attention kernel: correctness, then frame-attention timing with / without (flags bit 21)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k fmha -x > gpurun_out/ci_r3y.log 2>&1
tail -5 gpurun_out/ci_r3y.log
timeout 300 python tools/fmha_variants.py 0 0x200000 2>&1 | tee gpurun_out/fmha_variants_r3y.jsonl
