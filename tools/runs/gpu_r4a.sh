#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "fmha_tail or fmha_variants" -x 2>&1 | tail -3
timeout 200 python tools/fmha_tail_trace.py 2>&1 | tee gpurun_out/fmha_tail_trace.txt
timeout 300 python tools/fmha_variants.py 0 0x200000 2>&1 | tee gpurun_out/fmha_variants_r4a.jsonl
