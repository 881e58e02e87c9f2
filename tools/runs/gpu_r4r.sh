#!/bin/bash
# FMA-pipe share of the exponentials (pair kernel variants 3 / 4 / 7 = 0 / 2 / 3 of 8 column pairs) re-checked on the persistent kernel
mkdir -p gpurun_out
timeout 300 python tools/fmha_variants.py 0 $((256 | (3 << 9))) $((256 | (4 << 9))) $((256 | (7 << 9))) 2>&1 | tee gpurun_out/fmha_variants_r4r.jsonl
