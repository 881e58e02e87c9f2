#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/fmha_variants.py $((256 | (7 << 9))) $((256 | (4 << 9))) $((256 | (6 << 9))) $((256 | (3 << 9))) 2>&1 | tee gpurun_out/fmha_variants_r4s.jsonl
