#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/fmha_variants.py 0 8192 144 152 32912 16384 > gpurun_out/fmha_variants_r3f.jsonl 2>&1; cat gpurun_out/fmha_variants_r3f.jsonl
