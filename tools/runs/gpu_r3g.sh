#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/fmha_variants.py 8192 144 16384 16392 16400 16408 > gpurun_out/fmha_variants_r3g.jsonl 2>&1; cat gpurun_out/fmha_variants_r3g.jsonl | cut -c1-400
