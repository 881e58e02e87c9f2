#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/fmha_variants.py 0 8192 144 16384 65552 > gpurun_out/fmha_variants_r3m.jsonl 2>&1; cat gpurun_out/fmha_variants_r3m.jsonl | cut -c1-400
