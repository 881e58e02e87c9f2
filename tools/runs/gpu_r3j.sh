#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/fmha_variants.py 0 8192 65536 65544 65552 65560 > gpurun_out/fmha_variants_r3j.jsonl 2>&1; cat gpurun_out/fmha_variants_r3j.jsonl | cut -c1-400
