#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/fmha_variants.py 0 144 3840 2304 > gpurun_out/fmha_variants_r3q.jsonl 2>&1; cat gpurun_out/fmha_variants_r3q.jsonl | cut -c1-300
