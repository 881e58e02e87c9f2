#!/bin/bash
# d=64 (decoder) attention: FMA-pipe share of the exponentials, one thread per row (flags 128 | np << 3) against the defaults
mkdir -p gpurun_out
timeout 400 python tools/fmha_variants.py 0 $((128 | 8)) $((128 | 16)) $((128 | 24)) $((128 | 32)) $((16384)) 2>&1 | tee gpurun_out/fmha_variants_r4v.jsonl
