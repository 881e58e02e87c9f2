#!/bin/bash
# in-step A/B of the attention decomposition: default (persistent clusters + key split) vs the round-2 decomposition (flags bits 17 + 20), interleaved
mkdir -p gpurun_out
for rnd in 1 2 3; do
  for fl in 0 0x120000; do
    VIST3A_FMHA_FLAGS=$fl timeout 600 python bench.py --no-decoder --no-cpu-baseline --steps 50 --warmup 5 2> /dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.readline())
print('flags $fl round $rnd', round(d['value'], 3), round(d['ms_per_step'], 3), d['clocks']['sm_mhz'], d['roofline']['attention']['self']['ms'], d['roofline']['attention']['cross']['ms'])
"
  done
done 2>&1 | tee gpurun_out/ab_step_fmha_r4c.txt
