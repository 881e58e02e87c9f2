#!/bin/bash
# same-box A/B: AdaLN vectors per block (30 launches, HEAD) vs all blocks in B launches (new), alternating
mkdir -p gpurun_out
for rnd in 1 2 3; do
  for v in head new; do
    cp tools/ab/wan_dit_$v.py vist3a_b200/wan_dit.py
    timeout 600 python bench.py --no-decoder --no-cpu-baseline --steps 50 --warmup 5 2> /dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.readline())
print('$v', round(d['value'], 3), round(d['ms_per_step'], 3), d['clocks']['sm_mhz'], d['gpu_launches'])
"
  done
done 2>&1 | tee gpurun_out/ab_step_modulation_r4q.txt
