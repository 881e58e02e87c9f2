#!/bin/bash
# AdaLN vectors of all blocks in B launches per forward: DiT parity tests, then the bench line
mkdir -p gpurun_out
tools/gpu_ci.sh tests/test_dit_gpu.py tests/test_parity_full_gpu.py > gpurun_out/ci_r4p.log 2>&1
grep -h "passed\|failed\|rc=\|Error" gpurun_out/ci_r4p.log | tail -6
for i in 1 2; do
timeout 600 python bench.py --no-decoder --no-cpu-baseline --steps 50 --warmup 5 2> /dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.readline())
print(round(d['value'], 3), round(d['ms_per_step'], 3), d['clocks']['sm_mhz'], d['gpu_launches'])
"
done
