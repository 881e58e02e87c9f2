#!/bin/bash
# default = pair variant 4 (2 of 8 exponentials on the FMA pipe) from 512 keys on: tests, then the step against variant 7 everywhere (alternating)
mkdir -p gpurun_out
tools/gpu_ci.sh tests/test_kernels_gpu.py tests/test_dit_gpu.py tests/test_parity_full_gpu.py > gpurun_out/ci_r4t.log 2>&1
grep -h "passed\|failed\|rc=\|Error" gpurun_out/ci_r4t.log | tail -6
for rnd in 1 2 3; do
  for fl in 0 3840; do
    VIST3A_FMHA_FLAGS=$fl timeout 600 python bench.py --no-decoder --no-cpu-baseline --steps 50 --warmup 5 2> /dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.readline())
print('flags $fl', round(d['value'], 3), round(d['ms_per_step'], 3), d['clocks']['sm_mhz'], d['roofline']['attention']['self']['achieved'])
"
  done
done 2>&1 | tee gpurun_out/ab_step_fmha_np_r4t.txt
