#!/bin/bash
# in-step A/B under the power cap: A-operand multicast (less L2->SM traffic) / staged epilogue vs default
mkdir -p gpurun_out
for i in 1 2; do for fl in 0 8 16; do VIST3A_GEMM_FLAGS=$fl timeout 300 python bench.py --no-cpu-baseline --no-decoder --steps 40 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('gemm_flags', $fl, d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"; done; done | tee gpurun_out/ab_gemm_flags_r3n.txt
