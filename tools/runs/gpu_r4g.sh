#!/bin/bash
# same-box A/B of two key-split plans: r-plan (equal chunks of whole units, 1-3 short waves) vs balanced ranges (one equal range of the laid-out tail per cluster)
mkdir -p gpurun_out
for r in 1 2 3; do for l in tools/ab/lib_rplan.so tools/ab/lib_balanced.so; do VIST3A_AB_LIB=$l timeout 200 python tools/ab_fmha_lib.py; done; done 2>&1 | tee gpurun_out/ab_fmha_plans_r4g.txt
