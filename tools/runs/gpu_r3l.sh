#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do for l in tools/ab/lib_A.so tools/ab/lib_V1.so tools/ab/lib_V3.so; do VIST3A_AB_LIB=$l python tools/ab_norms.py; done; done 2>&1 | tee gpurun_out/ab_norms_r3l.jsonl
