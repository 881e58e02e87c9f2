#!/bin/bash
mkdir -p gpurun_out
tools/ubench/exp_half64 > gpurun_out/ubench_exp_half64.txt 2>&1; cat gpurun_out/ubench_exp_half64.txt
timeout 600 python tools/fmha_variants.py 130 146 > gpurun_out/fmha_variants_r2m.jsonl 2>&1; echo "rc=$?"; cat gpurun_out/fmha_variants_r2m.jsonl
