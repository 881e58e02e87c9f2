#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/fmha_variants.py 0 2 3 130 138 146 154 > gpurun_out/fmha_variants_r2l.jsonl 2>&1; echo "rc=$?"; cat gpurun_out/fmha_variants_r2l.jsonl
for fl in 3 130 146; do timeout 120 python tools/fmha_trace.py 2 12 4096 4096 128 $fl > gpurun_out/fmha_trace_r2l_d128_f$fl.txt 2>&1; tail -22 gpurun_out/fmha_trace_r2l_d128_f$fl.txt; done
