#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/fmha_variants.py 0 128 136 144 152 130 138 146 154 > gpurun_out/fmha_variants_r2o.jsonl 2>&1; echo "rc=$?"; cat gpurun_out/fmha_variants_r2o.jsonl
for fl in 136 138; do timeout 120 python tools/fmha_trace.py 2 12 4096 4096 128 $fl > gpurun_out/fmha_trace_r2o_d128_f$fl.txt 2>&1; tail -9 gpurun_out/fmha_trace_r2o_d128_f$fl.txt; done
timeout 120 python tools/fmha_trace.py 1 16 13377 13377 64 136 > gpurun_out/fmha_trace_r2o_d64_f136.txt 2>&1; tail -9 gpurun_out/fmha_trace_r2o_d64_f136.txt
