#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/fmha_variants.py 0 1 128 136 144 152 160 > gpurun_out/fmha_variants_r2k.jsonl 2>&1; echo "rc=$?"; cat gpurun_out/fmha_variants_r2k.jsonl
for fl in 0 128 144; do timeout 120 python tools/fmha_trace.py 2 12 4096 4096 128 $fl > gpurun_out/fmha_trace_r2k_d128_f$fl.txt 2>&1; tail -22 gpurun_out/fmha_trace_r2k_d128_f$fl.txt; done
for fl in 0 128 152; do timeout 120 python tools/fmha_trace.py 1 16 13377 13377 64 $fl > gpurun_out/fmha_trace_r2k_d64_f$fl.txt 2>&1; tail -12 gpurun_out/fmha_trace_r2k_d64_f$fl.txt; done
