#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/fmha_variants.py 0 8192 16384 16392 16400 16408 > gpurun_out/fmha_variants_r2s.jsonl 2>&1; echo "rc=$?"; cat gpurun_out/fmha_variants_r2s.jsonl
timeout 120 python tools/fmha_trace.py 1 16 13377 13377 64 16400 > gpurun_out/fmha_trace_r2s_d64.txt 2>&1; tail -9 gpurun_out/fmha_trace_r2s_d64.txt
