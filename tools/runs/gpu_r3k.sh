#!/bin/bash
mkdir -p gpurun_out
for fl in 65552 65560; do timeout 120 python tools/fmha_trace.py 2 12 4096 4096 128 $fl > gpurun_out/fmha_trace_r3k_d128_f$fl.txt 2>&1; tail -9 gpurun_out/fmha_trace_r3k_d128_f$fl.txt; done
for fl in 65536 16384 144; do timeout 120 python tools/fmha_trace.py 1 16 13377 13377 64 $fl > gpurun_out/fmha_trace_r3k_d64_f$fl.txt 2>&1; tail -9 gpurun_out/fmha_trace_r3k_d64_f$fl.txt; done
