#!/bin/bash
mkdir -p gpurun_out
tools/ubench/exp_half64 > gpurun_out/ubench_exp_half64_v2.txt 2>&1; cat gpurun_out/ubench_exp_half64_v2.txt
for fl in 4224 4226; do timeout 120 python tools/fmha_trace.py 2 12 4096 4096 128 $fl > gpurun_out/fmha_trace_r2n_d128_f$fl.txt 2>&1; tail -22 gpurun_out/fmha_trace_r2n_d128_f$fl.txt; done
timeout 120 python tools/fmha_trace.py 1 16 13377 13377 64 4224 > gpurun_out/fmha_trace_r2n_d64_f4224.txt 2>&1; tail -12 gpurun_out/fmha_trace_r2n_d64_f4224.txt
