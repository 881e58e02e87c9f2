#!/bin/bash
mkdir -p gpurun_out
for v in 0 2; do timeout 120 python tools/fmha_pair_trace.py $v > gpurun_out/pair_trace_v$v.txt 2>&1; echo rc=$?; cat gpurun_out/pair_trace_v$v.txt | tail -30; done
