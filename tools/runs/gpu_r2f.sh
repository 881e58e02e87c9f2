#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/fmha_pair_check.py --iters 10 > gpurun_out/fmha_pair_check3.log 2>&1
echo "pair check rc=$?"; grep -E '"shape": \[(2, 12|1, 40)|ALL OK|FAILED|"ok": false|rror' gpurun_out/fmha_pair_check3.log | cut -c1-330
for v in 0 2; do timeout 120 python tools/fmha_pair_trace.py $v > gpurun_out/pair_trace3_v$v.txt 2>&1; echo rc=$?; tail -3 gpurun_out/pair_trace3_v$v.txt; done
