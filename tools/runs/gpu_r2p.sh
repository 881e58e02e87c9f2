#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/fmha_pair_check.py --variants 0,3,4,7 > gpurun_out/fmha_pair_check_r2p.log 2>&1; echo "rc=$?"; grep -v '"shape": \[1, 2\|"shape": \[2, 3' gpurun_out/fmha_pair_check_r2p.log | cut -c1-400; grep -c '"ok": true' gpurun_out/fmha_pair_check_r2p.log; grep -c '"ok": false' gpurun_out/fmha_pair_check_r2p.log
for v in 0 4; do timeout 120 python tools/fmha_pair_trace.py $v > gpurun_out/pair_trace_r2p_v$v.txt 2>&1; tail -3 gpurun_out/pair_trace_r2p_v$v.txt; done
timeout 300 python tools/fmha_variants.py 0 144 > gpurun_out/fmha_variants_r2p.jsonl 2>&1; cat gpurun_out/fmha_variants_r2p.jsonl
