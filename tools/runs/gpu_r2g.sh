#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/fmha_pair_check.py --iters 10 --variants 2,4,5,6,7 > gpurun_out/fmha_pair_check4.log 2>&1
echo "pair check rc=$?"; grep -E '"shape": \[(2, 12|1, 40)|ALL OK|FAILED|"ok": false|rror' gpurun_out/fmha_pair_check4.log | cut -c150-330
