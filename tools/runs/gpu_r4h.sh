#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/fmha_split_check.py 2>&1 | tee gpurun_out/fmha_split_check_r4h.txt | tail -6
