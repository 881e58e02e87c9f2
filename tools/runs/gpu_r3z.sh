#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/fmha_tail_trace.py 2>&1 | tee gpurun_out/fmha_tail_trace.txt
