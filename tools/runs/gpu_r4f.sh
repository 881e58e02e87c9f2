#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/fmha_pair_overhead.py 0 2>&1 | tee gpurun_out/fmha_pair_overhead_r4f.txt
timeout 200 python tools/fmha_pair_overhead.py 0x20000 2>&1 | tee -a gpurun_out/fmha_pair_overhead_r4f.txt
