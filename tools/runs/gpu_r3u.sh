#!/bin/bash
# persistent clusters in the pair attention kernel: correctness, overhead stamps, A/B timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k fmha -x > gpurun_out/ci_r3u.log 2>&1
tail -5 gpurun_out/ci_r3u.log
timeout 200 python tools/fmha_pair_overhead.py 0 2>&1 | tee gpurun_out/fmha_pair_overhead_persistent.txt
timeout 200 python tools/fmha_pair_overhead.py 0x100000 2>&1 | tee -a gpurun_out/fmha_pair_overhead_persistent.txt
timeout 300 python tools/fmha_split_check.py 2>&1 | tee gpurun_out/fmha_split_check_r3u.txt | tail -12
