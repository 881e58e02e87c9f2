#!/bin/bash
mkdir -p gpurun_out
tools/gpu_ci.sh tests/test_kernels_gpu.py tests/test_dit_gpu.py tests/test_parity_full_gpu.py > gpurun_out/ci_r3o.log 2>&1
grep -h "passed\|failed\|rc=\|Error" gpurun_out/ci_r3o.log | tail -8
# decoder: multicast in situ
for i in 1 2; do for fl in 0 8; do VIST3A_GEMM_FLAGS=$fl python tools/decoder_profile.py --iters 5 2>/dev/null | head -1 | sed "s/^/gemm_flags $fl: /"; done; done | tee gpurun_out/ab_decoder_multicast_r3o.txt
for i in 1 2; do timeout 300 python bench.py --no-cpu-baseline --no-decoder --steps 40 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('dit default now', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"; done
