#!/bin/bash
mkdir -p gpurun_out
tools/gpu_ci.sh tests/test_vae_gpu.py > gpurun_out/ci_r2h.log 2>&1
grep -h "passed\|failed\|rc=\|rel-L2\|Error\|error" gpurun_out/ci_r2h.log | tail -30
timeout 300 python tools/vae_bench.py --detail > gpurun_out/vae_bench_r2h.txt 2>&1; echo "vae bench rc=$?"; tail -45 gpurun_out/vae_bench_r2h.txt
