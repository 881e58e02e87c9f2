#!/bin/bash
mkdir -p gpurun_out
tools/gpu_ci.sh tests/test_vae_gpu.py > gpurun_out/ci_r3a.log 2>&1
grep -h "passed\|failed\|rc=\|Error\|rel-L2" gpurun_out/ci_r3a.log | tail -12
