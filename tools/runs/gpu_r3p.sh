#!/bin/bash
mkdir -p gpurun_out
tools/gpu_ci.sh tests/test_kernels_gpu.py > gpurun_out/ci_r3p.log 2>&1
grep -h "passed\|failed\|rc=\|Error\|FAILED" gpurun_out/ci_r3p.log | tail -20
