#!/bin/bash
mkdir -p gpurun_out
tools/gpu_ci.sh tests/test_kernels_gpu.py tests/test_decoder_gpu.py tests/test_teacher_gpu.py > gpurun_out/ci_r3r.log 2>&1
grep -h "passed\|failed\|rc=\|Error\|FAILED" gpurun_out/ci_r3r.log | tail -8
python tools/decoder_profile.py --detail 2>/dev/null | head -8
