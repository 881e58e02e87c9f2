#!/bin/bash
mkdir -p gpurun_out
tools/gpu_ci.sh tests/test_vae_gpu.py tests/test_decoder_gpu.py > gpurun_out/ci_r2j.log 2>&1
grep -h "passed\|failed\|rc=\|Error\|assert" gpurun_out/ci_r2j.log | tail -30
python tools/decoder_profile.py > gpurun_out/decoder_profile_r2j.txt 2>&1; head -40 gpurun_out/decoder_profile_r2j.txt
