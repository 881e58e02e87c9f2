#!/bin/bash
mkdir -p gpurun_out
tools/gpu_ci.sh tests/test_decoder_kernels_gpu.py tests/test_decoder_gpu.py tests/test_kernels_gpu.py tests/test_teacher_gpu.py > gpurun_out/ci_r3e.log 2>&1
grep -h "passed\|failed\|rc=\|Error" gpurun_out/ci_r3e.log | tail -10
python tools/decoder_profile.py --detail > gpurun_out/decoder_profile_detail_r3e.txt 2>&1; head -14 gpurun_out/decoder_profile_detail_r3e.txt
timeout 300 python tools/fmha_variants.py 0 8192 > gpurun_out/fmha_variants_r3e.jsonl 2>&1; cat gpurun_out/fmha_variants_r3e.jsonl
