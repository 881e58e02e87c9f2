#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_epilogue_bench.py > gpurun_out/gemm_epilogue_r2x.jsonl 2>&1; cat gpurun_out/gemm_epilogue_r2x.jsonl
tools/gpu_ci.sh tests/test_kernels_gpu.py tests/test_decoder_kernels_gpu.py > gpurun_out/ci_r2x.log 2>&1
grep -h "passed\|failed\|rc=\|Error" gpurun_out/ci_r2x.log | tail -8
