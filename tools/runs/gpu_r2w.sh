#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_epilogue_bench.py > gpurun_out/gemm_epilogue_r2w.jsonl 2>&1; cat gpurun_out/gemm_epilogue_r2w.jsonl
