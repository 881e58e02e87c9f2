#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_trace.py > gpurun_out/gemm_trace_r2y.txt 2>&1; cat gpurun_out/gemm_trace_r2y.txt
