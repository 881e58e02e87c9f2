#!/bin/bash
mkdir -p gpurun_out
python tools/ab_norms.py 2>&1 | tee gpurun_out/ab_norms_r3i_floor.jsonl
