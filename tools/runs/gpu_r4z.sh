#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decoder_gpu.py -q -m gpu -x -k "graph or pipeline" 2>&1 | tail -4
