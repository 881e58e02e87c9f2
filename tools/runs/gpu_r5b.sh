#!/bin/bash
timeout 300 python -m pytest tests/test_decoder_gpu.py tests/test_kernels_gpu.py -q -m gpu -x -k "graph or pipeline or key_split" 2>&1 | tail -3
