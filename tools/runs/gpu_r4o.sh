#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "fmha" -x 2>&1 | tail -5
