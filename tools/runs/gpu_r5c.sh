#!/bin/bash
# one process for two test files (as the round-end driver runs the suite): caches and graphs of one file must not disturb the next
timeout 70 python -m pytest tests/test_decoder_gpu.py tests/test_kernels_gpu.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
