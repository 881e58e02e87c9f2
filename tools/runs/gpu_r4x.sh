#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/decoder_graph_try.py 2>&1 | tail -8 | tee gpurun_out/decoder_graph_try.txt
