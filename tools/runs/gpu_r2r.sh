#!/bin/bash
mkdir -p gpurun_out
(VIST3A_FMHA_FLAGS=0 python tools/decoder_numerics_ab.py; VIST3A_FMHA_FLAGS=8192 python tools/decoder_numerics_ab.py; VIST3A_FMHA_FLAGS=128 python tools/decoder_numerics_ab.py) > gpurun_out/decoder_numerics_ab.txt 2>&1; cat gpurun_out/decoder_numerics_ab.txt | tail -20
