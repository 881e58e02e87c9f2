#!/bin/bash
mkdir -p gpurun_out
python tools/decoder_profile.py --detail > gpurun_out/decoder_profile_detail_r2u.txt 2>&1; head -70 gpurun_out/decoder_profile_detail_r2u.txt
