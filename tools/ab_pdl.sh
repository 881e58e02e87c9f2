#!/bin/bash
# A/B of programmatic dependent launch on the GPU box:  tools/ab_pdl.sh   (writes gpurun_out/ab_pdl_*.json)
mkdir -p gpurun_out
for pdl in 1 0 1 0; do
  VIST3A_PDL=$pdl timeout 600 python bench.py --no-cpu-baseline --no-decoder --steps 30 --warmup 3 > gpurun_out/ab_pdl_${pdl}_$RANDOM.json 2> gpurun_out/ab_pdl_${pdl}.err
done
for pdl in 1 0; do
  VIST3A_PDL=$pdl timeout 600 python tools/decoder_profile.py > gpurun_out/ab_pdl_dec_${pdl}.log 2>&1
done
grep -h ms_per_step gpurun_out/ab_pdl_*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['ms_per_step'], d['value'], d['roofline']['by_kernel']['gemm_tcgen05']['ms'], d['clocks'])
"
tail -3 gpurun_out/ab_pdl_dec_*.log
