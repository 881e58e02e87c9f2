#!/bin/bash
mkdir -p gpurun_out
tools/gpu_ci.sh tests/test_kernels_gpu.py tests/test_decoder_kernels_gpu.py tests/test_decoder_gpu.py tests/test_dit_gpu.py tests/test_vae_gpu.py > gpurun_out/ci_r2v.log 2>&1
grep -h "passed\|failed\|rc=\|Error" gpurun_out/ci_r2v.log | tail -20
python tools/decoder_profile.py --detail > gpurun_out/decoder_profile_detail_r2v.txt 2>&1; head -12 gpurun_out/decoder_profile_detail_r2v.txt
timeout 600 python bench.py --no-cpu-baseline --no-vae > gpurun_out/bench_r2v.json 2> gpurun_out/bench_r2v.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2v.json"))
g = d["gaussians"]
print(d["value"], d["ms_per_step"], d["e2e"]["value"], g["decoder_ms"], d["roofline"]["by_kernel"]["gemm_tcgen05"], d["clocks"])
PY
