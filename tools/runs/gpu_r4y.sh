#!/bin/bash
# stitched decoder replayed from a CUDA graph: tests, then the bench line
mkdir -p gpurun_out
tools/gpu_ci.sh tests/test_decoder_gpu.py tests/test_parity_full_gpu.py > gpurun_out/ci_r4y.log 2>&1
grep -h "passed\|failed\|rc=\|Error\|error" gpurun_out/ci_r4y.log | tail -8
timeout 900 python bench.py > gpurun_out/bench_r4y.json 2> gpurun_out/bench_r4y.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r4y.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r4y.json"))
g = d["gaussians"]
print(d["value"], d["ms_per_step"], d["e2e"]["value"], {k: g.get(k) for k in ("decoder_ms", "vae_decode_ms", "e2e_prompt_ms", "e2e_gaussians_per_sec", "decoder_gaussians_per_sec", "decoder_launches_per_forward", "decoder_cuda_graph")})
PY
