#!/bin/bash
mkdir -p gpurun_out
tools/gpu_ci.sh tests/test_vae_gpu.py > gpurun_out/ci_r2i.log 2>&1
grep -h "passed\|failed\|rc=\|rel-L2\|Error" gpurun_out/ci_r2i.log | tail -30
timeout 300 python tools/vae_bench.py > gpurun_out/vae_bench_r2i.txt 2>&1; echo "vae bench rc=$?"; head -12 gpurun_out/vae_bench_r2i.txt
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2i.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2i.json"))
g = d["gaussians"]
print({k: g[k] for k in ("decoder_ms", "vae_decode_ms", "e2e_prompt_ms", "e2e_gaussians_per_sec", "decoder_gaussians_per_sec")}, d["value"], d["e2e"]["value"])
PY
