#!/bin/bash
mkdir -p gpurun_out
tools/gpu_ci.sh > gpurun_out/ci_r2q.log 2>&1
grep -h "passed\|failed\|rc=\|Error" gpurun_out/ci_r2q.log | tail -30
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r2q.json 2> gpurun_out/bench_r2q.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2q.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2q.json"))
g = d["gaussians"]
print({k: g.get(k) for k in ("decoder_ms", "vae_decode_ms", "e2e_prompt_ms", "e2e_gaussians_per_sec", "decoder_gaussians_per_sec")}, d["value"], d["ms_per_step"], d["e2e"]["value"])
print(d["roofline"])
print(d.get("clocks"))
PY
timeout 300 python tools/fmha_variants.py 0 8192 > gpurun_out/fmha_variants_r2q.jsonl 2>&1; cat gpurun_out/fmha_variants_r2q.jsonl
