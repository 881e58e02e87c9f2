#!/bin/bash
# after the attention-kernel rework (persistent clusters, key split, staged output): every GPU test, smoke, the default bench line
mkdir -p gpurun_out
tools/gpu_ci.sh > gpurun_out/ci_r4w.log 2>&1
grep -h "passed\|failed\|rc=\|Error" gpurun_out/ci_r4w.log | tail -24
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r4w.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_r4w.log
timeout 900 python bench.py > gpurun_out/bench_r4w.json 2> gpurun_out/bench_r4w.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r4w.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r4w.json"))
g = d["gaussians"]
print(d["value"], d["ms_per_step"], d["e2e"]["value"], {k: g.get(k) for k in ("decoder_ms", "vae_decode_ms", "e2e_prompt_ms", "e2e_gaussians_per_sec", "decoder_gaussians_per_sec")})
print(d["clocks"], d["gpu_launches"], d["roofline"])
PY
