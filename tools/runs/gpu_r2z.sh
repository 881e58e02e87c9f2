#!/bin/bash
mkdir -p gpurun_out
tools/gpu_ci.sh > gpurun_out/ci_r2z.log 2>&1
grep -h "passed\|failed\|rc=\|Error" gpurun_out/ci_r2z.log | tail -24
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r2z.json 2> gpurun_out/bench_r2z.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2z.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2z.json"))
g = d["gaussians"]
print(d["value"], d["ms_per_step"], d["e2e"]["value"], {k: g.get(k) for k in ("decoder_ms", "vae_decode_ms", "e2e_prompt_ms", "e2e_gaussians_per_sec", "decoder_gaussians_per_sec")})
print({k: (round(v["ms"], 3), v["launches"]) for k, v in d["roofline"]["by_kernel"].items()}, d["clocks"])
PY
