#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r4k.json 2> gpurun_out/bench_r4k.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r4k.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r4k.json"))
g = d["gaussians"]
print(d["value"], d["ms_per_step"], d["e2e"]["value"], {k: g.get(k) for k in ("decoder_ms", "vae_decode_ms", "e2e_prompt_ms", "e2e_gaussians_per_sec", "decoder_gaussians_per_sec")})
print(d["clocks"], d["gpu_launches"], d["roofline"]["attention"])
PY
