#!/bin/bash
# final pass of the round: every GPU test, smoke(), the default bench line, the reference and torch arms (short)
mkdir -p gpurun_out
tools/gpu_ci.sh > gpurun_out/ci_final.log 2>&1
grep -h "passed\|failed\|rc=\|Error" gpurun_out/ci_final.log | tail -24
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke_final.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_final.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err; echo "reference arm rc=$?"; cut -c1-600 gpurun_out/bench_final_reference.json
timeout 600 python bench.py --impl torch --steps 10 --warmup 3 > gpurun_out/bench_final_torch.json 2> gpurun_out/bench_final_torch.err; echo "torch arm rc=$?"; cut -c1-300 gpurun_out/bench_final_torch.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_final.json"))
g = d["gaussians"]
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["cpu_baseline"], {k: g.get(k) for k in ("decoder_ms", "vae_decode_ms", "e2e_prompt_ms", "e2e_gaussians_per_sec", "decoder_gaussians_per_sec", "cpu_baseline")})
print(d["clocks"], d["gpu_launches"])
PY
