#!/bin/bash
# final build: ncu --set full of the attention launches, then the default bench line
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/targets_r4u -f \
    python tools/ncu_targets.py --only fmha_dit_self,fmha_dit_cross > gpurun_out/ncu_targets_r4u.log 2>&1
echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/targets_r4u.ncu-rep > gpurun_out/targets_r4u_summary.txt 2> /dev/null
grep "kernel:\|gpu__time_duration\|sm__cycles_elapsed.max \|tensor_cycles_active" gpurun_out/targets_r4u_summary.txt
timeout 900 python bench.py > gpurun_out/bench_r4u.json 2> gpurun_out/bench_r4u.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r4u.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r4u.json"))
g = d["gaussians"]
print(d["value"], d["ms_per_step"], d["e2e"]["value"], {k: g.get(k) for k in ("decoder_ms", "vae_decode_ms", "e2e_prompt_ms", "e2e_gaussians_per_sec", "decoder_gaussians_per_sec")})
print(d["clocks"], d["gpu_launches"], d["roofline"]["attention"])
PY
