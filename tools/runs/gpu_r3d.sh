#!/bin/bash
# BASELINE configs[4] corner: VIST3A-1.3B, 4 prompts batched per GPU on 8 GPUs (32 prompts in flight)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --prompts-per-gpu 4 --steps 10 --warmup 3 --no-cpu-baseline --decoder-iters 2 > gpurun_out/bench_r3d_n8_b4.json 2> gpurun_out/bench_r3d_n8_b4.err; echo "rc=$?"; tail -4 gpurun_out/bench_r3d_n8_b4.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r3d_n8_b4.json"))
g = d["gaussians"]
print(d["value"], d["ms_per_step"], d["e2e"]["value"], {k: g.get(k) for k in ("decoder_ms", "gather_ms", "gather", "vae_decode_ms", "e2e_prompt_ms", "e2e_gaussians_per_sec")})
PY
