#!/bin/bash
# BASELINE configs[3]: VIST3A-14B, 21 views, prompts sharded over 8 GPUs with the NCCL Gaussian all-gather
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --model 14b --views 21 --steps 10 --warmup 3 --no-cpu-baseline --decoder-iters 2 > gpurun_out/bench_r3c_14b_21v_n8.json 2> gpurun_out/bench_r3c_14b_21v_n8.err; echo "rc=$?"; tail -4 gpurun_out/bench_r3c_14b_21v_n8.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r3c_14b_21v_n8.json"))
g = d["gaussians"]
print(d["value"], d["ms_per_step"], d["e2e"]["value"], {k: g.get(k) for k in ("decoder_ms", "gather_ms", "gather", "vae_decode_ms", "e2e_prompt_ms", "e2e_gaussians_per_sec")})
PY
