#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2t_n2.json 2> gpurun_out/bench_r2t_n2.err; echo "rc=$?"; tail -5 gpurun_out/bench_r2t_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2t_n2.json"))
g = d["gaussians"]
print(d["value"], d["ms_per_step"], d["e2e"]["value"], {k: g.get(k) for k in ("decoder_ms", "gather_ms", "gather", "e2e_prompt_ms", "e2e_gaussians_per_sec")})
PY
