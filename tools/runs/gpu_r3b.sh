#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_epilogue_bench.py > gpurun_out/gemm_epilogue_r3b.jsonl 2>&1; cut -c1-1500 gpurun_out/gemm_epilogue_r3b.jsonl
tools/gpu_ci.sh tests/test_kernels_gpu.py tests/test_decoder_kernels_gpu.py tests/test_vae_gpu.py tests/test_dit_gpu.py tests/test_decoder_gpu.py > gpurun_out/ci_r3b.log 2>&1
grep -h "passed\|failed\|rc=\|Error" gpurun_out/ci_r3b.log | tail -12
timeout 300 python tools/vae_bench.py > gpurun_out/vae_bench_r3b.txt 2>&1; head -8 gpurun_out/vae_bench_r3b.txt
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r3b.json 2> gpurun_out/bench_r3b.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r3b.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r3b.json"))
g = d["gaussians"]
print(d["value"], d["ms_per_step"], d["e2e"]["value"], {k: g.get(k) for k in ("decoder_ms", "vae_decode_ms", "e2e_prompt_ms", "e2e_gaussians_per_sec", "decoder_gaussians_per_sec")})
print({k: (round(v["ms"], 3), v["launches"]) for k, v in d["roofline"]["by_kernel"].items()}, d["clocks"])
PY
