#!/bin/bash
# Round-2 first GPU pass (run under gpurun, 1 GPU): all GPU tests incl. the BASELINE-size parity tests, the default bench line, the
# torch comparator arm, per-shape kernel bench next to cuBLAS / SDPA, and ONE ncu --set full run over every hot kernel of the shipped library.
tag=${1:-r2a}
mkdir -p gpurun_out
(timeout 25 python -m pip download diffusers==0.33.1 -d /tmp/dl > gpurun_out/pip_download_${tag}.log 2>&1; echo "rc=$?" >> gpurun_out/pip_download_${tag}.log)
tools/gpu_ci.sh > gpurun_out/ci_${tag}.log 2>&1
echo "ci rc=$?"
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_${tag}.json
python bench.py --impl torch --steps 10 --warmup 3 > gpurun_out/bench_torch_${tag}.json 2> gpurun_out/bench_torch_${tag}.err
echo "torch arm rc=$?"; cat gpurun_out/bench_torch_${tag}.json
python tools/kernel_bench.py --iters 20 > gpurun_out/kernel_bench_${tag}.jsonl 2> gpurun_out/kernel_bench_${tag}.err
echo "kernel_bench rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/targets_${tag} -f \
    python tools/ncu_targets.py > gpurun_out/ncu_targets_${tag}.log 2>&1
echo "ncu targets rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gaussian_epilogue|bias_act_t|depth_to_space|pose_to_cameras' -c 6 -o gpurun_out/decoder_misc_${tag} -f \
    python tools/decoder_profile.py --ncu-all > gpurun_out/ncu_decoder_misc_${tag}.log 2>&1
echo "ncu decoder misc rc=$?"
grep -h "passed\|failed\|rc=" gpurun_out/ci_${tag}.log | tail -40
ls -la gpurun_out/*_${tag}*
