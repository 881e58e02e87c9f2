#!/bin/bash
# Run on the GPU box (under gpurun): each test file in its own process with a timeout, so that a
# device-side trap in one kernel does not take the remaining files down with it.
#   tools/gpu_ci.sh [pytest files...]   (default: every tests/test_*gpu*.py)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.csv 2>&1
files=("$@")
if [ ${#files[@]} -eq 0 ]; then files=(tests/test_*gpu*.py); fi
rc_all=0
for f in "${files[@]}"; do
  name=$(basename "$f" .py)
  echo "=== $f"
  timeout 900 python -m pytest "$f" -m gpu -q --no-header -s -p no:cacheprovider 2>&1 | tail -60 | tee "gpurun_out/${name}.log"
  rc=${PIPESTATUS[0]}
  echo "rc=$rc" | tee -a "gpurun_out/${name}.log"
  [ $rc -ne 0 ] && rc_all=1
done
exit $rc_all
