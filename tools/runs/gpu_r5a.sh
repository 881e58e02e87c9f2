#!/bin/bash
# last pass of the round: every GPU test + smoke() on the final tree
mkdir -p gpurun_out
tools/gpu_ci.sh > gpurun_out/ci_r5a.log 2>&1
grep -h "passed\|failed\|rc=\|Error" gpurun_out/ci_r5a.log | tail -24
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r5a.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r5a.log
