#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/pair_one.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from vist3a_b200 import ops
var = int(sys.argv[1])
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = (torch.randn(2, 4096, 12, 128, device="cuda", generator=g).bfloat16() for _ in range(3))
o = torch.empty_like(q)
for _ in range(3):
    ops.fmha(q, k, v, out=o, flags=256 | (var << 9))
torch.cuda.synchronize()
PY
for var in 0 1; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fmha_pair -s 2 -c 1 -o gpurun_out/pair_v${var} -f python /tmp/pair_one.py $var > gpurun_out/ncu_pair_v${var}.log 2>&1
echo "ncu pair v$var rc=$?"
done
