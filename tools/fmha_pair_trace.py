"""Cycle timeline of the CTA-pair attention kernel (leader CTA of cluster 0): per 128-key step and query tile, when the MMA warp saw P ready /
finished issuing, and when the tile's first softmax warp saw S ready / had S in registers / had the row max / finished the exponentials /
arrived.  Run on the GPU box:  python tools/fmha_pair_trace.py [variant]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vist3a_b200 import _lib, ops  # noqa: E402

var = int(sys.argv[1]) if len(sys.argv) > 1 else 0
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = (torch.randn(2, 4096, 12, 128, device="cuda", generator=g).bfloat16() for _ in range(3))
o = torch.empty_like(q)
for _ in range(2):
    ops.fmha(q, k, v, out=o, flags=256 | (var << 9))
buf = torch.zeros(64 * 2 * 8, dtype=torch.int64, device="cuda")
lib.v3a_debug_fmha_pair_trace.argtypes = [C.c_void_p]
lib.v3a_debug_fmha_pair_trace(buf.data_ptr())
ops.fmha(q, k, v, out=o, flags=256 | (var << 9))
torch.cuda.synchronize()
lib.v3a_debug_fmha_pair_trace(None)
t = buf.cpu().view(64, 2, 8)
t0 = int(t[0, 0, 2])
names = ["P seen", "MMA issued", "S seen", "S loaded", "max done", "exp done", "arrived"]
print(f"variant {var}; cycles relative to the first 'S seen' of tile 0")
for j in list(range(0, 4)) + list(range(12, 20)):
    for i in range(2):
        row = t[j, i]
        if int(row[2]) == 0:
            continue
        print(f"step {j:2d} tile {i}: " + "  ".join(f"{n} {int(row[s]) - t0:7d}" for s, n in enumerate(names)))
# steady-state summary
import statistics
for i in range(2):
    per = [int(t[j + 1, i, 2] - t[j, i, 2]) for j in range(8, 28) if int(t[j + 1, i, 2]) and int(t[j, i, 2])]
    if per:
        ph = {n: statistics.mean(int(t[j, i, b] - t[j, i, a]) for j in range(8, 28)) for n, a, b in
              (("S seen -> loaded", 2, 3), ("loaded -> max", 3, 4), ("max -> exp done", 4, 5), ("exp done -> arrived", 5, 6), ("arrived -> P seen by MMA", 6, 0),
               ("P seen -> MMA issued", 0, 1))}
        nxt = statistics.mean(int(t[j + 1, i, 2] - t[j, i, 1]) for j in range(8, 27))
        print(f"tile {i}: period {statistics.mean(per):.0f} cycles per 128-key step;", "; ".join(f"{n} {v:.0f}" for n, v in ph.items()), f"; MMA issued -> next S seen {nxt:.0f}")
