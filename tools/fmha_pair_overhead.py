"""Where a cluster of the CTA-pair attention kernel spends its time outside the steady-state key loop: %globaltimer stamps of every cluster's
leader CTA (fmha_pair_sm100.cu PAIR_STAMP), grouped by SM to show the gap between consecutive clusters on the same SM pair.
Run on the GPU box:  python tools/fmha_pair_overhead.py [flags]"""
import ctypes as C
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vist3a_b200 import _lib, ops  # noqa: E402

flags = int(sys.argv[1], 0) if len(sys.argv) > 1 else (1 << 17)
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = (torch.randn(2, 4096, 12, 128, device="cuda", generator=g).bfloat16() for _ in range(3))
o = torch.empty_like(q)
for _ in range(3):
    ops.fmha(q, k, v, out=o, flags=flags)
NCL = 512
buf = torch.zeros(1024 + NCL * 8, dtype=torch.int64, device="cuda")
lib.v3a_debug_fmha_pair_trace.argtypes = [C.c_void_p]
lib.v3a_debug_fmha_pair_trace(buf.data_ptr())
ops.fmha(q, k, v, out=o, flags=flags)
torch.cuda.synchronize()
lib.v3a_debug_fmha_pair_trace(None)
t = buf.cpu()[1024:].view(NCL, 8)
cl = [r.tolist() + [int(r[7]) >> 16, idx // 74] for idx, r in enumerate(t) if int(r[0])]
for r in cl:
    r[7] &= 0xffff
t0 = min(r[6] for r in cl)
print(f"flags {flags:#x}: {len(cl)} items, kernel span (first entry -> last store) {(max(r[5] for r in cl) - t0) / 1000:.1f} us")
names = ["item start->Q requested", "Q requested->first S", "first S->last P", "last P->PV done", "PV done->stored", "total"]
cols = [[r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[5] - r[0]] for r in cl]
for n, c in zip(names, zip(*cols)):
    print(f"  {n:18s} median {statistics.median(c) / 1000:7.2f} us   min {min(c) / 1000:7.2f}   max {max(c) / 1000:7.2f}")
by_sm = {}
for r in cl:
    by_sm.setdefault(r[7], []).append(r)
gaps = []
for sm, rs in by_sm.items():
    rs.sort(key=lambda r: r[0])
    gaps += [b[2] - a[4] for a, b in zip(rs, rs[1:])]
if gaps:
    print(f"  tensor pipe idle between items on the same SM (last P V done -> next item's first S): median {statistics.median(gaps) / 1000:.2f} us  min {min(gaps) / 1000:.2f}  max {max(gaps) / 1000:.2f}  ({len(by_sm)} leader SMs)")
starts = sorted(r[0] - t0 for r in cl)
print("  item start times (us), every", len(by_sm), "th:", [round(x / 1000, 1) for x in starts[::len(by_sm)]])
# by position in the cluster's item list: duration of the key loop per key step
for pos in sorted(set(r[9] for r in cl)):
    rs = [r for r in cl if r[9] == pos]
    per = [(r[3] - r[2]) / 1000 / max(r[8] - 0.5, 0.5) for r in rs]
    print(f"  item #{pos}: {len(rs)} items, key steps {min(r[8] for r in rs)}..{max(r[8] for r in rs)}, first S -> last P per step: median {statistics.median(per):.2f} us"
          f"  (min {min(per):.2f}, max {max(per):.2f});  Q requested -> output stored: median {statistics.median((r[5] - r[1]) / 1000 for r in rs):.1f} us;"
          f"  output stored at median {statistics.median((r[5] - t0) / 1000 for r in rs):.1f} us, max {max((r[5] - t0) / 1000 for r in rs):.1f}")
