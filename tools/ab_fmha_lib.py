"""Two builds of the library, one process each (VIST3A_AB_LIB=path): CUDA-graph device time of the attention call at a few head_dim-128 shapes.
    for r in 1 2 3; do for l in tools/ab/lib_A.so tools/ab/lib_B.so; do VIST3A_AB_LIB=$l python tools/ab_fmha_lib.py; done; done"""
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vist3a_b200 import _lib  # noqa: E402

if os.environ.get("VIST3A_AB_LIB"):
    _lib.LIB_PATH = Path(os.environ["VIST3A_AB_LIB"]).resolve()
from vist3a_b200 import ops  # noqa: E402
from tools.fmha_variants import timeit  # noqa: E402

res = {"lib": os.path.basename(os.environ.get("VIST3A_AB_LIB", "default"))}
for name, B, H, Lq, Lk in (("dit_self_1.3b", 2, 12, 4096, 4096), ("dit_self_14b", 2, 40, 4096, 4096), ("dit_self_14b_21v", 2, 40, 6144, 6144), ("b1", 1, 12, 4096, 4096),
                           ("small", 1, 12, 1024, 4096), ("dit_cross", 2, 12, 4096, 512)):
    q = torch.randn(B, Lq, H, 128, device="cuda").bfloat16()
    k = torch.randn(B, Lk, H, 128, device="cuda").bfloat16()
    v = torch.randn(B, Lk, H, 128, device="cuda").bfloat16()
    o = torch.empty_like(q)
    res[name] = [round(4 * B * H * Lq * Lk * 128 / timeit(lambda: ops.fmha(q, k, v, out=o)) / 1e9, 1) for _ in range(2)]
print(json.dumps(res), flush=True)
