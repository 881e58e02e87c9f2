"""Phases of the CUDA-core tail-row CTAs of the attention kernel (fmha_sm100.cu fmha_tail_rows) next to the duration of the tensor-core CTAs
of the same launch (clock64 stamps, v3a_debug_fmha_trace).  GPU box:  python tools/fmha_tail_trace.py [B H Lq Lk]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vist3a_b200 import _lib, ops  # noqa: E402

B, H, Lq, Lk = [int(x) for x in sys.argv[1:5]] if len(sys.argv) >= 5 else (13, 16, 1029, 1029)
lib = _lib.load()
q, k, v = (torch.randn(B, L, H, 64, device="cuda").bfloat16() for L in (Lq, Lk, Lk))
o = torch.empty_like(q)
for fl in (1 << 21, 0):
    for _ in range(3):
        ops.fmha(q, k, v, out=o, flags=fl)
    gx = (Lq - (Lq % 256 if fl else 0) + 255) // 256
    main = B * H * gx
    extra = -(-(B * H) // (gx * H)) * gx * H if fl else 0
    tr = torch.zeros(main + extra, 32, dtype=torch.int64, device="cuda")
    lib.v3a_debug_fmha_trace.argtypes = [ctypes.c_void_p]
    lib.v3a_debug_fmha_trace(tr.data_ptr())
    ops.fmha(q, k, v, out=o, flags=fl)
    torch.cuda.synchronize()
    lib.v3a_debug_fmha_trace(None)
    t = tr.cpu()
    m = t[:main]
    dur = (m[:, 11] - m[:, 0]).double()
    print(f"flags {fl:#x}: {main} tensor-core CTAs, entry -> exit median {dur.median().item():.0f} cycles, min {dur.min().item():.0f}, max {dur.max().item():.0f}")
    if not fl:
        last = dur.view(B * H, gx)[:, -1]
        print(f"   the last query block of each (batch, head) (one tile, {Lq % 256} rows): median {last.median().item():.0f} cycles")
    if extra:
        tl = t[main:main + B * H]
        names = ["entry -> K in smem", "scores", "softmax", "wait for V", "P V", "reduce + store"]
        for i, n in enumerate(names):
            c = (tl[:, i + 1] - tl[:, i]).double()
            print(f"   tail CTA  {n:20s} median {c.median().item():8.0f}  max {c.max().item():8.0f} cycles")
        c = (tl[:, 6] - tl[:, 0]).double()
        print(f"   tail CTA  total                median {c.median().item():8.0f}  max {c.max().item():8.0f} cycles")
