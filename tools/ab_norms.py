"""A/B of two builds of the library on the norm kernels (one process per build, alternated by the calling script):
    VIST3A_AB_LIB=tools/ab/lib_A_old.so python tools/ab_norms.py"""
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


def timeit(fn, iters=30):
    """device time per call inside a CUDA graph of `iters` calls (eager timing of these 10-20 us kernels measures the host's launch
    overhead: python + ctypes cost about as much per call as the kernel runs)"""
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fn()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(iters):
                fn()
    torch.cuda.synchronize()
    gr.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        gr.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (5 * iters) * 1e3


g = torch.Generator(device="cuda").manual_seed(0)
M, D = 8192, 1536
x = torch.randn(M, D, device="cuda", generator=g).bfloat16()
mod = torch.randn(2, 6, D, device="cuda", generator=g)
h = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
res = {}
res["ln_dit_mod_us"] = timeit(lambda: ops.layernorm(x, mul=mod[:, 1], add=mod[:, 0], mul_bstride=6 * D, add_bstride=6 * D, rows_per_batch=4096, eps=1e-6, out=h))
qkv = torch.randn(M, 3 * D, device="cuda", generator=g).bfloat16()
wqk = torch.randn(2 * D, device="cuda", generator=g)
ang = torch.rand(4096, 64, device="cuda", generator=g)
cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
res["rmsnorm_rope_us"] = timeit(lambda: ops.rmsnorm_rope_(qkv[:, :2 * D], wqk, 128, eps=1e-6, cos=cos, sin=sin, nseg=2))
Md, C = 13377, 1024
xf = torch.randn(Md, C, device="cuda", generator=g)
lw, lb = torch.randn(C, device="cuda", generator=g), torch.randn(C, device="cuda", generator=g)
hd = torch.empty(Md, C, device="cuda", dtype=torch.bfloat16)
res["ln_dec_f32_us"] = timeit(lambda: ops.layernorm(xf, mul=lw, add=lb, eps=1e-5, out=hd))
res["torch_copy_bf16_8192x1536_us"] = timeit(lambda: h.copy_(x))
res["torch_copy_f32_to_bf16_13377x1024_us"] = timeit(lambda: hd.copy_(xf))
print(json.dumps({"lib": os.environ.get("VIST3A_AB_LIB", "default"), **{k: round(v, 2) for k, v in res.items()}}))
