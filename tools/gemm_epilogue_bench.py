"""Cost of the GEMM epilogue variants on the decoder / DiT shapes that carry a gate + residual: the same GEMM with (a) bf16 output, (b) fp32
output, (c) + bias, (d) + fp32 residual, (e) + gate, CUDA events, L2 flushed between launches.  GPU box."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vist3a_b200 import ops  # noqa: E402


def timeit(fn, iters=15):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


for name, M, N, K in (("dec_proj", 13377, 1024, 1024), ("dec_fc2", 13377, 1024, 4096), ("dit_o", 8192, 1536, 1536), ("dit_ffn2", 8192, 1536, 8960), ("dit_qkv", 8192, 4608, 1536), ("dit_ffn1", 8192, 8960, 1536)):
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = torch.randn(N, K, device="cuda").bfloat16() * 0.05
    bias = torch.randn(N, device="cuda")
    gate = torch.randn(N, device="cuda")
    res32 = torch.randn(M, N, device="cuda")
    res16 = res32.bfloat16()
    o32 = torch.empty(M, N, device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    rows = {}
    for tag, fn in (("bf16", lambda: ops.gemm(a, w, out=o16)),
                    ("bf16+bias+res16", lambda: ops.gemm(a, w, bias, residual=res16, out=o16)),
                    ("fp32", lambda: ops.gemm(a, w, out=o32)),
                    ("fp32+bias", lambda: ops.gemm(a, w, bias, out=o32)),
                    ("fp32+bias+res", lambda: ops.gemm(a, w, bias, residual=res32, out=o32)),
                    ("fp32+bias+gate+res", lambda: ops.gemm(a, w, bias, gate=gate, residual=res32, out=o32)),
                    ("fp32+bias+gate+res_inplace", lambda: ops.gemm(a, w, bias, gate=gate, residual=o32, out=o32)),
                    ("STAGED bf16", lambda: ops.gemm(a, w, out=o16, staged=True)),
                    ("STAGED bf16+bias+res16", lambda: ops.gemm(a, w, bias, residual=res16, out=o16, staged=True)),
                    ("STAGED fp32", lambda: ops.gemm(a, w, out=o32, staged=True)),
                    ("STAGED fp32+bias+gate+res", lambda: ops.gemm(a, w, bias, gate=gate, residual=res32, out=o32, staged=True)),
                    ("bf16 (again)", lambda: ops.gemm(a, w, out=o16))):
        ms = timeit(fn)
        rows[tag] = {"us": round(ms * 1e3, 1), "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}
    print(json.dumps({"name": name, "M": M, "N": N, "K": K, **rows}), flush=True)
