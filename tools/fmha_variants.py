"""Correctness + timing of the attention kernel variants selected by vist3a_fmha_args.flags (bit0: ONE thread per query row (default two),
bit1: 128-key aliased steps at d=128).  Run on the GPU box."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vist3a_b200 import ops  # noqa: E402


def timeit(fn, iters=10):
    """device time per call inside a CUDA graph of `iters` calls: an eager  record / call / record  with an idle GPU counts the host's launch path
    (python + ctypes + three tensor-map encodes, 15-20 us) into every sample -- 30 % of a 46 us cross-attention kernel"""
    for _ in range(3):
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
    for _ in range(3):
        gr.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (3 * iters)


def main():
    flag_list = [int(x, 0) for x in sys.argv[1:]] or [0, 1, 2, 3, 4, 5, 6]
    res = {fl: {"flags": fl, "rel_l2": []} for fl in flag_list}
    for flags in flag_list:
        for (B, H, Lq, Lk, D) in ((1, 2, 300, 333, 128), (2, 3, 1029, 1029, 64), (1, 2, 257, 129, 128), (1, 4, 640, 1100, 64)):
            if D == 64 and (flags & 2) or D == 64 and (flags & 256):
                continue
            g = torch.Generator(device="cuda").manual_seed(B + Lq + D + flags)
            q = torch.randn(B, Lq, H, D, device="cuda", generator=g).bfloat16()
            k = torch.randn(B, Lk, H, D, device="cuda", generator=g).bfloat16()
            v = torch.randn(B, Lk, H, D, device="cuda", generator=g).bfloat16()
            q[:, : Lq // 2] *= 4  # large logits on half the rows: exercises the lazy rescale
            o = ops.fmha(q, k, v, flags=flags)
            ref = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2).float(), k.transpose(1, 2).float(), v.transpose(1, 2).float()).transpose(1, 2)
            res[flags]["rel_l2"].append(round(float((o.float() - ref).norm() / ref.norm()), 5))
    # timing: shapes outermost, the variants interleaved and measured in TWO rounds (clocks ramp during a process: whichever variant is
    # measured first looks slower on the short kernels -- the first version of this tool compared variants measured minutes apart)
    for name, B, H, Lq, Lk, D in (("dit_self", 2, 12, 4096, 4096, 128), ("dit_cross", 2, 12, 4096, 512, 128), ("dec_frame", 13, 16, 1029, 1029, 64),
                                  ("dec_global", 1, 16, 13377, 13377, 64)):
        q = torch.randn(B, Lq, H, D, device="cuda").bfloat16()
        k = torch.randn(B, Lk, H, D, device="cuda").bfloat16()
        v = torch.randn(B, Lk, H, D, device="cuda").bfloat16()
        o = torch.empty_like(q)
        for rnd in range(2):
            for flags in flag_list:
                if D == 64 and (flags & 2) or D == 64 and (flags & 256):
                    continue
                ms = timeit(lambda: ops.fmha(q, k, v, out=o, flags=flags))
                res[flags].setdefault(name, []).append(round(4 * B * H * Lq * Lk * D / ms / 1e9, 1))
    for fl in flag_list:
        print(json.dumps(res[fl]), flush=True)


if __name__ == "__main__":
    main()
