"""Correctness + timing of the CTA-pair attention kernel (fmha_pair_sm100.cu, flags bit 8) against fp32 SDPA and the single-CTA kernel.
Run on the GPU box:  python tools/fmha_pair_check.py [--variants 0,1,2,3,4] [--iters 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vist3a_b200 import ops  # noqa: E402


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="0,1,2,3,4,5,6,7")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-time", action="store_true")
    a = ap.parse_args()
    variants = [int(v) for v in a.variants.split(",")]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ok_all = True
    shapes = [(1, 2, 128, 128), (1, 2, 256, 384), (2, 3, 300, 77), (1, 2, 129, 1000), (2, 12, 4096, 4096), (2, 12, 4096, 512), (1, 40, 6144, 6144)]
    for B, H, Lq, Lk in shapes:
        g = torch.Generator(device="cuda").manual_seed(Lq * 7 + Lk)
        qkv = torch.randn(B, max(Lq, Lk), 3, H, 128, device="cuda", generator=g).bfloat16()
        q, k, v = qkv[:, :Lq, 0], qkv[:, :Lk, 1], qkv[:, :Lk, 2]
        rs = torch.rand(B * Lq, device="cuda", generator=g) + 0.5
        ref = torch.nn.functional.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2)).transpose(1, 2)
        qs = (q.float() * rs.view(B, Lq, 1, 1))
        ref_rs = torch.nn.functional.scaled_dot_product_attention(qs.transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2)).transpose(1, 2)
        base = ops.fmha(q, k, v)
        for var in variants:
            fl = 256 | (var << 9)
            out = ops.fmha(q, k, v, flags=fl)
            out_rs = ops.fmha(q, k, v, flags=fl, q_row_scale=rs)
            torch.cuda.synchronize()
            err = float((out.float() - ref).abs().max())
            rel = float((out.float() - ref).norm() / ref.norm())
            err_rs = float((out_rs.float() - ref_rs).abs().max())
            rel_base = float((base.float() - ref).norm() / ref.norm())
            ok = err < 2e-2 and rel < 8e-3 and err_rs < 3e-2 and bool(torch.isfinite(out).all())
            ok_all &= ok
            r = {"shape": [B, H, Lq, Lk], "variant": var, "max_abs": err, "rel_l2": rel, "rel_l2_single_cta": rel_base, "max_abs_row_scale": err_rs, "ok": ok}
            if not a.no_time and Lq >= 4096:
                flops = 4.0 * B * H * Lq * Lk * 128
                o = torch.empty_like(q)
                ms = timeit(lambda: ops.fmha(q, k, v, out=o, flags=fl), a.iters, flush)
                ms0 = timeit(lambda: ops.fmha(q, k, v, out=o), a.iters, flush)
                r.update(ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1), single_cta_ms=round(ms0, 4), single_cta_tflops=round(flops / ms0 / 1e9, 1))
            print(json.dumps(r), flush=True)
    print("ALL OK" if ok_all else "FAILED", flush=True)
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
