"""Per-shape device timing of the GEMM / attention kernels at the DiT and decoder shapes, next to the
torch library kernels (cuBLAS / SDPA) as a same-box comparator.  Run on the GPU box."""
import argparse
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vist3a_b200 import ops  # noqa: E402


def timeit(fn, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
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
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-torch", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--fmha-flags", type=int, default=0)
    a = ap.parse_args()
    flush = None if a.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = []
    gemms = [("dit_qkv", 8192, 4608, 1536), ("dit_o", 8192, 1536, 1536), ("dit_ffn1", 8192, 8960, 1536),
             ("dit_ffn2", 8192, 1536, 8960), ("dit_qkv_b1", 4096, 4608, 1536), ("dec_qkv", 13377, 3072, 1024),
             ("dec_proj", 13377, 1024, 1024), ("dec_fc1", 13377, 4096, 1024), ("dec_fc2", 13377, 1024, 4096)]
    for name, M, N, K in gemms:
        if a.only and a.only not in name and a.only != "gemm":
            continue
        x = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
        bias = torch.randn(N, device="cuda")
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        r = {"name": name, "M": M, "N": N, "K": K}
        for tag, two, b176, mc in (("2cta", True, False, False), ("2cta_mc", True, False, True), ("2cta_bn176", True, True, False), ("1cta", False, False, False)):
            ms = timeit(lambda: ops.gemm(x, w, bias, out=out, two_cta=two, bn176=b176, multicast=mc), a.iters, flush=flush)
            r[tag + "_ms"] = round(ms, 4)
            r[tag + "_tflops"] = round(2 * M * N * K / ms / 1e9, 1)
        if not a.no_torch:
            ms = timeit(lambda: torch.nn.functional.linear(x, w, bias.bfloat16()), a.iters, flush=flush)
            r["cublas_ms"], r["cublas_tflops"] = round(ms, 4), round(2 * M * N * K / ms / 1e9, 1)
        print(json.dumps(r), flush=True)
        res.append(r)
    fm = [("dit_self", 2, 12, 4096, 4096, 128), ("dit_cross", 2, 12, 4096, 512, 128), ("dec_frame", 13, 16, 1029, 1029, 64),
          ("dec_global", 1, 16, 13377, 13377, 64)]
    for name, B, H, Lq, Lk, D in fm:
        if a.only and a.only not in name and a.only != "fmha":
            continue
        q = torch.randn(B, Lq, H, D, device="cuda").bfloat16()
        k = torch.randn(B, Lk, H, D, device="cuda").bfloat16()
        v = torch.randn(B, Lk, H, D, device="cuda").bfloat16()
        o = torch.empty_like(q)
        fl = 4 * B * H * Lq * Lk * D
        ms = timeit(lambda: ops.fmha(q, k, v, out=o, flags=a.fmha_flags), a.iters, flush=flush)
        r = {"name": name, "B": B, "H": H, "Lq": Lq, "Lk": Lk, "D": D, "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)}
        if not a.no_torch:
            qt, kt, vt = (t.transpose(1, 2) for t in (q, k, v))
            ms = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt), a.iters, flush=flush)
            r["sdpa_ms"], r["sdpa_tflops"] = round(ms, 4), round(fl / ms / 1e9, 1)
            try:
                from flash_attn import flash_attn_func
                ms = timeit(lambda: flash_attn_func(q, k, v), a.iters, flush=flush)
                r["fa2_ms"], r["fa2_tflops"] = round(ms, 4), round(fl / ms / 1e9, 1)
            except Exception as ex:  # noqa
                r["fa2"] = str(ex)[:60]
        print(json.dumps(r), flush=True)
        res.append(r)
    # voxelised fusion at the BASELINE size (13 views x 448^2 points, 83 features + confidence per row): HBM-bound index work;
    # algorithmic bytes = every point row read once + every voxel row written once
    if not a.only or a.only in ("voxel", "voxel_fusion"):
        n, c = 13 * 448 * 448, 83
        g = torch.Generator(device="cuda").manual_seed(3)
        for name, spread in (("voxel_sparse", 1.0), ("voxel_dense", 0.05)):   # ~1 point per voxel / many points per voxel
            pts = (torch.randn(n, 3, device="cuda", generator=g) * spread).contiguous()
            rows = torch.randn(n, c + 1, device="cuda", generator=g)
            o = ops.voxel_fusion(pts, rows, rows[:, c], 0.002, feat_dim=c)
            m = o["n_voxels"]
            ms = timeit(lambda: ops.voxel_fusion(pts, rows, rows[:, c], 0.002, feat_dim=c), a.iters, flush=flush)
            by = 4.0 * (n * (3 + c + 1) + m * (3 + c))
            r = {"name": name, "points": n, "voxels": m, "ms": round(ms, 4), "algorithmic_GB": round(by / 1e9, 3), "GBps": round(by / ms / 1e6, 1),
                 "note": "includes the host read of the voxel count (stream sync) and workspace allocation"}
            print(json.dumps(r), flush=True)
            res.append(r)


if __name__ == "__main__":
    main()
