"""DRAFT checker (see tools/next/README.md) for tools/next/gemm_conv_temporal_taps.patch: after

    git apply tools/next/gemm_conv_temporal_taps.patch && python __graft_entry__.py

compares the causal 3x3x3 / (3,1,1) implicit-GEMM convolution (ops.gemm(conv=dict(..., kt=3))) on cuda:0 with the per-operation
reference oracle/wan_vae_plan.py:conv_gemm (which composes to the pinned VAE oracle), at the channel counts the conv mode accepts
(C_in % 64 == 0: the 384 / 192-channel layers).  Also re-checks a plain 2-D conv (kt = 1) for regressions."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import wan_vae_plan as PL  # noqa: E402
from vist3a_b200 import ops  # noqa: E402
from vist3a_b200.wan_vae_layout import conv3d_weight_to_taps  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    ok = True
    for (T, H, W, ci, co, taps) in ((4, 16, 16, 64, 64, (3, 3, 3)), (7, 24, 40, 192, 192, (3, 3, 3)), (3, 16, 32, 128, 256, (3, 1, 1)),
                                    (13, 32, 32, 384, 192, (3, 3, 3)), (5, 16, 16, 64, 128, (1, 3, 3))):
        kt, kh, kw = taps
        x = torch.randn(T, H, W, ci, generator=g).bfloat16()
        w5 = (torch.randn(co, ci, kt, kh, kw, generator=g) / (ci * kt * kh * kw) ** 0.5).bfloat16()
        b = 0.1 * torch.randn(co, generator=g)
        wt = conv3d_weight_to_taps(w5.float())
        want = PL.conv_gemm(x.float(), wt, b, taps)                                   # fp32 accumulation of bf16 operands
        got = ops.gemm(x.to(dev), wt.bfloat16().to(dev), bias=b.to(dev), conv=dict(kh=kh, kw=kw, pad=kh // 2, pad_x=kw // 2, kt=kt))
        torch.cuda.synchronize()
        err = float((got.float().cpu().reshape(want.shape) - want).abs().max())
        tol = 2e-2 * float(want.abs().max())
        good = err <= tol
        ok &= good
        print(f"conv T={T} {H}x{W} {ci}->{co} taps={taps}: max |err| {err:.3e} (tol {tol:.1e}) {'ok' if good else 'FAIL'}")
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
