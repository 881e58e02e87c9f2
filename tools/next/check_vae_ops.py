"""DRAFT checker (see tools/next/README.md): builds tools/next/vae_ops.cu and compares each kernel on cuda:0 with its per-operation
reference in oracle/wan_vae_plan.py (which composes to the pinned VAE oracle).  Not collected by pytest; run it by hand:

    gpurun --timeout 300 -- 'python tools/next/check_vae_ops.py'
"""
import ctypes as C
import os
import subprocess
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import wan_vae_plan as PL  # noqa: E402


def build():
    out = os.path.join(HERE, "_build", "libvae_next.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(HERE, "vae_ops.cu")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                        "-shared", "-o", out, src], check=True)
    lib = C.CDLL(out)
    vp, ll, i32, f32 = C.c_void_p, C.c_longlong, C.c_int, C.c_float
    lib.v3a_next_rmsnorm_nhwc.argtypes = [vp, vp, vp, ll, i32, i32]
    lib.v3a_next_softmax_rows.argtypes = [vp, vp, ll, i32, f32]
    lib.v3a_next_time_interleave.argtypes = [vp, vp, ll, ll, i32]
    lib.v3a_next_transpose_bf16.argtypes = [vp, vp, ll, i32]
    return lib


def main():
    lib = build()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    ok = True

    def report(name, got, want, tol):
        nonlocal ok
        err = float((got.float().cpu() - want.float()).abs().max())
        good = err <= tol
        ok &= good
        print(f"{name:40s} max |err| {err:.3e}  (tol {tol:.1e})  {'ok' if good else 'FAIL'}")

    # RMS-norm (+ SiLU) over channels, the widths of the VAE
    for Cc in (96, 192, 384, 16):
        for silu in (1, 0):
            x = (torch.randn(1000, Cc, generator=g) * 2).bfloat16()
            gamma = 1 + 0.1 * torch.randn(Cc, generator=g)
            want = PL.rmsnorm_silu(x.float(), gamma, silu=bool(silu), rb=True)
            xd, gd = x.to(dev), gamma.to(dev)
            yd = torch.empty_like(xd)
            assert lib.v3a_next_rmsnorm_nhwc(xd.data_ptr(), gd.data_ptr(), yd.data_ptr(), x.shape[0], Cc, silu) == 0
            torch.cuda.synchronize()
            report(f"rmsnorm C={Cc} silu={silu}", yd, want, 2e-2)   # bf16 output: one ulp at |y| < 4 is 1.6e-2

    # row softmax (logits fp32 -> probabilities bf16), the mid-block attention shapes (L = 4096 at 512 x 512) and a short row
    for rows, L in ((64, 4096), (7, 36), (3, 1024)):
        s = torch.randn(rows, L, generator=g) * 3
        scale = 384 ** -0.5
        want = torch.softmax(s * scale, dim=-1)
        sd = s.to(dev)
        pd = torch.empty(rows, L, dtype=torch.bfloat16, device=dev)
        assert lib.v3a_next_softmax_rows(sd.data_ptr(), pd.data_ptr(), rows, L, scale) == 0
        torch.cuda.synchronize()
        report(f"softmax rows={rows} L={L}", pd, want, 4e-3 * float(want.max()) + 1e-6)
        report(f"softmax row sums rows={rows} L={L}", pd.float().sum(-1), torch.ones(rows), 2e-2)

    # time interleave of the temporal up-sampling conv
    T, P, Cc = 3, 50, 192
    y = torch.randn(T, P, 2 * Cc, generator=g).bfloat16()
    want = y.reshape(T, P, 2, Cc).permute(0, 2, 1, 3).reshape(2 * T, P, Cc)          # oracle/wan_vae_plan.py:upsample
    yd = y.to(dev)
    od = torch.empty(2 * T, P, Cc, dtype=torch.bfloat16, device=dev)
    assert lib.v3a_next_time_interleave(yd.data_ptr(), od.data_ptr(), T, P, Cc) == 0
    torch.cuda.synchronize()
    report("time interleave", od, want, 0.0)

    # transposes
    for R, Cc in ((4096, 384), (1000, 3), (77, 16)):
        a = torch.randn(R, Cc, generator=g).bfloat16()
        ad = a.to(dev)
        od = torch.empty(Cc, R, dtype=torch.bfloat16, device=dev)
        assert lib.v3a_next_transpose_bf16(ad.data_ptr(), od.data_ptr(), R, Cc) == 0
        torch.cuda.synchronize()
        report(f"transpose {R}x{Cc}", od, a.t(), 0.0)
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
