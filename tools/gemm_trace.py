"""Where a GEMM CTA's time goes (clock64 counters written by the kernel when v3a_debug_gemm_trace is armed): producer waiting for a free
stage, MMA warp waiting for operands / for a free accumulator stage, epilogue waiting for an accumulator.  GPU box."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vist3a_b200 import _lib, ops  # noqa: E402


def main():
    lib = _lib.load()
    lib.v3a_debug_gemm_trace.argtypes = [ctypes.c_void_p]
    shapes = [("dit_qkv", 8192, 4608, 1536, None), ("dit_o", 8192, 1536, 1536, None), ("dit_ffn1", 8192, 8960, 1536, "gelu_tanh"),
              ("dit_ffn2", 8192, 1536, 8960, None), ("dec_proj_f32res", 13377, 1024, 1024, "res"), ("dec_fc1", 13377, 4096, 1024, "gelu"),
              ("dec_proj_plain", 13377, 1024, 1024, None), ("dec_proj_f32", 13377, 1024, 1024, "f32"), ("dit_o_f32res", 8192, 1536, 1536, "res"),
              ("dit_o_f32", 8192, 1536, 1536, "f32")]
    for name, M, N, K, ep in shapes:
        x = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
        bias = torch.randn(N, device="cuda")
        kw = {}
        if ep == "res":
            res = torch.randn(M, N, device="cuda")
            kw = dict(residual=res, out=res, gate=torch.ones(N, device="cuda"), round_linear=True)
        elif ep == "f32":
            kw = dict(out=torch.empty(M, N, device="cuda"))
        elif ep:
            kw = dict(act=ep)
        for two in (True,):
            for _ in range(3):
                ops.gemm(x, w, bias, two_cta=two, **kw)
            tr = torch.zeros(148, 8, dtype=torch.int64, device="cuda")
            lib.v3a_debug_gemm_trace(tr.data_ptr())
            ops.gemm(x, w, bias, two_cta=two, **kw)
            torch.cuda.synchronize()
            lib.v3a_debug_gemm_trace(None)
            t = tr.cpu().double()
            t = t[t[:, 5] > 0]
            lead = t[t[:, 2] > 0]
            med = lambda c, tt=t: tt[:, c].median().item()  # noqa: E731
            print(f"{name:16s} {'2cta' if two else '1cta'} M{M} N{N} K{K}: producer total {med(0):9.0f} wait-empty {100 * med(1) / med(0):5.1f}% | "
                  f"mma total {med(2, lead):9.0f} wait-operands {100 * med(3, lead) / med(2, lead):5.1f}% wait-accum-stage {100 * med(4, lead) / med(2, lead):5.1f}% | "
                  f"epilogue total {med(5):9.0f} wait-accum {100 * med(6) / med(5):5.1f}%")


if __name__ == "__main__":
    main()
