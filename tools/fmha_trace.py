"""Per-CTA timeline of the attention kernel (clock64 stamps written by the kernel when v3a_debug_fmha_trace is armed).  GPU box."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vist3a_b200 import _lib, ops  # noqa: E402

NAMES = ["entry", "setup", "pdl", "tma_q_k0", "q_full", "qk0_issued", "s0_ready", "p0_done", "loop_end", "pv_last", "epi_done", "exit_sync"]


def main():
    B, H, Lq, Lk, D = [int(x) for x in sys.argv[1:6]] if len(sys.argv) >= 6 else (2, 12, 4096, 512, 128)
    flags = int(sys.argv[6]) if len(sys.argv) >= 7 else 0
    lib = _lib.load()
    q = torch.randn(B, Lq, H, D, device="cuda").bfloat16()
    k = torch.randn(B, Lk, H, D, device="cuda").bfloat16()
    v = torch.randn(B, Lk, H, D, device="cuda").bfloat16()
    o = torch.empty_like(q)
    for _ in range(3):
        ops.fmha(q, k, v, out=o, flags=flags)
    rows_per_cta = 128 if flags & 65536 else 256
    ctas = B * H * ((Lq + rows_per_cta - 1) // rows_per_cta)
    tr = torch.zeros(ctas, 32, dtype=torch.int64, device="cuda")
    lib.v3a_debug_fmha_trace.argtypes = [ctypes.c_void_p]
    lib.v3a_debug_fmha_trace(tr.data_ptr())
    ops.fmha(q, k, v, out=o, flags=flags)
    torch.cuda.synchronize()
    lib.v3a_debug_fmha_trace(None)
    t = tr.cpu()
    rel = (t[:, :12] - t[:, :1]).double()
    print(f"B{B} H{H} Lq{Lq} Lk{Lk} D{D} flags {flags}: {ctas} CTAs; cycles since CTA entry (median / p90 over CTAs)")
    for i, n in enumerate(NAMES):
        col = rel[:, i]
        print(f"  {n:12s} {col.median().item():10.0f} {col.quantile(0.9).item():10.0f}")
    steps = (Lk + 63) // 64 if (D == 128 and not (flags & (2 | 16384 | 65536))) else (Lk + 127) // 128
    for i, n in ((16, "softmax: wait S"), (17, "softmax: tmem ld"), (18, "softmax: max+xchg+rescale"), (19, "softmax: exp+pack+st issue"),
                 (20, "softmax: st wait+arrive"), (21, "mma warp: wait P"), (22, "mma warp: issue PV+QK")):
        col = t[:, i].double() / steps
        print(f"  per step  {n:28s} {col.median().item():8.0f} {col.quantile(0.9).item():8.0f}")
    # per-SM: gap between the exit of one CTA and the entry of the next on the same SM
    sm = t[:, 12]
    gaps = []
    for s_ in sm.unique():
        rows = t[sm == s_]
        rows = rows[rows[:, 0].argsort()]
        for a, b in zip(rows[:-1], rows[1:]):
            gaps.append(int(b[0] - a[11]))
    if gaps:
        g = torch.tensor(gaps).double()
        print(f"  gap exit_sync -> next CTA entry on the same SM: median {g.median().item():.0f}, p90 {g.quantile(0.9).item():.0f} cycles ({len(gaps)} pairs)")


if __name__ == "__main__":
    main()
