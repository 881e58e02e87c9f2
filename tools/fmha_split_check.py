"""Work decomposition of the CTA-pair attention kernel (fmha_pair_sm100.cu): correctness against fp32 SDPA with and without the key split of the
tail, then device time (CUDA graph, variants interleaved) of: default | no key split (flags bit 17) | one cluster per unit, no split (bits 17 + 20)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vist3a_b200 import ops  # noqa: E402
from tools.fmha_variants import timeit  # noqa: E402

NOSPLIT = 1 << 17


def main():
    out = {"check": []}
    for (B, H, Lq, Lk) in ((1, 2, 600, 2048), (1, 15, 1024, 4096), (2, 3, 700, 2000), (1, 12, 4096, 4096), (2, 12, 4096, 4096)):
        g = torch.Generator(device="cuda").manual_seed(Lq + Lk)
        q = torch.randn(B, Lq, H, 128, device="cuda", generator=g).bfloat16()
        k = torch.randn(B, Lk, H, 128, device="cuda", generator=g).bfloat16()
        v = torch.randn(B, Lk, H, 128, device="cuda", generator=g).bfloat16()
        q[:, : Lq // 2] *= 3
        k[:, Lk - 100:] *= 2   # the largest logits sit in the LAST key chunk
        rs = (torch.rand(B * Lq, device="cuda", generator=g) + 0.5).float()
        ref = torch.nn.functional.scaled_dot_product_attention((q.float() * rs.view(B, Lq, 1, 1)).transpose(1, 2), k.transpose(1, 2).float(),
                                                               v.transpose(1, 2).float()).transpose(1, 2)
        rec = {"shape": [B, H, Lq, Lk]}
        for name, fl in (("split", 0), ("nosplit", NOSPLIT)):
            o = ops.fmha(q, k, v, flags=fl, q_row_scale=rs)
            rec[name] = round(float((o.float() - ref).norm() / ref.norm()), 5)
        out["check"].append(rec)
        print(json.dumps(rec), flush=True)
    if "--no-time" in sys.argv:
        return
    # persistent clusters + balanced key split of the tail (0) | no key split (bit 17) | neither (bit 20): the round-2 kernel's decomposition
    variants = (("persist+split", 0), ("persist", NOSPLIT), ("neither", NOSPLIT | (1 << 20)), ("persist+split, merge kernel skipped", 1 << 22))
    for name, B, H, Lq, Lk in (("dit_self_1.3b", 2, 12, 4096, 4096), ("dit_cross", 2, 12, 4096, 512), ("dit_self_14b", 2, 40, 4096, 4096), ("dit_self_14b_21v", 2, 40, 6144, 6144),
                               ("b1", 1, 12, 4096, 4096), ("small", 1, 12, 1024, 4096)):
        q = torch.randn(B, Lq, H, 128, device="cuda").bfloat16()
        k = torch.randn(B, Lk, H, 128, device="cuda").bfloat16()
        v = torch.randn(B, Lk, H, 128, device="cuda").bfloat16()
        o = torch.empty_like(q)
        rec = {"shape": name}
        for rnd in range(4):
            for vn, fl in variants:
                ms = timeit(lambda: ops.fmha(q, k, v, out=o, flags=fl))
                rec.setdefault(vn, []).append(round(4 * B * H * Lq * Lk * 128 / ms / 1e9, 1))
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
