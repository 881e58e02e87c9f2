"""Fixed cost vs per-step cost of the attention kernel: times d=128 attention over a sweep of key lengths and CTA counts (GPU box)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vist3a_b200 import ops  # noqa: E402
from tools.fmha_variants import timeit  # noqa: E402


def main():
    flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    D = 128
    for (B, H, Lq, Lk) in [(2, 12, 4096, 64), (2, 12, 4096, 128), (2, 12, 4096, 256), (2, 12, 4096, 512), (2, 12, 4096, 1024), (2, 12, 4096, 2048),
                           (2, 12, 4096, 4096), (1, 37, 1024, 4096), (1, 37, 2048, 4096), (1, 37, 4096, 4096), (1, 37, 4096, 512), (1, 37, 8192, 8192)]:
        q = torch.randn(B, Lq, H, D, device="cuda").bfloat16()
        k = torch.randn(B, Lk, H, D, device="cuda").bfloat16()
        v = torch.randn(B, Lk, H, D, device="cuda").bfloat16()
        o = torch.empty_like(q)
        ms = timeit(lambda: ops.fmha(q, k, v, out=o, flags=flags))
        ctas = B * H * ((Lq + 255) // 256)
        print(json.dumps({"B": B, "H": H, "Lq": Lq, "Lk": Lk, "ctas": ctas, "waves": round(ctas / 148, 2), "us": round(ms * 1e3, 1),
                          "tflops": round(4 * B * H * Lq * Lk * D / ms / 1e9, 1)}), flush=True)


if __name__ == "__main__":
    main()
