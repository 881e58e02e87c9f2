"""Times WanVAEDecoderB200.decode at the BASELINE clip (latent [1,16,4,64,64] -> 13 x 512 x 512) with a per-kernel-class breakdown.
Run on the GPU box:  python tools/vae_bench.py [--iters 5] [--detail]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vist3a_b200 import ops  # noqa: E402
from vist3a_b200.wan_vae import WanVAEDecoderB200, random_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--detail", action="store_true")
    ap.add_argument("--ncu", action="store_true", help="one warm decode between cudaProfilerStart/Stop (ncu --profile-from-start off)")
    a = ap.parse_args()
    m = WanVAEDecoderB200.from_state_dict(random_state_dict(), None, "cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    z = torch.randn(1, 16, 4, 64, 64, device="cuda", generator=g)
    for _ in range(2):
        out = m.decode(z, return_dict=False)[0]
    torch.cuda.synchronize()
    if a.ncu:
        torch.cuda.profiler.start()
        m.decode(z, return_dict=False)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.iters):
        out = m.decode(z, return_dict=False)[0]
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / a.iters
    print(f"VAE decode 13 x 512 x 512: {ms:.2f} ms  ({29.5 / ms * 1e3:.0f} TFLOP/s of the 29.5 TFLOP algorithmic), out {tuple(out.shape)}, "
          f"peak mem {torch.cuda.max_memory_allocated() / 1e9:.1f} GB")
    with ops.OpTimer() as t:
        m.decode(z, return_dict=False)
    agg = t.summary(detail=a.detail)
    tot = sum(d["ms"] for d in agg.values())
    print(f"# sum of per-call device times {tot:.2f} ms")
    for k, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:40]:
        tf = d["flops"] / d["ms"] / 1e9 if d["ms"] > 0 else 0
        print(f"{k:70s} {d['launches']:4d} launches {d['ms']:8.3f} ms {tf:8.1f} TFLOP/s {d['bytes'] / d['ms'] / 1e6 if d['ms'] > 0 else 0:9.1f} GB/s")


if __name__ == "__main__":
    main()
