"""Times the full-size stitched decoder (13 views x 448x448 unless --views/--hw say otherwise) on cuda:0 and prints a
per-kernel-class breakdown (CUDA events around every C-ABI call; random-init weights, synthetic inputs).

    python tools/decoder_profile.py [--latent-frames 4] [--detail] [--iters 3]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vist3a_b200 import ops  # noqa: E402
from vist3a_b200.stitched_decoder import DecoderConfig, StitchVAE3DB200, random_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--latent-frames", type=int, default=4)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--detail", action="store_true")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--ncu", action="store_true", help="one warm forward between cudaProfilerStart/Stop (ncu --profile-from-start off)")
    ap.add_argument("--ncu-all", action="store_true", help="two forwards and exit (ncu -k filters the kernel)")
    ap.add_argument("--voxelize", action="store_true", help="voxelised-fusion branch (voxel_size 0.002)")
    ap.add_argument("--images", action="store_true", help="un-stitched image -> 3DGS path (patch embedding + 24 DINO blocks) instead of the stitched latent path")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    cfg = DecoderConfig(voxelize=a.voxelize, patch_embed=a.images, dino_blocks=24 if a.images else 22)
    sd = random_state_dict(cfg, 0, "cuda")
    m = StitchVAE3DB200.from_state_dict(sd, cfg, "cuda")
    del sd
    V = (a.latent_frames - 1) * 4 + 1
    g = torch.Generator(device="cuda").manual_seed(1)
    lat = torch.randn(a.batch, 16, a.latent_frames, 64, 64, device="cuda", generator=g)
    img = torch.rand(a.batch, 3, V, 448, 448, device="cuda", generator=g) * 2 - 1
    fwd = (lambda: m.forward_images(((img + 1) / 2).permute(0, 2, 1, 3, 4).contiguous())) if a.images else (lambda: m.forward_with_latent(lat, img))
    if a.images:
        img01 = ((img + 1) / 2).permute(0, 2, 1, 3, 4).contiguous()
        fwd = lambda: m.forward_images(img01)  # noqa: E731
    for _ in range(2):
        o = fwd()
    torch.cuda.synchronize()
    if a.ncu_all:
        return
    if a.ncu:
        torch.cuda.profiler.start()
        o = fwd()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.iters):
        o = fwd()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / a.iters
    N = a.batch * V * 448 * 448
    if a.voxelize:
        print(f"voxels: {o.gaussians.means.shape[1]} of {V * 448 * 448} pixels (ratio {o.infos['voxelize_ratio']:.4f})")
    print(f"decoder {V} views: {ms:.2f} ms/forward, {N / ms * 1e3 / 1e6:.1f} M Gaussians/s, peak mem {torch.cuda.max_memory_allocated() / 1e9:.1f} GB")
    with ops.OpTimer() as t:
        fwd()
    agg = t.summary(detail=a.detail)
    tot = sum(d["ms"] for d in agg.values())
    print(f"# sum of per-call device times {tot:.2f} ms")
    for k, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:60]:
        tf = d["flops"] / d["ms"] / 1e9 if d["ms"] > 0 else 0
        gb = d["bytes"] / d["ms"] / 1e6 if d["ms"] > 0 else 0
        print(f"# {k:64s} {d['launches']:5d} launches {d['ms']:9.3f} ms {tf:9.1f} TFLOP/s {gb:9.1f} GB/s")
    print(json.dumps({"decoder_ms": ms, "views": V, "gaussians_per_s": N / ms * 1e3}))


if __name__ == "__main__":
    main()
