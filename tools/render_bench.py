"""Rasteriser timing at the BASELINE size: the 2.6 M Gaussians a 13-view decoder forward produces (or synthetic ones with --synthetic),
rendered into 448x448 views.  Run on the GPU box.  Prints per-view time, intersections and the kernel-class breakdown."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vist3a_b200 import _lib, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--voxelize", action="store_true")
    a = ap.parse_args()
    from vist3a_b200.stitched_decoder import DecoderConfig, StitchVAE3DB200, random_state_dict

    cfg = DecoderConfig(voxelize=a.voxelize)
    m = StitchVAE3DB200.from_state_dict(random_state_dict(cfg, 0, "cuda"), cfg, "cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    lat = torch.randn(1, 16, 4, 64, 64, device="cuda", generator=g)
    img = torch.rand(1, 3, 13, 448, 448, device="cuda", generator=g) * 2 - 1
    o = m.forward_with_latent(lat, img)
    gs = o.gaussians
    n = gs.means.shape[1]
    c2w = o.pred_context_pose["extrinsic"][0].cpu()
    Kn = o.pred_context_pose["intrinsic"][0].cpu()
    del m
    torch.cuda.empty_cache()
    W = H = 448
    means, cov, opac, harm = (t[0].float().contiguous() for t in (gs.means, gs.covariances, gs.opacities, gs.harmonics))
    res = []
    for v in range(min(a.views, c2w.shape[0])):
        w2c = torch.linalg.inv(c2w[v])
        K = Kn[v].clone()
        K[0] *= W
        K[1] *= H
        for _ in range(2):
            r = ops.gs_render(means, cov, opac, harm, w2c, K, W, H, background=(1.0, 1.0, 1.0))
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = ops.gs_render(means, cov, opac, harm, w2c, K, W, H, background=(1.0, 1.0, 1.0))
        e.record()
        torch.cuda.synchronize()
        res.append(dict(view=v, ms=round(s.elapsed_time(e), 3), n_isect=r["n_isect"], launches=_lib.launch_count() - n0, alpha_mean=round(float(r["alpha"].mean()), 4)))
        print(json.dumps(res[-1]), flush=True)
    ms = sorted(x["ms"] for x in res)[len(res) // 2]
    # algorithmic bytes of the projection pass (the HBM-bound part): means 12 + covariances 36 + opacity 4 + SH 300 read, 40 + 8 + 4 written per Gaussian
    print(json.dumps({"gaussians": n, "median_ms_per_view": ms, "views_per_s": round(1e3 / ms, 1), "projection_algorithmic_GB": round(n * 404 / 1e9, 3)}))


if __name__ == "__main__":
    main()
