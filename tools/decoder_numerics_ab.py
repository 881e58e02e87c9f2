"""A/B of the decoder's numerics under two attention variants (VIST3A_FMHA_FLAGS): per-field rel-L2 against the fp32 oracle on the tiny golden cases."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import decoder_ref as D  # noqa: E402
import test_decoder_gpu as T  # noqa: E402

g = torch.load(T.GOLD)["cases"]
for case in ("v5_56", "v9_112_b2"):
    c = g[case]
    for seed in (c["weight_seed"], 101, 202):
        sd = D.init_state_dict(D.TINY, seed=seed)
        lat, img = D.synthetic_inputs(D.TINY, views_latent=c["latent_frames"], latent_hw=c["latent_hw"], image_hw=c["image_hw"], batch=c["batch"], seed=c["input_seed"])
        ref = D.decoder_forward(sd, D.TINY, lat, img, resolution=c["resolution"])
        m = T._engine(sd, D.TINY, c["resolution"])
        out = T._as_dict(m.forward_with_latent(lat.cuda(), img.cuda()))
        floor = T._autocast_floor(D, sd, D.TINY, lat, img, c["resolution"], ref)
        print(case, seed, os.environ.get("VIST3A_FMHA_FLAGS", "0"), {k: f"{T._rel(out[k], ref[k]):.2e}/{floor[k]:.2e}" for k in ("means", "scene_scale", "extrinsic", "intrinsic", "last_pred_pose_enc", "depth", "rotations")})
