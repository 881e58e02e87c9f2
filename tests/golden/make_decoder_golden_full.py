"""Generates tests/golden/decoder_full_13v.pt: the REAL reference decoder (/root/reference models/stitched_model.py:
StitchVAE3D.forward_with_latent, imported through oracle/ref_loader.py) at the BASELINE size -- full widths (1024-dim tokens, 22 + 48 blocks,
DPT heads), 13 views x 448x448 = 2 609 152 Gaussians -- on seeded synthetic weights / inputs, fp32 on the CPU.  Build container only
(about two minutes and ~25 GB of RAM):

    python tests/golden/make_decoder_golden_full.py

The weights are not stored (oracle.decoder_ref.init_state_dict(FULL, seed) regenerates them bit-identically anywhere).  Kept: every
STRIDE-th Gaussian of every field, every 7th depth pixel, the camera outputs, and float64 checksums of the full tensors.
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import decoder_ref as D  # noqa: E402
from oracle import ref_loader as RL  # noqa: E402

WSEED, ISEED, STRIDE = 1, 3, 2039   # 2 609 152 / 2039 = 1280 Gaussians kept


def main():
    model = RL.load_reference(RL.FULL, resolution=512, seed=0)
    sd = D.init_state_dict(D.FULL, seed=WSEED)
    missing = [k for k in model.load_state_dict(sd, strict=False).missing_keys if not k.startswith("diffusion_vae")]
    assert not missing, missing
    lat, img = D.synthetic_inputs(D.FULL, views_latent=4, latent_hw=64, image_hw=448, seed=ISEED)
    t0 = time.time()
    with torch.no_grad():
        ref = RL.outputs_to_dict(model.forward_with_latent(lat, feedforward_image=img))
    print(f"reference forward: {time.time() - t0:.1f} s")
    keep = {}
    for k, v in ref.items():
        if k in ("means", "covariances", "harmonics", "opacities", "scales", "rotations"):
            keep["checksum_" + k] = v.double().sum(dim=1).float()
            keep["abs_checksum_" + k] = v.double().abs().sum(dim=1).float()
            v = v[:, ::STRIDE]
        if k == "depth":
            keep["checksum_depth"] = v.double().sum(dim=(2, 3, 4)).float()
            v = v[:, :, ::7, ::7]
        keep[k] = v.clone()
    out = {"weight_seed": WSEED, "input_seed": ISEED, "stride": STRIDE, "depth_stride": 7, "outputs": keep}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "decoder_full_13v.pt")
    torch.save(out, path)
    print({k: tuple(v.shape) for k, v in keep.items()})
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
