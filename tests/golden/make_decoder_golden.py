"""Generates tests/golden/decoder_tiny.pt by running the REAL reference decoder
(/root/reference models/stitched_model.py:StitchVAE3D.forward_with_latent, imported through
oracle/ref_loader.py) on seeded synthetic weights/inputs.  Run in the build container only:

    python tests/golden/make_decoder_golden.py

The weights are NOT stored: oracle.decoder_ref.init_state_dict(TINY, seed) regenerates them
bit-identically from an integer stream on any machine; the fixture holds inputs' seeds and the
reference's outputs (Gaussians subsampled every `stride`-th and depth every 3rd pixel, plus full-tensor checksums, to keep the file small).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import decoder_ref as D  # noqa: E402
from oracle import ref_loader as RL  # noqa: E402

CASES = [
    # name, weight seed, input seed, latent frames, latent hw, image hw, resolution, batch
    ("v5_56", 3, 5, 2, 8, 56, 64, 1, 11),
    ("v9_112_b2", 4, 6, 3, 16, 112, 128, 2, 211),
]


def main():
    out = {"cases": {}}
    for name, wseed, iseed, T, lhw, ihw, res, B, STRIDE in CASES:
        model = RL.load_reference(RL.TINY, resolution=res, seed=0)
        sd = D.init_state_dict(D.TINY, seed=wseed)
        missing = [k for k in model.load_state_dict(sd, strict=False).missing_keys if not k.startswith("diffusion_vae")]
        assert not missing, missing
        lat, img = D.synthetic_inputs(D.TINY, views_latent=T, latent_hw=lhw, image_hw=ihw, batch=B, seed=iseed)
        with torch.no_grad():
            ref = RL.outputs_to_dict(model.forward_with_latent(lat, feedforward_image=img))
        keep = {}
        for k, v in ref.items():
            if k in ("means", "covariances", "harmonics", "opacities", "scales", "rotations"):
                v = v[:, ::STRIDE]
            if k == "depth":
                keep["checksum_depth"] = v.double().sum(dim=(2, 3, 4)).float()
                v = v[:, :, ::3, ::3]
            keep[k] = v.clone()
        keep["checksum_means"] = ref["means"].double().sum(dim=1).float()
        keep["checksum_harmonics"] = ref["harmonics"].double().abs().sum().float().reshape(1)
        out["cases"][name] = {"weight_seed": wseed, "input_seed": iseed, "latent_frames": T, "latent_hw": lhw, "image_hw": ihw,
                              "resolution": res, "batch": B, "stride": STRIDE, "outputs": keep}
        print(name, {k: tuple(v.shape) for k, v in keep.items()})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "decoder_tiny.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
