"""Generates tests/golden/voxel_fusion.pt by running the REAL reference `EncoderAnySplat.voxelizaton_with_fusion`
(/root/reference third_party_model/anysplat/src/model/encoder/anysplat.py:298-335, imported through oracle/ref_loader.py; the two
torch_scatter calls are served by the stand-in documented there) on seeded inputs, and the reference's voxelize=True forward on the
tiny decoder.  Run in the build container only:

    python tests/golden/make_voxel_golden.py

Inputs are regenerated from the seeds by `voxel_case_inputs` (also used by the tests); the fixture holds the reference's outputs.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import decoder_ref as D  # noqa: E402
from oracle import ref_loader as RL  # noqa: E402

# name, seed, views, C, H, W, point spread, voxel_size, conf scale
CASES = [
    ("spread", 1, 2, 83, 16, 16, 0.3, 0.05, 3.0),        # mostly 1-3 points per voxel
    ("dense", 2, 3, 83, 12, 20, 0.05, 0.05, 8.0),        # many points per voxel, large confidence range
    ("halfway", 3, 2, 7, 16, 16, 1.0, 0.25, 1.0),        # coordinates on exact .5 multiples (round-half-even), negative cells
    ("single", 4, 1, 83, 8, 8, 1e-4, 0.5, 1.0),          # every point in one voxel
]
FWD = dict(weight_seed=3, input_seed=5, latent_frames=2, latent_hw=8, image_hw=56, resolution=64, voxel_size=0.03)


def voxel_case_inputs(seed, V, C, H, W, spread, voxel_size, conf_scale):
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(V, C, H, W, generator=g)
    pts = torch.randn(V, 3, H, W, generator=g) * spread
    if spread == 1.0:  # "halfway": snap to multiples of voxel_size / 2 so that p / voxel_size hits x.5 exactly
        pts = (pts / (voxel_size / 2)).round() * (voxel_size / 2)
    conf = torch.randn(V, H, W, generator=g) * conf_scale
    return feat, pts, conf


def main():
    model = RL.load_reference(RL.TINY, resolution=64, seed=0, voxelize=True, voxel_size=FWD["voxel_size"])
    enc = model.stitched_3d_model.encoder
    out = {"cases": {}, "forward": None}
    for name, seed, V, C, H, W, spread, vs, cs in CASES:
        feat, pts, conf = voxel_case_inputs(seed, V, C, H, W, spread, vs, cs)
        vp, vf = enc.voxelizaton_with_fusion(feat, pts, vs, conf=conf)
        vox = (pts.permute(0, 2, 3, 1).flatten(0, 2) / vs).round().int()
        uniq, inv, cnt = torch.unique(vox, dim=0, return_inverse=True, return_counts=True)
        out["cases"][name] = dict(args=(seed, V, C, H, W, spread, vs, cs), voxel_pts=vp.clone(), voxel_feats=vf.clone(), unique=uniq,
                                  inverse=inv.int(), counts=cnt.int())
        print(name, tuple(vp.shape), tuple(vf.shape), int(cnt.max()))
    sd = D.init_state_dict(D.TINY, seed=FWD["weight_seed"])
    missing = [k for k in model.load_state_dict(sd, strict=False).missing_keys if not k.startswith("diffusion_vae")]
    assert not missing, missing
    lat, img = D.synthetic_inputs(D.TINY, views_latent=FWD["latent_frames"], latent_hw=FWD["latent_hw"], image_hw=FWD["image_hw"],
                                  seed=FWD["input_seed"])
    with torch.no_grad():
        ref = RL.outputs_to_dict(model.forward_with_latent(lat, feedforward_image=img))
    keep = {k: ref[k][:, ::7].clone() for k in ("means", "covariances", "harmonics", "opacities", "scales", "rotations")}
    keep["n_voxels"] = torch.tensor([ref["means"].shape[1]])
    keep["checksum_means"] = ref["means"].double().sum(dim=1).float()
    out["forward"] = dict(FWD, outputs=keep)
    print("forward: voxels", int(keep["n_voxels"]), "of", 5 * 56 * 56)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "voxel_fusion.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
