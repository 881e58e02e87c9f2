"""GPU parity of the un-stitched AnySplat encoder (image -> 3D Gaussians: DINOv2 patch embedding as an im2col permutation + tcgen05
GEMM, all DINO blocks, aggregator, heads; `StitchVAE3DB200.forward_images`, every op through the C ABI) against the fp32 CPU oracle
(oracle/decoder_ref.py:teacher_forward, pinned bit-exactly to the real reference's EncoderAnySplat.forward) and the golden vectors
the REAL reference produced (tests/golden/teacher_tiny.pt).  Tolerances as in tests/test_decoder_gpu.py (bf16 transformer, TF32
heads): relative L2 < 2e-2, or 2 x the error of the reference's own GPU numerics on the same inputs for camera-conditioned outputs."""
import os

import pytest
import torch

from test_decoder_gpu import FLOOR_MULT, GAUSS, KEYS, REL_TOL, _as_dict, _check, _rel  # noqa: F401

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "teacher_tiny.pt")


def _engine(sd, ocfg, **kw):
    from vist3a_b200.stitched_decoder import DecoderConfig, StitchVAE3DB200

    cfg = DecoderConfig(embed_dim=ocfg.embed_dim, num_heads=ocfg.num_heads, dino_blocks=ocfg.dino_blocks, agg_depth=ocfg.agg_depth,
                        cam_heads=ocfg.cam_heads, cam_trunk=ocfg.cam_trunk, dpt_features=ocfg.dpt_features,
                        dpt_out_channels=ocfg.dpt_out_channels, pos_grid=ocfg.pos_grid, patch=ocfg.patch, sh_degree=ocfg.sh_degree,
                        inter_layers=ocfg.inter_layers, patch_embed=True, **kw)
    return StitchVAE3DB200.from_state_dict(sd, cfg, device="cuda:0")


def _floor(D, sd, ocfg, img, ref):
    with torch.device("cuda"):
        auto = D.teacher_forward({k: v.cuda() for k, v in sd.items()}, ocfg, img.cuda(), gpu_autocast=True)
    return {k: _rel(auto[k], ref[k]) for k in KEYS}


def test_patch_embed_im2col_matches_unfold():
    """the patch-embedding operand: bf16-rounded image, normalised with the model's (bf16-rounded) ImageNet statistics, (c, py, px) order"""
    from vist3a_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(0)
    for dt in (torch.float32, torch.bfloat16):
        img = torch.rand(3, 3, 56, 70, device="cuda", generator=g).to(dt)
        A = ops.patch_embed_im2col(img, 14, 592)
        mean = torch.tensor(ops.IMAGENET_MEAN, device="cuda").view(1, 3, 1, 1)
        std = torch.tensor(ops.IMAGENET_STD, device="cuda").view(1, 3, 1, 1)
        x = (img.to(torch.bfloat16).float() - mean) / std
        want = torch.nn.functional.unfold(x, kernel_size=14, stride=14).transpose(1, 2).reshape(-1, 588).bfloat16()
        assert A.shape == (3 * 4 * 5, 592)
        assert torch.equal(A[:, :588], want) and bool((A[:, 588:] == 0).all())
    with pytest.raises(RuntimeError):
        ops.patch_embed_im2col(torch.rand(1, 3, 50, 56, device="cuda"), 14, 592)   # H not a multiple of the patch size


def test_rgb01_views_layout():
    from vist3a_b200 import ops

    img = torch.rand(2, 3, 3, 8, 12, device="cuda")
    out = ops.rgb01_views_to_nhwc4pad(img)
    assert out.shape == (6, 8, 20, 4)
    assert torch.equal(out[:, :, 3:15, :3], img.reshape(6, 3, 8, 12).permute(0, 2, 3, 1))
    assert bool((out[:, :, :3] == 0).all()) and bool((out[:, :, 15:] == 0).all()) and bool((out[..., 3] == 0).all())


@pytest.mark.parametrize("case", ["v3_56", "v4_84_b2"])
def test_tiny_teacher_matches_reference_golden_and_oracle(case):
    from oracle import decoder_ref as D
    from vist3a_b200 import _lib

    g = torch.load(GOLD)["cases"][case]
    sd = D.init_state_dict(D.TINY_TEACHER, seed=g["weight_seed"])
    img = D.synthetic_images(g["views"], g["image_hw"], batch=g["batch"], seed=g["input_seed"])
    n0 = _lib.launch_count()
    m = _engine(sd, D.TINY_TEACHER)
    out = _as_dict(m.forward_images(img.cuda()))
    torch.cuda.synchronize()
    assert _lib.launch_count() - n0 > 500
    want, st = g["outputs"], g["stride"]
    errs = {k: _rel(out[k][:, ::st], want[k]) for k in GAUSS}
    errs["depth"] = _rel(out["depth"][:, :, ::3, ::3], want["depth"])
    for k in ("extrinsic", "intrinsic", "last_pred_pose_enc", "scene_scale", "pred_pose_enc_0", "pred_pose_enc_3"):
        errs[k] = _rel(out[k], want[k])
    ref = D.teacher_forward(sd, D.TINY_TEACHER, img)
    for k in KEYS:
        errs["oracle_" + k] = _rel(out[k], ref[k])
    _check(case, errs, _floor(D, sd, D.TINY_TEACHER, img, ref))


def test_full_width_teacher_small_views():
    """released widths (ViT-L/14 patch embedding, 24 DINO blocks, 48 aggregator blocks) on 4 views x 112x112"""
    from oracle import decoder_ref as D

    sd = D.init_state_dict(D.FULL_TEACHER, seed=2)
    img = D.synthetic_images(4, 112, seed=3)
    ref = D.teacher_forward(sd, D.FULL_TEACHER, img)
    m = _engine(sd, D.FULL_TEACHER)
    out = _as_dict(m.forward_images(img.cuda()))
    _check("full-width teacher", {k: _rel(out[k], ref[k]) for k in KEYS}, _floor(D, sd, D.FULL_TEACHER, img, ref))
    with pytest.raises(RuntimeError, match="forward_images"):
        m.forward_with_latent(torch.zeros(1, 16, 1, 8, 8), torch.zeros(1, 3, 1, 112, 112))


def test_teacher_full_size_properties_and_voxelize():
    """13 views x 448x448 through the un-stitched path with voxelised fusion on: finite, consistent outputs"""
    from vist3a_b200.stitched_decoder import DecoderConfig, StitchVAE3DB200, random_state_dict

    cfg = DecoderConfig(patch_embed=True, dino_blocks=24, voxelize=True)
    m = StitchVAE3DB200.from_state_dict(random_state_dict(cfg, 0, "cuda"), cfg, "cuda")
    img = torch.rand(1, 13, 3, 448, 448, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    o = m.forward_images(img)
    n = o.gaussians.means.shape[1]
    assert 1 < n <= 13 * 448 * 448 and abs(o.infos["voxelize_ratio"] - n / (13 * 448 * 448)) < 1e-9
    for t in (o.gaussians.means, o.gaussians.covariances, o.gaussians.harmonics, o.gaussians.opacities, o.depth_dict["depth"]):
        assert bool(torch.isfinite(t).all())
    assert float((o.gaussians.rotations.norm(dim=-1) - 1).abs().max()) < 1e-4
    assert o.depth_dict["depth"].shape == (1, 13, 448, 448, 1)
