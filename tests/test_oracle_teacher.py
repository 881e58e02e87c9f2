"""Pins the un-stitched (image -> 3DGS) oracle, oracle/decoder_ref.py:teacher_forward, against golden vectors produced by the REAL
reference's EncoderAnySplat.forward (tests/golden/teacher_tiny.pt, script tests/golden/make_teacher_golden.py) and, where
/root/reference is mounted, against the reference imported live with its own constructors' init."""
import os

import pytest
import torch

from oracle import decoder_ref as D
from oracle import ref_loader as RL

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "teacher_tiny.pt")
GAUSS = ("means", "covariances", "harmonics", "opacities", "scales", "rotations")


def _cmp(name, got, want, tol=2e-5):
    err = float((got.double() - want.double()).abs().max())
    scale = float(want.double().abs().max()) + 1e-30
    assert err <= tol * scale + 1e-9, f"{name}: max abs err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("case", ["v3_56", "v4_84_b2"])
def test_teacher_oracle_reproduces_reference_golden(case):
    g = torch.load(GOLD)["cases"][case]
    sd = D.init_state_dict(D.TINY_TEACHER, seed=g["weight_seed"])
    img = D.synthetic_images(g["views"], g["image_hw"], batch=g["batch"], seed=g["input_seed"])
    out = D.teacher_forward(sd, D.TINY_TEACHER, img)
    want, st = g["outputs"], g["stride"]
    for k in GAUSS:
        _cmp(k, out[k][:, ::st], want[k])
    _cmp("depth", out["depth"][:, :, ::3, ::3], want["depth"])
    for k in ("extrinsic", "intrinsic", "last_pred_pose_enc", "scene_scale", "pred_pose_enc_0", "pred_pose_enc_3"):
        _cmp(k, out[k], want[k])
    _cmp("checksum_means", out["means"].double().sum(dim=1).float(), want["checksum_means"], tol=1e-4)
    _cmp("checksum_depth", out["depth"].double().sum(dim=(2, 3, 4)).float(), want["checksum_depth"], tol=1e-4)


@pytest.mark.skipif(not RL.available(), reason="/root/reference is only mounted in the build container")
def test_teacher_oracle_matches_live_reference_with_its_own_init():
    enc = RL.load_teacher(RL.TINY, seed=7)       # the reference constructors' own random init (incl. the bf16-rounded ImageNet buffers)
    sd = {D.E + k: v.detach().clone() for k, v in enc.state_dict().items()}
    assert {k: tuple(v.shape) for k, v in sd.items()} == D.param_shapes(D.TINY_TEACHER)
    torch.manual_seed(3)
    img = torch.rand(1, 3, 3, 56, 56)
    with torch.no_grad():
        ref = RL.outputs_to_dict(enc(img))
    got = D.teacher_forward(sd, D.TINY_TEACHER, img)
    for k in GAUSS + ("depth", "extrinsic", "intrinsic", "scene_scale"):
        _cmp(k, got[k], ref[k])
    # the model normalises with the bf16-rounded ImageNet statistics (anysplat.py:144 casts the aggregator's buffers)
    assert enc.aggregator._resnet_mean.flatten().tolist() == list(D.IMAGENET_MEAN)
    assert enc.aggregator._resnet_std.flatten().tolist() == list(D.IMAGENET_STD)


def test_teacher_manifests_match():
    from vist3a_b200 import stitched_decoder as SD

    assert SD.param_shapes(SD.DecoderConfig(patch_embed=True, dino_blocks=24)) == D.param_shapes(D.FULL_TEACHER)
