"""Host-side formats next to the hot path: PLY writer layout and checkpoint / LoRA folding."""
import math
import os
import sys

import numpy as np
import pytest
import torch

from vist3a_b200.checkpoint import apply_stitched_checkpoint, fold_loralib
from vist3a_b200.ply import export_ply, ply_attributes, read_ply

REF = "/root/reference"


def test_ply_layout_and_values(tmp_path):
    g = torch.Generator().manual_seed(0)
    n, d_sh = 257, 25
    means, scales = torch.randn(n, 3, generator=g), torch.rand(n, 3, generator=g) * 0.1 + 1e-3
    rot = torch.randn(n, 4, generator=g)
    sh, op = torch.randn(n, 3, d_sh, generator=g), torch.rand(n, generator=g)
    p = export_ply(means, scales, rot, sh, op, tmp_path / "a" / "g.ply")
    names, rec = read_ply(p)
    assert names == ply_attributes(0) == ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2", "opacity", "scale_0", "scale_1",
                                           "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    assert rec.shape == (n, 17)
    assert np.array_equal(rec[:, :3], means.numpy()) and not rec[:, 3:6].any()
    assert np.array_equal(rec[:, 6:9], sh[..., 0].numpy()) and np.array_equal(rec[:, 9], op.numpy())
    assert np.allclose(rec[:, 10:13], scales.log().numpy())
    qn = (rot / rot.norm(dim=-1, keepdim=True)).numpy()
    wxyz = rec[:, 13:17]
    sign = np.sign((wxyz[:, [1, 2, 3, 0]] * qn).sum(-1, keepdims=True))      # q and -q are the same rotation
    assert np.allclose(wxyz[:, [1, 2, 3, 0]] * sign, qn, atol=1e-5)           # stored as (w, x, y, z), unit norm
    names2, rec2 = read_ply(export_ply(means, scales, rot, sh, op, tmp_path / "full.ply", save_sh_dc_only=False))
    assert len(names2) == 17 + 3 * (d_sh - 1) and np.array_equal(rec2[:, 9:9 + 72], sh[..., 1:].flatten(1).numpy())
    header = open(p, "rb").read(64)
    assert header.startswith(b"ply\nformat binary_little_endian 1.0\nelement vertex 257\n")


def test_fold_loralib_linear_and_conv_known_answer():
    g = torch.Generator().manual_seed(1)
    W = torch.randn(6, 5, generator=g)
    A, B = torch.randn(2, 5, generator=g), torch.randn(6, 2, generator=g)
    Wc = torch.randn(4, 3, 3, 3, generator=g)                      # conv: lora_A [r*k, in*k], lora_B [out*k, r*k]
    Ac, Bc = torch.randn(2 * 3, 3 * 3, generator=g), torch.randn(4 * 3, 2 * 3, generator=g)
    sd = {"m.enc.fc.weight": W, "m.enc.conv.weight": Wc, "m.enc.fc.bias": torch.zeros(6)}
    out = fold_loralib(sd, {"enc.fc.lora_A": A, "enc.fc.lora_B": B, "enc.conv.lora_A": Ac, "enc.conv.lora_B": Bc}, alpha=32, prefix="m.")
    assert torch.allclose(out["m.enc.fc.weight"], W + (B @ A) * (32 / 2))
    assert torch.allclose(out["m.enc.conv.weight"], Wc + (Bc @ Ac).view(Wc.shape) * (32 / 2))
    assert out["m.enc.fc.bias"] is sd["m.enc.fc.bias"] and sd["m.enc.fc.weight"] is W     # input dict untouched
    x = torch.randn(7, 5, generator=g)
    assert torch.allclose(x @ out["m.enc.fc.weight"].t(), x @ W.t() + (x @ A.t() @ B.t()) * 16, atol=1e-5)   # == unmerged LoRA forward
    with pytest.raises(KeyError):
        fold_loralib(sd, {"enc.nope.lora_A": A, "enc.nope.lora_B": B}, alpha=32, prefix="m.")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "utils", "lora_util")), reason="/root/reference is only mounted in the build container")
def test_fold_matches_reference_lora_layers_in_eval_mode():
    sys.path.insert(0, REF)
    try:
        from utils.lora_util.layers import Conv2d as RefConv
        from utils.lora_util.layers import Linear as RefLinear
    finally:
        sys.path.remove(REF)
    torch.manual_seed(2)
    lin = RefLinear(16, 24, r=4, lora_alpha=32)
    conv = RefConv(8, 12, kernel_size=3, r=4, lora_alpha=32, padding=1)
    with torch.no_grad():
        lin.lora_B.normal_()
        conv.lora_B.normal_() if hasattr(conv, "lora_B") else conv.conv.lora_B.normal_()
    sd = {"a.weight": lin.weight.detach().clone(), "b.weight": (conv.conv.weight if hasattr(conv, "conv") else conv.weight).detach().clone()}
    lora = {"a.lora_A": lin.lora_A.detach(), "a.lora_B": lin.lora_B.detach(), "b.lora_A": conv.lora_A.detach(), "b.lora_B": conv.lora_B.detach()}
    folded = fold_loralib(sd, lora, alpha=32)
    x = torch.randn(3, 16)
    xi = torch.randn(2, 8, 10, 10)
    lin.train()
    conv.train()                                    # unmerged LoRA branch
    assert torch.allclose(torch.nn.functional.linear(x, folded["a.weight"], lin.bias), lin(x), atol=1e-5)
    refc = conv(xi)
    bias = (conv.conv.bias if hasattr(conv, "conv") else conv.bias)
    assert torch.allclose(torch.nn.functional.conv2d(xi, folded["b.weight"], bias, padding=1), refc, atol=1e-4)


def test_apply_stitched_checkpoint_layout():
    sd = {"stitching_layer.weight": torch.zeros(4, 16, 5, 3, 3), "stitching_layer.bias": torch.zeros(4),
          "stitched_3d_model.encoder.aggregator.patch_embed.cls_token": torch.zeros(1, 1, 4),
          "stitched_3d_model.encoder.aggregator.patch_embed.register_tokens": torch.zeros(1, 4, 4),
          "stitched_3d_model.encoder.aggregator.patch_embed.mask_token": torch.zeros(1, 4),
          "stitched_3d_model.encoder.aggregator.frame_blocks.0.attn.qkv.weight": torch.zeros(12, 4),
          "stitched_3d_model.encoder.aggregator.frame_blocks.0.attn.qkv.bias": torch.zeros(12)}
    ck = {"lora": {"encoder.aggregator.frame_blocks.0.attn.qkv.lora_A": torch.ones(2, 4), "encoder.aggregator.frame_blocks.0.attn.qkv.lora_B": torch.ones(12, 2),
                   "encoder.aggregator.frame_blocks.0.attn.qkv.bias": torch.full((12,), 3.0)},
          "stitching_layer": {"weight": torch.ones(4, 16, 5, 3, 3), "bias": torch.ones(4)}, "cls_token": torch.ones(1, 1, 4),
          "register_tokens": torch.ones(1, 4, 4), "mask_token": torch.ones(1, 4)}
    out = apply_stitched_checkpoint(sd, ck, lora_alpha=32)
    q = "stitched_3d_model.encoder.aggregator.frame_blocks.0.attn.qkv."
    assert torch.equal(out[q + "weight"], torch.full((12, 4), 2 * 16.0)) and torch.equal(out[q + "bias"], torch.full((12,), 3.0))
    assert float(out["stitching_layer.weight"].sum()) == 4 * 16 * 45 and float(out["stitched_3d_model.encoder.aggregator.patch_embed.cls_token"].sum()) == 4


# ------------------------------------------------------------------------------------------------
# load_stitching_model(args) mirror: the two argument mini-grammars and the block renumbering
# ------------------------------------------------------------------------------------------------
def test_conv_spec_and_lora_grammars_match_the_reference():
    import pytest

    from vist3a_b200 import loader as LD

    spec = LD.parse_conv_spec("conv3d_k5x3x3_o1024_s1x2x2_p2x1x1")
    assert (spec.dim, spec.out_channels, spec.kernel_size, spec.stride, spec.padding, spec.dilation) == (3, 1024, (5, 3, 3), (1, 2, 2), (2, 1, 1), 1)
    assert LD.parse_conv_spec("conv2d_k3_o64") == LD.ConvSpec(2, 64, 3, 1, 0, 1)
    with pytest.raises(ValueError):
        LD.parse_conv_spec("conv4d_k3_o8")
    lc = LD.parse_lora_mode("r8,a16,d0.05,f0")
    assert (lc.r, lc.alpha, lc.dropout, lc.fan_in_fan_out, lc.bias) == (8, 16, 0.05, False, "lora_only")
    lc = LD.parse_lora_mode("r4,a32,bnone,tqkv|proj,enc,fix_head")
    assert (lc.r, lc.alpha, lc.bias, lc.target_modules, lc.finetune_encoder, lc.freeze_head) == (4, 32, "none", ("qkv", "proj"), True, True)
    with pytest.raises(ValueError):
        LD.parse_lora_mode("r8,zz")
    with pytest.raises(ValueError):
        LD.parse_lora_mode("bsome")
    # against the reference's own parsers where the reference tree is mounted (build container)
    import os
    import sys

    if os.path.isdir("/root/reference/models"):
        sys.path.insert(0, "/root/reference")
        try:
            from models.stitching_layer_builder import parse_conv_spec as ref_conv
            from utils.lora_util.utils import parse_lora_mode as ref_lora
        except Exception:
            return
        finally:
            sys.path.remove("/root/reference")
        for txt in ("conv3d_k5x3x3_o1024_s1x2x2_p2x1x1", "conv2d_k3_o64", "conv1d_k7_o8_s2_p3_d2", "conv3d_k3x3x3_o32_s2_p1"):
            a, b = LD.parse_conv_spec(txt), ref_conv(txt)
            assert (a.dim, a.out_channels, a.kernel_size, a.stride, a.padding, a.dilation) == (b.dim, b.out_channels, b.kernel_size, b.stride, b.padding, b.dilation)
        for txt in ("r8,a16,d0.05,f0", "r4,a32,bnone,tqkv|proj,enc,fix_head", "r16,a1,f1,ball"):
            a, b = LD.parse_lora_mode(txt), ref_lora(txt)
            for f in ("r", "alpha", "dropout", "bias", "target_modules", "fan_in_fan_out", "finetune_encoder", "freeze_head"):
                assert getattr(a, f) == getattr(b, f), (txt, f)


def test_unstitched_checkpoint_is_renumbered_or_refused():
    import pytest

    from vist3a_b200.checkpoint import apply_stitched_checkpoint, renumber_stitched_blocks

    pe = "stitched_3d_model.encoder.aggregator.patch_embed."
    sd = {pe + "patch_embed.proj.weight": torch.zeros(8, 3, 14, 14), pe + "patch_embed.proj.bias": torch.zeros(8), pe + "cls_token": torch.zeros(1, 1, 8)}
    for i in range(6):
        sd[pe + f"blocks.{i}.attn.qkv.weight"] = torch.full((24, 8), float(i))
    ck = {"lora": {"encoder.aggregator.patch_embed.blocks.0.attn.qkv.lora_A": torch.ones(2, 8),
                   "encoder.aggregator.patch_embed.blocks.0.attn.qkv.lora_B": torch.ones(24, 2)},
          "stitching_layer": {"weight": torch.zeros(8, 16, 5, 3, 3), "bias": torch.zeros(8)}}
    with pytest.raises(ValueError):          # 24-block numbering + patch-embedding conv: refused without the stitch index
        apply_stitched_checkpoint(sd, ck, lora_alpha=2, lora_r=2)
    rn = renumber_stitched_blocks(sd, 2)
    assert pe + "patch_embed.proj.weight" not in rn and pe + "blocks.4.attn.qkv.weight" not in rn
    assert float(rn[pe + "blocks.0.attn.qkv.weight"][0, 0]) == 2.0 and float(rn[pe + "blocks.3.attn.qkv.weight"][0, 0]) == 5.0
    out = apply_stitched_checkpoint(sd, ck, lora_alpha=2, lora_r=2, stitched_layer_index=2)
    # the LoRA delta of stitched blocks.0 lands on ORIGINAL block 2: 2 + (alpha / r) * (B @ A) = 2 + 2
    assert float(out[pe + "blocks.0.attn.qkv.weight"][0, 0]) == 4.0 and float(out[pe + "blocks.1.attn.qkv.weight"][0, 0]) == 3.0


def test_load_stitching_model_argument_errors():
    import types

    import pytest

    from vist3a_b200 import loader as LD

    args = types.SimpleNamespace(feedforward_model="dust3r", video_model="wan", stitching_layer_location="enc_blocks_2",
                                 stitching_layer_config="conv3d_k5x3x3_o1024_s1x2x2_p2x1x1", resolution=512, initialization_weight_path=None,
                                 lora_config="r8,a16,d0.05,f0", checkpoint_path=None)
    with pytest.raises(NotImplementedError):
        LD.load_stitching_model(args, feedforward_state_dict={})
    args.feedforward_model = "anysplat"
    args.video_model = "cogvideo"
    with pytest.raises(NotImplementedError):
        LD.load_stitching_model(args, feedforward_state_dict={})
    args.video_model = "wan"
    args.stitching_layer_location = "dec_blocks_1"
    with pytest.raises(NotImplementedError):
        LD.load_stitching_model(args, feedforward_state_dict={})
    args.stitching_layer_location = "enc_blocks_2"
    args.stitching_layer_config = "conv3d_k3x3x3_o1024_s1x2x2_p1x1x1"
    with pytest.raises(NotImplementedError):
        LD.load_stitching_model(args, feedforward_state_dict={})
    args.stitching_layer_config = "not_a_conv"
    with pytest.raises(ValueError):
        LD.load_stitching_model(args, feedforward_state_dict={})
    args.stitching_layer_config = "conv3d_k5x3x3_o1024_s1x2x2_p2x1x1"
    with pytest.raises(FileNotFoundError):    # no network path: the AnySplat weights must be local
        LD.load_stitching_model(args)
