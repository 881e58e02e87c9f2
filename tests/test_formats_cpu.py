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
