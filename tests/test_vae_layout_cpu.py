"""Load-time weight layouts of the VAE convolution path (vist3a_b200/wan_vae_layout.py) against torch's convolutions: the GEMM form
(explicit gather on the CPU here) must reproduce WanCausalConv3d (utils/wan_utils.py:96-147) and the up-sampling conv (:226-238)."""
import torch
import torch.nn.functional as F

from vist3a_b200.wan_vae_layout import conv3d_weight_to_taps, upsample_conv_weight_to_parity


def _gather(x, kt, kh, kw):
    """x [T, H, W, C] -> A [T*H*W, kt*kh*kw*C] with the module's out-of-bounds rule (zeros; time padded in front only)"""
    T, H, W, C = x.shape
    xp = F.pad(x, (0, 0, kw // 2, kw // 2, kh // 2, kh // 2, kt - 1, 0))
    cols = [xp[dt:dt + T, dh:dh + H, dw:dw + W] for dt in range(kt) for dh in range(kh) for dw in range(kw)]
    return torch.stack(cols, dim=3).reshape(T * H * W, kt * kh * kw * C)


def test_causal_conv3d_as_tap_major_gemm():
    g = torch.Generator().manual_seed(0)
    for (kt, kh, kw) in ((3, 3, 3), (3, 1, 1), (1, 1, 1)):
        x = torch.randn(5, 6, 7, 4, generator=g, dtype=torch.float64)                 # T, H, W, C
        w = torch.randn(8, 4, kt, kh, kw, generator=g, dtype=torch.float64)
        b = torch.randn(8, generator=g, dtype=torch.float64)
        ref = F.conv3d(F.pad(x.permute(3, 0, 1, 2)[None], (kw // 2, kw // 2, kh // 2, kh // 2, 2 * (kt // 2), 0)), w, b)[0].permute(1, 2, 3, 0)
        out = (_gather(x, kt, kh, kw) @ conv3d_weight_to_taps(w).t() + b).reshape(5, 6, 7, 8)
        assert torch.allclose(out, ref, atol=1e-12)


def test_upsample_conv_on_the_low_resolution_map():
    g = torch.Generator().manual_seed(1)
    n, h, w_, ci, co = 2, 5, 6, 4, 3
    x = torch.randn(n, h, w_, ci, generator=g, dtype=torch.float64)
    w = torch.randn(co, ci, 3, 3, generator=g, dtype=torch.float64)
    b = torch.randn(co, generator=g, dtype=torch.float64)
    up = x.permute(0, 3, 1, 2).repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    ref = F.conv2d(up, w, b, padding=1).permute(0, 2, 3, 1)                           # [n, 2h, 2w, co]
    wt, bt = upsample_conv_weight_to_parity(w, b)
    assert wt.shape == (4 * co, 9 * ci) and bt.shape == (4 * co,)
    y = _gather(x, 1, 3, 3) @ wt.t() + bt                                              # [n*h*w, (ph, pw, co)]
    # depth-to-space, k = 2 (vist3a_depth_to_space: column (dy*k + dx)*C + c)
    out = y.reshape(n, h, w_, 2, 2, co).permute(0, 1, 3, 2, 4, 5).reshape(n, 2 * h, 2 * w_, co)
    assert torch.allclose(out, ref, atol=1e-12)
    # per parity only a 2x2 block of the low-resolution neighbourhood carries weight
    blocks = wt.reshape(2, 2, co, 3, 3, ci).abs().sum(dim=(2, 5)) > 0
    assert all(int(blocks[ph, pw].sum()) == 4 for ph in range(2) for pw in range(2))


def test_padded_input_channels():
    """96-channel activations stored 128 wide: the padded tap-major weight reproduces the convolution on the padded layout"""
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 5, 6, 96, generator=g, dtype=torch.float64)
    w = torch.randn(8, 96, 3, 3, 3, generator=g, dtype=torch.float64)
    ref = _gather(x, 3, 3, 3) @ conv3d_weight_to_taps(w).t()
    xp = F.pad(x, (0, 32))                                             # pixel stride 128, channels 96..127 zero
    wt = conv3d_weight_to_taps(w, c_in_pad=128)
    assert wt.shape == (8, 27 * 128)
    assert torch.allclose(_gather(xp, 3, 3, 3) @ wt.t(), ref, atol=1e-12)
    # garbage in the padding channels is harmless as long as it is finite: it meets zero weights
    xp[..., 96:] = 7.0
    assert torch.allclose(_gather(xp, 3, 3, 3) @ wt.t(), ref, atol=1e-12)
