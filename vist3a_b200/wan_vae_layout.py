"""Weight layouts of the Wan-2.1 VAE for the implicit-GEMM convolution path (load-time tensor re-layout only; the device path that
consumes them is the next row of DESIGN.md §8 and is not built yet).  Reference layers: `utils/wan_utils.py:96-147` (WanCausalConv3d),
`:226-238` (WanResample up-sampling: nearest-exact 2x + Conv2d 3x3).

A convolution becomes  out[pixel, :] = sum_k A[pixel, k] * Wt[:, k]  with k = tap * C_in + c, tap = (dt * kh + dh) * kw + dw, where the
A row of an output pixel (t, h, w) is the input at (t + dt - (kt - 1), h + dh - kh // 2, w + dw - kw // 2), zero outside the clip (the
causal two-frame front padding and the spatial zero padding are the same out-of-bounds rule).
"""
from __future__ import annotations

import torch


def conv3d_weight_to_taps(w: torch.Tensor, c_in_pad: int | None = None) -> torch.Tensor:
    """[C_out, C_in, kt, kh, kw] (nn.Conv3d) or [C_out, C_in, kh, kw] (nn.Conv2d, kt = 1) -> [C_out, kt*kh*kw*C_in], tap-major K.
    `c_in_pad` > C_in appends zero input channels per tap: activations kept with a wider pixel stride (e.g. the 96-channel layers stored
    128 wide so that a tap is a whole number of 64-channel k-blocks) read their padding as zeros times zero weights."""
    if w.dim() == 4:
        w = w.unsqueeze(2)
    if w.dim() != 5:
        raise ValueError("conv3d_weight_to_taps: expected a Conv3d / Conv2d weight")
    co, ci, kt, kh, kw = w.shape
    wt = w.permute(0, 2, 3, 4, 1)
    if c_in_pad is not None:
        if c_in_pad < ci:
            raise ValueError("conv3d_weight_to_taps: c_in_pad is smaller than C_in")
        wt = torch.nn.functional.pad(wt, (0, c_in_pad - ci))
        ci = c_in_pad
    return wt.reshape(co, kt * kh * kw * ci).contiguous()


def upsample_conv_weight_to_parity(w: torch.Tensor, bias: torch.Tensor | None = None):
    """Nearest 2x up-sampling followed by a zero-padded 3x3 convolution, restated on the LOW-resolution map.

    Output pixel (2i + ph, 2j + pw) reads, through tap (dh, dw), the high-resolution pixel (2i + ph + dh - 1, 2j + pw + dw - 1), i.e. the
    low-resolution pixel (i + floor((ph + dh - 1) / 2), j + floor((pw + dw - 1) / 2)): per parity the three taps of an axis fall onto two
    low-resolution neighbours and their weights add.  Returns ([4*C_out, 9*C_in] tap-major over the low-resolution 3x3 neighbourhood,
    rows ordered (ph, pw, c_out) = the column order `vist3a_depth_to_space` scatters with k = 2, and the bias repeated per parity)."""
    if w.dim() != 4 or w.shape[2:] != (3, 3):
        raise ValueError("upsample_conv_weight_to_parity: expected a [C_out, C_in, 3, 3] weight")
    co, ci = w.shape[:2]
    out = w.new_zeros(2, 2, co, 3, 3, ci)           # ph, pw, c_out, a (row offset + 1), b (column offset + 1), c_in
    for ph in range(2):
        for dh in range(3):
            a = (ph + dh - 1) // 2 + 1              # floor division: -1 // 2 = -1
            for pw in range(2):
                for dw in range(3):
                    b = (pw + dw - 1) // 2 + 1
                    out[ph, pw, :, a, b, :] += w[:, :, dh, dw]
    wt = out.reshape(4 * co, 9 * ci).contiguous()
    return wt, (None if bias is None else bias.repeat(4).contiguous())
