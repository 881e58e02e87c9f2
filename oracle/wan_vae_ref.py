"""ORACLE (test infrastructure, never imported by vist3a_b200): CPU fp32 restatement of the Wan-2.1 video VAE that sits either side of
the hot path in the reference -- `pipe.vae.decode(latents)` between the denoiser and the stitched decoder (`inference_t23d.py:104-114`,
its frames are what `feedforward_image` is resized from, :116-123) and `diffusion_vae.encode(images).latent_dist.sample()` in front of the
stitching layer (`models/stitched_model.py:123-137,140-157`).  The arithmetic is vendored in the reference: `utils/wan_utils.py:96-1180`
(`WanCausalConv3d` :96-147, `WanRMS_norm` :150-184, `WanResample` :202-330, `WanResidualBlock` :333-425, `WanAttentionBlock` :428-475,
`WanMidBlock` :478-531, `WanEncoder3d` :534-662, `WanUpBlock` :665-742, `WanDecoder3d` :745-901, `AutoencoderKLWan._encode` :1021-1048,
`._decode` :1078-1117, latent statistics :925-960).

Parity: PINNED -- `tests/test_oracle_vae.py` runs the reference's own modules and chunk loops (oracle/ref_loader.py:LiveWanVAE) on the same
weights and inputs, and `tests/golden/wan_vae_tiny.pt` holds vectors they produced.

Form.  The reference walks the clip in chunks (1 frame, then 4 at a time when encoding; one latent frame at a time when decoding) and
carries the last two frames of every causal convolution's input from chunk to chunk.  That procedure is a whole-clip computation, stated
here directly (this is also the form a device implementation wants -- one convolution per layer over all frames):
  * a 3x3x3 causal convolution with its cache == the convolution over the whole clip with two zero frames in front      (:139-147, :376-394)
  * temporal down-sampling (`downsample3d`): the first frame passes through, the frames after it are the stride-2, kernel-3, un-padded
    convolution over the whole clip (windows (0,1,2), (2,3,4), ...)                                                       (:316-330)
  * temporal up-sampling (`upsample3d`): the first frame passes through ("Rep": no time convolution), frames 1.. go through the causal
    kernel-3 convolution to 2 C channels with frame 0 REPLACED BY ZEROS in its window, and the two channel halves are interleaved in time
                                                                                                                         (:257-308)
Everything else is per frame (RMS norm over channels, SiLU, the mid-block's single-head attention over the H x W tokens of a frame,
nearest 2x up-sampling + 3x3 conv, zero-pad (0,1,0,1) + stride-2 3x3 conv).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class WanVaeConfig:
    """constructor arguments of AutoencoderKLWan (utils/wan_utils.py:916-924; defaults = Wan 2.1)"""
    base_dim: int = 96
    z_dim: int = 16
    dim_mult: tuple = (1, 2, 4, 4)
    num_res_blocks: int = 2
    temporal_downsample: tuple = (False, True, True)


WAN_VAE = WanVaeConfig()
TINY_VAE = WanVaeConfig(base_dim=8, z_dim=4)

# utils/wan_utils.py:925-960 (per latent channel); the pipeline de-normalises with  latents * std + mean  before decoding (inference_t23d.py:105-113)
LATENTS_MEAN = (-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503,
                -0.2921)
LATENTS_STD = (2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160)


# --------------------------------------------------------------------------------------------------
# parameter manifest (state-dict keys of the reference modules == diffusers' AutoencoderKLWan keys)
# --------------------------------------------------------------------------------------------------
def _res_shapes(prefix, cin, cout):
    s = {f"{prefix}.norm1.gamma": (cin, 1, 1, 1), f"{prefix}.conv1.weight": (cout, cin, 3, 3, 3), f"{prefix}.conv1.bias": (cout,),
         f"{prefix}.norm2.gamma": (cout, 1, 1, 1), f"{prefix}.conv2.weight": (cout, cout, 3, 3, 3), f"{prefix}.conv2.bias": (cout,)}
    if cin != cout:
        s[f"{prefix}.conv_shortcut.weight"] = (cout, cin, 1, 1, 1)
        s[f"{prefix}.conv_shortcut.bias"] = (cout,)
    return s


def _mid_shapes(prefix, c):
    s = {}
    s.update(_res_shapes(f"{prefix}.resnets.0", c, c))
    s.update({f"{prefix}.attentions.0.norm.gamma": (c, 1, 1), f"{prefix}.attentions.0.to_qkv.weight": (3 * c, c, 1, 1),
              f"{prefix}.attentions.0.to_qkv.bias": (3 * c,), f"{prefix}.attentions.0.proj.weight": (c, c, 1, 1),
              f"{prefix}.attentions.0.proj.bias": (c,)})
    s.update(_res_shapes(f"{prefix}.resnets.1", c, c))
    return s


def encoder_layout(cfg: WanVaeConfig):
    """the down_blocks list as the reference builds it (:575-590): [(kind, index, cin, cout)], kind in res / down2d / down3d"""
    dims = [cfg.base_dim * u for u in (1,) + tuple(cfg.dim_mult)]
    out, idx = [], 0
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(cfg.num_res_blocks):
            out.append(("res", idx, cin, cout))
            idx += 1
            cin = cout
        if i != len(cfg.dim_mult) - 1:
            out.append(("down3d" if cfg.temporal_downsample[i] else "down2d", idx, cout, cout))
            idx += 1
    return out, dims[-1]


def decoder_layout(cfg: WanVaeConfig):
    """up_blocks as the reference builds them (:795-822): [(cin, cout, upsample mode or None)]"""
    dm = tuple(cfg.dim_mult)
    dims = [cfg.base_dim * u for u in (dm[-1],) + dm[::-1]]
    t_up = tuple(cfg.temporal_downsample)[::-1]
    out = []
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        if i > 0:
            cin = cin // 2
        mode = None if i == len(dm) - 1 else ("up3d" if t_up[i] else "up2d")
        out.append((cin, cout, mode))
    return out, dims[0]


def param_shapes(cfg: WanVaeConfig) -> dict:
    s = {}
    z = cfg.z_dim
    s["encoder.conv_in.weight"], s["encoder.conv_in.bias"] = (cfg.base_dim, 3, 3, 3, 3), (cfg.base_dim,)
    layout, ctop = encoder_layout(cfg)
    for kind, idx, cin, cout in layout:
        p = f"encoder.down_blocks.{idx}"
        if kind == "res":
            s.update(_res_shapes(p, cin, cout))
        else:
            s[f"{p}.resample.1.weight"], s[f"{p}.resample.1.bias"] = (cout, cout, 3, 3), (cout,)
            if kind == "down3d":
                s[f"{p}.time_conv.weight"], s[f"{p}.time_conv.bias"] = (cout, cout, 3, 1, 1), (cout,)
    s.update(_mid_shapes("encoder.mid_block", ctop))
    s["encoder.norm_out.gamma"] = (ctop, 1, 1, 1)
    s["encoder.conv_out.weight"], s["encoder.conv_out.bias"] = (2 * z, ctop, 3, 3, 3), (2 * z,)
    s["quant_conv.weight"], s["quant_conv.bias"] = (2 * z, 2 * z, 1, 1, 1), (2 * z,)
    s["post_quant_conv.weight"], s["post_quant_conv.bias"] = (z, z, 1, 1, 1), (z,)
    ups, c0 = decoder_layout(cfg)
    s["decoder.conv_in.weight"], s["decoder.conv_in.bias"] = (c0, z, 3, 3, 3), (c0,)
    s.update(_mid_shapes("decoder.mid_block", c0))
    for i, (cin, cout, mode) in enumerate(ups):
        for j in range(cfg.num_res_blocks + 1):
            s.update(_res_shapes(f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout))
        if mode is not None:
            p = f"decoder.up_blocks.{i}.upsamplers.0"
            s[f"{p}.resample.1.weight"], s[f"{p}.resample.1.bias"] = (cout // 2, cout, 3, 3), (cout // 2,)
            if mode == "up3d":
                s[f"{p}.time_conv.weight"], s[f"{p}.time_conv.bias"] = (2 * cout, cout, 3, 1, 1), (2 * cout,)
    clast = ups[-1][1]
    s["decoder.norm_out.gamma"] = (clast, 1, 1, 1)
    s["decoder.conv_out.weight"], s["decoder.conv_out.bias"] = (3, clast, 3, 3, 3), (3,)
    return s


def init_state_dict(cfg: WanVaeConfig, seed: int = 0) -> dict:
    """seeded synthetic weights (no checkpoint is reachable offline): conv weights N(0, 1/fan_in) so activations stay O(1) through the
    stack, small biases, gammas around 1"""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(cfg).items():
        if k.endswith("gamma"):
            sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            sd[k] = 0.05 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            sd[k] = torch.randn(shp, generator=g) / fan_in ** 0.5
    return sd


def synthetic_clip(frames: int = 5, hw: int = 32, batch: int = 1, seed: int = 0) -> torch.Tensor:
    """[B, 3, T, H, W] in [-1, 1]"""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, 3, frames, hw, hw, generator=g) * 2.0 - 1.0


# --------------------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------------------
def causal_conv3d(x, w, b):
    """WanCausalConv3d with padding = k // 2 (:96-147) over a whole clip: zeros left/right/top/bottom, 2 (k_t // 2) zero frames in front"""
    kt, kh, kw = w.shape[2:]
    x = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2, 2 * (kt // 2), 0))
    return F.conv3d(x, w, b)


def rms_norm(x, gamma):
    """WanRMS_norm (:150-184): x / max(||x||_2 over channels, 1e-12) * sqrt(C) * gamma"""
    return F.normalize(x, dim=1) * (x.shape[1] ** 0.5) * gamma


def _per_frame(x, fn):
    b, c, t, h, w = x.shape
    y = fn(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w))
    return y.view(b, t, *y.shape[1:]).permute(0, 2, 1, 3, 4)


def residual_block(x, sd, p):
    """WanResidualBlock (:333-425)"""
    h = F.conv3d(x, sd[f"{p}.conv_shortcut.weight"], sd[f"{p}.conv_shortcut.bias"]) if f"{p}.conv_shortcut.weight" in sd else x
    y = causal_conv3d(F.silu(rms_norm(x, sd[f"{p}.norm1.gamma"])), sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"])
    y = causal_conv3d(F.silu(rms_norm(y, sd[f"{p}.norm2.gamma"])), sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"])
    return y + h


def attention_block(x, sd, p):
    """WanAttentionBlock (:428-475): per frame, one head of width C over the H x W positions"""
    def one(f):                                   # [B T, C, H, W]
        n, c, h, w = f.shape
        qkv = F.conv2d(rms_norm(f, sd[f"{p}.norm.gamma"]), sd[f"{p}.to_qkv.weight"], sd[f"{p}.to_qkv.bias"])
        q, k, v = qkv.reshape(n, 3 * c, h * w).transpose(1, 2).chunk(3, dim=-1)   # [n, hw, c] each
        a = torch.softmax(q @ k.transpose(1, 2) / c ** 0.5, dim=-1) @ v
        return F.conv2d(a.transpose(1, 2).reshape(n, c, h, w), sd[f"{p}.proj.weight"], sd[f"{p}.proj.bias"])
    return x + _per_frame(x, one)


def mid_block(x, sd, p):
    """WanMidBlock (:478-531): resnet, attention, resnet"""
    x = residual_block(x, sd, f"{p}.resnets.0")
    x = attention_block(x, sd, f"{p}.attentions.0")
    return residual_block(x, sd, f"{p}.resnets.1")


def downsample(x, sd, p, temporal: bool):
    """WanResample 'downsample2d' / 'downsample3d' (:240-247, :311-330)"""
    x = _per_frame(x, lambda f: F.conv2d(F.pad(f, (0, 1, 0, 1)), sd[f"{p}.resample.1.weight"], sd[f"{p}.resample.1.bias"], stride=2))
    if temporal and x.shape[2] > 1:
        x = torch.cat([x[:, :, :1], F.conv3d(x, sd[f"{p}.time_conv.weight"], sd[f"{p}.time_conv.bias"], stride=(2, 1, 1))], dim=2)
    return x


def upsample(x, sd, p, temporal: bool):
    """WanResample 'upsample2d' / 'upsample3d' (:226-238, :257-315)"""
    if temporal and x.shape[2] > 1:
        b, c, t, h, w = x.shape
        y = causal_conv3d(x[:, :, 1:], sd[f"{p}.time_conv.weight"], sd[f"{p}.time_conv.bias"])       # [b, 2c, t-1, h, w]; frame 0 is not seen
        y = y.reshape(b, 2, c, t - 1, h, w).permute(0, 2, 3, 1, 4, 5).reshape(b, c, 2 * (t - 1), h, w)  # halves alternate in time
        x = torch.cat([x[:, :, :1], y], dim=2)
    # nearest-exact at scale 2 = every pixel repeated 2 x 2
    return _per_frame(x, lambda f: F.conv2d(f.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3), sd[f"{p}.resample.1.weight"],
                                            sd[f"{p}.resample.1.bias"], padding=1))


# --------------------------------------------------------------------------------------------------
# encoder / decoder
# --------------------------------------------------------------------------------------------------
def encode_moments(sd: dict, cfg: WanVaeConfig, x: torch.Tensor) -> torch.Tensor:
    """AutoencoderKLWan._encode (:1021-1048): clip [B, 3, 1 + 4k, H, W] in [-1, 1] -> [B, 2 z, 1 + k, H/8, W/8] = mean | logvar"""
    if (x.shape[2] - 1) % 4 != 0:
        raise ValueError("encode: the clip must hold 1 + 4k frames")   # the reference's chunk loop silently drops the remainder
    x = causal_conv3d(x, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"])
    layout, _ = encoder_layout(cfg)
    for kind, idx, _cin, _cout in layout:
        p = f"encoder.down_blocks.{idx}"
        x = residual_block(x, sd, p) if kind == "res" else downsample(x, sd, p, temporal=(kind == "down3d"))
    x = mid_block(x, sd, "encoder.mid_block")
    x = causal_conv3d(F.silu(rms_norm(x, sd["encoder.norm_out.gamma"])), sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"])
    return F.conv3d(x, sd["quant_conv.weight"], sd["quant_conv.bias"])


def posterior(moments: torch.Tensor, noise: torch.Tensor | None = None) -> torch.Tensor:
    """diffusers DiagonalGaussianDistribution (un-vendored; diffusers==0.33.1 models/autoencoders/vae.py): mean, logvar = chunk(2, dim=1);
    logvar clamped to [-30, 20]; sample = mean + exp(0.5 logvar) * noise; mode = mean.  The reference samples
    (`.latent_dist.sample()`, models/stitched_model.py:134): pass the noise explicitly; None gives the mode."""
    mean, logvar = moments.chunk(2, dim=1)
    if noise is None:
        return mean
    return mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * noise


def decode(sd: dict, cfg: WanVaeConfig, z: torch.Tensor) -> torch.Tensor:
    """AutoencoderKLWan._decode (:1078-1117): latent [B, z, T', h, w] -> frames [B, 3, 1 + 4 (T' - 1), 8h, 8w], clamped to [-1, 1]"""
    x = F.conv3d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    x = causal_conv3d(x, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"])
    x = mid_block(x, sd, "decoder.mid_block")
    ups, _ = decoder_layout(cfg)
    for i, (_cin, _cout, mode) in enumerate(ups):
        for j in range(cfg.num_res_blocks + 1):
            x = residual_block(x, sd, f"decoder.up_blocks.{i}.resnets.{j}")
        if mode is not None:
            x = upsample(x, sd, f"decoder.up_blocks.{i}.upsamplers.0", temporal=(mode == "up3d"))
    x = causal_conv3d(F.silu(rms_norm(x, sd["decoder.norm_out.gamma"])), sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"])
    return x.clamp(-1.0, 1.0)


def denormalise_latents(latents: torch.Tensor) -> torch.Tensor:
    """inference_t23d.py:105-113: latents / (1 / std) + mean, per channel"""
    mean = torch.tensor(LATENTS_MEAN, dtype=latents.dtype).view(1, -1, 1, 1, 1)
    std = torch.tensor(LATENTS_STD, dtype=latents.dtype).view(1, -1, 1, 1, 1)
    return latents / (1.0 / std) + mean


def frames_to_feedforward(samples: torch.Tensor, hw: int = 448) -> torch.Tensor:
    """inference_t23d.py:116-123: decoded frames [B, 3, T, H, W] -> trilinear (align_corners=False) resize to [B, 3, T, 448, 448]"""
    return F.interpolate(samples, (samples.shape[2], hw, hw), mode="trilinear", align_corners=False)


def flops(cfg: WanVaeConfig, frames: int, hw: int) -> dict:
    """dense multiply-add work (2 flops each) of encode and decode for a clip of `frames` x hw x hw"""
    def conv(cout, cin, k, t, h, w):
        return 2.0 * cout * cin * k * t * h * w
    enc = conv(cfg.base_dim, 3, 27, frames, hw, hw)
    t, s = frames, hw
    layout, ctop = encoder_layout(cfg)
    for kind, _idx, cin, cout in layout:
        if kind == "res":
            enc += conv(cout, cin, 27, t, s, s) + conv(cout, cout, 27, t, s, s) + (conv(cout, cin, 1, t, s, s) if cin != cout else 0.0)
        else:
            s //= 2
            enc += conv(cout, cout, 9, t, s, s)
            if kind == "down3d":
                t = 1 + (t - 1) // 2
                enc += conv(cout, cout, 3, t - 1, s, s)
    def mid(c, t, s):
        return 4 * conv(c, c, 27, t, s, s) + conv(3 * c, c, 1, t, s, s) + conv(c, c, 1, t, s, s) + t * 4.0 * (s * s) ** 2 * c
    enc += mid(ctop, t, s) + conv(2 * cfg.z_dim, ctop, 27, t, s, s) + conv(2 * cfg.z_dim, 2 * cfg.z_dim, 1, t, s, s)
    tl, sl = t, s
    ups, c0 = decoder_layout(cfg)
    dec = conv(cfg.z_dim, cfg.z_dim, 1, tl, sl, sl) + conv(c0, cfg.z_dim, 27, tl, sl, sl) + mid(c0, tl, sl)
    for cin, cout, mode in ups:
        for j in range(cfg.num_res_blocks + 1):
            ci = cin if j == 0 else cout
            dec += conv(cout, ci, 27, tl, sl, sl) + conv(cout, cout, 27, tl, sl, sl) + (conv(cout, ci, 1, tl, sl, sl) if ci != cout else 0.0)
        if mode is not None:
            if mode == "up3d" and tl > 1:
                dec += conv(2 * cout, cout, 3, tl - 1, sl, sl)
                tl = 1 + 2 * (tl - 1)
            sl *= 2
            dec += conv(cout // 2, cout, 9, tl, sl, sl)
    dec += conv(3, ups[-1][1], 27, tl, sl, sl)
    return {"encode": enc, "decode": dec, "latent_frames": t, "latent_hw": s}
