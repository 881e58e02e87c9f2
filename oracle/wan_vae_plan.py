"""ORACLE-side executable specification (test infrastructure, never imported by vist3a_b200) of the DEVICE decomposition planned for the
Wan VAE decode and encode: the same network as oracle/wan_vae_ref.py:decode (pinned to `utils/wan_utils.py:1078-1117`), restated
operation by operation in the layouts the kernels will use, so that every future kernel has a per-op reference with its exact data layout
and the composition is already known to equal the pinned oracle (tests/test_oracle_vae_plan.py).  Nothing here runs on a GPU.

Layouts: one clip, activations [T, H, W, C] (NDHWC, batch 1), bf16 in HBM between operations (`round_bf16=True` emulates those
roundings; accumulation stays fp32), weights from vist3a_b200.wan_vae_layout.  Planned operations:
  conv_gemm        implicit-GEMM convolution: A row of output pixel (t, h, w) = taps (dt, dh, dw) x C_in gathered at
                   (t + dt - (kt - 1), h + dh - kh // 2, w + dw - kw // 2), zero outside (TMA out-of-bounds fill); bias, optional residual
  rmsnorm_silu     x / max(||x||_2 over C, 1e-12) * sqrt(C) * gamma, then SiLU (the prologue of every residual conv)   (:150-184, :366-372)
  attention        per frame, one head of width C over H x W positions: two GEMMs around a row softmax                  (:428-475)
  time_upsample    conv_gemm with taps (3, 1, 1) over frames 1.., channel halves interleaved in time                    (:257-308)
  upsample_conv    nearest 2x + 3x3 conv as ONE low-resolution GEMM with N = 4 C_out (parity decomposition) + depth-to-space (:226-238)
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from oracle import wan_vae_ref as V
from vist3a_b200.wan_vae_layout import conv3d_weight_to_taps, upsample_conv_weight_to_parity


def _r(t: torch.Tensor, on: bool) -> torch.Tensor:
    return t.bfloat16().float() if on else t


def gather_taps(x: torch.Tensor, kt: int, kh: int, kw: int) -> torch.Tensor:
    """[T, H, W, C] -> [T*H*W, kt*kh*kw*C]: what the TMA producer stages, tap-major"""
    T, H, W, C = x.shape
    xp = F.pad(x, (0, 0, kw // 2, kw // 2, kh // 2, kh // 2, kt - 1, 0))
    cols = [xp[dt:dt + T, dh:dh + H, dw:dw + W] for dt in range(kt) for dh in range(kh) for dw in range(kw)]
    return torch.stack(cols, dim=3).reshape(T * H * W, kt * kh * kw * C)


def conv_gemm(x, wt, bias, taps, residual=None, rb=False):
    T, H, W, _ = x.shape
    y = gather_taps(x, *taps) @ wt.t() + bias
    y = y.reshape(T, H, W, wt.shape[0])
    if residual is not None:
        y = y + residual
    return _r(y, rb)


def rmsnorm_silu(x, gamma, silu=True, rb=False):
    c = x.shape[-1]
    y = x / x.norm(dim=-1, keepdim=True).clamp_min(1e-12) * (c ** 0.5) * gamma
    return _r(F.silu(y) if silu else y, rb)


def prepare(sd: dict, cfg: V.WanVaeConfig, rb: bool = False) -> dict:
    """load-time re-layout: conv weights tap-major, up-sampling convs parity-decomposed, gammas flat; bf16 weights when rb"""
    out = {}
    for k, v in sd.items():
        if not (k.startswith("decoder.") or k.startswith("post_quant_conv.")):
            continue
        if k.endswith("gamma"):
            out[k] = v.reshape(-1)
        elif k.endswith(".resample.1.weight"):
            wt, bt = upsample_conv_weight_to_parity(v, sd[k[:-6] + "bias"])
            out[k], out[k[:-6] + "bias"] = _r(wt, rb), bt
        elif k.endswith(".resample.1.bias"):
            continue
        elif k.endswith("weight"):
            out[k] = _r(conv3d_weight_to_taps(v), rb)
        else:
            out[k] = v
    return out


def _taps_of(p, key, sd_raw):
    w = sd_raw[key]
    return tuple(w.shape[2:]) if w.dim() == 5 else (1,) + tuple(w.shape[2:])


def decode_plan(sd_raw: dict, cfg: V.WanVaeConfig, z: torch.Tensor, round_bf16: bool = False) -> torch.Tensor:
    """latent [1, z, T', h, w] -> frames [1, 3, 1 + 4 (T' - 1), 8h, 8w] through the planned operations"""
    if z.shape[0] != 1:
        raise ValueError("decode_plan: one clip at a time")
    rb = round_bf16
    P = prepare(sd_raw, cfg, rb)

    def conv(x, name, residual=None):
        return conv_gemm(x, P[f"{name}.weight"], P[f"{name}.bias"], _taps_of(P, f"{name}.weight", sd_raw), residual, rb)

    def res_block(x, p):
        h = conv(x, f"{p}.conv_shortcut") if f"{p}.conv_shortcut.weight" in P else x
        y = conv(rmsnorm_silu(x, P[f"{p}.norm1.gamma"], rb=rb), f"{p}.conv1")
        return conv(rmsnorm_silu(y, P[f"{p}.norm2.gamma"], rb=rb), f"{p}.conv2", residual=h)

    def attention(x, p):
        T, H, W, C = x.shape
        n = rmsnorm_silu(x, P[f"{p}.norm.gamma"], silu=False, rb=rb)
        qkv = _r(n.reshape(T, H * W, C) @ P[f"{p}.to_qkv.weight"].t() + P[f"{p}.to_qkv.bias"], rb)      # [T, HW, 3C]
        q, k, v = qkv.split(C, dim=-1)
        pr = _r(torch.softmax(q @ k.transpose(1, 2) / C ** 0.5, dim=-1), rb)                                # [T, HW, HW]
        a = _r(pr @ v, rb)
        y = a @ P[f"{p}.proj.weight"].t() + P[f"{p}.proj.bias"]
        return _r(y.reshape(T, H, W, C) + x, rb)

    def upsample(x, p, temporal):
        T, H, W, C = x.shape
        if temporal and T > 1:
            y = conv(x[1:], f"{p}.time_conv")                                             # [T-1, H, W, 2C]
            y = y.reshape(T - 1, H, W, 2, C).permute(0, 3, 1, 2, 4).reshape(2 * (T - 1), H, W, C)   # halves alternate in time
            x = torch.cat([x[:1], y], dim=0)
            T = x.shape[0]
        y = gather_taps(x, 1, 3, 3) @ P[f"{p}.resample.1.weight"].t() + P[f"{p}.resample.1.bias"]  # [T*H*W, (ph, pw, C/2)]
        co = y.shape[1] // 4
        return _r(y.reshape(T, H, W, 2, 2, co).permute(0, 1, 3, 2, 4, 5).reshape(T, 2 * H, 2 * W, co), rb)   # depth-to-space, k = 2

    x = _r(z[0].permute(1, 2, 3, 0), rb)                                                  # [T', h, w, z]
    x = conv(x, "post_quant_conv")
    x = conv(x, "decoder.conv_in")
    x = res_block(x, "decoder.mid_block.resnets.0")
    x = attention(x, "decoder.mid_block.attentions.0")
    x = res_block(x, "decoder.mid_block.resnets.1")
    ups, _ = V.decoder_layout(cfg)
    for i, (_cin, _cout, mode) in enumerate(ups):
        for j in range(cfg.num_res_blocks + 1):
            x = res_block(x, f"decoder.up_blocks.{i}.resnets.{j}")
        if mode is not None:
            x = upsample(x, f"decoder.up_blocks.{i}.upsamplers.0", temporal=(mode == "up3d"))
    x = rmsnorm_silu(x, P["decoder.norm_out.gamma"], rb=rb)
    y = gather_taps(x, 3, 3, 3) @ P["decoder.conv_out.weight"].t() + P["decoder.conv_out.bias"]   # fp32 out, clamped
    T, H, W, _ = x.shape
    return y.reshape(T, H, W, 3).clamp(-1.0, 1.0).permute(3, 0, 1, 2)[None]


# --------------------------------------------------------------------------------------------------
# encode (diffusion_vae.encode in front of StitchVAE3D.forward: models/stitched_model.py:123-157; utils/wan_utils.py:1021-1048)
# --------------------------------------------------------------------------------------------------
def gather_taps_stride2(x: torch.Tensor) -> torch.Tensor:
    """ZeroPad2d((0, 1, 0, 1)) + 3x3 stride-2 conv input (:240-247): output pixel (i, j) of frame t reads (2i + dh, 2j + dw), zero at
    index H / W (a TMA box with element strides (1, 2, 2, 1) and no offset).  [T, H, W, C] -> [T*(H/2)*(W/2), 9*C]"""
    T, H, W, C = x.shape
    xp = F.pad(x, (0, 0, 0, 1, 0, 1))
    cols = [xp[:, dh:dh + H:2, dw:dw + W:2] for dh in range(3) for dw in range(3)]
    return torch.stack(cols, dim=3).reshape(T * (H // 2) * (W // 2), 9 * C)


def gather_time_stride2(x: torch.Tensor) -> torch.Tensor:
    """temporal down-sampling (:316-330) over the whole clip: output frame k >= 1 reads frames 2(k-1) + dt, dt = 0..2 (no padding).
    [T, H, W, C], T = 1 + 2m -> [m*H*W, 3*C]"""
    T, H, W, C = x.shape
    m = (T - 1) // 2
    cols = [x[dt:dt + 2 * m:2] for dt in range(3)]
    return torch.stack(cols, dim=3).reshape(m * H * W, 3 * C)


def encode_plan(sd: dict, cfg: V.WanVaeConfig, clip: torch.Tensor, round_bf16: bool = False) -> torch.Tensor:
    """clip [1, 3, 1 + 4k, H, W] -> moments [1, 2 z, 1 + k, H/8, W/8]; `quant_conv` (1x1x1 after conv_out) folded into conv_out at load"""
    if clip.shape[0] != 1 or (clip.shape[2] - 1) % 4:
        raise ValueError("encode_plan: one clip of 1 + 4k frames at a time")
    rb = round_bf16

    def wt(name):
        return _r(conv3d_weight_to_taps(sd[f"{name}.weight"]), rb)

    def taps(name):
        w = sd[f"{name}.weight"]
        return tuple(w.shape[2:]) if w.dim() == 5 else (1,) + tuple(w.shape[2:])

    def conv(x, name, residual=None):
        return conv_gemm(x, wt(name), sd[f"{name}.bias"], taps(name), residual, rb)

    def g(name):
        return sd[name].reshape(-1)

    def res_block(x, p):
        h = conv(x, f"{p}.conv_shortcut") if f"{p}.conv_shortcut.weight" in sd else x
        y = conv(rmsnorm_silu(x, g(f"{p}.norm1.gamma"), rb=rb), f"{p}.conv1")
        return conv(rmsnorm_silu(y, g(f"{p}.norm2.gamma"), rb=rb), f"{p}.conv2", residual=h)

    def attention(x, p):
        T, H, W, C = x.shape
        n = rmsnorm_silu(x, g(f"{p}.norm.gamma"), silu=False, rb=rb)
        qkv = _r(n.reshape(T, H * W, C) @ wt(f"{p}.to_qkv").t() + sd[f"{p}.to_qkv.bias"], rb)
        q, k, v = qkv.split(C, dim=-1)
        a = _r(_r(torch.softmax(q @ k.transpose(1, 2) / C ** 0.5, dim=-1), rb) @ v, rb)
        return _r((a @ wt(f"{p}.proj").t() + sd[f"{p}.proj.bias"]).reshape(T, H, W, C) + x, rb)

    def downsample(x, p, temporal):
        T, H, W, C = x.shape
        y = _r((gather_taps_stride2(x) @ wt(f"{p}.resample.1").t() + sd[f"{p}.resample.1.bias"]).reshape(T, H // 2, W // 2, C), rb)
        if temporal and T > 1:
            z = gather_time_stride2(y) @ wt(f"{p}.time_conv").t() + sd[f"{p}.time_conv.bias"]
            y = torch.cat([y[:1], _r(z.reshape((T - 1) // 2, H // 2, W // 2, C), rb)], dim=0)
        return y

    x = _r(clip[0].permute(1, 2, 3, 0), rb)
    x = conv(x, "encoder.conv_in")
    layout, _ = V.encoder_layout(cfg)
    for kind, idx, _cin, _cout in layout:
        p = f"encoder.down_blocks.{idx}"
        x = res_block(x, p) if kind == "res" else downsample(x, p, temporal=(kind == "down3d"))
    x = res_block(x, "encoder.mid_block.resnets.0")
    x = attention(x, "encoder.mid_block.attentions.0")
    x = res_block(x, "encoder.mid_block.resnets.1")
    x = rmsnorm_silu(x, g("encoder.norm_out.gamma"), rb=rb)
    # quant_conv o conv_out as one convolution: W' = Wq Wc (per tap), b' = Wq bc + bq
    wq = sd["quant_conv.weight"].reshape(sd["quant_conv.weight"].shape[0], -1)
    wc = sd["encoder.conv_out.weight"]
    w_fold = torch.einsum("oc,cikhw->oikhw", wq, wc)
    b_fold = wq @ sd["encoder.conv_out.bias"] + sd["quant_conv.bias"]
    y = gather_taps(x, 3, 3, 3) @ _r(conv3d_weight_to_taps(w_fold), rb).t() + b_fold     # fp32 out
    T, H, W, _ = x.shape
    return y.reshape(T, H, W, -1).permute(3, 0, 1, 2)[None]
