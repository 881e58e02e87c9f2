"""ORACLE (test infrastructure, never imported by vist3a_b200): CPU restatement of the Wan-2.1 DiT.

PARITY UNPINNED.  The arithmetic of this half of the hot path lives in `diffusers==0.33.1`
(`/root/reference/requirements.txt:20`), which is neither vendored in the reference tree nor
installed/installable here (no network), and the reference holds no test or golden vector for it.
What follows restates the published algorithm of diffusers 0.33.1
`models/transformers/transformer_wan.py` (WanRotaryPosEmbed, WanTimeTextImageEmbedding,
WanAttnProcessor2_0, WanTransformerBlock, WanTransformer3DModel.forward) as summarised in
SURVEY.md Appendix A, anchored on the reference's own call sites:
  - pipeline call            /root/reference/inference_t23d.py:94-103
  - direct transformer call  /root/reference/train_vdm.py:557-562, 598-603
  - LoRA targets (r=8, a=16) /root/reference/train_vdm.py:370-388
What pins it instead: state-dict key names and shapes identical to the diffusers checkpoint (so a
real checkpoint loads unchanged), the published parameter counts (1.3B -> 1.419 B, 14B -> 14.29 B;
see `param_count`), and analytic identities checked in tests/test_oracle_dit.py.

Everything runs in fp32 (RoPE in fp64, as the original does) on CPU with plain torch ops.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class WanConfig:
    """Field names follow WanTransformer3DModel.config (diffusers 0.33.1)."""

    patch_size: Tuple[int, int, int] = (1, 2, 2)
    num_attention_heads: int = 12
    attention_head_dim: int = 128
    in_channels: int = 16
    out_channels: int = 16
    text_dim: int = 4096
    freq_dim: int = 256
    ffn_dim: int = 8960
    num_layers: int = 30
    cross_attn_norm: bool = True
    eps: float = 1e-6
    rope_max_seq_len: int = 1024

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim


WAN_1_3B = WanConfig()
WAN_14B = WanConfig(num_attention_heads=40, ffn_dim=13824, num_layers=40)
# small configuration with the same structure (head_dim stays 128): seconds on CPU
WAN_TINY = WanConfig(num_attention_heads=2, ffn_dim=512, num_layers=2, text_dim=128)


def param_shapes(cfg: WanConfig) -> Dict[str, Tuple[int, ...]]:
    """State-dict manifest: key -> shape, in diffusers naming."""
    D, Fd = cfg.inner_dim, cfg.ffn_dim
    pt, ph, pw = cfg.patch_size
    s: Dict[str, Tuple[int, ...]] = {
        "patch_embedding.weight": (D, cfg.in_channels, pt, ph, pw),
        "patch_embedding.bias": (D,),
        "condition_embedder.time_embedder.linear_1.weight": (D, cfg.freq_dim),
        "condition_embedder.time_embedder.linear_1.bias": (D,),
        "condition_embedder.time_embedder.linear_2.weight": (D, D),
        "condition_embedder.time_embedder.linear_2.bias": (D,),
        "condition_embedder.time_proj.weight": (6 * D, D),
        "condition_embedder.time_proj.bias": (6 * D,),
        "condition_embedder.text_embedder.linear_1.weight": (D, cfg.text_dim),
        "condition_embedder.text_embedder.linear_1.bias": (D,),
        "condition_embedder.text_embedder.linear_2.weight": (D, D),
        "condition_embedder.text_embedder.linear_2.bias": (D,),
        "scale_shift_table": (1, 2, D),
        "proj_out.weight": (cfg.out_channels * pt * ph * pw, D),
        "proj_out.bias": (cfg.out_channels * pt * ph * pw,),
    }
    for i in range(cfg.num_layers):
        p = f"blocks.{i}."
        s[p + "scale_shift_table"] = (1, 6, D)
        for a in ("attn1", "attn2"):
            for l in ("to_q", "to_k", "to_v", "to_out.0"):
                s[p + f"{a}.{l}.weight"] = (D, D)
                s[p + f"{a}.{l}.bias"] = (D,)
            s[p + f"{a}.norm_q.weight"] = (D,)
            s[p + f"{a}.norm_k.weight"] = (D,)
        if cfg.cross_attn_norm:
            s[p + "norm2.weight"] = (D,)
            s[p + "norm2.bias"] = (D,)
        s[p + "ffn.net.0.proj.weight"] = (Fd, D)
        s[p + "ffn.net.0.proj.bias"] = (Fd,)
        s[p + "ffn.net.2.weight"] = (D, Fd)
        s[p + "ffn.net.2.bias"] = (D,)
    return s


def param_count(cfg: WanConfig) -> int:
    return sum(math.prod(v) for v in param_shapes(cfg).values())


def init_state_dict(cfg: WanConfig, seed: int = 0, *, bias_std: float = 0.0, dtype=torch.float32,
                    round_bf16: bool = True) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights: Linear/Conv weights N(0, 0.02), biases N(0, bias_std)
    (0 for the benchmark; tests use a non-zero std so bias paths are exercised), scale_shift_table
    = randn/sqrt(D), norm weights 1 (+N(0, bias_std)).  With round_bf16 the values are bf16-representable
    so the fp32 oracle and the bf16 engine start from identical numbers."""
    g = torch.Generator().manual_seed(seed)
    D = cfg.inner_dim
    sd: Dict[str, torch.Tensor] = {}
    for k, shp in param_shapes(cfg).items():
        if k.endswith("scale_shift_table"):
            t = torch.randn(shp, generator=g) / math.sqrt(D)
        elif "norm" in k and k.endswith(".weight"):
            t = torch.ones(shp) + (torch.randn(shp, generator=g) * bias_std if bias_std else 0.0)
        elif k.endswith(".bias"):
            t = torch.randn(shp, generator=g) * bias_std if bias_std else torch.zeros(shp)
        else:
            t = torch.randn(shp, generator=g) * 0.02
        if round_bf16:
            t = t.bfloat16().float()
        sd[k] = t.to(dtype)
    return sd


def init_lora(cfg: WanConfig, seed: int = 1, r: int = 8, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """PEFT-style adapter tensors for the 8 attention projections of every block
    (/root/reference/train_vdm.py:370-388), saved-file key convention (no adapter name)."""
    g = torch.Generator().manual_seed(seed)
    D = cfg.inner_dim
    out = {}
    for i in range(cfg.num_layers):
        for a in ("attn1", "attn2"):
            for l in ("to_q", "to_k", "to_v", "to_out.0"):
                base = f"base_model.model.blocks.{i}.{a}.{l}"
                out[base + ".lora_A.weight"] = (torch.randn(r, D, generator=g) * std).bfloat16().float()
                out[base + ".lora_B.weight"] = (torch.randn(D, r, generator=g) * std).bfloat16().float()
    return out


def fold_lora(sd: Dict[str, torch.Tensor], lora: Dict[str, torch.Tensor], lora_alpha: float = 16.0, r: int = 8):
    """W += (alpha / r) * B @ A  -- what PeftModel computes unmerged at run time (inference_t23d.py:74-77)."""
    out = dict(sd)
    scale = lora_alpha / r
    for k, A in lora.items():
        if not k.endswith("lora_A.weight") and not k.endswith("lora_A.default.weight"):
            continue
        kb = k.replace("lora_A", "lora_B")
        name = k.split(".lora_A")[0]
        name = name[len("base_model.model."):] if name.startswith("base_model.model.") else name
        out[name + ".weight"] = out[name + ".weight"].float() + scale * (lora[kb].float() @ A.float())
    return out


# ------------------------------------------------------------------------------------------------
# forward
# ------------------------------------------------------------------------------------------------
def rope_freqs(cfg: WanConfig, f: int, h: int, w: int, device=None) -> torch.Tensor:
    """complex128 [f*h*w, head_dim/2] (WanRotaryPosEmbed): head_dim split t/h/w = 44/42/42 at d=128."""
    hd = cfg.attention_head_dim
    h_dim = w_dim = 2 * (hd // 6)
    t_dim = hd - h_dim - w_dim
    tabs = []
    for dim in (t_dim, h_dim, w_dim):
        fr = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float64, device=device)[: dim // 2] / dim))
        ang = torch.outer(torch.arange(cfg.rope_max_seq_len, dtype=torch.float64, device=device), fr)
        tabs.append(torch.polar(torch.ones_like(ang), ang))
    ft = tabs[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1)
    fh = tabs[1][:h].view(1, h, 1, -1).expand(f, h, w, -1)
    fw = tabs[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)
    return torch.cat([ft, fh, fw], dim=-1).reshape(f * h * w, -1)


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """Timesteps(num_channels=dim, flip_sin_to_cos=True, downscale_freq_shift=0)."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t.float()[:, None] * torch.exp(exponent)[None]
    return torch.cat([emb.cos(), emb.sin()], dim=-1)


def _ln(x, weight=None, bias=None, eps=1e-6):
    return F.layer_norm(x.float(), (x.shape[-1],), weight, bias, eps)


def _rms(x, w, eps):
    return x * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + eps) * w


def _attention(sd, p, x, ctx, cfg: WanConfig, freqs: Optional[torch.Tensor]):
    H, hd = cfg.num_attention_heads, cfg.attention_head_dim
    q = F.linear(x, sd[p + "to_q.weight"], sd[p + "to_q.bias"])
    k = F.linear(ctx, sd[p + "to_k.weight"], sd[p + "to_k.bias"])
    v = F.linear(ctx, sd[p + "to_v.weight"], sd[p + "to_v.bias"])
    q = _rms(q, sd[p + "norm_q.weight"], cfg.eps)  # rms_norm_across_heads: over the full inner dim
    k = _rms(k, sd[p + "norm_k.weight"], cfg.eps)
    B = x.shape[0]
    q = q.view(B, -1, H, hd).transpose(1, 2)
    k = k.view(B, -1, H, hd).transpose(1, 2)
    v = v.view(B, -1, H, hd).transpose(1, 2)
    if freqs is not None:
        def rot(t):
            tc = torch.view_as_complex(t.to(torch.float64).unflatten(3, (-1, 2)))
            return torch.view_as_real(tc * freqs[None, None]).flatten(3, 4).to(t.dtype)
        q, k = rot(q), rot(k)
    o = F.scaled_dot_product_attention(q, k, v)
    o = o.transpose(1, 2).flatten(2, 3)
    return F.linear(o, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])


def condition_embed(sd, cfg: WanConfig, timestep: torch.Tensor, text: torch.Tensor):
    """WanTimeTextImageEmbedding: returns temb [B, D], timestep_proj [B, 6, D], text [B, Lt, D]."""
    p = "condition_embedder."
    ts = timestep_embedding(timestep, cfg.freq_dim)
    temb = F.linear(F.silu(F.linear(ts, sd[p + "time_embedder.linear_1.weight"], sd[p + "time_embedder.linear_1.bias"])),
                    sd[p + "time_embedder.linear_2.weight"], sd[p + "time_embedder.linear_2.bias"])
    tproj = F.linear(F.silu(temb), sd[p + "time_proj.weight"], sd[p + "time_proj.bias"]).unflatten(1, (6, -1))
    txt = F.linear(F.gelu(F.linear(text.float(), sd[p + "text_embedder.linear_1.weight"],
                                   sd[p + "text_embedder.linear_1.bias"]), approximate="tanh"),
                   sd[p + "text_embedder.linear_2.weight"], sd[p + "text_embedder.linear_2.bias"])
    return temb, tproj, txt


def block_forward(sd, i: int, cfg: WanConfig, x, txt, tproj, freqs, stream_dtype=None):
    """WanTransformerBlock.forward (AdaLN-zero; chunk order shift, scale, gate, c_shift, c_scale, c_gate).
    stream_dtype: the dtype the residual stream is kept in.  diffusers computes every gated residual add in fp32 and casts the sum back
    with `.type_as(hidden_states)`, so a bf16 pipeline carries a bf16 stream; None = no rounding (the fp32 oracle)."""
    p = f"blocks.{i}."
    cast = (lambda t: t.to(stream_dtype)) if stream_dtype is not None else (lambda t: t)
    sh1, sc1, g1, sh2, sc2, g2 = (sd[p + "scale_shift_table"].float() + tproj.float()).chunk(6, dim=1)
    h = cast(_ln(x, eps=cfg.eps) * (1 + sc1) + sh1)
    x = cast(x.float() + _attention(sd, p + "attn1.", h, h, cfg, freqs) * g1) if stream_dtype is not None else x + _attention(sd, p + "attn1.", h, h, cfg, freqs) * g1
    if cfg.cross_attn_norm:
        h = cast(_ln(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"], cfg.eps))
    else:
        h = x
    x = cast(x + _attention(sd, p + "attn2.", h, txt, cfg, None))
    h = cast(_ln(x, eps=cfg.eps) * (1 + sc2) + sh2)
    f = F.linear(F.gelu(F.linear(h, sd[p + "ffn.net.0.proj.weight"], sd[p + "ffn.net.0.proj.bias"]), approximate="tanh"),
                 sd[p + "ffn.net.2.weight"], sd[p + "ffn.net.2.bias"])
    return cast(x.float() + f.float() * g2) if stream_dtype is not None else x + f * g2


@torch.no_grad()
def wan_forward(sd: Dict[str, torch.Tensor], cfg: WanConfig, hidden_states: torch.Tensor, timestep: torch.Tensor,
                encoder_hidden_states: torch.Tensor, num_layers: Optional[int] = None, cast_fp32: bool = True,
                stream_dtype=None) -> torch.Tensor:
    """WanTransformer3DModel.forward(hidden_states [B,C,T,H,W], timestep [B], encoder_hidden_states [B,Lt,text_dim]).
    cast_fp32=False runs on the tensors as given (any device / dtype): bench.py's `--impl torch` comparator executes this same graph with
    bf16 weights under CUDA autocast, i.e. through torch's cuBLAS / SDPA kernels.  stream_dtype=torch.bfloat16 additionally keeps the
    residual stream in bf16 between blocks, as a bf16 diffusers pipeline does (see block_forward)."""
    if cast_fp32:
        sd = {k: v.float() for k, v in sd.items()}
    B, C, T, H, W = hidden_states.shape
    pt, ph, pw = cfg.patch_size
    f, h, w = T // pt, H // ph, W // pw
    freqs = rope_freqs(cfg, f, h, w, device=hidden_states.device)
    x = F.conv3d(hidden_states.float() if cast_fp32 else hidden_states.to(sd["patch_embedding.weight"].dtype), sd["patch_embedding.weight"],
                 sd["patch_embedding.bias"], stride=cfg.patch_size)
    x = x.flatten(2).transpose(1, 2)  # [B, L, D], token order (t, h, w)
    temb, tproj, txt = condition_embed(sd, cfg, timestep, encoder_hidden_states)
    if stream_dtype is not None:
        x = x.to(stream_dtype)
    for i in range(cfg.num_layers if num_layers is None else num_layers):
        x = block_forward(sd, i, cfg, x, txt, tproj, freqs, stream_dtype)
    shift, scale = (sd["scale_shift_table"] + temb[:, None]).chunk(2, dim=1)
    x = _ln(x, eps=cfg.eps) * (1 + scale) + shift
    x = F.linear(x, sd["proj_out.weight"], sd["proj_out.bias"])
    x = x.reshape(B, f, h, w, pt, ph, pw, -1).permute(0, 7, 1, 4, 2, 5, 3, 6)
    return x.flatten(6, 7).flatten(4, 5).flatten(2, 3)


@torch.no_grad()
def single_block(sd, cfg: WanConfig, hidden_states, timestep, encoder_hidden_states, layer: int = 0):
    """BASELINE.json configs[0]: one WanTransformerBlock on the patch-embedded latent (plumbing case)."""
    sd = {k: v.float() for k, v in sd.items()}
    B, C, T, H, W = hidden_states.shape
    pt, ph, pw = cfg.patch_size
    freqs = rope_freqs(cfg, T // pt, H // ph, W // pw)
    x = F.conv3d(hidden_states.float(), sd["patch_embedding.weight"], sd["patch_embedding.bias"], stride=cfg.patch_size)
    x = x.flatten(2).transpose(1, 2)
    _, tproj, txt = condition_embed(sd, cfg, timestep, encoder_hidden_states)
    return block_forward(sd, layer, cfg, x, txt, tproj, freqs)


def synthetic_inputs(cfg: WanConfig, batch: int = 1, frames: int = 4, hw: int = 64, text_len: int = 512,
                     text_valid: Optional[int] = None, seed: int = 0):
    """SURVEY §8d inputs: latent randn -> bf16, text randn -> bf16 with rows >= text_valid zeroed."""
    g = torch.Generator().manual_seed(seed)
    lat = torch.randn(batch, cfg.in_channels, frames, hw, hw, generator=g).bfloat16()
    txt = torch.randn(batch, text_len, cfg.text_dim, generator=g).bfloat16()
    if text_valid is not None:
        txt[:, text_valid:] = 0
    return lat, txt
