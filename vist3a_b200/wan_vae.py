"""Wan-2.1 VAE decode on the sm_100a kernels -- drop-in for `pipe.vae.decode(latents, return_dict=False)[0]`
(/root/reference/inference_t23d.py:104-114; arithmetic vendored at utils/wan_utils.py:745-901 WanDecoder3d, :1078-1117
AutoencoderKLWan._decode), the step between the denoiser and the stitched decoder whose frames become `feedforward_image`.

The reference walks the clip one latent frame at a time and carries two cached frames per causal convolution; that procedure is a
whole-clip computation (oracle/wan_vae_ref.py states and pins it), and the device path runs it that way: ONE implicit-GEMM convolution
per layer over all frames.

Data layout in HBM: activations NDHWC bf16, one clip [T, H, W, ld]; ld = channels rounded up to 64 (the 96-channel layers at 512x512 are
stored 128 wide so that a convolution tap is a whole number of 64-channel k-blocks of the GEMM's TMA producer; padding channels of every
convolution operand are zero and meet zero weights).  Weights are re-laid-out once at load: [C_out, taps * ld] tap-major bf16
(`wan_vae_layout.conv3d_weight_to_taps`), the 3x3 convolution behind each nearest-2x up-sampling as a parity-decomposed [4 C_out, 9 C_in]
matrix over the LOW-resolution map (`upsample_conv_weight_to_parity`: the up-sampled tensor is never written).
Kernels: vist3a_gemm conv mode with `kt` temporal taps (zero padding and the causal front padding are TMA out-of-bounds fill),
vist3a_vae_rmsnorm (+SiLU), vist3a_softmax_rows / vist3a_transpose_bf16 (mid-block attention as two GEMMs), vist3a_time_interleave,
vist3a_depth_to_space2_bf16, vist3a_latent_to_ndhwc, vist3a_vae_frames_out.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from . import ops
from .wan_vae_layout import conv3d_weight_to_taps, upsample_conv_weight_to_parity

# AutoencoderKLWan constructor defaults (utils/wan_utils.py:916-924)
WAN_VAE_CONFIG = dict(base_dim=96, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temporal_downsample=(False, True, True))


def _ld(c: int) -> int:
    return (c + 63) // 64 * 64


def decoder_layout(cfg) -> Tuple[list, int]:
    """up_blocks as the reference builds them (utils/wan_utils.py:795-822): [(c_in, c_out, upsample mode or None)], top width"""
    dm = tuple(cfg.dim_mult)
    dims = [cfg.base_dim * u for u in (dm[-1],) + dm[::-1]]
    t_up = tuple(cfg.temporal_downsample)[::-1]
    out = []
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        if i > 0:
            cin = cin // 2
        out.append((cin, cout, None if i == len(dm) - 1 else ("up3d" if t_up[i] else "up2d")))
    return out, dims[0]


def random_state_dict(device="cuda", seed=0):
    """random-init weights of the released architecture (shapes of AutoencoderKLWan's decoder half), generated on the device"""
    cfg = SimpleNamespace(**WAN_VAE_CONFIG)
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}

    def conv(name, co, ci, *k):
        fan = ci
        for x in k:
            fan *= x
        sd[name + ".weight"] = torch.randn((co, ci) + k, device=device, generator=g) * fan ** -0.5
        sd[name + ".bias"] = torch.randn(co, device=device, generator=g) * 0.02

    def res(p, ci, co):
        sd[p + ".norm1.gamma"] = torch.ones(ci, 1, 1, 1, device=device)
        conv(p + ".conv1", co, ci, 3, 3, 3)
        sd[p + ".norm2.gamma"] = torch.ones(co, 1, 1, 1, device=device)
        conv(p + ".conv2", co, co, 3, 3, 3)
        if ci != co:
            conv(p + ".conv_shortcut", co, ci, 1, 1, 1)

    z = cfg.z_dim
    ups, c0 = decoder_layout(cfg)
    conv("post_quant_conv", z, z, 1, 1, 1)
    conv("decoder.conv_in", c0, z, 3, 3, 3)
    for r in (0, 1):
        res(f"decoder.mid_block.resnets.{r}", c0, c0)
    a = "decoder.mid_block.attentions.0"
    sd[a + ".norm.gamma"] = torch.ones(c0, 1, 1, device=device)
    conv(a + ".to_qkv", 3 * c0, c0, 1, 1)
    conv(a + ".proj", c0, c0, 1, 1)
    for i, (cin, cout, mode) in enumerate(ups):
        for j in range(cfg.num_res_blocks + 1):
            res(f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
        if mode is not None:
            p = f"decoder.up_blocks.{i}.upsamplers.0"
            conv(p + ".resample.1", cout // 2, cout, 3, 3)
            if mode == "up3d":
                conv(p + ".time_conv", 2 * cout, cout, 3, 1, 1)
    sd["decoder.norm_out.gamma"] = torch.ones(ups[-1][1], 1, 1, 1, device=device)
    conv("decoder.conv_out", 3, ups[-1][1], 3, 3, 3)
    return sd


class WanVAEDecoderB200(torch.nn.Module):
    """B200-native decode half of AutoencoderKLWan (inference)."""

    def __init__(self, config=None, device="cuda"):
        super().__init__()
        cfg = dict(WAN_VAE_CONFIG)
        if config is not None:
            cfg.update(config if isinstance(config, dict) else {k: getattr(config, k) for k in WAN_VAE_CONFIG if hasattr(config, k)})
        self.config = SimpleNamespace(**cfg)
        self._dev = torch.device(device)
        self.w: Dict[str, torch.Tensor] = {}

    @property
    def device(self):
        return self._dev

    @property
    def dtype(self):
        return torch.bfloat16

    def eval(self):
        return self

    def _apply(self, fn):
        return self

    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor], config=None, device="cuda"):
        m = cls(config, device)
        m.load_weights(sd)
        return m

    # ------------------------------------------------------------------ weights
    def load_weights(self, sd: Dict[str, torch.Tensor]):
        dev = self._dev
        w: Dict[str, torch.Tensor] = {}
        self._taps: Dict[str, Tuple[int, int, int]] = {}

        def conv(name, n_pad=None):
            wt = sd[name + ".weight"].detach().float()
            b = sd[name + ".bias"].detach().float()
            taps = tuple(wt.shape[2:]) if wt.dim() == 5 else (1,) + tuple(wt.shape[2:])
            self._taps[name] = taps
            if taps == (1, 1, 1):
                m = wt.reshape(wt.shape[0], wt.shape[1])            # plain GEMM over the channels (K = C_in exactly)
            else:
                m = conv3d_weight_to_taps(wt, c_in_pad=_ld(wt.shape[1]))
            if n_pad is not None and n_pad > m.shape[0]:
                m = F.pad(m, (0, 0, 0, n_pad - m.shape[0]))
                b = F.pad(b, (0, n_pad - b.shape[0]))
            w[name + ".w"], w[name + ".b"] = m.to(dev, torch.bfloat16).contiguous(), b.to(dev).contiguous()

        def gamma(name):
            w[name] = sd[name].detach().float().reshape(-1).to(dev).contiguous()

        def res(p):
            gamma(p + ".norm1.gamma")
            conv(p + ".conv1")
            gamma(p + ".norm2.gamma")
            conv(p + ".conv2")
            if p + ".conv_shortcut.weight" in sd:
                conv(p + ".conv_shortcut")

        z = self.config.z_dim
        # post_quant_conv (1x1x1 over the z latent channels) as a GEMM over the 64-wide zero-padded latent rows
        pq = sd["post_quant_conv.weight"].detach().float().reshape(z, z)
        z8 = (z + 7) // 8 * 8                                                # output rows padded to the GEMM's 8-column granularity (zero rows)
        w["post_quant_conv.w"] = F.pad(pq, (0, _ld(z) - z, 0, z8 - z)).to(dev, torch.bfloat16).contiguous()
        w["post_quant_conv.b"] = F.pad(sd["post_quant_conv.bias"].detach().float(), (0, z8 - z)).to(dev).contiguous()
        conv("decoder.conv_in")
        for r in (0, 1):
            res(f"decoder.mid_block.resnets.{r}")
        a = "decoder.mid_block.attentions.0"
        gamma(a + ".norm.gamma")
        for n in ("to_qkv", "proj"):
            wt = sd[f"{a}.{n}.weight"].detach().float()
            w[f"{a}.{n}.w"] = wt.reshape(wt.shape[0], wt.shape[1]).to(dev, torch.bfloat16).contiguous()
            w[f"{a}.{n}.b"] = sd[f"{a}.{n}.bias"].detach().float().to(dev).contiguous()
        ups, _ = decoder_layout(self.config)
        for i, (_cin, cout, mode) in enumerate(ups):
            for j in range(self.config.num_res_blocks + 1):
                res(f"decoder.up_blocks.{i}.resnets.{j}")
            if mode is not None:
                p = f"decoder.up_blocks.{i}.upsamplers.0"
                wt, bt = upsample_conv_weight_to_parity(sd[p + ".resample.1.weight"].detach().float(), sd[p + ".resample.1.bias"].detach().float())
                if cout % 64:   # activation rows are _ld(cout) wide: zero weights meet the padding channels of every tap
                    wt = F.pad(wt.reshape(wt.shape[0], 9, cout), (0, _ld(cout) - cout)).reshape(wt.shape[0], 9 * _ld(cout))
                w[p + ".resample.w"], w[p + ".resample.b"] = wt.to(dev, torch.bfloat16).contiguous(), bt.to(dev).contiguous()
                if mode == "up3d":
                    conv(p + ".time_conv")
        gamma("decoder.norm_out.gamma")
        conv("decoder.conv_out", n_pad=8)    # 3 output channels: the fp32 result rows are padded to 8 (16-byte row pieces)
        self.w = w

    def weight_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.w.values())

    # ------------------------------------------------------------------ layers
    def _conv(self, x: torch.Tensor, name: str, c_in: int, *, residual=None, out_dtype=torch.bfloat16) -> torch.Tensor:
        """x [T, H, W, ld_in] bf16 (channels [0, c_in) valid, padding zero when the convolution has taps) -> [T, H, W, ld(C_out)]"""
        wt, b = self.w[name + ".w"], self.w[name + ".b"]
        kt, kh, kw = self._taps[name]
        T, H, W, ld = x.shape
        N = wt.shape[0]
        ldo = _ld(N) if out_dtype == torch.bfloat16 else N
        out = torch.empty((T, H, W, ldo), dtype=out_dtype, device=x.device)
        o2 = out.view(-1, ldo)[:, :N]
        r2 = None if residual is None else residual.view(-1, residual.shape[-1])[:, :N]
        if (kt, kh, kw) == (1, 1, 1):
            ops.gemm(x.view(-1, ld)[:, :c_in], wt, b, out=o2, residual=r2)
        else:
            ops.gemm(x, wt, b, conv=dict(kh=kh, kw=kw, pad=kh // 2, kt=kt), out=o2, residual=r2)
        return out

    def _res_block(self, x: torch.Tensor, p: str, cin: int, cout: int) -> torch.Tensor:
        """WanResidualBlock (utils/wan_utils.py:333-425): norm-SiLU-conv, norm-SiLU-conv, + (1x1x1 shortcut or identity)"""
        w = self.w
        h = self._conv(x, p + ".conv_shortcut", cin) if (p + ".conv_shortcut.w") in w else x
        y = self._conv(ops.vae_rmsnorm(x, w[p + ".norm1.gamma"], cin), p + ".conv1", cin)
        return self._conv(ops.vae_rmsnorm(y, w[p + ".norm2.gamma"], cout), p + ".conv2", cout, residual=h)

    def _attention(self, x: torch.Tensor, p: str, C: int) -> torch.Tensor:
        """WanAttentionBlock (:428-475): per frame, one head of width C over the H*W positions: logits GEMM (fp32) -> row softmax (bf16) ->
        P V GEMM against the transposed values, then the 1x1 projection + residual"""
        w = self.w
        T, H, W, ld = x.shape
        HW = H * W
        n = ops.vae_rmsnorm(x, w[p + ".norm.gamma"], C, silu=False)
        # ragged H*W (not a multiple of 8: the GEMM's N / 16-byte row granularity): keys and values are read HW8 >= HW rows deep -- the rows
        # behind a frame belong to the next frame, behind the last frame to 8 zero rows -- and the softmax masks the padding columns (P = 0)
        HW8 = (HW + 7) // 8 * 8
        qkv = (torch.empty if HW8 == HW else torch.zeros)((T * HW + HW8 - HW, 3 * C), dtype=torch.bfloat16, device=x.device)
        ops.gemm(n.view(-1, ld)[:, :C], w[p + ".to_qkv.w"], w[p + ".to_qkv.b"], out=qkv[:T * HW])  # [T*HW, 3C] bf16
        att = torch.empty((T * HW, C), dtype=torch.bfloat16, device=x.device)
        logits = torch.empty((HW, HW8), dtype=torch.float32, device=x.device)
        probs = torch.empty((HW, HW8), dtype=torch.bfloat16, device=x.device)
        for t in range(T):
            f = qkv[t * HW:t * HW + HW8]
            ops.gemm(f[:HW, :C], f[:, C:2 * C], out=logits)                                      # q k^T
            ops.softmax_rows(logits, C ** -0.5, out=probs, valid=HW)
            ops.gemm(probs, ops.transpose_bf16(f[:, 2 * C:]), out=att[t * HW:(t + 1) * HW])      # P v
        out = torch.empty_like(x)
        ops.gemm(att, w[p + ".proj.w"], w[p + ".proj.b"], residual=x.view(-1, ld)[:, :C], out=out.view(-1, ld)[:, :C])
        return out

    def _upsample(self, x: torch.Tensor, p: str, C: int, temporal: bool) -> torch.Tensor:
        """WanResample upsample2d / upsample3d (:202-308): [time_conv on frames 1.. (frame 0 passes through) + channel halves interleaved in
        time,] nearest 2x + 3x3 conv as ONE low-resolution GEMM with N = 4 * C/2 output columns + depth-to-space"""
        w = self.w
        T, H, W, ld = x.shape
        if temporal and T > 1:
            y = self._conv(x[1:], p + ".time_conv", C)                                          # [T-1, H, W, 2C]
            xn = (torch.empty if ld == C else torch.zeros)((2 * T - 1, H, W, ld), dtype=torch.bfloat16, device=x.device)   # padding channels: finite (they meet zero weights)
            xn[0].copy_(x[0])
            ops.time_interleave(y, xn[1:], C)
            x, T = xn, 2 * T - 1
        wt, b = w[p + ".resample.w"], w[p + ".resample.b"]
        co = wt.shape[0] // 4
        y = torch.empty((T * H * W, 4 * co), dtype=torch.bfloat16, device=x.device)
        ops.gemm(x, wt, b, conv=dict(kh=3, kw=3, pad=1), out=y)
        return ops.depth_to_space2_bf16(y, T, H, W, co, _ld(co))

    # ------------------------------------------------------------------ decode
    @torch.no_grad()
    def decode_clip(self, z: torch.Tensor) -> torch.Tensor:
        """z [z_dim, T', h, w] (de-normalised latent) -> frames [3, 1 + 4 (T' - 1), 8h, 8w] fp32 in [-1, 1]"""
        if not self.w:
            raise RuntimeError("weights not loaded: use WanVAEDecoderB200.from_state_dict(...)")
        cfg, w = self.config, self.w
        zc = cfg.z_dim
        x = ops.latent_to_ndhwc(z.to(self._dev), _ld(zc))                                        # [T', h, w, 64]
        T, h, wd, ldz = x.shape
        pq = torch.zeros((T, h, wd, ldz), dtype=torch.bfloat16, device=self._dev)                # padding channels stay zero (conv_in operand)
        ops.gemm(x.view(-1, ldz), w["post_quant_conv.w"], w["post_quant_conv.b"], out=pq.view(-1, ldz)[:, :w["post_quant_conv.w"].shape[0]])
        ups, c0 = decoder_layout(cfg)
        x = self._conv(pq, "decoder.conv_in", zc)
        x = self._res_block(x, "decoder.mid_block.resnets.0", c0, c0)
        x = self._attention(x, "decoder.mid_block.attentions.0", c0)
        x = self._res_block(x, "decoder.mid_block.resnets.1", c0, c0)
        c = c0
        for i, (cin, cout, mode) in enumerate(ups):
            for j in range(cfg.num_res_blocks + 1):
                x = self._res_block(x, f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
            c = cout
            if mode is not None:
                x = self._upsample(x, f"decoder.up_blocks.{i}.upsamplers.0", cout, temporal=(mode == "up3d"))
                c = cout // 2
        y = self._conv(ops.vae_rmsnorm(x, w["decoder.norm_out.gamma"], c), "decoder.conv_out", c, out_dtype=torch.float32)
        T, H, W, _ = x.shape
        return ops.vae_frames_out(y.view(T * H * W, -1), T, H, W)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True):
        """diffusers' `AutoencoderKLWan.decode(latents, return_dict=False)[0]` call surface: z [B, z_dim, T', h, w] -> [B, 3, T, 8h, 8w]"""
        if z.dim() != 5 or z.shape[1] != self.config.z_dim:
            raise ValueError(f"decode: expected latents [B, {self.config.z_dim}, T, h, w], got {tuple(z.shape)}")
        out = torch.stack([self.decode_clip(z[b]) for b in range(z.shape[0])], 0)
        if return_dict:
            return SimpleNamespace(sample=out)
        return (out,)
