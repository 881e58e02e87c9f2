"""Wan-2.1 DiT denoiser on the sm_100a kernels -- drop-in for `pipe.transformer`.

Mirrors the call surface of diffusers' `WanTransformer3DModel` as the reference uses it
(/root/reference/inference_t23d.py:73-80,94-103 through WanPipeline; direct calls at
/root/reference/train_vdm.py:557-562,598-603):

    out, = model(hidden_states[B,16,T,H,W], timestep[B], encoder_hidden_states[B,Lt,4096], return_dict=False)

plus `.config`, `.dtype`, `.device`, `.eval()`.  Weights are ingested from a diffusers-keyed state
dict (`from_state_dict`); a PEFT LoRA adapter (train_vdm.py:370-388: r=8, alpha=16 on the eight
attention projections) is folded into the base weights at load, so no rank-r GEMMs run per step.

Data layout in HBM: the residual stream is one bf16 [B*L, D] matrix (token order t,h,w as
`flatten(2).transpose(1,2)` gives); q/k/v of self-attention live in one fused [B*L, 3D] buffer
written by a single N=3D GEMM and are consumed in place by RMSNorm+RoPE and attention (strided
views, no transposes); text K/V of all layers are computed once per prompt and cached.
dtype policy = the reference's CUDA-autocast trace: bf16 GEMM operands and
residual stream, fp32 LayerNorm / RMSNorm statistics, fp32 AdaLN vectors and gated residual adds,
bf16 rounding of each Linear output before gating.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import ops


# WanTransformer3DModel.config values of the two released sizes (diffusers 0.33.1)
WAN_1_3B_CONFIG = dict(patch_size=(1, 2, 2), num_attention_heads=12, attention_head_dim=128, in_channels=16, out_channels=16,
                       text_dim=4096, freq_dim=256, ffn_dim=8960, num_layers=30, cross_attn_norm=True, eps=1e-6,
                       rope_max_seq_len=1024)
WAN_14B_CONFIG = dict(WAN_1_3B_CONFIG, num_attention_heads=40, ffn_dim=13824, num_layers=40)


def param_shapes(cfg: dict) -> Dict[str, tuple]:
    """diffusers-keyed parameter manifest of WanTransformer3DModel (what `from_state_dict` ingests)."""
    D, Fd = cfg["num_attention_heads"] * cfg["attention_head_dim"], cfg["ffn_dim"]
    pt, ph, pw = cfg["patch_size"]
    po = cfg["out_channels"] * pt * ph * pw
    s: Dict[str, tuple] = {"patch_embedding.weight": (D, cfg["in_channels"], pt, ph, pw), "patch_embedding.bias": (D,),
                           "scale_shift_table": (1, 2, D), "proj_out.weight": (po, D), "proj_out.bias": (po,)}

    def lin(name, n, k):
        s[name + ".weight"], s[name + ".bias"] = (n, k), (n,)

    ce = "condition_embedder."
    lin(ce + "time_embedder.linear_1", D, cfg["freq_dim"])
    lin(ce + "time_embedder.linear_2", D, D)
    lin(ce + "time_proj", 6 * D, D)
    lin(ce + "text_embedder.linear_1", D, cfg["text_dim"])
    lin(ce + "text_embedder.linear_2", D, D)
    for i in range(cfg["num_layers"]):
        b = f"blocks.{i}."
        s[b + "scale_shift_table"] = (1, 6, D)
        for a in ("attn1.", "attn2."):
            for l in ("to_q", "to_k", "to_v", "to_out.0"):
                lin(b + a + l, D, D)
            s[b + a + "norm_q.weight"], s[b + a + "norm_k.weight"] = (D,), (D,)
        if cfg["cross_attn_norm"]:
            s[b + "norm2.weight"], s[b + "norm2.bias"] = (D,), (D,)
        lin(b + "ffn.net.0.proj", Fd, D)
        lin(b + "ffn.net.2", D, Fd)
    return s


def random_state_dict(cfg: dict, seed: int = 0, device="cuda") -> Dict[str, torch.Tensor]:
    """Random-init weights of the named architecture for benchmarking (no checkpoint is reachable offline): Linear/Conv
    N(0, 0.02) in bf16, zero biases, scale_shift_table = randn / sqrt(D), RMSNorm/LayerNorm weights 1."""
    g = torch.Generator(device=device).manual_seed(seed)
    D = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    sd = {}
    for k, shp in param_shapes(cfg).items():
        if k.endswith("scale_shift_table"):
            sd[k] = torch.randn(shp, device=device, generator=g) / math.sqrt(D)
        elif "norm" in k and k.endswith(".weight"):
            sd[k] = torch.ones(shp, device=device)
        elif k.endswith(".bias"):
            sd[k] = torch.zeros(shp, device=device)
        else:
            sd[k] = (torch.randn(shp, device=device, generator=g) * 0.02).bfloat16()
    return sd


def _rope_tables(head_dim: int, f: int, h: int, w: int, max_len: int, device) -> tuple:
    """cos/sin [f*h*w, head_dim/2] fp32 of WanRotaryPosEmbed (fp64 angles; split 44/42/42 at d=128)."""
    h_dim = w_dim = 2 * (head_dim // 6)
    t_dim = head_dim - h_dim - w_dim
    parts = []
    for dim, n, shape in ((t_dim, f, (f, 1, 1)), (h_dim, h, (1, h, 1)), (w_dim, w, (1, 1, w))):
        inv = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float64)[: dim // 2] / dim))
        ang = torch.outer(torch.arange(n, dtype=torch.float64), inv)  # [n, dim/2]
        parts.append(ang.view(*shape, -1).expand(f, h, w, -1))
    ang = torch.cat(parts, dim=-1).reshape(f * h * w, head_dim // 2)
    return ang.cos().float().contiguous().to(device), ang.sin().float().contiguous().to(device)


class WanTransformer3DModelB200(torch.nn.Module):
    """B200-native WanTransformer3DModel (inference only)."""

    def __init__(self, config, device="cuda"):
        super().__init__()
        cfg = config if isinstance(config, dict) else {k: getattr(config, k) for k in (
            "patch_size", "num_attention_heads", "attention_head_dim", "in_channels", "out_channels", "text_dim",
            "freq_dim", "ffn_dim", "num_layers", "cross_attn_norm", "eps", "rope_max_seq_len")}
        self.config = SimpleNamespace(**cfg)
        if tuple(self.config.patch_size) != (1, 2, 2):
            raise NotImplementedError(f"patch_size {self.config.patch_size}: the patchify kernel implements Wan's (1,2,2)")
        if self.config.attention_head_dim not in (64, 128):
            raise NotImplementedError("attention_head_dim must be 64 or 128")
        self._dev = torch.device(device)
        self.w: Dict[str, torch.Tensor] = {}
        self._rope_cache = {}
        self._text_cache = None
        self._ws = {}

    # ------------------------------------------------------------------ module-like surface
    @property
    def dtype(self):
        return torch.bfloat16

    @property
    def device(self):
        return self._dev

    def eval(self):
        return self

    def _apply(self, fn):  # .to()/.cuda() on the owning pipeline must not move or re-type our buffers
        return self

    # ------------------------------------------------------------------ weights
    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor], config, *, lora: Optional[Dict[str, torch.Tensor]] = None,
                        lora_alpha: float = 16.0, lora_r: Optional[int] = None, device="cuda"):
        m = cls(config, device)
        m.load_weights(sd, lora=lora, lora_alpha=lora_alpha, lora_r=lora_r)
        return m

    def load_weights(self, sd, *, lora=None, lora_alpha=16.0, lora_r=None):
        c, dev = self.config, self._dev
        D = c.num_attention_heads * c.attention_head_dim

        def lin_w(name):
            wt = sd[name + ".weight"].to(torch.float32)
            if lora is not None:
                for pre in ("base_model.model.", ""):
                    for mid in (".lora_A.weight", ".lora_A.default.weight"):
                        ka = pre + name + mid
                        if ka in lora:
                            A = lora[ka].float()
                            Bm = lora[ka.replace("lora_A", "lora_B")].float()
                            r = lora_r or A.shape[0]
                            wt = wt + (lora_alpha / r) * (Bm @ A)
            return wt

        def W(*names):  # bf16 [sum N, K] on device
            return torch.cat([lin_w(n) for n in names], 0).to(dev, torch.bfloat16).contiguous()

        def Bv(*names):  # fp32 [sum N]
            return torch.cat([sd[n + ".bias"].float() for n in names], 0).to(dev).contiguous()

        def V32(name):
            return sd[name].float().to(dev).contiguous()

        w = {}
        w["patch.w"] = sd["patch_embedding.weight"].float().reshape(D, -1).to(dev, torch.bfloat16).contiguous()
        w["patch.b"] = V32("patch_embedding.bias")
        p = "condition_embedder."
        for short, full in (("t1", "time_embedder.linear_1"), ("t2", "time_embedder.linear_2"), ("tp", "time_proj"),
                            ("x1", "text_embedder.linear_1"), ("x2", "text_embedder.linear_2")):
            w[short + ".w"], w[short + ".b"] = W(p + full), Bv(p + full)
        w["out.table"] = V32("scale_shift_table").reshape(2, D).contiguous()
        w["out.w"], w["out.b"] = W("proj_out"), Bv("proj_out")
        # the AdaLN tables of all blocks in one [layers, 6, D] tensor: one modulation launch per batch row covers every block of a forward
        w["tables"] = torch.stack([V32(f"blocks.{i}.scale_shift_table").reshape(6, D) for i in range(c.num_layers)]).contiguous()
        for i in range(c.num_layers):
            b = f"blocks.{i}."
            k = f"b{i}."
            w[k + "table"] = w["tables"][i]
            w[k + "qkv.w"] = W(b + "attn1.to_q", b + "attn1.to_k", b + "attn1.to_v")
            w[k + "qkv.b"] = Bv(b + "attn1.to_q", b + "attn1.to_k", b + "attn1.to_v")
            w[k + "nqk1"] = torch.cat([V32(b + "attn1.norm_q.weight"), V32(b + "attn1.norm_k.weight")]).contiguous()
            w[k + "o1.w"], w[k + "o1.b"] = W(b + "attn1.to_out.0"), Bv(b + "attn1.to_out.0")
            w[k + "q2.w"], w[k + "q2.b"] = W(b + "attn2.to_q"), Bv(b + "attn2.to_q")
            w[k + "kv2.w"] = W(b + "attn2.to_k", b + "attn2.to_v")
            w[k + "kv2.b"] = Bv(b + "attn2.to_k", b + "attn2.to_v")
            # cross-attention: (q rinv_q w_q) . (k rinv_k w_k) = rinv_q * q . (k rinv_k (w_k w_q)): the query norm weight is folded
            # into the cached text keys, the per-row factor rinv_q scales the logits inside the attention kernel
            w[k + "nk2q2"] = (V32(b + "attn2.norm_k.weight") * V32(b + "attn2.norm_q.weight")).contiguous()
            w[k + "o2.w"], w[k + "o2.b"] = W(b + "attn2.to_out.0"), Bv(b + "attn2.to_out.0")
            if c.cross_attn_norm:
                w[k + "n2.w"], w[k + "n2.b"] = V32(b + "norm2.weight"), V32(b + "norm2.bias")
            w[k + "f1.w"], w[k + "f1.b"] = W(b + "ffn.net.0.proj"), Bv(b + "ffn.net.0.proj")
            w[k + "f2.w"], w[k + "f2.b"] = W(b + "ffn.net.2"), Bv(b + "ffn.net.2")
        self.w = w
        self._text_cache = None

    def weight_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.w.values())

    # ------------------------------------------------------------------ per-prompt text state
    def encode_text(self, encoder_hidden_states: torch.Tensor):
        """Text MLP (4096 -> D -> D, GELU-tanh) and the K/V projections (+RMSNorm on K) of every
        cross-attention layer: constant across the 50 steps of one prompt."""
        c, w = self.config, self.w
        B, Lt, _ = encoder_hidden_states.shape
        D = c.num_attention_heads * c.attention_head_dim
        x = encoder_hidden_states.to(self._dev, torch.bfloat16).reshape(B * Lt, -1)
        t1 = ops.gemm(x, w["x1.w"], w["x1.b"], act="gelu_tanh")
        txt = ops.gemm(t1, w["x2.w"], w["x2.b"])
        kv = torch.empty((c.num_layers, B * Lt, 2 * D), dtype=torch.bfloat16, device=self._dev)
        for i in range(c.num_layers):
            k = f"b{i}."
            ops.gemm(txt, w[k + "kv2.w"], w[k + "kv2.b"], out=kv[i])
            ops.rmsnorm_rope_(kv[i][:, :D], w[k + "nk2q2"], c.attention_head_dim, eps=c.eps)
        return SimpleNamespace(kv=kv, B=B, Lt=Lt)

    TEXT_CACHE_SLOTS = 4  # cond + uncond of the current prompt stay cached while WanPipeline alternates them (2 calls per step)

    def _text_state(self, enc: torch.Tensor):
        """Per-prompt text state, cached on the IDENTITY of the embedding tensor (the entry keeps a reference to it, so the caching
        allocator cannot hand its address to another prompt's embeddings) plus its in-place version counter; least recently used of
        TEXT_CACHE_SLOTS entries is dropped."""
        cache = self._text_cache if self._text_cache is not None else []
        for n, (t, ver, st) in enumerate(cache):
            if t is enc and ver == enc._version:
                if n:
                    cache.insert(0, cache.pop(n))
                return st
        st = self.encode_text(enc)
        cache.insert(0, (enc, enc._version, st))
        del cache[self.TEXT_CACHE_SLOTS:]
        self._text_cache = cache
        return st

    # ------------------------------------------------------------------ forward
    def _workspace(self, B, L):
        key = (B, L)
        ws = self._ws.get(key)
        if ws is None:
            c = self.config
            D = c.num_attention_heads * c.attention_head_dim
            M, dev, bf = B * L, self._dev, torch.bfloat16
            ws = SimpleNamespace(
                x=torch.empty((M, D), dtype=bf, device=dev), h=torch.empty((M, D), dtype=bf, device=dev),
                qkv=torch.empty((M, 3 * D), dtype=bf, device=dev), q2=torch.empty((M, D), dtype=bf, device=dev),
                att=torch.empty((M, D), dtype=bf, device=dev), ffn=torch.empty((M, c.ffn_dim), dtype=bf, device=dev),
                mod=torch.empty((B, c.num_layers, 6, D), dtype=torch.float32, device=dev),
                mod2=torch.empty((B, 2, D), dtype=torch.float32, device=dev),
                po=torch.empty((M, self.w["out.w"].shape[0]), dtype=bf, device=dev),
                rq=torch.empty((M,), dtype=torch.float32, device=dev))
            self._ws[key] = ws  # every shape stays alive: a captured CUDA graph (DenoiseEngine) holds the pointers of its workspace
        return ws

    @torch.no_grad()
    def forward(self, hidden_states, timestep, encoder_hidden_states, encoder_hidden_states_image=None,
                return_dict: bool = False, attention_kwargs=None, text_state=None):
        if encoder_hidden_states_image is not None:
            raise NotImplementedError("image-conditioned Wan (I2V) is outside the VIST3A text-to-3D path")
        if not self.w:
            raise RuntimeError("weights not loaded: use WanTransformer3DModelB200.from_state_dict(...)")
        c, w = self.config, self.w
        H_, hd = c.num_attention_heads, c.attention_head_dim
        D = H_ * hd
        B, C_, T, Hh, Ww = hidden_states.shape
        f, h, wd = T, Hh // 2, Ww // 2
        L = f * h * wd
        M = B * L
        in_dtype = hidden_states.dtype
        if hidden_states.dtype not in (torch.bfloat16, torch.float32):
            hidden_states = hidden_states.to(torch.bfloat16)
        hidden_states = hidden_states.to(self._dev)
        ts = text_state if text_state is not None else self._text_state(encoder_hidden_states)
        if ts.B != B:
            raise ValueError(f"encoder_hidden_states batch {ts.B} != hidden_states batch {B}")
        rk = (f, h, wd)
        if rk not in self._rope_cache:
            self._rope_cache[rk] = _rope_tables(hd, f, h, wd, c.rope_max_seq_len, self._dev)
        cos, sin = self._rope_cache[rk]
        ws = self._workspace(B, L)

        # condition embedder (M = B rows: weight-streaming kernels)
        tstep = torch.as_tensor(timestep).to(self._dev).reshape(-1)
        if tstep.numel() == 1 and B > 1:
            tstep = tstep.expand(B)
        tfeat = ops.timestep_features(tstep, c.freq_dim)
        t1 = ops.skinny_linear(tfeat, w["t1.w"], w["t1.b"], act="silu", out_dtype=torch.float32)
        temb = ops.skinny_linear(t1, w["t2.w"], w["t2.b"], out_dtype=torch.float32)
        tproj = ops.skinny_linear(temb, w["tp.w"], w["tp.b"], pre_act="silu", out_dtype=torch.float32)  # [B, 6D]

        # patch embedding: gather 2x2 patches -> K=64 GEMM (+bias)
        a = ops.patchify(hidden_states)
        x = ops.gemm(a, w["patch.w"], w["patch.b"], out=ws.x)

        # AdaLN vectors of every block: mod[b, i] = [shift1, 1+scale1, gate1, shift2, 1+scale2, gate2] = table_i + timestep_proj[b].  The
        # kernel adds a [6, D] "table" to a batch of [6, D] vectors: here the batch runs over the blocks and the timestep projection of row b
        # plays the table -- B launches per forward instead of one per block (30 kernel boundaries less on the step)
        NL = c.num_layers
        for b_ in range(B):
            ops.modulation(tproj[b_].view(6, D), w["tables"].view(NL, 6 * D), nvec=6, broadcast=False, one_plus_mask=0b010010, out=ws.mod[b_])
        mbs = NL * 6 * D   # batch stride of the per-block views below

        for i in range(c.num_layers):
            k = f"b{i}."
            mod = ws.mod[:, i]
            # --- self-attention
            ops.layernorm(x, mul=mod[:, 1], add=mod[:, 0], mul_bstride=mbs, add_bstride=mbs, rows_per_batch=L,
                          eps=c.eps, out=ws.h)
            # (multicast=True: clusters of two CTA pairs share the activation rows by TMA multicast where the tile grid allows it -- slower in
            #  isolation, -0.8 % on the power-capped step: 27.93 -> 27.69 ms in an alternating same-box A/B, tools/runs/gpu_r3n.sh)
            ops.gemm(ws.h, w[k + "qkv.w"], w[k + "qkv.b"], out=ws.qkv, multicast=True)
            ops.rmsnorm_rope_(ws.qkv[:, :2 * D], w[k + "nqk1"], hd, eps=c.eps, cos=cos, sin=sin, nseg=2)
            qkv5 = ws.qkv.view(B, L, 3, H_, hd)
            ops.fmha(qkv5[:, :, 0], qkv5[:, :, 1], qkv5[:, :, 2], out=ws.att.view(B, L, H_, hd))
            ops.gemm(ws.att, w[k + "o1.w"], w[k + "o1.b"], gate=mod[:, 2], gate_bstride=mbs, rows_per_batch=L,
                     residual=x, out=x, round_linear=True, multicast=True)
            # --- cross-attention
            if c.cross_attn_norm:
                ops.layernorm(x, mul=w[k + "n2.w"], add=w[k + "n2.b"], eps=c.eps, out=ws.h)
                hq = ws.h
            else:
                hq = x
            ops.gemm(hq, w[k + "q2.w"], w[k + "q2.b"], out=ws.q2, multicast=True)
            ops.row_rinv(ws.q2, eps=c.eps, out=ws.rq)
            kv5 = ts.kv[i].view(B, ts.Lt, 2, H_, hd)
            ops.fmha(ws.q2.view(B, L, H_, hd), kv5[:, :, 0], kv5[:, :, 1], out=ws.att.view(B, L, H_, hd), q_row_scale=ws.rq)
            ops.gemm(ws.att, w[k + "o2.w"], w[k + "o2.b"], residual=x, out=x, round_linear=True, multicast=True)
            # --- feed-forward
            ops.layernorm(x, mul=mod[:, 4], add=mod[:, 3], mul_bstride=mbs, add_bstride=mbs, rows_per_batch=L,
                          eps=c.eps, out=ws.h)
            ops.gemm(ws.h, w[k + "f1.w"], w[k + "f1.b"], act="gelu_tanh", out=ws.ffn, multicast=True)
            ops.gemm(ws.ffn, w[k + "f2.w"], w[k + "f2.b"], gate=mod[:, 5], gate_bstride=mbs, rows_per_batch=L,
                     residual=x, out=x, round_linear=True, multicast=True)

        # output head: mod2 = [shift, 1+scale] from table + temb
        ops.modulation(w["out.table"], temb, nvec=2, broadcast=True, one_plus_mask=0b10, out=ws.mod2)
        ops.layernorm(x, mul=ws.mod2[:, 1], add=ws.mod2[:, 0], mul_bstride=2 * D, add_bstride=2 * D, rows_per_batch=L,
                      eps=c.eps, out=ws.h)
        ops.gemm(ws.h, w["out.w"], w["out.b"], out=ws.po)
        out_dtype = in_dtype if in_dtype in (torch.bfloat16, torch.float32) else torch.bfloat16
        out = ops.unpatchify(ws.po, B, c.out_channels, T, Hh, Ww, out_dtype=out_dtype)
        if out.dtype != in_dtype:
            out = out.to(in_dtype)
        if return_dict:
            return SimpleNamespace(sample=out)
        return (out,)
