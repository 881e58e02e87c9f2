"""Tensor-level wrappers of the C-ABI kernels.  PyTorch is used for device memory and streams only:
every wrapper validates layout, passes raw device pointers + the current CUDA stream to
libvist3a_sm100.so and returns the (pre-allocated or new) output tensor.  No op has a fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib as L

_DT = {torch.bfloat16: L.DTYPE_BF16, torch.float32: L.DTYPE_F32}
ACT = {None: L.ACT_NONE, "none": L.ACT_NONE, "gelu_tanh": L.ACT_GELU_TANH, "gelu": L.ACT_GELU_ERF,
       "gelu_erf": L.ACT_GELU_ERF, "silu": L.ACT_SILU, "relu": L.ACT_RELU}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}; the kernels take bfloat16 or float32") from None


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vist3a_b200 ops run on a CUDA (sm_100a) device only; got a CPU tensor")


def _rows2d(t: torch.Tensor) -> torch.Tensor:
    """View as [rows, cols] with unit inner stride (no copy unless the layout forces one)."""
    if t.dim() != 2:
        t = t.reshape(-1, t.shape[-1])
    if t.stride(-1) != 1:
        t = t.contiguous()
    return t


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, act=None,
         out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None,
         gate: Optional[torch.Tensor] = None, gate_bstride: int = 0, rows_per_batch: int = 0,
         residual: Optional[torch.Tensor] = None, round_linear: bool = False, round_gate: bool = False,
         two_cta: Optional[bool] = None, residual2: Optional[torch.Tensor] = None, post_act=None,
         cmap: Optional[tuple] = None, rmap: Optional[tuple] = None, conv: Optional[dict] = None, bn176: bool = False, multicast: bool = False,
         staged: bool = False) -> torch.Tensor:
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T); see vist3a_gemm in include/vist3a_sm100.h.
    cmap / rmap = (rows_per_group, group_stride, group_offset) row maps of out(+residual2) / residual;
    conv = dict(kh, kw, pad) with `a` an NHWC [n, h, w, c] tensor: implicit-GEMM convolution (stride 1).  Optional conv keys
    pad_x, geom=(n, h, w, c_in) and strides=(pixel, row, image) in elements describe overlapping windows over a physically
    padded image (see vist3a_conv in the header)."""
    _need_cuda(a, w, bias, gate, residual, residual2, out)
    w2 = _rows2d(w)
    N, K2 = w2.shape
    if conv is not None:
        if a.dim() != 4 or not a.is_contiguous():
            raise ValueError("gemm(conv): a must be a contiguous NHWC [n, h, w, c] tensor")
        n_img, h_in, w_in, c_in = conv.get("geom", a.shape)
        kh, kw, pad = conv["kh"], conv["kw"], conv["pad"]
        pad_x = conv.get("pad_x", pad)
        h_out, w_out = h_in + 2 * pad - kh + 1, w_in + 2 * pad_x - kw + 1
        kt = int(conv.get("kt", 1))   # > 1: causal temporal taps over the image (frame) index, one clip per call
        M, K = n_img * h_out * w_out, kt * kh * kw * c_in
        a2 = a
    else:
        a2 = _rows2d(a)
        M, K = a2.shape
    if a2.dtype != w2.dtype:
        raise TypeError(f"gemm: A is {a2.dtype} but W is {w2.dtype}")
    if K != K2:
        raise ValueError(f"gemm: K mismatch {K} vs {K2}")
    if out is None:
        if cmap is not None:
            raise ValueError("gemm: a row-mapped output must be pre-allocated")
        out = torch.empty((M, N), dtype=out_dtype or a2.dtype, device=a.device)
    o2 = out if out.dim() == 2 else out.view(-1, out.shape[-1])
    if o2.stride(-1) != 1 or o2.shape[1] != N or (cmap is None and o2.shape[0] != M):
        raise ValueError("gemm: out must be [M, N] with unit inner stride")
    for v, n in ((bias, "bias"), (gate, "gate")):
        if v is not None and v.dtype != torch.float32:
            raise TypeError(f"gemm: {n} must be float32")
    r2 = None
    if residual is not None:
        r2 = residual if residual.dim() == 2 else residual.view(-1, residual.shape[-1])
        if r2.dtype != o2.dtype or r2.stride(-1) != 1:
            raise TypeError("gemm: residual must have the output dtype and unit inner stride")
        if r2.dim() == 1:
            r2 = r2.view(1, -1)
    args = L.GemmArgs()
    args.A, args.W, args.C = a2.data_ptr(), w2.data_ptr(), o2.data_ptr()
    args.bias, args.gate, args.residual = _ptr(bias), _ptr(gate), _ptr(r2)
    if residual2 is not None:
        q2 = residual2 if residual2.dim() == 2 else residual2.view(-1, residual2.shape[-1])
        if q2.dtype != o2.dtype or q2.stride(-1) != 1 or q2.stride(0) != o2.stride(0):
            raise TypeError("gemm: residual2 must share the output's dtype and row stride")
        args.residual2 = q2.data_ptr()
    args.M, args.N, args.K = M, N, K
    args.lda, args.ldw, args.ldc = (0 if conv is not None else a2.stride(0)), w2.stride(0), o2.stride(0)
    if cmap is not None:
        args.cmap.rpg, args.cmap.gstride, args.cmap.goff = cmap
    if rmap is not None:
        args.rmap.rpg, args.rmap.gstride, args.rmap.goff = rmap
    if conv is not None:
        args.conv.enabled, args.conv.kh, args.conv.kw, args.conv.pad_y, args.conv.pad_x = 1, kh, kw, pad, pad_x
        args.conv.n_img, args.conv.h, args.conv.w, args.conv.c_in = n_img, h_in, w_in, c_in
        args.conv.kt = kt
        if "strides" in conv:
            args.conv.pix_stride, args.conv.row_stride, args.conv.img_stride = conv["strides"]
    args.post_act = ACT[post_act]
    args.ldr = r2.stride(0) if r2 is not None else 0
    args.rows_per_batch = rows_per_batch if rows_per_batch > 0 else M
    args.gate_bstride = gate_bstride
    args.in_dtype, args.out_dtype = _dt(a2), _dt(o2)
    args.act = ACT[act]
    args.round_linear, args.round_gate = int(round_linear), int(round_gate)
    if two_cta is None:
        two_cta = M >= 2048
    args.flags = ((L.GEMM_FLAG_2CTA if two_cta else L.GEMM_FLAG_1CTA) | (L.GEMM_FLAG_BN176 if bn176 else 0) | (L.GEMM_FLAG_MULTICAST if multicast else 0)
                  | (L.GEMM_FLAG_STAGED if staged else 0) | _GEMM_EXTRA_FLAGS)
    L.check(L.load().vist3a_gemm(C.byref(args), _stream()))
    return out


_GEMM_EXTRA_FLAGS = int(os.environ.get("VIST3A_GEMM_FLAGS", "0"), 0)   # A/B switch (tools/): OR-ed into every GEMM call's flags (8 = multicast, 16 = staged epilogue)
# A/B switch for measurements and numerics studies (tools/): flags used when the caller passes none (e.g. 8192 = the former default variants)
_FMHA_DEFAULT_FLAGS = int(os.environ.get("VIST3A_FMHA_FLAGS", "0"), 0)


_FMHA_WS_BYTES: dict = {}
_FMHA_WS: dict = {}


def fmha(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, scale: Optional[float] = None,
         out: Optional[torch.Tensor] = None, flags: int = 0, q_row_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Non-causal attention.  q [B, Lq, H, D], k/v [B, Lkv, H, D] (any strides with unit inner stride,
    e.g. slices of a fused QKV buffer); returns [B, Lq, H, D] bf16.  q_row_scale [B*Lq] fp32 (optional) multiplies the
    logits of each query row (all heads): the per-row RMSNorm factor of cross-attention queries (see vist3a_row_rinv)."""
    _need_cuda(q, k, v, out)
    B, Lq, H, D = q.shape
    Lk = k.shape[1]
    if k.shape != (B, Lk, H, D) or v.shape != (B, Lk, H, D):
        raise ValueError(f"fmha: shape mismatch q {tuple(q.shape)} k {tuple(k.shape)} v {tuple(v.shape)}")
    for t in (q, k, v):
        if t.dtype != torch.bfloat16 or t.stride(-1) != 1:
            raise TypeError("fmha: q, k, v must be bfloat16 with unit inner stride")
    if out is None:
        out = torch.empty((B, Lq, H, D), dtype=torch.bfloat16, device=q.device)
    a = L.FmhaArgs()
    a.Q, a.K, a.V, a.O = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.batch, a.heads, a.len_q, a.len_kv, a.head_dim = B, H, Lq, Lk, D
    a.q_bs, a.q_rs, a.q_hs = q.stride(0), q.stride(1), q.stride(2)
    a.k_bs, a.k_rs, a.k_hs = k.stride(0), k.stride(1), k.stride(2)
    a.v_bs, a.v_rs, a.v_hs = v.stride(0), v.stride(1), v.stride(2)
    a.o_bs, a.o_rs, a.o_hs = out.stride(0), out.stride(1), out.stride(2)
    a.scale = float(scale if scale is not None else D ** -0.5)
    a.flags = flags if flags else _FMHA_DEFAULT_FLAGS
    if q_row_scale is not None:
        if q_row_scale.dtype != torch.float32 or not q_row_scale.is_contiguous() or q_row_scale.numel() != B * Lq or not q_row_scale.is_cuda:
            raise TypeError("fmha: q_row_scale must be a contiguous CUDA float32 tensor of B*Lq elements")
        a.q_row_scale = q_row_scale.data_ptr()
    st = _stream()
    if D == 128 and Lk >= 512:
        # scratch of the key-split last wave (vist3a_fmha_workspace_bytes): one buffer per (device, stream), reused by every call on that stream
        key = (q.device.index, B, H, Lq, Lk, a.flags)
        need = _FMHA_WS_BYTES.get(key)
        if need is None:
            need = int(L.load().vist3a_fmha_workspace_bytes(C.byref(a)))
            if need < 0:
                L.check(need)
            _FMHA_WS_BYTES[key] = need
        if need:
            if torch.cuda.is_current_stream_capturing():
                # a buffer of the graph's own pool, like any temporary of captured code: the cache below must neither hand a graph-pool
                # buffer to eager callers nor bake one eager buffer into several graphs that may replay on different streams
                ws = torch.empty(need, dtype=torch.uint8, device=q.device)
            else:
                skey = (q.device.index, st or 0)
                ws = _FMHA_WS.get(skey)
                if ws is None or ws.numel() < need:
                    ws = torch.empty(need, dtype=torch.uint8, device=q.device)
                    _FMHA_WS.pop(skey, None)
                    _FMHA_WS[skey] = ws
                    while len(_FMHA_WS) > 8:   # streams come and go: keep the most recent few (a dropped buffer is freed stream-ordered)
                        _FMHA_WS.pop(next(iter(_FMHA_WS)))
            a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    L.check(L.load().vist3a_fmha_fwd(C.byref(a), st))
    return out


def layernorm(x: torch.Tensor, *, mul: Optional[torch.Tensor] = None, add: Optional[torch.Tensor] = None,
              mul_bstride: int = 0, add_bstride: int = 0, rows_per_batch: int = 0, eps: float = 1e-6,
              out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None, mul_plus_one: bool = False,
              in_map: Optional[tuple] = None, out_map: Optional[tuple] = None, rows: Optional[int] = None) -> torch.Tensor:
    """out[r] = LN(x[r]) * mul[b] + add[b], b = r // rows_per_batch (fp32 statistics).
    in_map / out_map = (rows_per_group, group_stride, group_offset) select memory rows of x / out; `rows` = rows to process."""
    _need_cuda(x, mul, add, out)
    x2 = _rows2d(x)
    dim = x2.shape[1]
    if rows is None:
        rows = x2.shape[0]
    if out is None:
        out = torch.empty((rows, dim), dtype=out_dtype or x2.dtype, device=x.device)
    o2 = out if out.dim() == 2 else out.view(-1, out.shape[-1])
    for v in (mul, add):
        if v is not None and v.dtype != torch.float32:
            raise TypeError("layernorm: mul/add must be float32")
    L.check(L.load().vist3a_layernorm(x2.data_ptr(), _dt(x2), x2.stride(0), o2.data_ptr(), _dt(o2), o2.stride(0), rows,
                                      dim, rows_per_batch if rows_per_batch > 0 else rows, _ptr(mul), mul_bstride,
                                      _ptr(add), add_bstride, eps, int(mul_plus_one),
                                      C.byref(L.RowMap(*in_map)) if in_map else None,
                                      C.byref(L.RowMap(*out_map)) if out_map else None, _stream()))
    return out


def rmsnorm_rope_(x: torch.Tensor, weight: torch.Tensor, head_dim: int, *, eps: float = 1e-6,
                  cos: Optional[torch.Tensor] = None, sin: Optional[torch.Tensor] = None, nseg: int = 1) -> torch.Tensor:
    """In place on a bf16 [rows, nseg*dim] matrix (may be a column slice of a wider buffer).  nseg > 1: the columns are nseg
    independent segments (q | k of a fused qkv buffer), each normalised over its own `dim` columns with weight[s*dim:(s+1)*dim]."""
    _need_cuda(x, weight, cos, sin)
    if x.dim() != 2 or x.dtype != torch.bfloat16 or x.stride(1) != 1:
        raise TypeError("rmsnorm_rope_: x must be a 2-D bfloat16 tensor with unit inner stride")
    rows, dim = x.shape
    if dim % nseg or weight.numel() != dim:
        raise ValueError("rmsnorm_rope_: weight must hold one vector per segment")
    dim //= nseg
    rope_len = cos.shape[0] if cos is not None else 0
    L.check(L.load().vist3a_rmsnorm_rope(x.data_ptr(), x.stride(0), rows, dim, head_dim, weight.data_ptr(), eps,
                                         _ptr(cos), _ptr(sin), rope_len, nseg, dim if nseg > 1 else 0, _stream()))
    return x


def row_rinv(x: torch.Tensor, eps: float = 1e-6, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """rsqrt(mean(x^2, -1) + eps) per row of a bf16 [rows, dim] matrix -> fp32 [rows]."""
    _need_cuda(x, out)
    if x.dim() != 2 or x.dtype != torch.bfloat16 or x.stride(1) != 1:
        raise TypeError("row_rinv: x must be a 2-D bfloat16 tensor with unit inner stride")
    if out is None:
        out = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device)
    L.check(L.load().vist3a_row_rinv(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], eps, out.data_ptr(), _stream()))
    return out


def modulation(table: torch.Tensor, mod: torch.Tensor, *, nvec: int, broadcast: bool, one_plus_mask: int,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[b, j, :] = table[j, :] + mod[b, (j,) :] (+1 where bit j of one_plus_mask is set); fp32."""
    _need_cuda(table, mod, out)
    B = mod.shape[0]
    dim = table.shape[-1]
    if out is None:
        out = torch.empty((B, nvec, dim), dtype=torch.float32, device=mod.device)
    L.check(L.load().vist3a_modulation(table.data_ptr(), mod.data_ptr(), _dt(mod), int(broadcast), out.data_ptr(), B,
                                       nvec, dim, one_plus_mask, _stream()))
    return out


def skinny_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, pre_act=None, act=None,
                  out_dtype: Optional[torch.dtype] = None, out: Optional[torch.Tensor] = None,
                  gate: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = residual + gate * act(pre_act(x) @ w^T + bias) for M <= 16 rows (weight-streaming, HBM bound)."""
    _need_cuda(x, w, bias, out, gate, residual)
    x2 = _rows2d(x)
    M, K = x2.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype or x2.dtype, device=x.device)
    if bias is not None and bias.dtype != torch.float32:
        raise TypeError("skinny_linear: bias must be float32")
    for m0 in range(0, M, 16):  # the kernel keeps 16 token rows in registers per pass over the weights
        m1 = min(M, m0 + 16)
        res = residual[m0:m1] if residual is not None else None
        L.check(L.load().vist3a_skinny_linear(x2[m0:m1].data_ptr(), _dt(x2), x2.stride(0), w.data_ptr(), _dt(w), w.stride(0),
                                              _ptr(bias), out[m0:m1].data_ptr(), _dt(out), out.stride(0), m1 - m0, N, K, ACT[pre_act],
                                              ACT[act], _ptr(gate), _ptr(res), residual.stride(0) if residual is not None else 0,
                                              _stream()))
    return out


def timestep_features(t: torch.Tensor, dim: int, out_dtype=torch.float32) -> torch.Tensor:
    _need_cuda(t)
    t = t.to(torch.float32).contiguous()
    out = torch.empty((t.shape[0], dim), dtype=out_dtype, device=t.device)
    L.check(L.load().vist3a_timestep_features(t.data_ptr(), out.data_ptr(), _dt(out), t.shape[0], dim, _stream()))
    return out


def patchify(x: torch.Tensor) -> torch.Tensor:
    """x [B, C, T, H, W] -> [B*T*(H/2)*(W/2), 4C] bf16 (Wan patch (1,2,2), k = c*4 + dy*2 + dx)."""
    _need_cuda(x)
    x = x.contiguous()
    B, Cc, T, H, W = x.shape
    out = torch.empty((B * T * (H // 2) * (W // 2), 4 * Cc), dtype=torch.bfloat16, device=x.device)
    L.check(L.load().vist3a_patchify(x.data_ptr(), _dt(x), out.data_ptr(), B, Cc, T, H, W, _stream()))
    return out


def unpatchify(p: torch.Tensor, B: int, Cc: int, T: int, H: int, W: int, out_dtype=torch.bfloat16) -> torch.Tensor:
    _need_cuda(p)
    out = torch.empty((B, Cc, T, H, W), dtype=out_dtype, device=p.device)
    L.check(L.load().vist3a_unpatchify(p.data_ptr(), _dt(p), p.stride(0), out.data_ptr(), _dt(out), B, Cc, T, H, W,
                                       _stream()))
    return out


def cfg_combine(cond: torch.Tensor, uncond: torch.Tensor, guidance: float, out: Optional[torch.Tensor] = None):
    _need_cuda(cond, uncond, out)
    cond, uncond = cond.contiguous(), uncond.contiguous()
    if out is None:
        out = torch.empty(cond.shape, dtype=torch.float32, device=cond.device)
    L.check(L.load().vist3a_cfg_combine(cond.data_ptr(), uncond.data_ptr(), _dt(cond), guidance, out.data_ptr(),
                                        cond.numel(), _stream()))
    return out


def axpby_n(out: torch.Tensor, terms, coeffs) -> torch.Tensor:
    """out = sum_i coeffs[i] * terms[i] (fp32, elementwise; out may alias a term)."""
    _need_cuda(out, *terms)
    n = len(terms)
    for t in terms:
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != out.numel():
            raise TypeError("axpby_n: terms must be contiguous float32 of the output's size")
    ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in terms])
    cf = (C.c_float * n)(*[float(c) for c in coeffs])
    L.check(L.load().vist3a_axpby_n(out.data_ptr(), n, ptrs, cf, out.numel(), _stream()))
    return out



# ------------------------------------------------------------------------------------------------
# stitched-decoder kernels
# ------------------------------------------------------------------------------------------------
def im2col_stitch(latent: torch.Tensor) -> torch.Tensor:
    """latent [B, C, T, h, w] -> bf16 [B*V*(h/2)*(w/2), C*45] (T-upsample + replicate pad fused), V = 4(T-1)+1."""
    _need_cuda(latent)
    latent = latent.contiguous()
    B, Cc, T, h, w = latent.shape
    V = (T - 1) * 4 + 1
    A = torch.empty((B * V * (h // 2) * (w // 2), Cc * 45), dtype=torch.bfloat16, device=latent.device)
    L.check(L.load().vist3a_im2col_stitch(latent.data_ptr(), _dt(latent), A.data_ptr(), B, Cc, T, h, w, _stream()))
    return A


def rgb_to_nhwc4pad(image: torch.Tensor) -> torch.Tensor:
    """image [B, 3, V, H, W] in [-1, 1] -> zero-padded RGB0 image [B*V, H, W+8, 4] fp32 in [0, 1] (3 zero pixels left, 5 right)."""
    _need_cuda(image)
    image = image.contiguous()
    B, Cc, V, H, W = image.shape
    if Cc != 3:
        raise ValueError("rgb_to_nhwc4pad: expected 3 colour channels")
    out = torch.empty((B * V, H, W + 8, 4), dtype=torch.float32, device=image.device)
    L.check(L.load().vist3a_rgb_to_nhwc4pad(image.data_ptr(), _dt(image), out.data_ptr(), B, V, H, W, _stream()))
    return out


def rgb01_views_to_nhwc4pad(image: torch.Tensor) -> torch.Tensor:
    """image [B, V, 3, H, W] in [0, 1] -> zero-padded RGB0 image [B*V, H, W+8, 4] fp32 (un-stitched image -> 3DGS path)."""
    _need_cuda(image)
    image = image.contiguous()
    B, V, Cc, H, W = image.shape
    if Cc != 3:
        raise ValueError("rgb01_views_to_nhwc4pad: expected 3 colour channels")
    out = torch.empty((B * V, H, W + 8, 4), dtype=torch.float32, device=image.device)
    L.check(L.load().vist3a_rgb01_views_to_nhwc4pad(image.data_ptr(), _dt(image), out.data_ptr(), B, V, H, W, _stream()))
    return out


# ImageNet statistics of AS/.../vggt/models/aggregator.py:29-30 as the reference model holds them: EncoderAnySplat casts the aggregator and its
# buffers to bfloat16 (AS/model/encoder/anysplat.py:144), i.e. (0.485, 0.456, 0.406) / (0.229, 0.224, 0.225) rounded to bf16
IMAGENET_MEAN, IMAGENET_STD = (0.484375, 0.455078125, 0.40625), (0.228515625, 0.2236328125, 0.224609375)


def patch_embed_im2col(image: torch.Tensor, patch: int, k_pad: int) -> torch.Tensor:
    """image [n, 3, H, W] in [0, 1] -> bf16 [n * (H/patch) * (W/patch), k_pad] operand of the patch-embedding GEMM
    (ImageNet-normalised, k = c*p*p + py*p + px); see vist3a_patch_embed_im2col."""
    _need_cuda(image)
    image = image.contiguous()
    n, Cc, H, W = image.shape
    if Cc != 3:
        raise ValueError("patch_embed_im2col: expected 3 colour channels")
    A = torch.empty((n * (H // patch) * (W // patch), k_pad), dtype=torch.bfloat16, device=image.device)
    mean, std = (C.c_float * 3)(*IMAGENET_MEAN), (C.c_float * 3)(*IMAGENET_STD)
    L.check(L.load().vist3a_patch_embed_im2col(image.data_ptr(), _dt(image), A.data_ptr(), k_pad, n, H, W, patch, mean, std, _stream()))
    return A


def im2col_nhwc(x: torch.Tensor, kh: int, kw: int, stride: int, pad: int, k_pad: Optional[int] = None) -> torch.Tensor:
    """x NHWC fp32 -> [n*ho*wo, k_pad] fp32 with column (dy*kw+dx)*C + c (zero beyond kh*kw*C)."""
    _need_cuda(x)
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise TypeError("im2col_nhwc: x must be contiguous float32 NHWC")
    n, h, w, Cc = x.shape
    ho, wo = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
    ld = k_pad or kh * kw * Cc
    A = torch.empty((n * ho * wo, ld), dtype=torch.float32, device=x.device)
    L.check(L.load().vist3a_im2col_nhwc(x.data_ptr(), A.data_ptr(), ld, n, h, w, Cc, kh, kw, stride, pad, _stream()))
    return A


def qknorm_rope2d_(qkv: torch.Tensor, heads: int, qw, qb, kw, kb, cos_tab, sin_tab, *, tokens_per_view: int, n_special: int,
                   grid_w: int, eps: float = 1e-5) -> torch.Tensor:
    """in place on bf16 [rows, 3*heads*64]: LayerNorm(64) on every q / k head + 2-D RoPE."""
    _need_cuda(qkv, qw, qb, kw, kb, cos_tab, sin_tab)
    if qkv.dim() != 2 or qkv.dtype != torch.bfloat16 or qkv.stride(1) != 1:
        raise TypeError("qknorm_rope2d_: qkv must be 2-D bfloat16 with unit inner stride")
    L.check(L.load().vist3a_qknorm_rope2d(qkv.data_ptr(), qkv.stride(0), qkv.shape[0], heads, qw.data_ptr(), qb.data_ptr(),
                                          kw.data_ptr(), kb.data_ptr(), eps, cos_tab.data_ptr(), sin_tab.data_ptr(),
                                          cos_tab.shape[0], tokens_per_view, n_special, grid_w, _stream()))
    return qkv


def bilinear_nhwc(x: torch.Tensor, h_out: int, w_out: int, *, add: Optional[torch.Tensor] = None,
                  pos_x: Optional[torch.Tensor] = None, pos_y: Optional[torch.Tensor] = None) -> torch.Tensor:
    """align_corners=True bilinear resize of NHWC fp32 (+ optional same-shape `add` and separable pos-embed tables)."""
    _need_cuda(x, add, pos_x, pos_y)
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise TypeError("bilinear_nhwc: x must be contiguous float32 NHWC")
    n, h, w, Cc = x.shape
    out = torch.empty((n, h_out, w_out, Cc), dtype=torch.float32, device=x.device)
    L.check(L.load().vist3a_bilinear_nhwc(x.data_ptr(), out.data_ptr(), n, h, w, h_out, w_out, Cc, _ptr(add), _ptr(pos_x),
                                          _ptr(pos_y), _stream()))
    return out


def depth_to_space(x: torch.Tensor, n: int, h: int, w: int, Cc: int, k: int) -> torch.Tensor:
    """[n*h*w, k*k*C] (col (dy*k+dx)*C + c) -> NHWC [n, h*k, w*k, C]."""
    _need_cuda(x)
    out = torch.empty((n, h * k, w * k, Cc), dtype=torch.float32, device=x.device)
    L.check(L.load().vist3a_depth_to_space(x.data_ptr(), out.data_ptr(), n, h, w, Cc, k, _stream()))
    return out


def attention_small(qkv: torch.Tensor, B: int, Lq: int, H: int, D: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 attention for L <= 32: qkv [B*L, 3*H*D] -> [B*L, H*D]."""
    _need_cuda(qkv, out)
    if qkv.dtype != torch.float32 or not qkv.is_contiguous():
        raise TypeError("attention_small: qkv must be contiguous float32")
    if out is None:
        out = torch.empty((B * Lq, H * D), dtype=torch.float32, device=qkv.device)
    elif not out.is_contiguous() or out.shape != (B * Lq, H * D):
        raise ValueError("attention_small: out must be contiguous [B*L, H*D]")
    L.check(L.load().vist3a_attention_small(qkv.data_ptr(), out.data_ptr(), B, Lq, H, D, D ** -0.5, _stream()))
    return out


def fma_rows(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a * b + c on [rows, dim] fp32 views (row strides allowed)."""
    _need_cuda(a, b, c, out)
    rows, dim = a.shape
    if out is None:
        out = torch.empty((rows, dim), dtype=torch.float32, device=a.device)
    L.check(L.load().vist3a_fma_rows(out.data_ptr(), out.stride(0), a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0),
                                     c.data_ptr(), c.stride(0), rows, dim, _stream()))
    return out


def linear_tokens16(x16: torch.Tensor, M: int, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, act=None,
                    gate: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.Linear for M <= 16 tokens, HBM-bound on the weights: the weight matrix [N, K] is streamed by TMA as the A operand of
    the tcgen05 GEMM (TF32 for fp32 weights) against the 16-row padded token matrix x16 [16, K]; a transposed epilogue kernel
    adds bias / activation / LayerScale / residual.  x16 may also be a 32-row buffer (up to 32 tokens: 21-view scenes).
    Returns y [rows(x16), N] (rows >= M are left untouched)."""
    _need_cuda(x16, w, bias, gate, residual, out)
    TP = x16.shape[0]
    if TP not in (16, 32) or not x16.is_contiguous() or M > TP:
        raise ValueError("linear_tokens16: x16 must be a contiguous [16 or 32, K] buffer holding M <= rows token rows")
    N, K = w.shape
    # split-K by reshaping: W [N, K] is the same memory as [N*S, K/S] (row n*S + s = K-slice s of row n), x16 the same as [TP*S, K/S]; the
    # GEMM then has N*S/128 CTAs streaming weights instead of N/128 (16 for a 2048-row matrix on 148 SMs) and bias_act_t sums the
    # diagonal blocks ct[n*S + s, m*S + s].  S: smallest power of two that gives >= 96 CTAs, K/S a multiple of 32 floats, TP*S <= 256.
    S = 1
    while w.is_contiguous() and (N * S) // 128 < 96 and S < 8 and (K // (2 * S)) % 32 == 0 and TP * 2 * S <= 256:
        S *= 2
    if S > 1:
        ct = gemm(w.view(N * S, K // S), x16.view(TP * S, K // S), out_dtype=torch.float32, two_cta=False)   # [N*S, TP*S]
    else:
        ct = gemm(w, x16, out_dtype=torch.float32, two_cta=False)  # [N, TP] = W x^T
    if out is None:
        out = torch.zeros((TP, N), dtype=torch.float32, device=x16.device)
    L.check(L.load().vist3a_bias_act_t(ct.data_ptr(), ct.stride(0), _ptr(bias), ACT[act], _ptr(gate), _ptr(residual),
                                       residual.stride(0) if residual is not None else 0, out.data_ptr(), out.stride(0), M, N, S,
                                       _stream()))
    return out


def pose_to_cameras(pose_raw: torch.Tensor, H: int, W: int, cameras: bool = True):
    """pose_raw [S, 9] fp32 -> dict(pose_act [S,9], extr [S,3,4], intr [S,3,3], c2w [S,4,4], intr_norm [S,3,3])."""
    _need_cuda(pose_raw)
    pose_raw = pose_raw.contiguous()
    S = pose_raw.shape[0]
    dev = pose_raw.device
    out = {"pose_act": torch.empty((S, 9), dtype=torch.float32, device=dev)}
    if cameras:
        out.update(extr=torch.empty((S, 3, 4), dtype=torch.float32, device=dev), intr=torch.empty((S, 3, 3), dtype=torch.float32, device=dev),
                   c2w=torch.empty((S, 4, 4), dtype=torch.float32, device=dev), intr_norm=torch.empty((S, 3, 3), dtype=torch.float32, device=dev))
    L.check(L.load().vist3a_pose_to_cameras(pose_raw.data_ptr(), out["pose_act"].data_ptr(), _ptr(out.get("extr")), _ptr(out.get("intr")),
                                            _ptr(out.get("c2w")), _ptr(out.get("intr_norm")), S, H, W, _stream()))
    return out


GAUSSIAN_FIELDS = (("means", 3), ("scales", 3), ("rotations", 4), ("opacities", 1), ("harmonics", None), ("covariances", 9))


def alloc_gaussian_fields(P: int, d_sh: int, device) -> dict:
    """One flat fp32 buffer holding every per-Gaussian output field-major (means | scales | rotations | opacities | harmonics |
    covariances, each field contiguous over the P Gaussians), plus the per-field views the kernels write.  The multi-GPU gather sends
    the buffer as it is (vist3a_b200.t23d.all_gather_gaussians): no packing or unpacking copies."""
    widths = [(k, 3 * d_sh if w is None else w) for k, w in GAUSSIAN_FIELDS]
    flat = torch.empty((P * sum(w for _, w in widths),), dtype=torch.float32, device=device)
    o, off = {"packed": flat}, 0
    for k, w in widths:
        o[k] = flat[off:off + P * w]
        off += P * w
    o["means"], o["scales"], o["rotations"] = o["means"].view(P, 3), o["scales"].view(P, 3), o["rotations"].view(P, 4)
    o["harmonics"], o["covariances"] = o["harmonics"].view(P, 3, d_sh), o["covariances"].view(P, 3, 3)
    return o


def gaussian_epilogue(depth_feat: torch.Tensor, depth_w: torch.Tensor, depth_b: float, gs_raw: torch.Tensor, extr: torch.Tensor,
                      intr: torch.Tensor, sh_mask: torch.Tensor, S: int, H: int, W: int):
    """fused depth activation + unprojection + Gaussian adapter; see vist3a_gaussian_epilogue."""
    _need_cuda(depth_feat, depth_w, gs_raw, extr, intr, sh_mask)
    P = S * H * W
    d_sh = sh_mask.shape[0]
    dev = gs_raw.device
    f32 = torch.float32
    o = alloc_gaussian_fields(P, d_sh, dev)
    o["depth"], o["scene_sum"] = torch.empty((P,), dtype=f32, device=dev), torch.zeros((1,), dtype=f32, device=dev)
    L.check(L.load().vist3a_gaussian_epilogue(depth_feat.data_ptr(), depth_feat.stride(0), depth_w.shape[0], depth_w.data_ptr(),
                                              float(depth_b), gs_raw.data_ptr(), gs_raw.stride(0), extr.data_ptr(), intr.data_ptr(),
                                              sh_mask.data_ptr(), d_sh, S, H, W, o["depth"].data_ptr(), o["means"].data_ptr(),
                                              o["scales"].data_ptr(), o["rotations"].data_ptr(), o["opacities"].data_ptr(),
                                              o["harmonics"].data_ptr(), o["covariances"].data_ptr(), o["scene_sum"].data_ptr(), _stream()))
    return o


def gaussian_adapter(pts: torch.Tensor, feats: torch.Tensor, sh_mask: torch.Tensor):
    """Gaussian adapter on given positions (voxelize=True branch); pts [P,3], feats [P, >= 8 + 3 d_sh] rows; see vist3a_gaussian_adapter."""
    _need_cuda(pts, feats, sh_mask)
    P = pts.shape[0]
    d_sh = sh_mask.shape[0]
    dev, f32 = pts.device, torch.float32
    if pts.dtype != f32 or feats.dtype != f32 or not pts.is_contiguous() or feats.stride(1) != 1:
        raise ValueError("gaussian_adapter: fp32 contiguous points and unit-stride fp32 feature rows expected")
    o = alloc_gaussian_fields(P, d_sh, dev)
    if P > 0:
        L.check(L.load().vist3a_gaussian_adapter(pts.data_ptr(), feats.data_ptr(), feats.stride(0), sh_mask.data_ptr(), d_sh, P,
                                                 o["means"].data_ptr(), o["scales"].data_ptr(), o["rotations"].data_ptr(),
                                                 o["opacities"].data_ptr(), o["harmonics"].data_ptr(), o["covariances"].data_ptr(), _stream()))
    return o


def voxel_fusion(pts: torch.Tensor, feats: torch.Tensor, conf: torch.Tensor, voxel_size: float, *, feat_dim: Optional[int] = None,
                 want_index: bool = False):
    """Voxelised fusion (see vist3a_voxel_fusion): pts [N,3] fp32 contiguous, feats [N, >= feat_dim] unit-stride fp32 rows, conf a
    1-D fp32 view of N values (may be a strided column of the feature rows).  Returns dict(pts [M,3], feats [M,C], n_voxels = M
    [, inverse [N] int32, counts [M] int32]); reading M synchronises the stream (the output size is data dependent, as
    torch.unique's is in the reference)."""
    _need_cuda(pts, feats, conf)
    N = pts.shape[0]
    Cc = int(feat_dim if feat_dim is not None else feats.shape[1])
    f32, dev = torch.float32, pts.device
    if pts.dtype != f32 or feats.dtype != f32 or conf.dtype != f32 or not pts.is_contiguous() or feats.stride(1) != 1 or conf.dim() != 1:
        raise ValueError("voxel_fusion: fp32 contiguous points, unit-stride fp32 feature rows and a 1-D fp32 confidence view expected")
    if feats.shape[0] != N or conf.shape[0] != N:
        raise ValueError("voxel_fusion: pts / feats / conf disagree on the number of points")
    lib = L.load()
    ws_bytes = int(lib.vist3a_voxel_fusion_workspace_bytes(N))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    vp = torch.empty((N, 3), dtype=f32, device=dev)
    vf = torch.empty((N, Cc), dtype=f32, device=dev)
    inv = torch.empty((N,), dtype=torch.int32, device=dev) if want_index else None
    cnt = torch.empty((N,), dtype=torch.int32, device=dev) if want_index else None
    nv = torch.zeros((1,), dtype=torch.int64, device=dev)
    L.check(lib.vist3a_voxel_fusion(pts.data_ptr(), feats.data_ptr(), feats.stride(0), Cc, conf.data_ptr(), conf.stride(0), N, float(voxel_size),
                                    vp.data_ptr(), vf.data_ptr(), _ptr(inv), _ptr(cnt), nv.data_ptr(), ws.data_ptr(), ws_bytes, _stream()))
    M = int(nv.item())
    if M < 0:
        raise L.Vist3aError(L.ERR_UNSUPPORTED, "voxel_fusion: voxel coordinate ranges need more than 64 key bits")
    out = dict(pts=vp[:M], feats=vf[:M], n_voxels=M)
    if want_index:
        out["inverse"], out["counts"] = inv, cnt[:M]
    return out


def gs_render(means: torch.Tensor, covariances: torch.Tensor, opacities: torch.Tensor, harmonics: torch.Tensor, viewmat, K, W: int, H: int, *,
              sh_degree: int = 4, background=(0.0, 0.0, 0.0), near_plane: float = 1e-10, far_plane: float = 1e10, radius_clip: float = 0.1,
              eps2d: float = 0.3):
    """Rasterise N Gaussians into one view (see vist3a_gs_project / vist3a_gs_rasterize): means [N,3], covariances [N,3,3], opacities [N],
    harmonics [N,3,d_sh] fp32 on the device; viewmat 4x4 world->camera and K 3x3 (pixels) as nested lists / CPU tensors.  Returns
    dict(rgb [H,W,3] unclamped, depth [H,W], alpha [H,W], n_isect).  Reading the intersection count synchronises the stream."""
    _need_cuda(means, covariances, opacities, harmonics)
    f32, dev = torch.float32, means.device
    N, d_sh = means.shape[0], harmonics.shape[-1]
    for t_, shp in ((means, (N, 3)), (covariances, (N, 3, 3)), (opacities, (N,)), (harmonics, (N, 3, d_sh))):
        if t_.dtype != f32 or not t_.is_contiguous() or tuple(t_.shape) != shp:
            raise ValueError(f"gs_render: contiguous fp32 tensor of shape {shp} expected, got {tuple(t_.shape)} {t_.dtype}")
    vm = (C.c_float * 16)(*[float(v) for v in torch.as_tensor(viewmat, dtype=f32).reshape(-1).tolist()])
    kk = (C.c_float * 9)(*[float(v) for v in torch.as_tensor(K, dtype=f32).reshape(-1).tolist()])
    bg = (C.c_float * 3)(*[float(v) for v in background])
    lib = L.load()
    pws = torch.empty((int(lib.vist3a_gs_project_workspace_bytes(N)),), dtype=torch.uint8, device=dev)
    n_is = torch.zeros((1,), dtype=torch.int64, device=dev)
    L.check(lib.vist3a_gs_project(means.data_ptr(), covariances.data_ptr(), opacities.data_ptr(), harmonics.data_ptr(), d_sh, int(sh_degree), N, vm, kk,
                                  W, H, float(near_plane), float(far_plane), float(radius_clip), float(eps2d), pws.data_ptr(), pws.numel(),
                                  n_is.data_ptr(), _stream()))
    n_isect = int(n_is.item())
    rws = torch.empty((int(lib.vist3a_gs_rasterize_workspace_bytes(n_isect, W, H)),), dtype=torch.uint8, device=dev)
    out = dict(rgb=torch.empty((H, W, 3), dtype=f32, device=dev), depth=torch.empty((H, W), dtype=f32, device=dev),
               alpha=torch.empty((H, W), dtype=f32, device=dev), n_isect=n_isect)
    L.check(lib.vist3a_gs_rasterize(pws.data_ptr(), N, n_isect, W, H, bg, rws.data_ptr(), rws.numel(), out["rgb"].data_ptr(), out["depth"].data_ptr(),
                                    out["alpha"].data_ptr(), _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# per-launch device timing (bench.py roofline): CUDA events recorded on the launching stream around
# every call of the wrapped op, with its algorithmic FLOPs / bytes.
# ------------------------------------------------------------------------------------------------
class OpTimer:
    """with OpTimer() as t: ...run the path eagerly...; t.summary() -> {kernel class: {ms, launches, flops, bytes}}"""

    def __init__(self):
        self.records = []
        self._saved = {}

    def _wrap(self, name, fn, cost):
        def wrapped(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = fn(*a, **k)
            e.record()
            fl, by, tag = cost(out, a, k)
            self.records.append((tag or name, s, e, fl, by))
            return out
        return wrapped

    def __enter__(self):
        import sys

        mod = sys.modules[__name__]

        def gemm_cost(out, args, kw):
            a, w = args[0], args[1]
            N, K = w.shape
            M = out.numel() // N if kw.get("cmap") is None else (a.numel() // a.shape[-1])
            if kw.get("conv") is not None:
                cv = kw["conv"]
                n_, h_, w_, _ = cv.get("geom", a.shape)
                M = n_ * (h_ + 2 * cv["pad"] - cv["kh"] + 1) * (w_ + 2 * cv.get("pad_x", cv["pad"]) - cv["kw"] + 1)
                N = out.shape[-1] if out.dim() == 2 else N
            by = M * K * a.element_size() + N * K * w.element_size() + M * N * out.element_size()
            ep = ("+" + kw["act"] if kw.get("act") else "") + ("+gate" if kw.get("gate") is not None else "") + (
                "+res" if kw.get("residual") is not None else "")
            kind = ("tf32" if a.dtype == torch.float32 else "bf16") + ("conv" if kw.get("conv") is not None else "")
            return 2.0 * M * N * K, by, f"gemm_tcgen05|{kind} {M}x{N}x{K}{ep}"

        def fmha_cost(out, args, kw):
            q, kk = args[0], args[1]
            B, Lq, H, D = q.shape
            Lk = kk.shape[1]
            return 4.0 * B * H * Lq * Lk * D, 2 * (2 * B * Lq * H * D + 2 * B * Lk * H * D), f"fmha_tcgen05|{B}x{H}x{Lq}x{Lk}x{D}"

        def io_cost(tag):
            def f(out, args, kw):
                outs = list(out.values()) if isinstance(out, dict) else [out]
                ts = [t for t in list(args) + list(kw.values()) + outs if isinstance(t, torch.Tensor)]
                return 0.0, sum(t.numel() * t.element_size() for t in ts), tag
            return f

        def voxel_cost(out, args, kw):  # algorithmic: every point row read once, every voxel row written once
            n, m, c = args[0].shape[0], out["n_voxels"], out["feats"].shape[1]
            return 0.0, 4.0 * (n * (3 + c + 1) + m * (3 + c)), "voxel_fusion"

        table = {"voxel_fusion": voxel_cost, "gaussian_adapter": io_cost("gaussian_adapter"), "gemm": gemm_cost, "fmha": fmha_cost, "layernorm": io_cost("layernorm"),
                 "rmsnorm_rope_": io_cost("rmsnorm_rope"), "row_rinv": io_cost("rmsnorm_rope"), "modulation": io_cost("small"), "skinny_linear": io_cost("small"),
                 "timestep_features": io_cost("small"), "patchify": io_cost("small"), "unpatchify": io_cost("small"),
                 "cfg_combine": io_cost("small"), "axpby_n": io_cost("small"), "im2col_stitch": io_cost("im2col"),
                 "im2col_nhwc": io_cost("im2col"), "rgb_to_nhwc4pad": io_cost("small"), "rgb01_views_to_nhwc4pad": io_cost("small"),
                 "patch_embed_im2col": io_cost("im2col"), "qknorm_rope2d_": io_cost("qknorm_rope2d"), "bilinear_nhwc": io_cost("bilinear"),
                 "depth_to_space": io_cost("depth_to_space"), "attention_small": io_cost("small"), "fma_rows": io_cost("small"),
                 "pose_to_cameras": io_cost("small"), "linear_tokens16": io_cost("linear_tokens16"), "gaussian_epilogue": io_cost("gaussian_epilogue"),
                 "vae_rmsnorm": io_cost("vae_rmsnorm"), "softmax_rows": io_cost("softmax_rows"), "time_interleave": io_cost("vae_layout"),
                 "transpose_bf16": io_cost("vae_layout"), "depth_to_space2_bf16": io_cost("vae_layout"), "latent_to_ndhwc": io_cost("small"),
                 "vae_frames_out": io_cost("vae_layout"), "resize_planes": io_cost("vae_layout")}
        for name, cost in table.items():
            self._saved[name] = getattr(mod, name)
            setattr(mod, name, self._wrap(name, self._saved[name], cost))
        return self

    def __exit__(self, *exc):
        import sys

        mod = sys.modules[__name__]
        for name, fn in self._saved.items():
            setattr(mod, name, fn)
        torch.cuda.synchronize()
        return False

    def summary(self, detail: bool = False):
        """aggregate by kernel class (default) or by kernel class | shape+epilogue (detail=True)"""
        agg = {}
        for tag, s, e, fl, by in self.records:
            d = agg.setdefault(tag if detail else tag.split("|")[0], {"ms": 0.0, "launches": 0, "flops": 0.0, "bytes": 0.0})
            d["ms"] += s.elapsed_time(e)
            d["launches"] += 1
            d["flops"] += fl
            d["bytes"] += by
        return agg


# ------------------------------------------------------------------------------------------------
# Wan VAE decode kernels (NDHWC bf16 activations; see vist3a_b200/wan_vae.py)
# ------------------------------------------------------------------------------------------------
def vae_rmsnorm(x: torch.Tensor, gamma: torch.Tensor, C: int, *, silu: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [..., ldx] bf16 (channels [0, C) valid) -> [..., ldy] bf16: RMS norm over the C channels * gamma (+ SiLU); channels [C, ldy) = 0."""
    _need_cuda(x, gamma, out)
    if x.dtype != torch.bfloat16 or not x.is_contiguous() or gamma.dtype != torch.float32:
        raise TypeError("vae_rmsnorm: contiguous bfloat16 activations and a float32 gamma expected")
    ldx = x.shape[-1]
    if out is None:
        out = torch.empty_like(x)
    rows = x.numel() // ldx
    L.check(L.load().vist3a_vae_rmsnorm(x.data_ptr(), ldx, gamma.data_ptr(), out.data_ptr(), out.shape[-1], rows, C, int(silu), _stream()))
    return out


def softmax_rows(s: torch.Tensor, scale: float, out: Optional[torch.Tensor] = None, valid: Optional[int] = None) -> torch.Tensor:
    """softmax(scale * s) per row: fp32 [rows, L] -> bf16 [rows, L] (`out` may have a padded row stride); columns >= valid are padding (P = 0)."""
    _need_cuda(s, out)
    if s.dim() != 2 or s.dtype != torch.float32 or not s.is_contiguous():
        raise TypeError("softmax_rows: contiguous float32 [rows, L] expected")
    if out is None:
        out = torch.empty(s.shape, dtype=torch.bfloat16, device=s.device)
    if out.shape != s.shape or out.dtype != torch.bfloat16 or out.stride(1) != 1:
        raise TypeError("softmax_rows: out must be bfloat16 [rows, L] with unit inner stride")
    L.check(L.load().vist3a_softmax_rows(s.data_ptr(), out.data_ptr(), s.shape[0], s.shape[1], s.shape[1] if valid is None else int(valid), out.stride(0),
                                        float(scale), _stream()))
    return out


def time_interleave(y: torch.Tensor, out: torch.Tensor, C: Optional[int] = None) -> torch.Tensor:
    """y [T, H, W, ldy] bf16 (channels [0, 2C) valid) -> out [2T, H, W, ldo] (channels [0, C) written): out[2t + half] = y[t, ..., half*C:(half+1)*C].
    C defaults to ldy / 2 (dense rows)."""
    _need_cuda(y, out)
    T, H, W, ldy = y.shape
    C = ldy // 2 if C is None else C
    if (not y.is_contiguous() or not out.is_contiguous() or out.shape[:3] != (2 * T, H, W) or out.shape[3] < C or ldy < 2 * C
            or y.dtype != torch.bfloat16 or out.dtype != torch.bfloat16):
        raise ValueError("time_interleave: y [T, H, W, >= 2C] and out [2T, H, W, >= C] contiguous bfloat16 expected")
    L.check(L.load().vist3a_time_interleave(y.data_ptr(), ldy, out.data_ptr(), out.shape[3], T, H * W, C, _stream()))
    return out


def transpose_bf16(x: torch.Tensor) -> torch.Tensor:
    """[R, C] bf16 (unit inner stride, any row stride) -> [C, R] with the row stride rounded up to 8 elements (a 16-byte TMA stride)."""
    _need_cuda(x)
    if x.dim() != 2 or x.dtype != torch.bfloat16 or x.stride(1) != 1:
        raise TypeError("transpose_bf16: 2-D bfloat16 with unit inner stride expected")
    R = x.shape[0]
    out = torch.empty((x.shape[1], (R + 7) // 8 * 8), dtype=torch.bfloat16, device=x.device)[:, :R]
    L.check(L.load().vist3a_transpose_bf16(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), R, x.shape[1], _stream()))
    return out


def depth_to_space2_bf16(y: torch.Tensor, n: int, h: int, w: int, C: int, ldo: int) -> torch.Tensor:
    """[n*h*w, 4*C] bf16 (col (py*2+px)*C + c) -> NHWC [n, 2h, 2w, ldo] (channels [0, C) written)."""
    _need_cuda(y)
    if y.dtype != torch.bfloat16 or not y.is_contiguous() or y.shape != (n * h * w, 4 * C):
        raise ValueError("depth_to_space2_bf16: contiguous bfloat16 [n*h*w, 4*C] expected")
    out = torch.empty((n, 2 * h, 2 * w, ldo), dtype=torch.bfloat16, device=y.device)
    L.check(L.load().vist3a_depth_to_space2_bf16(y.data_ptr(), out.data_ptr(), n, h, w, C, ldo, _stream()))
    return out


def latent_to_ndhwc(z: torch.Tensor, ld: int) -> torch.Tensor:
    """z [C, T, h, w] (fp32 / bf16) -> [T, h, w, ld] bf16 with channels [C, ld) zero."""
    _need_cuda(z)
    z = z.contiguous()
    Cc, T, h, w = z.shape
    out = torch.empty((T, h, w, ld), dtype=torch.bfloat16, device=z.device)
    L.check(L.load().vist3a_latent_to_ndhwc(z.data_ptr(), _dt(z), out.data_ptr(), Cc, T * h * w, ld, _stream()))
    return out


def vae_frames_out(y: torch.Tensor, T: int, H: int, W: int) -> torch.Tensor:
    """conv_out result [T*H*W, ld] fp32 -> frames [3, T, H, W] fp32 clamped to [-1, 1]."""
    _need_cuda(y)
    if y.dtype != torch.float32 or not y.is_contiguous() or y.shape[0] != T * H * W:
        raise ValueError("vae_frames_out: contiguous float32 [T*H*W, ld] expected")
    out = torch.empty((3, T, H, W), dtype=torch.float32, device=y.device)
    L.check(L.load().vist3a_vae_frames_out(y.data_ptr(), y.shape[1], out.data_ptr(), T * H * W, _stream()))
    return out


def resize_planes(x: torch.Tensor, h_out: int, w_out: int) -> torch.Tensor:
    """[..., H, W] fp32 -> [..., h_out, w_out]: bilinear, half-pixel centres (align_corners=False), every leading index a plane."""
    _need_cuda(x)
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise TypeError("resize_planes: contiguous float32 expected")
    out = torch.empty(tuple(x.shape[:-2]) + (h_out, w_out), dtype=torch.float32, device=x.device)
    planes = x.numel() // (x.shape[-1] * x.shape[-2])
    L.check(L.load().vist3a_resize_planes(x.data_ptr(), out.data_ptr(), planes, x.shape[-2], x.shape[-1], h_out, w_out, _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# confidence-quantile branches of the stitched decoder (render_conf / opacity_conf)
# ------------------------------------------------------------------------------------------------
def depth_conf(feat: torch.Tensor, w: torch.Tensor, bias: float) -> torch.Tensor:
    """conf[p] = 1 + exp(feat[p] . w + bias): feat [P, C] fp32 rows (unit inner stride) -> [P] fp32."""
    _need_cuda(feat, w)
    if feat.dim() != 2 or feat.dtype != torch.float32 or feat.stride(1) != 1 or w.dtype != torch.float32:
        raise TypeError("depth_conf: float32 [P, C] rows and a float32 weight vector expected")
    out = torch.empty((feat.shape[0],), dtype=torch.float32, device=feat.device)
    L.check(L.load().vist3a_depth_conf(feat.data_ptr(), feat.stride(0), feat.shape[1], w.data_ptr(), float(bias), out.data_ptr(), feat.shape[0], _stream()))
    return out


def quantile(x: torch.Tensor, q: float) -> torch.Tensor:
    """torch.quantile(x.flatten(), q) (linear interpolation) as a 1-element device tensor; no host synchronisation."""
    _need_cuda(x)
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise TypeError("quantile: contiguous float32 expected")
    lib = L.load()
    n = x.numel()
    ws = torch.empty((int(lib.vist3a_quantile_workspace_bytes(n)),), dtype=torch.uint8, device=x.device)
    out = torch.empty((1,), dtype=torch.float32, device=x.device)
    L.check(lib.vist3a_quantile_f32(x.data_ptr(), n, float(q), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return out


def compact_rows(conf: torch.Tensor, threshold: torch.Tensor, feats: torch.Tensor, pts: torch.Tensor, *, feat_dim: int, use_threshold: bool = True,
                 want_damp: bool = False):
    """Rows with conf > threshold (all rows if not use_threshold), in order: dict(feats [M, C], pts [M, 3], damp [M] or None, count = M).
    Reading M synchronises the stream (the output size is data dependent, as the boolean-mask gather is in the reference)."""
    _need_cuda(conf, threshold, feats, pts)
    n = conf.shape[0]
    f32, dev = torch.float32, conf.device
    if conf.dtype != f32 or not conf.is_contiguous() or feats.dtype != f32 or feats.stride(1) != 1 or pts.dtype != f32 or not pts.is_contiguous() or feats.shape[0] != n:
        raise ValueError("compact_rows: fp32 contiguous conf / points and unit-stride fp32 feature rows expected")
    lib = L.load()
    ws = torch.empty((int(lib.vist3a_compact_rows_workspace_bytes(n)),), dtype=torch.uint8, device=dev)
    of = torch.empty((n, feat_dim), dtype=f32, device=dev)
    op = torch.empty((n, 3), dtype=f32, device=dev)
    od = torch.empty((n,), dtype=f32, device=dev) if want_damp else None
    cnt = torch.zeros((1,), dtype=torch.int64, device=dev)
    L.check(lib.vist3a_compact_rows(conf.data_ptr(), threshold.data_ptr(), int(use_threshold), n, feats.data_ptr(), feats.stride(0), feat_dim, pts.data_ptr(),
                                    of.data_ptr(), op.data_ptr(), _ptr(od), cnt.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    m = int(cnt.item())
    return dict(feats=of[:m], pts=op[:m], damp=None if od is None else od[:m], count=m)
