"""Tensor-level wrappers of the C-ABI kernels.  PyTorch is used for device memory and streams only:
every wrapper validates layout, passes raw device pointers + the current CUDA stream to
libvist3a_sm100.so and returns the (pre-allocated or new) output tensor.  No op has a fallback.
"""
from __future__ import annotations

import ctypes as C
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
         two_cta: Optional[bool] = None) -> torch.Tensor:
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T); see vist3a_gemm in include/vist3a_sm100.h."""
    _need_cuda(a, w, bias, gate, residual, out)
    a2 = _rows2d(a)
    w2 = _rows2d(w)
    if a2.dtype != w2.dtype:
        raise TypeError(f"gemm: A is {a2.dtype} but W is {w2.dtype}")
    M, K = a2.shape
    N, K2 = w2.shape
    if K != K2:
        raise ValueError(f"gemm: K mismatch {K} vs {K2}")
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype or a2.dtype, device=a.device)
    o2 = out if out.dim() == 2 else out.view(-1, out.shape[-1])
    if o2.stride(-1) != 1 or o2.shape[0] != M or o2.shape[1] != N:
        raise ValueError("gemm: out must be [M, N] with unit inner stride")
    for v, n in ((bias, "bias"), (gate, "gate")):
        if v is not None and v.dtype != torch.float32:
            raise TypeError(f"gemm: {n} must be float32")
    r2 = None
    if residual is not None:
        r2 = residual if residual.dim() == 2 else residual.view(-1, residual.shape[-1])
        if r2.dtype != o2.dtype or r2.stride(-1) != 1:
            raise TypeError("gemm: residual must have the output dtype and unit inner stride")
    args = L.GemmArgs()
    args.A, args.W, args.C = a2.data_ptr(), w2.data_ptr(), o2.data_ptr()
    args.bias, args.gate, args.residual = _ptr(bias), _ptr(gate), _ptr(r2)
    args.M, args.N, args.K = M, N, K
    args.lda, args.ldw, args.ldc = a2.stride(0), w2.stride(0), o2.stride(0)
    args.ldr = r2.stride(0) if r2 is not None else 0
    args.rows_per_batch = rows_per_batch if rows_per_batch > 0 else M
    args.gate_bstride = gate_bstride
    args.in_dtype, args.out_dtype = _dt(a2), _dt(o2)
    args.act = ACT[act]
    args.round_linear, args.round_gate = int(round_linear), int(round_gate)
    if two_cta is None:
        two_cta = M >= 2048
    args.flags = L.GEMM_FLAG_2CTA if two_cta else L.GEMM_FLAG_1CTA
    L.check(L.load().vist3a_gemm(C.byref(args), _stream()))
    return out


def fmha(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, scale: Optional[float] = None,
         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Non-causal attention.  q [B, Lq, H, D], k/v [B, Lkv, H, D] (any strides with unit inner stride,
    e.g. slices of a fused QKV buffer); returns [B, Lq, H, D] bf16."""
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
    a.flags = 0
    L.check(L.load().vist3a_fmha_fwd(C.byref(a), _stream()))
    return out


def layernorm(x: torch.Tensor, *, mul: Optional[torch.Tensor] = None, add: Optional[torch.Tensor] = None,
              mul_bstride: int = 0, add_bstride: int = 0, rows_per_batch: int = 0, eps: float = 1e-6,
              out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """out[r] = LN(x[r]) * mul[b] + add[b], b = r // rows_per_batch (fp32 statistics)."""
    _need_cuda(x, mul, add, out)
    x2 = _rows2d(x)
    rows, dim = x2.shape
    if out is None:
        out = torch.empty((rows, dim), dtype=out_dtype or x2.dtype, device=x.device)
    o2 = out if out.dim() == 2 else out.view(-1, out.shape[-1])
    for v in (mul, add):
        if v is not None and v.dtype != torch.float32:
            raise TypeError("layernorm: mul/add must be float32")
    L.check(L.load().vist3a_layernorm(x2.data_ptr(), _dt(x2), x2.stride(0), o2.data_ptr(), _dt(o2), o2.stride(0), rows,
                                      dim, rows_per_batch if rows_per_batch > 0 else rows, _ptr(mul), mul_bstride,
                                      _ptr(add), add_bstride, eps, _stream()))
    return out


def rmsnorm_rope_(x: torch.Tensor, weight: torch.Tensor, head_dim: int, *, eps: float = 1e-6,
                  cos: Optional[torch.Tensor] = None, sin: Optional[torch.Tensor] = None) -> torch.Tensor:
    """In place on a bf16 [rows, dim] matrix (may be a column slice of a wider buffer)."""
    _need_cuda(x, weight, cos, sin)
    if x.dim() != 2 or x.dtype != torch.bfloat16 or x.stride(1) != 1:
        raise TypeError("rmsnorm_rope_: x must be a 2-D bfloat16 tensor with unit inner stride")
    rows, dim = x.shape
    rope_len = cos.shape[0] if cos is not None else 0
    L.check(L.load().vist3a_rmsnorm_rope(x.data_ptr(), x.stride(0), rows, dim, head_dim, weight.data_ptr(), eps,
                                         _ptr(cos), _ptr(sin), rope_len, _stream()))
    return x


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
                  out_dtype: Optional[torch.dtype] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = act(pre_act(x) @ w^T + bias) for M <= 16 rows (weight-streaming, HBM bound)."""
    _need_cuda(x, w, bias, out)
    x2 = _rows2d(x)
    M, K = x2.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype or x2.dtype, device=x.device)
    if bias is not None and bias.dtype != torch.float32:
        raise TypeError("skinny_linear: bias must be float32")
    L.check(L.load().vist3a_skinny_linear(x2.data_ptr(), _dt(x2), x2.stride(0), w.data_ptr(), _dt(w), w.stride(0),
                                          _ptr(bias), out.data_ptr(), _dt(out), out.stride(0), M, N, K, ACT[pre_act],
                                          ACT[act], _stream()))
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
            M, K = a.reshape(-1, a.shape[-1]).shape
            N = w.shape[0]
            by = M * K * a.element_size() + N * K * w.element_size() + M * N * out.element_size()
            ep = ("+" + kw["act"] if kw.get("act") else "") + ("+gate" if kw.get("gate") is not None else "") + (
                "+res" if kw.get("residual") is not None else "")
            return 2.0 * M * N * K, by, f"gemm_tcgen05|{M}x{N}x{K}{ep}"

        def fmha_cost(out, args, kw):
            q, kk = args[0], args[1]
            B, Lq, H, D = q.shape
            Lk = kk.shape[1]
            return 4.0 * B * H * Lq * Lk * D, 2 * (2 * B * Lq * H * D + 2 * B * Lk * H * D), f"fmha_tcgen05|{B}x{H}x{Lq}x{Lk}x{D}"

        def io_cost(tag):
            def f(out, args, kw):
                ts = [t for t in args if isinstance(t, torch.Tensor)] + [out]
                return 0.0, sum(t.numel() * t.element_size() for t in ts), tag
            return f

        table = {"gemm": gemm_cost, "fmha": fmha_cost, "layernorm": io_cost("layernorm"),
                 "rmsnorm_rope_": io_cost("rmsnorm_rope"), "modulation": io_cost("small"), "skinny_linear": io_cost("small"),
                 "timestep_features": io_cost("small"), "patchify": io_cost("small"), "unpatchify": io_cost("small"),
                 "cfg_combine": io_cost("small"), "axpby_n": io_cost("small")}
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
