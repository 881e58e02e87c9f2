"""GPU parity of the individual sm_100a kernels (called through the C ABI) against plain fp32 PyTorch.

Tolerances: bf16 GEMM / attention outputs are compared with fp32 math on the same bf16-rounded
inputs; the bound is a few bf16 ulps of the output magnitude (stated per test).
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from vist3a_b200 import ops as _ops

    return _ops


def _rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _gelu_tanh(x):
    return torch.nn.functional.gelu(x, approximate="tanh")


@pytest.mark.parametrize("two_cta", [False, True])
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 256, 128), (4096, 1536, 1536), (1000, 520, 200), (77, 1536, 4096),
                                   (4096, 8960, 1536), (300, 64, 8960)])
def test_gemm_bf16_plain(ops, M, N, K, two_cta):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    ref = a.float() @ w.float().t()
    out = ops.gemm(a, w, out_dtype=torch.float32, two_cta=two_cta)
    torch.cuda.synchronize()
    assert _rel_l2(out, ref) < 2e-5  # fp32 accumulation of exact bf16 products
    out_bf = ops.gemm(a, w, two_cta=two_cta)
    assert _rel_l2(out_bf.float(), ref) < 4e-3  # one bf16 rounding


@pytest.mark.parametrize("act", ["gelu_tanh", "gelu_erf", "silu", "relu"])
def test_gemm_bias_act(ops, act):
    M, N, K = 512, 768, 256
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    pre = a.float() @ w.float().t() + b
    fn = {"gelu_tanh": _gelu_tanh, "gelu_erf": torch.nn.functional.gelu, "silu": torch.nn.functional.silu,
          "relu": torch.relu}[act]
    ref = fn(pre)
    out = ops.gemm(a, w, b, act=act, out_dtype=torch.float32)
    # tanh.approx / __expf in the epilogue: absolute error ~1e-3 of O(1) values
    assert float((out - ref).abs().max()) < 4e-3
    assert _rel_l2(out, ref) < 1e-3


def test_gemm_gate_residual_fp32_stream(ops):
    # DiT: x32 = x + gate[b] * bf16(linear);   aggregator: x32 = x32 + bf16(ls * bf16(linear))
    B, Lr, N, K = 2, 640, 512, 384
    g = torch.Generator(device="cuda").manual_seed(2)
    a = torch.randn(B * Lr, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    gate = torch.randn(B, N, device="cuda", generator=g)
    res = torch.randn(B * Lr, N, device="cuda", generator=g)
    lin = (a.float() @ w.float().t() + bias).bfloat16().float()
    ref = res + lin * gate.repeat_interleave(Lr, 0)
    out = ops.gemm(a, w, bias, gate=gate, gate_bstride=N, rows_per_batch=Lr, residual=res, round_linear=True,
                   out_dtype=torch.float32)
    assert float((out - ref).abs().max()) < 5e-2 and _rel_l2(out, ref) < 2e-3
    # in place on the residual
    res2 = res.clone()
    ops.gemm(a, w, bias, gate=gate, gate_bstride=N, rows_per_batch=Lr, residual=res2, out=res2, round_linear=True)
    assert torch.equal(res2, out)
    # bf16 stream with LayerScale (DINO blocks): x = x + bf16(ls * bf16(linear))
    ls = torch.randn(N, device="cuda", generator=g)
    resb = res.bfloat16()
    refb = (resb.float() + (lin * ls).bfloat16().float()).bfloat16()
    outb = ops.gemm(a, w, bias, gate=ls, gate_bstride=0, residual=resb, round_linear=True, round_gate=True)
    assert _rel_l2(outb.float(), refb.float()) < 4e-3


@pytest.mark.parametrize("M,N,K,f32out", [(8192, 1536, 512, False), (8192, 1528, 256, True), (8192, 1536, 1536, False), (4000, 1400, 320, True)])
def test_gemm_176_wide_tiles_match_256_wide(ops, M, N, K, f32out):
    """N ~ 1536 on 74 CTA pairs: the 176-wide column tiles of the A/B flag (4 full waves instead of 2.6 of 256-wide ones) give the same
    results, incl. the 16-column last chunk of a tile, the partial last tile, bias / gate / residual epilogues and both output types."""
    g = torch.Generator(device="cuda").manual_seed(N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    gate = torch.randn(2, N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    res = res if f32out else res.bfloat16()
    kw = dict(gate=gate, gate_bstride=N, rows_per_batch=M // 2, residual=res, round_linear=True, two_cta=True,
              out_dtype=torch.float32 if f32out else torch.bfloat16)
    o176 = ops.gemm(a, w, bias, bn176=True, **kw)
    o256 = ops.gemm(a, w, bias, **kw)
    assert torch.equal(o176, o256)   # same MMAs per element (K order identical), same epilogue arithmetic
    lin = (a.float() @ w.float().t() + bias).bfloat16().float()
    ref = res.float() + lin * gate.repeat_interleave(M // 2, 0)
    assert _rel_l2(o176.float(), ref) < (2e-3 if f32out else 6e-3)


@pytest.mark.parametrize("M,N,K,f32", [(8192, 1536, 1536, False), (8192, 4608, 512, False), (8192, 1536, 1024, True), (13377, 1024, 512, True)])
def test_gemm_multicast_and_staged_variants_are_bit_identical(ops, M, N, K, f32):
    """A-operand multicast between two CTA pairs (VIST3A_GEMM_FLAG_MULTICAST: the DiT block GEMMs use it) and the staged epilogue (A/B flag) run
    the same MMAs in the same K order and the same epilogue arithmetic as the default path: results are bit-identical, with and without the
    gate + residual epilogue, for both output types (256-bit direct stores vs the staging buffer)."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    gate = torch.randn(2, N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    res = res if f32 else res.bfloat16()
    dt = torch.float32 if f32 else torch.bfloat16
    for kw in (dict(), dict(gate=gate, gate_bstride=N, rows_per_batch=(M + 1) // 2, residual=res, round_linear=True), dict(act="gelu_tanh")):
        base = ops.gemm(a, w, bias, out_dtype=dt, two_cta=True, **kw)
        assert torch.equal(ops.gemm(a, w, bias, out_dtype=dt, two_cta=True, multicast=True, **kw), base)
        assert torch.equal(ops.gemm(a, w, bias, out_dtype=dt, two_cta=True, staged=True, **kw), base)
        assert torch.equal(ops.gemm(a, w, bias, out_dtype=dt, two_cta=True, staged=True, multicast=True, **kw), base)
    lin = (a.float() @ w.float().t() + bias)
    assert _rel_l2(ops.gemm(a, w, bias, out_dtype=dt, two_cta=True, multicast=True).float(), lin) < 4e-3


@pytest.mark.parametrize("M,N,K", [(1024, 256, 512), (500, 136, 72), (4096, 32, 1152)])
def test_gemm_tf32(ops, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    ref = a.double() @ w.double().t()
    out = ops.gemm(a, w)
    assert out.dtype == torch.float32
    assert _rel_l2(out, ref) < 1.5e-3  # tf32 operands: 10-bit mantissa


def test_gemm_strided_views(ops):
    # A is a column slice of a wider buffer; C is a column slice of a fused output
    g = torch.Generator(device="cuda").manual_seed(4)
    big = torch.randn(700, 1024, device="cuda", generator=g).bfloat16()
    a = big[:, 256:768]
    w = (torch.randn(384, 512, device="cuda", generator=g) / 22.0).bfloat16()
    outbuf = torch.zeros(700, 1152, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, out=outbuf[:, 384:768])
    ref = a.float() @ w.float().t()
    assert _rel_l2(outbuf[:, 384:768].float(), ref) < 4e-3
    assert float(outbuf[:, :384].abs().max()) == 0 and float(outbuf[:, 768:].abs().max()) == 0


def test_gemm_rejects_bad_arguments(ops):
    from vist3a_b200._lib import Vist3aError

    a = torch.randn(64, 60, device="cuda").bfloat16()  # K stride not 16-byte aligned
    w = torch.randn(64, 60, device="cuda").bfloat16()
    with pytest.raises(Vist3aError):
        ops.gemm(a, w)
    with pytest.raises(RuntimeError):
        ops.gemm(torch.randn(8, 8).bfloat16(), torch.randn(8, 8).bfloat16())


@pytest.mark.parametrize("B,H,Lq,Lk,D", [(1, 2, 128, 128, 128), (1, 2, 256, 384, 64), (2, 3, 300, 77, 128), (1, 4, 1029, 1029, 64),
                                         (1, 12, 4096, 4096, 128), (1, 12, 4096, 512, 128), (3, 16, 1029, 1029, 64)])
def test_fmha(ops, B, H, Lq, Lk, D):
    g = torch.Generator(device="cuda").manual_seed(Lq + Lk + D)
    # slices of one fused [B, L, 3, H, D] buffer, as the block code uses them
    qkv = torch.randn(B, max(Lq, Lk), 3, H, D, device="cuda", generator=g).bfloat16()
    q, k, v = qkv[:, :Lq, 0], qkv[:, :Lk, 1], qkv[:, :Lk, 2]
    ref = torch.nn.functional.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2),
                                                           v.float().transpose(1, 2)).transpose(1, 2)
    out = ops.fmha(q, k, v)
    torch.cuda.synchronize()
    assert out.shape == (B, Lq, H, D)
    err = float((out.float() - ref).abs().max())
    assert err < 2e-2, err  # P and O rounded to bf16; values are O(1)
    assert _rel_l2(out.float(), ref) < 8e-3


# every selectable attention variant (vist3a_fmha_args.flags; A/B measurements in tools/fmha_variants.py, tools/fmha_pair_check.py) computes the
# same function: 0 = default dispatch, 8192 = the former default (two threads per row, exact running maximum), 128 | np << 3 = speculative
# softmax with one thread per row, | 2 = aliased 128-key steps with the half hand-off of P, 16384 = speculative with two threads per row,
# 65536 | np << 3 = ONE query tile per CTA with two threads per row, 256 | v << 9 = CTA-pair kernel variant v (head_dim 128)
@pytest.mark.parametrize("flags", [0, 8192, 128, 128 | 16, 128 | 2 | 16, 16384, 65536, 65536 | 16, 256, 256 | (4 << 9), 256 | (7 << 9)])
@pytest.mark.parametrize("B,H,Lq,Lk,D", [(2, 3, 300, 333, 128), (1, 2, 257, 129, 128), (2, 3, 1029, 1029, 64), (1, 4, 640, 1100, 64)])
def test_fmha_variants_agree_with_sdpa(ops, flags, B, H, Lq, Lk, D):
    if D == 64 and (flags & (256 | 2)):
        pytest.skip("CTA-pair kernel and aliased 128-key steps exist for head_dim 128 only")
    g = torch.Generator(device="cuda").manual_seed(B + Lq + D)
    q = torch.randn(B, Lq, H, D, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, Lk, H, D, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Lk, H, D, device="cuda", generator=g).bfloat16()
    q[:, : Lq // 2] *= 4   # large logits on half the rows
    k[:, Lk // 2:] *= 2    # ... growing along the keys: the running maximum moves by more than the lazy-rescale threshold
    ref = torch.nn.functional.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2)).transpose(1, 2)
    out = ops.fmha(q, k, v, flags=flags)
    assert torch.isfinite(out).all()
    assert _rel_l2(out.float(), ref) < 6e-3, flags


@pytest.mark.parametrize("flags", [0, 8192, 128 | 16, 16384, 16384 | 16, 65536 | 16, 65536 | 24, 256 | (7 << 9)])
@pytest.mark.parametrize("D", [64, 128])
def test_fmha_outlier_keys_late_in_the_sequence(ops, flags, D):
    """A few keys far down the sequence whose scores exceed everything before them by more than 2^127 (after scaling): the speculative softmax
    has already exponentiated them against the stale maximum -- on the MUFU that is +inf, on the FMA-pipe path a clamped 2^127 -- and must take
    its exact path (rescale O and the row sum, redo the half-step); columns chosen to hit MUFU and polynomial positions alike."""
    if D == 64 and (flags & 256):
        pytest.skip("CTA-pair kernel: head_dim 128 only")
    g = torch.Generator(device="cuda").manual_seed(D + flags)
    B, H, Lq, Lk = 1, 2, 300, 700
    q = torch.randn(B, Lq, H, D, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, Lk, H, D, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Lk, H, D, device="cuda", generator=g).bfloat16()
    for j in (Lk - 3, Lk - 10, Lk - 21, Lk - 150, 400, 401, 402, 403):
        k[:, j] *= 300.0
    ref = torch.nn.functional.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2)).transpose(1, 2)
    out = ops.fmha(q, k, v, flags=flags)
    assert torch.isfinite(out).all()
    assert _rel_l2(out.float(), ref) < 6e-3, flags


@pytest.mark.parametrize("shape", [(1, 2, 600, 2048), (1, 15, 1024, 4096), (2, 3, 700, 2000), (1, 12, 2048, 4096), (2, 10, 4096, 1100), (1, 19, 2048, 1024)])
@pytest.mark.parametrize("flags", [0, 256 | (3 << 9), 256 | (4 << 9), 256])
def test_fmha_key_split_last_wave(ops, shape, flags):
    """head_dim 128 on CTA pairs: the query blocks of the grid's last, partly filled wave are laid end to end and cut along the keys into one
    equal range per cluster (partial results in a workspace, a merge kernel follows).  Shapes with fewer units than clusters (pieces of a few
    key steps, many per unit), with whole waves in front of the tail, with a ragged last key tile, with per-row logit scales and with the
    largest logits in the LAST piece; flags select the speculative / exact-maximum kernels.  Compared with fp32 SDPA and with the unsplit
    kernel (flags bit 17)."""
    from vist3a_b200 import _lib
    B, H, Lq, Lk = shape
    g = torch.Generator(device="cuda").manual_seed(Lq + Lk + flags)
    q = torch.randn(B, Lq, H, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, Lk, H, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Lk, H, 128, device="cuda", generator=g).bfloat16()
    q[:, : Lq // 2] *= 3
    k[:, Lk - 100:] *= 2
    rs = (torch.rand(B * Lq, device="cuda", generator=g) + 0.5).float()
    ref = torch.nn.functional.scaled_dot_product_attention((q.float() * rs.view(B, Lq, 1, 1)).transpose(1, 2), k.float().transpose(1, 2),
                                                           v.float().transpose(1, 2)).transpose(1, 2)
    n0 = _lib.launch_count()
    out = ops.fmha(q, k, v, flags=flags, q_row_scale=rs)
    assert _lib.launch_count() - n0 == 2, "attention kernel + merge kernel expected (the shape is chosen to split)"
    whole = ops.fmha(q, k, v, flags=flags | (1 << 17), q_row_scale=rs)
    assert torch.isfinite(out).all()
    assert _rel_l2(out.float(), ref) < 6e-3
    assert _rel_l2(out.float(), whole.float()) < 4e-3


@pytest.mark.parametrize("shape", [(2, 3, 1029, 1029, 64), (1, 2, 261, 700, 128), (1, 2, 264, 300, 64), (3, 5, 513, 1029, 64), (1, 16, 4101, 4200, 64)])
def test_fmha_tail_rows(ops, shape):
    """len_q a few rows more than a multiple of the CTA's 256 query rows (the decoder's frames: 1029 tokens).  Opt-in variant (flags bit 21,
    measured no faster than the default, see fmha_sm100.cu): the tensor-core CTAs cover the multiple, extra CTAs of the same launch compute the
    last 1..8 rows on the CUDA cores.  With per-row logit scales; against fp32 SDPA, the tail rows on their own, and against the default."""
    B, H, Lq, Lk, D = shape
    g = torch.Generator(device="cuda").manual_seed(Lq + Lk + D)
    q = torch.randn(B, Lq, H, D, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, Lk, H, D, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Lk, H, D, device="cuda", generator=g).bfloat16()
    q[:, -3:] *= 3
    rs = (torch.rand(B * Lq, device="cuda", generator=g) + 0.5).float()
    ref = torch.nn.functional.scaled_dot_product_attention((q.float() * rs.view(B, Lq, 1, 1)).transpose(1, 2), k.float().transpose(1, 2),
                                                           v.float().transpose(1, 2)).transpose(1, 2)
    out = torch.full((B, Lq + 2, H, D), 7.0, device="cuda", dtype=torch.bfloat16)   # two guard rows behind the output
    ops.fmha(q, k, v, out=out[:, :Lq], flags=1 << 21, q_row_scale=rs)
    default = ops.fmha(q, k, v, q_row_scale=rs)
    assert (out[:, Lq:] == 7.0).all(), "rows behind len_q were written"
    o = out[:, :Lq].float()
    t = Lq % 256
    assert _rel_l2(o, ref) < 6e-3
    if D == 64:
        assert _rel_l2(o[:, -t:], ref[:, -t:]) < 4e-3          # the CUDA-core rows (fp32 P: only the bf16 rounding of the output)
    assert _rel_l2(o[:, :-t], default.float()[:, :-t]) == 0.0  # the tensor-core rows do not change
    assert _rel_l2(o[:, -t:], default.float()[:, -t:]) < 6e-3


def test_fmha_pair_decompositions_are_bit_identical_on_whole_units(ops):
    """The CTA-pair kernel's work decomposition does not change the arithmetic of a query block that is processed over all keys: persistent
    clusters (default, key split off), one cluster per block (flags bit 20), and the automatic fall-back to one cluster per block when a
    cluster's item list would exceed its shared-memory table (4800 blocks > 60 x 74) give the same bits."""
    g = torch.Generator(device="cuda").manual_seed(11)
    B, H, Lq, Lk = 10, 60, 4096, 640
    q = torch.randn(B, Lq, H, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, Lk, H, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Lk, H, 128, device="cuda", generator=g).bfloat16()
    big = ops.fmha(q, k, v)                                    # 4800 query blocks: one cluster per block
    for b in (0, 7):
        persistent = ops.fmha(q[b:b + 1], k[b:b + 1], v[b:b + 1], flags=1 << 17)
        per_block = ops.fmha(q[b:b + 1], k[b:b + 1], v[b:b + 1], flags=(1 << 17) | (1 << 20))
        assert torch.equal(persistent, per_block)
        assert torch.equal(persistent, big[b:b + 1])
    ref = torch.nn.functional.scaled_dot_product_attention(q[:1].float().transpose(1, 2), k[:1].float().transpose(1, 2), v[:1].float().transpose(1, 2)).transpose(1, 2)
    assert _rel_l2(big[:1].float(), ref) < 6e-3


def test_fmha_large_scores(ops):
    # rows whose max moves by > 2^8 between kv tiles exercise the lazy-rescale path
    g = torch.Generator(device="cuda").manual_seed(5)
    B, H, L, D = 1, 2, 512, 64
    q = (torch.randn(B, L, H, D, device="cuda", generator=g) * 4).bfloat16()
    k = (torch.randn(B, L, H, D, device="cuda", generator=g) * 4).bfloat16()
    k[:, 300:] *= 3
    v = torch.randn(B, L, H, D, device="cuda", generator=g).bfloat16()
    ref = torch.nn.functional.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2),
                                                           v.float().transpose(1, 2)).transpose(1, 2)
    out = ops.fmha(q, k, v)
    assert torch.isfinite(out).all()
    assert _rel_l2(out.float(), ref) < 1e-2


@pytest.mark.parametrize("din,dout", [(torch.bfloat16, torch.bfloat16), (torch.float32, torch.bfloat16),
                                      (torch.float32, torch.float32)])
@pytest.mark.parametrize("dim", [1024, 1536, 2048])
def test_layernorm_modulated(ops, din, dout, dim):
    g = torch.Generator(device="cuda").manual_seed(6)
    B, Lr = 2, 333
    x = (torch.randn(B * Lr, dim, device="cuda", generator=g) * 2 + 0.5).to(din)
    mul = torch.randn(B, dim, device="cuda", generator=g)
    add = torch.randn(B, dim, device="cuda", generator=g)
    ref = torch.nn.functional.layer_norm(x.float(), (dim,), eps=1e-6) * mul.repeat_interleave(Lr, 0) + add.repeat_interleave(Lr, 0)
    out = ops.layernorm(x, mul=mul, add=add, mul_bstride=dim, add_bstride=dim, rows_per_batch=Lr, eps=1e-6, out_dtype=dout)
    tol = 1e-5 if dout == torch.float32 else 4e-3
    assert _rel_l2(out.float(), ref) < tol
    # affine LayerNorm: shared weight / bias
    out2 = ops.layernorm(x, mul=mul[0], add=add[0], eps=1e-5, out_dtype=dout)
    ref2 = torch.nn.functional.layer_norm(x.float(), (dim,), mul[0], add[0], eps=1e-5)
    assert _rel_l2(out2.float(), ref2) < tol


def test_rmsnorm_rope(ops):
    g = torch.Generator(device="cuda").manual_seed(7)
    L_, H, D = 96, 12, 128
    dim = H * D
    buf = torch.randn(2 * L_, 3 * dim, device="cuda", generator=g).bfloat16()
    x = buf[:, dim:2 * dim]  # the K third of a fused QKV buffer
    w = torch.randn(dim, device="cuda", generator=g)
    ang = torch.rand(L_, D // 2, device="cuda", generator=g) * 6.28
    cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
    xf = x.float()
    n = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6) * w
    n = n.view(2, L_, H, D // 2, 2)
    re, im = n[..., 0], n[..., 1]
    c, s = cos[None, :, None, :], sin[None, :, None, :]
    ref = torch.stack([re * c - im * s, re * s + im * c], -1).reshape(2 * L_, dim)
    keep = buf.clone()
    ops.rmsnorm_rope_(x, w, D, eps=1e-6, cos=cos, sin=sin)
    assert _rel_l2(x.float(), ref) < 4e-3
    assert torch.equal(buf[:, :dim], keep[:, :dim]) and torch.equal(buf[:, 2 * dim:], keep[:, 2 * dim:])
    # no-RoPE variant (cross-attention)
    y = keep[:, :dim].clone()
    ops.rmsnorm_rope_(y, w, D, eps=1e-6)
    yf = keep[:, :dim].float()
    assert _rel_l2(y.float(), yf * torch.rsqrt(yf.pow(2).mean(-1, keepdim=True) + 1e-6) * w) < 4e-3


def test_rmsnorm_rope_two_segments_one_launch(ops):
    """q and k of a fused qkv buffer normalised + rotated by ONE launch == two single-segment launches"""
    g = torch.Generator(device="cuda").manual_seed(8)
    L_, H, D = 64, 4, 128
    dim = H * D
    buf = torch.randn(2 * L_, 3 * dim, device="cuda", generator=g).bfloat16()
    wq, wk = torch.randn(dim, device="cuda", generator=g), torch.randn(dim, device="cuda", generator=g)
    ang = torch.rand(L_, D // 2, device="cuda", generator=g) * 6.28
    cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
    a, b = buf.clone(), buf.clone()
    ops.rmsnorm_rope_(a[:, :dim], wq, D, eps=1e-6, cos=cos, sin=sin)
    ops.rmsnorm_rope_(a[:, dim:2 * dim], wk, D, eps=1e-6, cos=cos, sin=sin)
    ops.rmsnorm_rope_(b[:, :2 * dim], torch.cat([wq, wk]), D, eps=1e-6, cos=cos, sin=sin, nseg=2)
    assert torch.equal(a, b)


def test_fmha_row_scale_is_query_rmsnorm(ops):
    """cross-attention: RMSNorm(q) w_q . k  ==  rinv_q * (q . (w_q k)) with rinv from the read-only row pass"""
    g = torch.Generator(device="cuda").manual_seed(9)
    B, H, Lq, Lk, D = 2, 3, 200, 77, 128
    dim = H * D
    q = (torch.randn(B * Lq, dim, device="cuda", generator=g) * 3).bfloat16()
    k = torch.randn(B, Lk, H, D, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Lk, H, D, device="cuda", generator=g).bfloat16()
    wq = 1 + 0.2 * torch.randn(dim, device="cuda", generator=g)
    rinv = ops.row_rinv(q, eps=1e-6)
    qf = q.float()
    assert _rel_l2(rinv, torch.rsqrt(qf.pow(2).mean(-1) + 1e-6)) < 1e-5
    qn = (qf * torch.rsqrt(qf.pow(2).mean(-1, keepdim=True) + 1e-6) * wq).view(B, Lq, H, D)
    ref = torch.nn.functional.scaled_dot_product_attention(qn.transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2)).transpose(1, 2)
    kw = (k.float() * wq.view(1, 1, H, D)).bfloat16()
    out = ops.fmha(q.view(B, Lq, H, D), kw, v, q_row_scale=rinv)
    assert _rel_l2(out.float(), ref) < 1.2e-2   # bf16 rounding of w_q k and of P


def test_modulation_and_timestep(ops):
    g = torch.Generator(device="cuda").manual_seed(8)
    B, D = 2, 1536
    table = torch.randn(6, D, device="cuda", generator=g)
    mod = torch.randn(B, 6 * D, device="cuda", generator=g)
    out = ops.modulation(table, mod, nvec=6, broadcast=False, one_plus_mask=0b010010)
    ref = table[None] + mod.view(B, 6, D)
    ref[:, 1] += 1
    ref[:, 4] += 1
    assert torch.allclose(out, ref, atol=1e-6)
    temb = torch.randn(B, D, device="cuda", generator=g)
    out2 = ops.modulation(table[:2].contiguous(), temb, nvec=2, broadcast=True, one_plus_mask=0b10)
    ref2 = table[None, :2] + temb[:, None]
    ref2[:, 1] += 1
    assert torch.allclose(out2, ref2, atol=1e-6)
    t = torch.tensor([999.0, 17.5], device="cuda")
    f = ops.timestep_features(t, 256)
    half = 128
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, device="cuda", dtype=torch.float32) / half)
    a = t[:, None] * freqs[None]
    ref3 = torch.cat([a.cos(), a.sin()], -1)
    assert float((f - ref3).abs().max()) < 2e-4


@pytest.mark.parametrize("M", [1, 2, 13])
def test_skinny_linear(ops, M):
    g = torch.Generator(device="cuda").manual_seed(9)
    K, N = 2048, 777
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / 45.0
    b = torch.randn(N, device="cuda", generator=g)
    out = ops.skinny_linear(x, w, b, pre_act="silu", act=None)
    ref = torch.nn.functional.silu(x).double() @ w.double().t() + b.double()
    assert _rel_l2(out, ref) < 1e-5
    wb = w.bfloat16()
    out2 = ops.skinny_linear(x.bfloat16(), wb, b, act="gelu_tanh", out_dtype=torch.float32)
    ref2 = _gelu_tanh(x.bfloat16().double() @ wb.double().t() + b.double())
    assert _rel_l2(out2, ref2) < 1e-4


def test_patchify_roundtrip(ops):
    g = torch.Generator(device="cuda").manual_seed(10)
    B, C_, T, H, W = 2, 16, 4, 64, 64
    x = torch.randn(B, C_, T, H, W, device="cuda", generator=g).bfloat16()
    a = ops.patchify(x)
    ref = x.view(B, C_, T, H // 2, 2, W // 2, 2).permute(0, 2, 3, 5, 1, 4, 6).reshape(B * T * (H // 2) * (W // 2), C_ * 4)
    assert torch.equal(a, ref)
    # unpatchify consumes proj_out rows laid out (dy, dx, c)
    p = torch.randn(B * T * 32 * 32, 64, device="cuda", generator=g).bfloat16()
    u = ops.unpatchify(p, B, C_, T, H, W)
    refu = p.view(B, T, 32, 32, 1, 2, 2, C_).permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(B, C_, T, H, W)
    assert torch.equal(u, refu)


def test_cfg_axpby(ops):
    g = torch.Generator(device="cuda").manual_seed(11)
    c = torch.randn(1, 16, 4, 64, 64, device="cuda", generator=g).bfloat16()
    u = torch.randn(1, 16, 4, 64, 64, device="cuda", generator=g).bfloat16()
    out = ops.cfg_combine(c, u, 6.0)
    assert torch.allclose(out, u.float() + 6.0 * (c.float() - u.float()), atol=1e-5)
    x = torch.randn_like(out)
    y = torch.randn_like(out)
    z = ops.axpby_n(torch.empty_like(out), [out, x, y], [0.5, -1.25, 2.0])
    assert torch.allclose(z, 0.5 * out - 1.25 * x + 2.0 * y, atol=1e-5)


def test_launch_counter(ops):
    from vist3a_b200 import _lib

    n0 = _lib.launch_count()
    ops.cfg_combine(torch.zeros(8, device="cuda"), torch.zeros(8, device="cuda"), 1.0)
    assert _lib.launch_count() == n0 + 1
