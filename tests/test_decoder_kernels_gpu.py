"""GPU parity of the decoder-side kernels (through the C ABI) against plain fp32 PyTorch / the CPU oracle.

Tolerances are stated per test: TF32 implicit-GEMM convolutions are compared with fp64 convolutions of
the same fp32 inputs (10-bit operand mantissa => ~1e-3 relative L2); elementwise kernels are fp32-exact
up to the rounding of one fused multiply-add.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from vist3a_b200 import ops as _ops

    return _ops


def _rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _conv_w(w):
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 64, 256), (3, 32, 32, 256, 256), (1, 37, 23, 32, 64), (2, 128, 128, 256, 128),
                                            (1, 448, 448, 128, 32), (13, 16, 16, 1024, 256)])
def test_conv3x3_implicit_gemm_tf32(ops, n, h, w, cin, cout):
    g = torch.Generator(device="cuda").manual_seed(n * 1000 + h + cin)
    x = torch.randn(n, h, w, cin, device="cuda", generator=g)
    wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / math.sqrt(9 * cin)
    b = torch.randn(cout, device="cuda", generator=g)
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), wt.double(), b.double(), padding=1).permute(0, 2, 3, 1)
    out = torch.empty(n, h, w, cout, device="cuda")
    ops.gemm(x, _conv_w(wt), b, conv=dict(kh=3, kw=3, pad=1), out=out.view(-1, cout))
    torch.cuda.synchronize()
    assert _rel_l2(out, ref) < 1.5e-3


def test_conv3x3_fused_epilogue(ops):
    """ResidualConvUnit tail: relu(conv(x) + b + res + res2)"""
    n, h, w, c = 2, 24, 40, 64
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(n, h, w, c, device="cuda", generator=g)
    wt = torch.randn(c, c, 3, 3, device="cuda", generator=g) / math.sqrt(9 * c)
    b = torch.randn(c, device="cuda", generator=g)
    r1 = torch.randn(n, h, w, c, device="cuda", generator=g)
    r2 = torch.randn(n, h, w, c, device="cuda", generator=g)
    ref = torch.relu(F.conv2d(x.permute(0, 3, 1, 2).double(), wt.double(), b.double(), padding=1).permute(0, 2, 3, 1) + r1 + r2)
    out = torch.empty(n, h, w, c, device="cuda")
    ops.gemm(x, _conv_w(wt), b, conv=dict(kh=3, kw=3, pad=1), out=out.view(-1, c), residual=r1.view(-1, c), residual2=r2.view(-1, c),
             post_act="relu")
    assert _rel_l2(out, ref) < 1.5e-3
    out2 = torch.empty(n, h, w, c, device="cuda")
    ops.gemm(x, _conv_w(wt), b, conv=dict(kh=3, kw=3, pad=1), out=out2.view(-1, c), act="relu")
    ref2 = torch.relu(F.conv2d(x.permute(0, 3, 1, 2).double(), wt.double(), b.double(), padding=1).permute(0, 2, 3, 1))
    assert _rel_l2(out2, ref2) < 1.5e-3


def test_conv_bf16_implicit_gemm(ops):
    n, h, w, cin, cout = 2, 20, 36, 128, 192
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / math.sqrt(9 * cin)).bfloat16()
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), wt.double(), padding=1).permute(0, 2, 3, 1)
    out = torch.empty(n, h, w, cout, device="cuda")
    ops.gemm(x, _conv_w(wt), conv=dict(kh=3, kw=3, pad=1), out=out.view(-1, cout))
    assert _rel_l2(out, ref) < 2e-5


def test_gemm_row_maps(ops):
    """stitching GEMM: rows of 3 groups x 16 land behind 5 special rows of each 21-row group; a [16, N] table is added to every group"""
    G, rpg, P, N, K = 3, 16, 21, 64, 72
    g = torch.Generator(device="cuda").manual_seed(7)
    a = torch.randn(G * rpg, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / 8).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    tab = torch.randn(rpg, N, device="cuda", generator=g)
    out = torch.full((G * P, N), 7.0, device="cuda")
    ops.gemm(a, w, b, out=out, residual=tab, rmap=(rpg, 0, 0), cmap=(rpg, P, 5))
    ref = (a.float() @ w.float().t() + b).view(G, rpg, N) + tab
    o3 = out.view(G, P, N)
    assert _rel_l2(o3[:, 5:], ref) < 1e-5
    assert bool((o3[:, :5] == 7.0).all())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_im2col_stitch(ops, dtype):
    B, C, T, h, w = 2, 16, 3, 12, 8
    g = torch.Generator(device="cuda").manual_seed(8)
    lat = torch.randn(B, C, T, h, w, device="cuda", generator=g).to(dtype)
    V = (T - 1) * 4 + 1
    up = F.interpolate(lat.float(), size=[V, h, w], mode="trilinear", align_corners=True)
    pad = F.pad(up, (1, 1, 1, 1, 2, 2), mode="replicate")
    # unfold: [B, C, V, h/2, w/2, 5, 3, 3]
    u = pad.unfold(2, 5, 1).unfold(3, 3, 2).unfold(4, 3, 2)
    ref = u.permute(0, 2, 3, 4, 1, 5, 6, 7).reshape(B * V * (h // 2) * (w // 2), C * 45)
    out = ops.im2col_stitch(lat)
    assert out.dtype == torch.bfloat16 and out.shape == ref.shape
    assert float((out.float() - ref).abs().max()) <= 2 ** -7 * float(ref.abs().max())  # one bf16 rounding
    # and the conv it feeds
    wt = torch.randn(32, C, 5, 3, 3, device="cuda", generator=g) / math.sqrt(C * 45)
    conv = F.conv3d(pad, wt, stride=(1, 2, 2)).permute(0, 2, 3, 4, 1).reshape(-1, 32)
    got = ops.gemm(out, wt.reshape(32, -1).bfloat16().contiguous(), out_dtype=torch.float32)
    assert _rel_l2(got, conv) < 6e-3


@pytest.mark.parametrize("kh,stride,pad,C,kpad", [(7, 1, 3, 3, 148), (3, 2, 1, 64, None)])
def test_im2col_nhwc(ops, kh, stride, pad, C, kpad):
    n, h, w = 2, 18, 14
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(n, h, w, C, device="cuda", generator=g)
    out = ops.im2col_nhwc(x, kh, kh, stride, pad, k_pad=kpad)
    cols = F.unfold(x.permute(0, 3, 1, 2), kh, padding=pad, stride=stride)  # [n, C*kh*kh, L], index c*kh*kh + dy*kh + dx
    ho = (h + 2 * pad - kh) // stride + 1
    wo = (w + 2 * pad - kh) // stride + 1
    ref = cols.view(n, C, kh * kh, ho * wo).permute(0, 3, 2, 1).reshape(n * ho * wo, kh * kh * C)
    assert torch.equal(out[:, :kh * kh * C], ref)
    if kpad:
        assert float(out[:, kh * kh * C:].abs().max()) == 0.0


@pytest.mark.parametrize("B,V,H,W,dtype", [(1, 2, 20, 36, torch.float32), (2, 3, 56, 56, torch.bfloat16), (1, 1, 224, 232, torch.float32)])
def test_rgb7x7_overlapping_window_conv(ops, B, V, H, W, dtype):
    """7x7 RGB input_merger without im2col: RGB0 image with physically padded rows, TMA boxes over overlapping windows"""
    g = torch.Generator(device="cuda").manual_seed(19)
    img = (torch.rand(B, 3, V, H, W, device="cuda", generator=g) * 2 - 1).to(dtype)
    wt = torch.randn(128, 3, 7, 7, device="cuda", generator=g) / math.sqrt(147)
    b = torch.randn(128, device="cuda", generator=g)
    x01 = (img.float().permute(0, 2, 1, 3, 4).reshape(B * V, 3, H, W) + 1) / 2
    ref = torch.relu(F.conv2d(x01.double(), wt.double(), b.double(), padding=3)).permute(0, 2, 3, 1)
    rgb = ops.rgb_to_nhwc4pad(img)
    assert rgb.shape == (B * V, H, W + 8, 4)
    assert torch.equal(rgb[:, :, 3:3 + W, :3], x01.permute(0, 2, 3, 1)) and float(rgb[:, :, :3].abs().max()) == 0 and float(rgb[..., 3].abs().max()) == 0
    mk = torch.zeros(128, 7, 8, 4, device="cuda")
    mk[:, :, :7, :3] = wt.permute(0, 2, 3, 1)
    out = torch.empty(B * V, H, W, 128, device="cuda")
    ops.gemm(rgb, mk.reshape(128, 224), b, act="relu", out=out.view(-1, 128),
             conv=dict(kh=7, kw=1, pad=3, pad_x=0, geom=(B * V, H, W, 32), strides=(4, (W + 8) * 4, H * (W + 8) * 4)))
    assert _rel_l2(out, ref) < 1.5e-3


def test_qknorm_rope2d_matches_oracle(ops):
    from oracle import decoder_ref as D

    BV, gh, gw, Hn = 3, 4, 6, 2
    P = gh * gw + 5
    C = Hn * 64
    g = torch.Generator().manual_seed(10)
    qkv = torch.randn(BV * P, 3 * C, generator=g).bfloat16()
    qw, qb, kw, kb = (1 + 0.1 * torch.randn(64, generator=g), 0.1 * torch.randn(64, generator=g), 1 + 0.1 * torch.randn(64, generator=g),
                      0.1 * torch.randn(64, generator=g))
    yy, xx = torch.meshgrid(torch.arange(gh), torch.arange(gw), indexing="ij")
    pos = torch.stack([yy.reshape(-1), xx.reshape(-1)], -1)[None].expand(BV, -1, -1) + 1
    pos = torch.cat([torch.zeros(BV, 5, 2, dtype=pos.dtype), pos], dim=1)
    q5 = qkv.float().view(BV, P, 3, Hn, 64).permute(2, 0, 3, 1, 4)
    q = D._rope2d(F.layer_norm(q5[0], (64,), qw, qb, 1e-5), pos)
    k = D._rope2d(F.layer_norm(q5[1], (64,), kw, kb, 1e-5), pos)
    inv = 1.0 / (100.0 ** (torch.arange(0, 32, 2).float() / 32))
    ang = torch.arange(40).float()[:, None] * inv[None]
    dev = qkv.cuda()
    ops.qknorm_rope2d_(dev, Hn, qw.cuda(), qb.cuda(), kw.cuda(), kb.cuda(), ang.cos().cuda().contiguous(), ang.sin().cuda().contiguous(),
                       tokens_per_view=P, n_special=5, grid_w=gw)
    got = dev.float().cpu().view(BV, P, 3, Hn, 64).permute(2, 0, 3, 1, 4)
    assert float((got[0] - q).abs().max()) < 3e-2 and _rel_l2(got[0], q) < 4e-3   # bf16 output rounding
    assert float((got[1] - k).abs().max()) < 3e-2 and _rel_l2(got[1], k) < 4e-3
    assert torch.equal(got[2], q5[2])  # v untouched


@pytest.mark.parametrize("hi,wi,ho,wo", [(16, 16, 32, 32), (256, 256, 448, 448), (9, 13, 20, 17)])
def test_bilinear_nhwc(ops, hi, wi, ho, wo):
    n, C = 2, 32
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(n, hi, wi, C, device="cuda", generator=g)
    add = torch.randn(n, ho, wo, C, device="cuda", generator=g)
    px = torch.randn(wo, C // 2, device="cuda", generator=g)
    py = torch.randn(ho, C // 2, device="cuda", generator=g)
    ref = F.interpolate(x.permute(0, 3, 1, 2), size=(ho, wo), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    out = ops.bilinear_nhwc(x, ho, wo)
    assert float((out - ref).abs().max()) < 2e-5
    pos = torch.cat([px[None].expand(ho, wo, -1), py[:, None].expand(ho, wo, -1)], -1)
    out2 = ops.bilinear_nhwc(x, ho, wo, add=add, pos_x=px, pos_y=py)
    assert float((out2 - (ref + add + pos)).abs().max()) < 3e-5


@pytest.mark.parametrize("k", [2, 4])
def test_conv_transpose_as_gemm_plus_depth_to_space(ops, k):
    n, h, w, C = 2, 6, 5, 32
    g = torch.Generator(device="cuda").manual_seed(12)
    x = torch.randn(n, h, w, C, device="cuda", generator=g)
    wt = torch.randn(C, C, k, k, device="cuda", generator=g) / math.sqrt(C)  # [in, out, kh, kw]
    b = torch.randn(C, device="cuda", generator=g)
    ref = F.conv_transpose2d(x.permute(0, 3, 1, 2).double(), wt.double(), b.double(), stride=k).permute(0, 2, 3, 1)
    wg = wt.permute(2, 3, 1, 0).reshape(k * k * C, C).contiguous()
    y = ops.gemm(x.view(-1, C), wg, b.repeat(k * k))
    out = ops.depth_to_space(y, n, h, w, C, k)
    assert out.shape == ref.shape and _rel_l2(out, ref) < 1.5e-3


def test_attention_small_and_fma_rows(ops):
    B, Lq, H, Dh = 2, 13, 4, 128
    g = torch.Generator(device="cuda").manual_seed(13)
    qkv = torch.randn(B * Lq, 3 * H * Dh, device="cuda", generator=g)
    q, k, v = qkv.view(B, Lq, 3, H, Dh).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * Lq, H * Dh)
    out = ops.attention_small(qkv, B, Lq, H, Dh)
    assert _rel_l2(out, ref) < 1e-5
    a, b, c = (torch.randn(7, 3 * 64, device="cuda", generator=g) for _ in range(3))
    got = ops.fma_rows(a[:, 64:128], b[:, :64], c[:, 128:])
    assert torch.allclose(got, a[:, 64:128] * b[:, :64] + c[:, 128:], atol=1e-6)


def test_skinny_linear_fp32_gate_residual(ops):
    M, N, K = 13, 200, 256
    g = torch.Generator(device="cuda").manual_seed(14)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / 16
    b, gate = torch.randn(N, device="cuda", generator=g), torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    out = ops.skinny_linear(x, w, b, gate=gate, residual=res, out_dtype=torch.float32)
    ref = res.double() + gate.double() * (x.double() @ w.double().t() + b.double())
    assert _rel_l2(out, ref) < 1e-6


@pytest.mark.parametrize("N,K", [(6144, 2048), (2048, 8192), (384, 128), (9, 1024)])
def test_linear_tokens16_weight_streaming(ops, N, K):
    """camera-head linears: weights are the streamed TMA operand (TF32), 13 tokens padded to 16"""
    M = 13
    g = torch.Generator(device="cuda").manual_seed(17)
    x16 = torch.zeros(16, K, device="cuda")
    x16[:M] = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    b, gate = torch.randn(N, device="cuda", generator=g), torch.randn(N, device="cuda", generator=g)
    res = torch.zeros(16, N, device="cuda")
    res[:M] = torch.randn(M, N, device="cuda", generator=g)
    ref = res[:M].double() + gate.double() * F.gelu(x16[:M].double() @ w.double().t() + b.double())
    out = ops.linear_tokens16(x16, M, w, b, act="gelu_erf", gate=gate, residual=res)
    assert out.shape == (16, N) and float(out[M:].abs().max()) == 0.0
    assert _rel_l2(out[:M], ref) < 1e-3          # tf32 operands
    res2 = res.clone()
    ops.linear_tokens16(x16, M, w, b, act="gelu_erf", gate=gate, residual=res2, out=res2)   # in place on the residual
    assert torch.equal(res2[:M], out[:M])


def test_pose_to_cameras_matches_oracle(ops):
    from oracle import decoder_ref as D

    g = torch.Generator().manual_seed(15)
    raw = torch.randn(1, 13, 9, generator=g)
    raw[..., 6] += 1.0
    raw[..., 7:] = raw[..., 7:].abs() + 0.3
    raw[0, 0, 7] = -0.2  # exercised relu on the FoV entries
    act = torch.cat([raw[..., :7], torch.relu(raw[..., 7:])], -1)
    extr, intr = D.pose_to_cameras(act, (448, 448))
    out = ops.pose_to_cameras(raw.cuda().view(13, 9), 448, 448)
    assert torch.allclose(out["pose_act"].cpu(), act[0], atol=1e-7)
    assert torch.allclose(out["extr"].cpu(), extr[0], atol=2e-6)
    assert _rel_l2(out["intr"].cpu(), intr[0]) < 1e-5
    pad = torch.tensor([0.0, 0, 0, 1]).view(1, 1, 4).repeat(13, 1, 1)
    c2w = torch.cat([extr[0], pad], dim=1).inverse()
    assert torch.allclose(out["c2w"].cpu(), c2w, atol=2e-5)
    inorm = torch.stack([intr[0, :, 0] / 448, intr[0, :, 1] / 448, intr[0, :, 2]], dim=1)
    assert _rel_l2(out["intr_norm"].cpu(), inorm) < 1e-5


def test_gaussian_epilogue_matches_oracle(ops):
    from oracle import decoder_ref as D

    cfg = D.FULL
    S, H, W, cd = 2, 12, 20, 32
    g = torch.Generator().manual_seed(16)
    dfeat = torch.randn(S * H * W, cd, generator=g).abs()
    dw = torch.randn(cd, generator=g) / cd
    db = 0.3
    raw = torch.randn(S * H * W, 84, generator=g)
    raw[:7, 1:4] = 25.0   # softplus threshold branch + clamp at 0.3 (needs 0.001*x > 0.3 => not reached; threshold only)
    raw[7:9, 1:4] = 1e4   # clamp
    pose = torch.randn(1, S, 9, generator=g)
    pose[..., 6] += 1.0
    pose[..., 7:] = pose[..., 7:].abs() + 0.4
    extr, intr = D.pose_to_cameras(pose, (H, W))
    depth = torch.exp(dfeat @ dw + db).view(S, H, W)
    pts = D.unproject(depth, extr[0], intr[0])
    ref = D.gaussian_adapter(cfg, pts.reshape(1, -1, 3), raw[None, :, :83])
    o = ops.gaussian_epilogue(dfeat.cuda(), dw.cuda(), db, raw.cuda(), extr[0].cuda().contiguous(), intr[0].cuda().contiguous(),
                              D.sh_mask(cfg).cuda(), S, H, W)
    assert _rel_l2(o["depth"].cpu(), depth.reshape(-1)) < 1e-5
    assert _rel_l2(o["means"].cpu(), ref["means"][0]) < 1e-5
    assert _rel_l2(o["scales"].cpu(), ref["scales"][0]) < 1e-5
    assert _rel_l2(o["rotations"].cpu(), ref["rotations"][0]) < 1e-5
    assert _rel_l2(o["opacities"].cpu(), ref["opacities"][0]) < 1e-5
    assert _rel_l2(o["harmonics"].cpu(), ref["harmonics"][0]) < 1e-6
    assert _rel_l2(o["covariances"].cpu(), ref["covariances"][0]) < 2e-5
    scene = float(o["scene_sum"]) / (S * H * W)
    assert abs(scene - float(pts.norm(dim=-1).mean())) < 1e-4 * scene
