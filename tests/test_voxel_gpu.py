"""GPU parity of the voxelised Gaussian fusion (vist3a_voxel_fusion / vist3a_gaussian_adapter through the C ABI) against the
oracle (oracle/decoder_ref.py:voxelize_with_fusion, pinned bit-exactly to the reference's voxelizaton_with_fusion,
AS/model/encoder/anysplat.py:298-335) and the golden vectors the real reference produced (tests/golden/voxel_fusion.pt).

Bar: voxel assignment, order, inverse index and counts are integer work -> bit-exact.  The fused positions / features are fp32
sums of softmax weights: same operation order as the reference's CPU path (stable sort keeps members in point order), differences
come from expf only -> tolerance 2e-6 absolute + 2e-6 relative per element (stated here; the reference's own GPU path sums with
atomics in arbitrary order)."""
import importlib.util
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "voxel_fusion.pt")
_spec = importlib.util.spec_from_file_location("make_voxel_golden", os.path.join(HERE, "golden", "make_voxel_golden.py"))
MG = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(MG)
GAUSS = ("means", "covariances", "harmonics", "opacities", "scales", "rotations")


def _close(name, got, want, atol=2e-6, rtol=2e-6):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    assert got.shape == want.shape, (name, got.shape, want.shape)
    err = (got - want).abs()
    bad = err > atol + rtol * want.abs()
    assert not bool(bad.any()), f"{name}: {int(bad.sum())} of {bad.numel()} elements off, max abs err {float(err.max()):.3e}"


def _flat(args):
    feat, pts, conf = MG.voxel_case_inputs(*args)
    C = feat.shape[1]
    return feat.permute(0, 2, 3, 1).reshape(-1, C).contiguous(), pts.permute(0, 2, 3, 1).reshape(-1, 3).contiguous(), conf.flatten().contiguous(), args[6]


@pytest.mark.parametrize("case", ["spread", "dense", "halfway", "single"])
def test_voxel_fusion_matches_reference_golden(case):
    from vist3a_b200 import ops

    g = torch.load(GOLD)["cases"][case]
    feats, pts, conf, vs = _flat(g["args"])
    o = ops.voxel_fusion(pts.cuda(), feats.cuda(), conf.cuda(), vs, want_index=True)
    assert o["n_voxels"] == g["voxel_pts"].shape[0]
    assert torch.equal(o["inverse"].cpu(), g["inverse"]) and torch.equal(o["counts"].cpu(), g["counts"])
    _close("voxel_pts", o["pts"], g["voxel_pts"])
    _close("voxel_feats", o["feats"], g["voxel_feats"])


@pytest.mark.parametrize("n,c,spread,vs", [(1, 83, 1.0, 0.002), (4097, 83, 0.5, 0.01), (200_000, 83, 2.0, 0.02), (65_536, 5, 40.0, 0.002),
                                           (50_000, 125, 1.0, 0.05)])
def test_voxel_fusion_matches_oracle(n, c, spread, vs):
    """ragged sizes (one point, one past a sort tile), wide coordinate ranges (more radix passes), widest feature rows"""
    from oracle import decoder_ref as D
    from vist3a_b200 import ops

    g = torch.Generator().manual_seed(n + c)
    pts = torch.randn(n, 3, generator=g) * spread
    rows = torch.randn(n, c + 1, generator=g)  # confidence = last column of the same rows (as in the decoder's raw maps)
    rows[:, c] *= 4
    vp, vf, inv, cnt = D.voxelize_with_fusion(rows[:, :c], pts, vs, rows[:, c])
    rg = rows.cuda()
    o = ops.voxel_fusion(pts.cuda(), rg, rg[:, c], vs, feat_dim=c, want_index=True)
    assert o["n_voxels"] == vp.shape[0]
    assert torch.equal(o["inverse"].cpu().long(), inv) and torch.equal(o["counts"].cpu().long(), cnt)
    _close("voxel_pts", o["pts"], vp)
    _close("voxel_feats", o["feats"], vf)


def test_voxel_fusion_degenerate_inputs():
    from vist3a_b200 import ops

    # all points identical -> one voxel holding the (equal-confidence) mean; zero-extent coordinate ranges (0 key bits, no sort pass)
    n = 10_000
    pts = torch.full((n, 3), 0.1234, device="cuda")
    feats = torch.arange(n, dtype=torch.float32, device="cuda").view(n, 1).repeat(1, 3).contiguous()
    conf = torch.zeros(n, device="cuda")
    o = ops.voxel_fusion(pts, feats, conf, 0.002, want_index=True)
    assert o["n_voxels"] == 1 and int(o["counts"][0]) == n and int(o["inverse"].max()) == 0
    assert abs(float(o["feats"][0, 0]) - (n - 1) / 2) < 1e-2 * n
    # huge ranges on every axis: 3 x 32 bits do not fit a 64-bit key -> reported, not mis-sorted
    big = torch.tensor([[-2.0e6, -2.0e6, -2.0e6], [2.0e6, 2.0e6, 2.0e6]], device="cuda")
    with pytest.raises(RuntimeError, match="64 key bits"):
        ops.voxel_fusion(big, feats[:2].contiguous(), conf[:2].contiguous(), 0.001)
    # CPU tensors are rejected (no fallback)
    with pytest.raises(RuntimeError):
        ops.voxel_fusion(pts.cpu(), feats.cpu(), conf.cpu(), 0.002)


def test_voxel_fusion_full_size_properties():
    """BASELINE size (13 views x 448^2 = 2 609 152 points, 83 features): properties that do not need the CPU oracle, plus
    torch.unique on the device as an independent check of the integer part."""
    from vist3a_b200 import ops

    n, c, vs = 13 * 448 * 448, 83, 0.002
    g = torch.Generator(device="cuda").manual_seed(7)
    # points on a few thousand surfaces so that voxels hold 1..many points
    pts = (torch.randn(n, 3, device="cuda", generator=g) * 0.05).contiguous()
    rows = torch.randn(n, c + 1, device="cuda", generator=g)
    rows[:, c] = 0.0  # equal confidences: every voxel is the plain mean of its members
    o = ops.voxel_fusion(pts, rows, rows[:, c], vs, feat_dim=c, want_index=True)
    m = o["n_voxels"]
    # IEEE division as on the reference's CPU path (the oracle): torch's CUDA `tensor / python_scalar` multiplies by the rounded
    # reciprocal instead, which moves a handful of points per million across a cell border
    cells = torch.div(pts, torch.full_like(pts, vs)).round().int()
    uniq, inv, cnt = torch.unique(cells, dim=0, return_inverse=True, return_counts=True)
    assert m == uniq.shape[0] and 0 < m < n
    assert torch.equal(o["inverse"].long(), inv) and torch.equal(o["counts"].long(), cnt)
    assert int(o["counts"].sum()) == n
    # every fused position lies in (or on the border of) its own cell
    assert bool((((o["pts"] / vs) - uniq.float()).abs() <= 0.5 + 1e-3).all())
    # checksum of checksums: sum_v count_v * mean_v == sum_i x_i  (weights are 1 / (count + 1e-6))
    lhs = (o["feats"].double() * cnt.double()[:, None]).sum(0)
    rhs = rows[:, :c].double().sum(0)
    assert float((lhs - rhs).abs().max()) < 1e-3 * float(rhs.abs().max() + n ** 0.5)
    # idempotence: fusing the fused voxels (one point per cell now, unless the mean left its cell) changes nothing but the 1e-6 weight bias
    o2 = ops.voxel_fusion(o["pts"].contiguous(), o["feats"].contiguous(), torch.zeros(m, device="cuda"), vs)
    assert abs(o2["n_voxels"] - m) <= 1e-3 * m


def _engine(sd, ocfg, resolution, **kw):
    from vist3a_b200.stitched_decoder import DecoderConfig, StitchVAE3DB200

    cfg = DecoderConfig(embed_dim=ocfg.embed_dim, num_heads=ocfg.num_heads, dino_blocks=ocfg.dino_blocks, agg_depth=ocfg.agg_depth,
                        cam_heads=ocfg.cam_heads, cam_trunk=ocfg.cam_trunk, dpt_features=ocfg.dpt_features,
                        dpt_out_channels=ocfg.dpt_out_channels, pos_grid=ocfg.pos_grid, patch=ocfg.patch, sh_degree=ocfg.sh_degree,
                        latent_channels=ocfg.latent_channels, inter_layers=ocfg.inter_layers, resolution=resolution, **kw)
    return StitchVAE3DB200.from_state_dict(sd, cfg, device="cuda:0")


@pytest.mark.parametrize("batch", [1, 2])
def test_decoder_voxelize_branch(batch):
    """forward_with_latent with DecoderConfig.voxelize: (i) the fusion + adapter stage equals the oracle's on exactly the per-pixel
    points / raw maps the engine produced (voxel membership is discontinuous in the points, so the stage is compared on identical
    inputs); (ii) the voxel count agrees with the real reference's full fp32 forward (golden) to within the cell-boundary flips the
    bf16 transformer causes; (iii) batch > 1 pads to the largest count with the reference's fill values."""
    from oracle import decoder_ref as D

    g = torch.load(GOLD)["forward"]
    o = D.TINY
    sd = D.init_state_dict(o, seed=g["weight_seed"])
    lat, img = D.synthetic_inputs(o, views_latent=g["latent_frames"], latent_hw=g["latent_hw"], image_hw=g["image_hw"], batch=batch,
                                  seed=g["input_seed"])
    eng = _engine(sd, o, g["resolution"], voxelize=True, voxel_size=g["voxel_size"])
    eng.keep_voxel_inputs = True
    out = eng.forward_with_latent(lat.cuda(), img.cuda())
    vin = eng.voxel_inputs
    C = o.raw_gs_dim
    want_n = []
    for b in range(batch):
        pts, raw = vin["pts"][b].cpu(), vin["raw"][b].cpu()
        vp, vf, _, _ = D.voxelize_with_fusion(raw[:, :C], pts, g["voxel_size"], raw[:, C])
        want_n.append(vp.shape[0])
        ref = D.gaussian_adapter(o, vp[None], vf[None])
        assert vin["counts"][b] == vp.shape[0]
        for k in GAUSS:
            got = getattr(out.gaussians, k)[b, :vp.shape[0]]
            _close(f"b{b}.{k}", got, ref[k][0], atol=1e-6, rtol=2e-5)
    nmax = max(want_n)
    assert out.gaussians.means.shape == (batch, nmax, 3)
    assert abs(out.infos["voxelize_ratio"] - nmax / (5 * 56 * 56)) < 1e-9
    if batch == 1:
        ref_n = int(g["outputs"]["n_voxels"])
        assert abs(want_n[0] - ref_n) <= 0.02 * ref_n, (want_n, ref_n)
    for b in range(batch):  # padded tail: points -1e4, opacity sigmoid(-1e10) = 0 (anysplat_stitched.py:448-455)
        if want_n[b] < nmax:
            assert bool((out.gaussians.means[b, want_n[b]:] == -1e4).all()) and bool((out.gaussians.opacities[b, want_n[b]:] == 0).all())
