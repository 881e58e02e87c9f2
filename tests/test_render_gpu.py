"""GPU parity of the 3D-Gaussian rasteriser (vist3a_gs_project + vist3a_gs_rasterize through the C ABI, vist3a_b200.renderer) against the CPU
oracle (oracle/gsplat_ref.py: restatement of gsplat 1.4.0's published algorithm, parity unpinned against the absent package; analytic known
answers in tests/test_oracle_render.py).  Tolerance: fp32 sums in a different order and __expf vs exp -> 2e-4 absolute on colours / alpha of
O(1); the number of tile intersections (integer work) must agree exactly."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _render_gpu(m, c, o, h, V, K, W, H, **kw):
    from vist3a_b200 import ops

    return ops.gs_render(m.cuda().contiguous(), c.cuda().contiguous(), o.cuda().contiguous(), h.cuda().contiguous(), V, K, W, H, **kw)


def _cmp(got, want, tol=2e-4):
    for k in ("rgb", "depth", "alpha"):
        err = float((got[k].cpu() - want[k]).abs().max())
        scale = max(1.0, float(want[k].abs().max()))
        assert err <= tol * scale, (k, err)


@pytest.mark.parametrize("n,W,H,deg,spread,yaw", [(1500, 80, 64, 4, 1.0, 0.0), (4000, 96, 96, 4, 2.5, 20.0), (600, 50, 37, 2, 0.6, -10.0), (1, 32, 32, 0, 0.1, 0.0),
                                                 (3000, 64, 64, 3, 0.3, 0.0)])
def test_rasteriser_matches_oracle(n, W, H, deg, spread, yaw):
    """ragged image sizes (partial tiles), off-screen and behind-camera Gaussians (wide spread + yaw), every SH degree class, dense overdraw"""
    from oracle import gsplat_ref as G

    m, c, o, h = G.random_scene(n, seed=n + W, spread=spread, sh_degree=deg)
    V, K = G.look_at_camera(W, H, fov_deg=55.0, shift=(0.05, -0.03, 0.1), yaw_deg=yaw)
    bg = (0.1, 0.2, 0.3)
    want = G.render(m, c, o, h, V, K, W, H, sh_degree=deg, background=bg)
    got = _render_gpu(m, c, o, h, V, K, W, H, sh_degree=deg, background=bg)
    assert got["n_isect"] == want["n_isect"]
    _cmp(got, want)


def test_rasteriser_edge_cases():
    from oracle import gsplat_ref as G

    W, H = 48, 48
    V, K = G.look_at_camera(W, H)
    m, c, o, h = G.random_scene(200, seed=5)
    # nothing visible: every Gaussian behind the camera -> pure background, alpha 0, no intersections
    mb = m.clone()
    mb[:, 2] = -mb[:, 2]
    r = _render_gpu(mb, c, o, h, V, K, W, H, background=(0.3, 0.6, 0.9))
    assert r["n_isect"] == 0 and float(r["alpha"].abs().max()) == 0.0
    assert torch.allclose(r["rgb"].cpu(), torch.tensor([0.3, 0.6, 0.9]).expand(H, W, 3))
    # one huge opaque Gaussian covers every tile: alpha capped at 0.999 at its centre
    big = _render_gpu(torch.tensor([[0.0, 0.0, 2.0]]), torch.eye(3)[None] * 4.0, torch.tensor([1.5]), h[:1], V, K, W, H)
    assert big["n_isect"] == 9 and abs(float(big["alpha"].max()) - 0.999) < 1e-6
    # CPU tensors are rejected (no fallback)
    from vist3a_b200 import ops
    with pytest.raises(RuntimeError):
        ops.gs_render(m, c, o, h, V, K, W, H)


def test_renderer_module_on_decoder_sized_scene():
    """DecoderSplattingB200.rendering_fn at the BASELINE size (2.6 M Gaussians, 448x448, 3 views): properties that need no oracle --
    finite, colour in [0, 1], alpha in [0, 0.9999], depth consistent with alpha, and identical results for a permuted Gaussian order
    up to the summation order of equal-depth ties (depths are distinct here)."""
    from oracle import gsplat_ref as G
    from vist3a_b200.renderer import DecoderSplattingB200
    from vist3a_b200.stitched_decoder import Gaussians

    n = 13 * 448 * 448
    g = torch.Generator(device="cuda").manual_seed(3)
    means = torch.cat([(torch.rand(n, 2, device="cuda", generator=g) - 0.5) * 3.0, 1.0 + 3.0 * torch.rand(n, 1, device="cuda", generator=g)], -1)
    s = 0.002 + 0.004 * torch.rand(n, 3, device="cuda", generator=g)
    cov = torch.diag_embed(s * s)
    opac = torch.rand(n, device="cuda", generator=g)
    harm = torch.randn(n, 3, 25, device="cuda", generator=g) * 0.2
    gs = Gaussians(means=means[None], covariances=cov[None], harmonics=harm[None], opacities=opac[None], scales=s[None], rotations=torch.zeros(1, n, 4, device="cuda"))
    W = H = 448
    V0, K = G.look_at_camera(W, H, fov_deg=60.0)
    c2w = torch.stack([torch.linalg.inv(G.look_at_camera(W, H, yaw_deg=a)[0]) for a in (-8.0, 0.0, 8.0)])[None]
    Kn = K.clone()
    Kn[0] /= W
    Kn[1] /= H
    out = DecoderSplattingB200((1.0, 1.0, 1.0)).rendering_fn(gs, c2w, Kn[None, None].expand(1, 3, 3, 3), image_shape=(H, W))
    assert out.color.shape == (1, 3, 3, H, W) and out.depth.shape == (1, 3, H, W)
    for t in (out.color, out.depth, out.alpha):
        assert bool(torch.isfinite(t).all())
    assert float(out.color.min()) >= 0.0 and float(out.color.max()) <= 1.0
    assert float(out.alpha.min()) >= 0.0 and float(out.alpha.max()) <= 1.0 - 1e-4 + 1e-6
    assert float(out.alpha.mean()) > 0.5                                     # the scene fills the view
    # camera-space depths lie in [1, 4] for the frontal view and in [1 cos 8 - 1.5 sin 8, 4 cos 8 + 1.5 sin 8] = [0.78, 4.17] for the yawed ones
    assert bool((out.depth <= 4.2 * out.alpha + 1e-4).all()) and bool((out.depth >= 0.75 * out.alpha - 1e-4).all())
    assert bool((out.depth[0, 1] <= 4.0 * out.alpha[0, 1] + 1e-4).all()) and bool((out.depth[0, 1] >= 1.0 * out.alpha[0, 1] - 1e-4).all())
    perm = torch.randperm(n, device="cuda", generator=g)
    gp = Gaussians(means=means[perm][None], covariances=cov[perm][None], harmonics=harm[perm][None], opacities=opac[perm][None], scales=s[perm][None],
                   rotations=torch.zeros(1, n, 4, device="cuda"))
    out2 = DecoderSplattingB200((1.0, 1.0, 1.0)).rendering_fn(gp, c2w[:, 1:2], Kn[None, None], image_shape=(H, W))
    assert float((out2.color[0, 0] - out.color[0, 1]).abs().max()) < 1e-3   # equal fp32 depths among 2.6 M Gaussians composite in index order


def test_text_to_3d_tail_decoder_then_interpolated_video_frames():
    """the tail of inference_t23d.py:133-155 on the tiny decoder: forward_with_latent -> Gaussians + predicted cameras -> the interpolated camera path
    of save_interpolated_video -> frames; checked against the oracle renderer on the engine's own Gaussians for one interpolated camera"""
    from oracle import decoder_ref as D
    from oracle import gsplat_ref as G
    from vist3a_b200.renderer import DecoderSplattingB200, interpolate_context_cameras
    from vist3a_b200.stitched_decoder import DecoderConfig, StitchVAE3DB200

    o = D.TINY
    sd = D.init_state_dict(o, seed=3)
    lat, img = D.synthetic_inputs(o, views_latent=2, latent_hw=8, image_hw=56, seed=5)
    cfg = DecoderConfig(embed_dim=o.embed_dim, num_heads=o.num_heads, dino_blocks=o.dino_blocks, cam_heads=o.cam_heads, dpt_out_channels=o.dpt_out_channels, resolution=64)
    out = StitchVAE3DB200.from_state_dict(sd, cfg, device="cuda:0").forward_with_latent(lat.cuda(), img.cuda())
    rend = DecoderSplattingB200((1.0, 1.0, 1.0))
    frames = rend.render_interpolated_views(out, image_shape=(56, 56), t=2)
    assert frames.color.shape == (1, 12, 3, 56, 56) and frames.depth.shape == (1, 12, 56, 56)      # (5 - 1)(2 + 1) frames
    assert bool(torch.isfinite(frames.color).all()) and float(frames.color.min()) >= 0 and float(frames.color.max()) <= 1
    ex, ix = interpolate_context_cameras(out.pred_context_pose["extrinsic"].float().cpu(), out.pred_context_pose["intrinsic"].float().cpu(), 2)
    k = 4                                                                                            # an interpolated (not a context) camera
    K = ix[0, k].clone()
    K[0] *= 56
    K[1] *= 56
    g = out.gaussians
    want = G.render(g.means[0].cpu(), g.covariances[0].cpu(), g.opacities[0].cpu(), g.harmonics[0].cpu(), torch.linalg.inv(ex[0, k]), K, 56, 56,
                    background=(1.0, 1.0, 1.0))
    assert float((frames.color[0, k].permute(1, 2, 0).cpu() - want["rgb"].clamp(0, 1)).abs().max()) < 5e-4
    assert float((frames.alpha[0, k].cpu() - want["alpha"]).abs().max()) < 5e-4
