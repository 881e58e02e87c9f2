"""Pins oracle/wan_vae_ref.py (the Wan-2.1 VAE either side of the hot path: utils/wan_utils.py:96-1180) against
  (i)  the reference's own modules and chunked encode / decode loops, imported live where /root/reference is mounted,
  (ii) golden vectors those produced (tests/golden/wan_vae_tiny.pt, script tests/golden/make_vae_golden.py) -- checked everywhere,
  (iii) structure: state-dict manifest of the full-size model, causality, clip-length mapping, posterior and latent statistics.
CPU only."""
import os

import pytest
import torch

from oracle import ref_loader as RL
from oracle import wan_vae_ref as V

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "wan_vae_tiny.pt")
TOL = 2e-5   # fp32, whole-clip convolutions vs the reference's chunked ones (different summation order inside the conv kernels)


@pytest.mark.skipif(not RL.available(), reason="/root/reference is not mounted (build container only)")
@pytest.mark.parametrize("frames,hw,batch", [(1, 16, 1), (5, 32, 2), (13, 24, 1), (9, 40, 1)])
def test_against_live_reference(frames, hw, batch):
    cfg = V.TINY_VAE
    live = RL.LiveWanVAE(base_dim=cfg.base_dim, z_dim=cfg.z_dim, seed=0)
    lsd = live.state_dict()
    shapes = V.param_shapes(cfg)
    assert set(shapes) == set(lsd)
    assert all(tuple(lsd[k].shape) == tuple(shapes[k]) for k in shapes)
    sd = V.init_state_dict(cfg, seed=7 + frames)
    live.load_state_dict(sd)           # strict: the manifest is the reference's
    clip = V.synthetic_clip(frames, hw, batch=batch, seed=frames)
    m_ref = live.encode_moments(clip)
    m = V.encode_moments(sd, cfg, clip)
    assert m.shape == m_ref.shape == (batch, 2 * cfg.z_dim, 1 + (frames - 1) // 4, hw // 8, hw // 8)
    assert (m - m_ref).abs().max().item() <= TOL
    z = V.posterior(m_ref, torch.randn(m_ref[:, : cfg.z_dim].shape, generator=torch.Generator().manual_seed(1)))
    y_ref = live.decode(z)
    y = V.decode(sd, cfg, z)
    assert y.shape == y_ref.shape == (batch, 3, frames, hw, hw)
    assert (y - y_ref).abs().max().item() <= TOL


def test_against_golden():
    g = torch.load(GOLDEN, weights_only=False)
    cfg = V.TINY_VAE
    assert len(g["cases"]) >= 3
    for name, c in g["cases"].items():
        sd = V.init_state_dict(cfg, seed=c["weight_seed"])
        clip = V.synthetic_clip(c["frames"], c["hw"], batch=c["batch"], seed=c["clip_seed"])
        m = V.encode_moments(sd, cfg, clip)
        assert (m - c["moments"]).abs().max().item() <= TOL, name
        y = V.decode(sd, cfg, c["latent"])
        assert (y - c["decoded"]).abs().max().item() <= TOL, name
        assert y.abs().max().item() <= 1.0


def test_full_size_manifest():
    """Wan-2.1 VAE: 126.9 M parameters; diffusers' AutoencoderKLWan key names (a released checkpoint loads unchanged)"""
    s = V.param_shapes(V.WAN_VAE)
    assert sum(torch.Size(v).numel() for v in s.values()) == 126_892_531
    assert s["encoder.conv_in.weight"] == (96, 3, 3, 3, 3)
    assert s["encoder.down_blocks.5.time_conv.weight"] == (192, 192, 3, 1, 1)           # first temporal down-sampling sits at level 1
    assert "encoder.down_blocks.2.time_conv.weight" not in s                             # level 0 is spatial only
    assert s["encoder.conv_out.weight"] == (32, 384, 3, 3, 3) and s["quant_conv.weight"] == (32, 32, 1, 1, 1)
    assert s["decoder.conv_in.weight"] == (384, 16, 3, 3, 3)
    assert s["decoder.up_blocks.0.upsamplers.0.time_conv.weight"] == (768, 384, 3, 1, 1)
    assert s["decoder.up_blocks.1.resnets.0.conv_shortcut.weight"] == (384, 192, 1, 1, 1)
    assert "decoder.up_blocks.2.upsamplers.0.time_conv.weight" not in s and "decoder.up_blocks.3.upsamplers.0.resample.1.weight" not in s
    assert s["decoder.conv_out.weight"] == (3, 96, 3, 3, 3)
    f = V.flops(V.WAN_VAE, 13, 512)
    assert f["latent_frames"] == 4 and f["latent_hw"] == 64
    assert 17e12 < f["encode"] < 18e12 and 29e12 < f["decode"] < 30e12


def test_causality_and_clip_mapping():
    cfg = V.TINY_VAE
    sd = V.init_state_dict(cfg, seed=11)
    clip = V.synthetic_clip(13, 16, seed=2)
    m = V.encode_moments(sd, cfg, clip)
    assert m.shape[2] == 4
    # frames 9..12 only reach latent frame 3; frame 0 reaches everything after it
    c2 = clip.clone()
    c2[:, :, 9:] = -c2[:, :, 9:]
    m2 = V.encode_moments(sd, cfg, c2)
    assert torch.equal(m[:, :, :3], m2[:, :, :3]) and not torch.equal(m[:, :, 3], m2[:, :, 3])
    # the latent of a one-frame clip is the first latent frame of any longer clip with the same first frame
    m1 = V.encode_moments(sd, cfg, clip[:, :, :1])
    assert (m1[:, :, 0] - m[:, :, 0]).abs().max().item() <= 1e-5
    with pytest.raises(ValueError):
        V.encode_moments(sd, cfg, clip[:, :, :6])
    z = torch.randn(1, cfg.z_dim, 4, 2, 2, generator=torch.Generator().manual_seed(3))
    y = V.decode(sd, cfg, z)
    assert y.shape == (1, 3, 13, 16, 16)
    z2 = z.clone()
    z2[:, :, 3] += 1.0
    y2 = V.decode(sd, cfg, z2)
    assert torch.equal(y[:, :, :9], y2[:, :, :9]) and not torch.equal(y[:, :, 9:], y2[:, :, 9:])
    y1 = V.decode(sd, cfg, z[:, :, :1])
    assert (y1[:, :, 0] - y[:, :, 0]).abs().max().item() <= 1e-5


def test_posterior_and_latent_statistics():
    m = torch.tensor([0.5, -1.0, 40.0, -50.0]).view(1, 4, 1, 1, 1)      # mean (0.5, -1), logvar (40, -50) -> clamped to (20, -30)
    assert torch.equal(V.posterior(m), m[:, :2])
    n = torch.tensor([1.0, -2.0]).view(1, 2, 1, 1, 1)
    s = V.posterior(m, n)
    exp = torch.tensor([0.5 + torch.exp(torch.tensor(10.0)).item(), -1.0 - 2.0 * torch.exp(torch.tensor(-15.0)).item()])
    assert torch.allclose(s.flatten(), exp, rtol=1e-6)
    lat = torch.zeros(1, 16, 1, 1, 1)
    assert torch.allclose(V.denormalise_latents(lat).flatten(), torch.tensor(V.LATENTS_MEAN))
    one = torch.ones(1, 16, 1, 1, 1)
    assert torch.allclose(V.denormalise_latents(one).flatten(), torch.tensor(V.LATENTS_MEAN) + torch.tensor(V.LATENTS_STD), rtol=1e-6)
    f = V.frames_to_feedforward(torch.rand(1, 3, 2, 16, 16), hw=14)
    assert f.shape == (1, 3, 2, 14, 14)


@pytest.mark.slow
@pytest.mark.skipif(not RL.available(), reason="/root/reference is not mounted (build container only)")
def test_against_live_reference_full_width():
    """the released widths (base 96, z 16, 126.9 M parameters) on a 5-frame 32 x 32 clip, ~15 s on 8 cores"""
    cfg = V.WAN_VAE
    live = RL.LiveWanVAE(seed=0)
    sd = V.init_state_dict(cfg, seed=9)
    live.load_state_dict(sd)
    clip = V.synthetic_clip(5, 32, seed=1)
    m_ref = live.encode_moments(clip)
    assert (V.encode_moments(sd, cfg, clip) - m_ref).abs().max().item() <= TOL
    z = V.posterior(m_ref)
    assert (V.decode(sd, cfg, z) - live.decode(z)).abs().max().item() <= TOL
