"""The device decomposition planned for the Wan VAE decode (oracle/wan_vae_plan.py: implicit-GEMM causal convs over NDHWC, RMS-norm + SiLU
prologue, parity-decomposed up-sampling convs, time interleave, attention as two GEMMs + softmax) composes to the pinned oracle
(oracle/wan_vae_ref.py:decode == utils/wan_utils.py:1078-1117), and the bf16 storage roundings it implies stay small.  CPU only."""
import pytest
import torch

from oracle import wan_vae_plan as PL
from oracle import wan_vae_ref as V


@pytest.mark.parametrize("seed,t_lat,hw", [(1, 1, 2), (2, 2, 4), (3, 4, 3)])
def test_planned_decomposition_equals_the_oracle(seed, t_lat, hw):
    cfg = V.TINY_VAE
    sd = V.init_state_dict(cfg, seed=seed)
    z = 1.5 * torch.randn(1, cfg.z_dim, t_lat, hw, hw, generator=torch.Generator().manual_seed(seed))
    ref = V.decode(sd, cfg, z)
    out = PL.decode_plan(sd, cfg, z)
    assert out.shape == ref.shape == (1, 3, 1 + 4 * (t_lat - 1), 8 * hw, 8 * hw)
    assert (out - ref).abs().max().item() <= 2e-5
    # bf16 activations / weights in HBM, fp32 accumulation: the tolerance a device test of this path will carry
    # (frames live in [-1, 1]; measured here 0.03-0.045 max, 0.004-0.006 mean on random weights)
    b = PL.decode_plan(sd, cfg, z, round_bf16=True)
    assert (b - ref).abs().max().item() <= 0.08 and (b - ref).abs().mean().item() <= 0.01


def test_time_interleave_order():
    """the two channel halves of the temporal up-sampling conv alternate in time, first half first (utils/wan_utils.py:304-306)"""
    T, H, W, C = 3, 2, 2, 4
    y = torch.arange(T * H * W * 2 * C, dtype=torch.float32).reshape(T, H, W, 2 * C)
    ref = y.permute(3, 0, 1, 2)[None]                                        # [1, 2C, T, H, W]
    ref = ref.reshape(1, 2, C, T, H, W)
    ref = torch.stack((ref[:, 0], ref[:, 1]), 3).reshape(1, C, 2 * T, H, W)  # the reference's three lines
    out = y.reshape(T, H, W, 2, C).permute(0, 3, 1, 2, 4).reshape(2 * T, H, W, C)
    assert torch.equal(out.permute(3, 0, 1, 2)[None], ref)


@pytest.mark.parametrize("seed,frames,hw", [(4, 1, 16), (5, 5, 16), (6, 13, 24)])
def test_planned_encode_equals_the_oracle(seed, frames, hw):
    """stride-2 gathers (ZeroPad2d((0,1,0,1)) + conv; whole-clip temporal down-sampling) and the conv_out / quant_conv fold"""
    cfg = V.TINY_VAE
    sd = V.init_state_dict(cfg, seed=seed)
    clip = V.synthetic_clip(frames, hw, seed=seed)
    ref = V.encode_moments(sd, cfg, clip)
    out = PL.encode_plan(sd, cfg, clip)
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= 2e-5
    b = PL.encode_plan(sd, cfg, clip, round_bf16=True)
    assert (b - ref).abs().max().item() <= 0.1 * max(1.0, ref.abs().max().item())
