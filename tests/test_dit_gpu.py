"""GPU parity of the DiT path (through vist3a_b200.WanTransformer3DModelB200 -> C ABI) against the
fp32 CPU oracle (oracle/wan_dit_ref.py) on identical seeded weights, latents and text embeddings.

Stated bf16 tolerance: the engine keeps the reference's CUDA-autocast dtype policy (bf16 GEMM
operands / residual stream / attention probabilities), so against an all-fp32 oracle the output
differs by accumulated bf16 rounding: rel-L2 <= 2e-2 for the model output (3 x the 4.6e-3 - 6e-3 measured
on the B200 for these cases), and no worse than 1.5x the error of the same oracle graph executed by
torch in bf16 autocast on this GPU (3.2e-3 - 3.8e-3).  The full 30-layer BASELINE forward and the 50-step
trajectory are in tests/test_parity_full_gpu.py.
"""
import pytest
import torch

from oracle import wan_dit_ref as R
from oracle.unipc_ref import denoise_loop

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


def _torch_bf16_forward(sd, cfg, lat, t, txt, num_layers=None):
    """the oracle graph run by torch on the GPU under bf16 autocast = the reference's own execution mode (wan_forward upcasts the weights
    to fp32; F.linear / conv / SDPA under autocast then run in bf16 as in the reference)"""
    sdg = {k: v.cuda() for k, v in sd.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return R.wan_forward(sdg, cfg, lat.cuda(), t.cuda(), txt.cuda(), num_layers=num_layers).float()


def _engine(sd, cfg, lora=None):
    from vist3a_b200.wan_dit import WanTransformer3DModelB200

    return WanTransformer3DModelB200.from_state_dict(sd, cfg, lora=lora)


@pytest.mark.parametrize("batch,frames,hw,text_len", [(1, 2, 16, 20), (2, 3, 16, 77), (1, 1, 32, 130)])
def test_tiny_dit_matches_oracle(batch, frames, hw, text_len):
    cfg = R.WAN_TINY
    sd = R.init_state_dict(cfg, seed=7, bias_std=0.05)
    lat, txt = R.synthetic_inputs(cfg, batch=batch, frames=frames, hw=hw, text_len=text_len, text_valid=text_len - 3, seed=1)
    t = torch.tensor([999.0, 321.0][:batch])
    ref = R.wan_forward(sd, cfg, lat, t, txt)
    m = _engine(sd, cfg)
    out = m(lat.cuda(), t.cuda(), txt.cuda(), return_dict=False)[0]
    assert out.dtype == lat.dtype and out.shape == lat.shape
    e_ours = _rel(out, ref)
    e_torch = _rel(_torch_bf16_forward(sd, cfg, lat, t, txt), ref)
    print(f"rel-L2 vs fp32 oracle: ours {e_ours:.3e}, torch-bf16-autocast {e_torch:.3e}")
    assert e_ours < TOL
    assert e_ours < 1.5 * e_torch + 1e-3


def test_lora_is_folded_at_load():
    cfg = R.WAN_TINY
    sd = R.init_state_dict(cfg, seed=8, bias_std=0.05)
    lora = R.init_lora(cfg, seed=9, std=0.05)
    lat, txt = R.synthetic_inputs(cfg, frames=2, hw=16, text_len=16, seed=2)
    t = torch.tensor([700.0])
    ref = R.wan_forward(R.fold_lora(sd, lora), cfg, lat, t, txt)
    base = R.wan_forward(sd, cfg, lat, t, txt)
    out = _engine(sd, cfg, lora=lora)(lat.cuda(), t.cuda(), txt.cuda(), return_dict=False)[0]
    assert _rel(out, ref) < TOL
    assert _rel(base, ref) > 5 * _rel(out, ref)  # the adapter matters and was applied


def test_1p3b_two_layers_full_sequence():
    """Real 1.3B widths (D=1536, 12 heads, F=8960) at the BASELINE sequence (L=4096, 512 text tokens), 2 layers."""
    import dataclasses

    cfg = dataclasses.replace(R.WAN_1_3B, num_layers=2)
    sd = R.init_state_dict(cfg, seed=0, bias_std=0.02)
    lat, txt = R.synthetic_inputs(cfg, frames=4, hw=64, text_len=512, text_valid=200, seed=0)
    t = torch.tensor([875.0])
    ref = R.wan_forward(sd, cfg, lat, t, txt)
    out = _engine(sd, cfg)(lat.cuda(), t.cuda(), txt.cuda(), return_dict=False)[0]
    e_ours = _rel(out, ref)
    e_torch = _rel(_torch_bf16_forward(sd, cfg, lat, t, txt), ref)
    print(f"1.3B x2 layers: ours {e_ours:.3e}, torch-bf16-autocast {e_torch:.3e}")
    assert e_ours < TOL and e_ours < 1.5 * e_torch + 1e-3


def test_14b_width_one_layer_21_views():
    """BASELINE configs[3] shapes: Wan-14B widths (D=5120, 40 heads, F=13824) at the 21-view sequence (latent [1,16,6,64,64],
    L=6144, 512 text tokens), 1 layer (the oracle's CPU forward of one 14B block takes ~20 s)."""
    import dataclasses

    cfg = dataclasses.replace(R.WAN_14B, num_layers=1)
    sd = R.init_state_dict(cfg, seed=5, bias_std=0.02)
    lat, txt = R.synthetic_inputs(cfg, frames=6, hw=64, text_len=512, text_valid=300, seed=6)
    t = torch.tensor([640.0])
    ref = R.wan_forward(sd, cfg, lat, t, txt)
    out = _engine(sd, cfg)(lat.cuda(), t.cuda(), txt.cuda(), return_dict=False)[0]
    e_ours = _rel(out, ref)
    print(f"14B x1 layer, L=6144: ours {e_ours:.3e}")
    assert out.shape == (1, 16, 6, 64, 64) and e_ours < TOL


def test_text_cache_tracks_inplace_updates():
    cfg = R.WAN_TINY
    sd = R.init_state_dict(cfg, seed=3, bias_std=0.02)
    m = _engine(sd, cfg)
    lat, txt = R.synthetic_inputs(cfg, frames=1, hw=16, text_len=8, seed=4)
    txt = txt.cuda()
    t = torch.tensor([100.0]).cuda()
    a = m(lat.cuda(), t, txt, return_dict=False)[0].clone()
    txt.mul_(-1.0)
    b = m(lat.cuda(), t, txt, return_dict=False)[0]
    ref_b = R.wan_forward(sd, cfg, lat, t.cpu(), txt.cpu())
    assert _rel(b, ref_b) < TOL and _rel(a, ref_b) > 3 * _rel(b, ref_b)


def test_denoise_loop_matches_oracle_trajectory():
    """8-step CFG sampling (UniPC flow, shift 5) of the tiny model: device loop (batched cond/uncond,
    precomputed coefficients) vs the oracle's diffusers-style loop."""
    from vist3a_b200.unipc import denoise

    cfg = R.WAN_TINY
    sd = R.init_state_dict(cfg, seed=5, bias_std=0.02)
    g = torch.Generator().manual_seed(6)
    noise = torch.randn(1, 16, 2, 16, 16, generator=g)
    _, tc = R.synthetic_inputs(cfg, text_len=12, text_valid=9, seed=7)
    _, tu = R.synthetic_inputs(cfg, text_len=12, text_valid=4, seed=8)
    ref = denoise_loop(lambda x, t, txt: R.wan_forward(sd, cfg, x, t, txt), noise, tc, tu, num_inference_steps=8,
                       guidance_scale=6.0, flow_shift=5.0)
    m = _engine(sd, cfg)
    out = denoise(m, noise, tc, tu, num_inference_steps=8, guidance_scale=6.0, flow_shift=5.0)
    out2 = denoise(m, noise, tc, tu, num_inference_steps=8, guidance_scale=6.0, flow_shift=5.0, batch_cfg=False)
    e = _rel(out, ref)
    print(f"8-step trajectory rel-L2 {e:.3e}; batched vs sequential CFG {_rel(out, out2):.3e}")
    assert e < 5e-2  # guidance 6 amplifies per-step bf16 error ~6x
    assert _rel(out, out2) < 2e-2
