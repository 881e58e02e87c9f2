"""CPU checks that pin the DiT / sampler oracle as far as it can be pinned without diffusers
(see the "PARITY UNPINNED" header of oracle/wan_dit_ref.py): published parameter counts, analytic
identities, LoRA folding, and host-side agreement of the product's UniPC coefficient form with the
tensor-style restatement.  Also BASELINE.json configs[0] (single 1.3B block on CPU)."""
import math

import numpy as np
import pytest
import torch

from oracle import wan_dit_ref as R
from oracle.unipc_ref import UniPCFlowRef


def test_param_counts_match_published_sizes():
    assert abs(R.param_count(R.WAN_1_3B) / 1e9 - 1.419) < 2e-3
    assert abs(R.param_count(R.WAN_14B) / 1e9 - 14.29) < 2e-2
    shapes = R.param_shapes(R.WAN_1_3B)
    assert shapes["blocks.0.ffn.net.0.proj.weight"] == (8960, 1536)  # non-gated GELU MLP
    assert shapes["patch_embedding.weight"] == (1536, 16, 1, 2, 2)
    assert shapes["blocks.29.scale_shift_table"] == (1, 6, 1536)


def test_rope_split_and_identity_at_origin():
    fr = R.rope_freqs(R.WAN_1_3B, 4, 32, 32)
    assert fr.shape == (4096, 64) and fr.dtype == torch.complex128
    assert torch.allclose(fr[0], torch.ones(64, dtype=torch.complex128))  # position (0,0,0) rotates nothing
    # token 1 moves only along w: the t (22) and h (21) groups stay 1
    assert torch.allclose(fr[1, :43], torch.ones(43, dtype=torch.complex128))
    assert not torch.allclose(fr[1, 43:], torch.ones(21, dtype=torch.complex128))


def test_rope_logits_depend_on_relative_position_only():
    """the defining property of rotary embeddings, per axis group: <rot(q, p), rot(k, p')> is a function of p - p' (t, h, w offsets)"""
    cfg = R.WAN_1_3B
    f, h, w = 3, 4, 5
    fr = R.rope_freqs(cfg, f, h, w)                                   # [f*h*w, 64]
    g = torch.Generator().manual_seed(0)
    q = torch.view_as_complex(torch.randn(64, 2, generator=g, dtype=torch.float64))
    k = torch.view_as_complex(torch.randn(64, 2, generator=g, dtype=torch.float64))
    L = ((q * fr)[:, None, :] * (k * fr).conj()[None, :, :]).sum(-1).real.view(f, h, w, f, h, w)   # real dot product of the rotated pairs
    assert torch.allclose(L[:-1, :, :, :-1], L[1:, :, :, 1:], atol=1e-9)             # shift both along t
    assert torch.allclose(L[:, :-1, :, :, :-1], L[:, 1:, :, :, 1:], atol=1e-9)       # along h
    assert torch.allclose(L[:, :, :-1, :, :, :-1], L[:, :, 1:, :, :, 1:], atol=1e-9)  # along w
    assert not torch.allclose(L[0, 0, 0, 0, 0, 0], L[0, 0, 0, 1, 2, 3])


def test_zero_gates_make_block_residual_only():
    cfg = R.WAN_TINY
    sd = R.init_state_dict(cfg, seed=3, bias_std=0.02)
    lat, txt = R.synthetic_inputs(cfg, frames=1, hw=8, text_len=5)
    t = torch.tensor([500.0])
    for k in list(sd):
        if k.startswith("blocks.0.") and (k.endswith("attn2.to_out.0.weight") or k.endswith("attn2.to_out.0.bias")):
            sd[k] = torch.zeros_like(sd[k])
    # gate rows (2, 5) of scale_shift_table + time_proj = 0  => block is the identity
    sd["condition_embedder.time_proj.weight"] = torch.zeros_like(sd["condition_embedder.time_proj.weight"])
    sd["condition_embedder.time_proj.bias"] = torch.zeros_like(sd["condition_embedder.time_proj.bias"])
    tab = sd["blocks.0.scale_shift_table"].clone()
    tab[:, 2] = 0
    tab[:, 5] = 0
    sd["blocks.0.scale_shift_table"] = tab
    x_in = torch.nn.functional.conv3d(lat.float(), sd["patch_embedding.weight"], sd["patch_embedding.bias"],
                                      stride=cfg.patch_size).flatten(2).transpose(1, 2)
    out = R.single_block(sd, cfg, lat, t, txt, layer=0)
    assert torch.allclose(out, x_in, atol=1e-6)


def test_unpatchify_inverts_patchify_layout():
    cfg = R.WAN_TINY
    sd = R.init_state_dict(cfg, seed=0)
    lat, txt = R.synthetic_inputs(cfg, frames=2, hw=8, text_len=4)
    out = R.wan_forward(sd, cfg, lat, torch.tensor([10.0]), txt)
    assert out.shape == lat.shape and torch.isfinite(out).all()


def test_lora_fold_equals_unmerged_branch():
    cfg = R.WAN_TINY
    sd = R.init_state_dict(cfg, seed=1, bias_std=0.02)
    lora = R.init_lora(cfg, seed=2, r=8, std=0.05)
    folded = R.fold_lora(sd, lora, lora_alpha=16.0, r=8)
    name = "blocks.1.attn1.to_q"
    x = torch.randn(7, cfg.inner_dim)
    A, B = lora[f"base_model.model.{name}.lora_A.weight"], lora[f"base_model.model.{name}.lora_B.weight"]
    unmerged = x @ sd[name + ".weight"].t() + 2.0 * (x @ A.t()) @ B.t()
    assert torch.allclose(x @ folded[name + ".weight"].t(), unmerged, atol=1e-5)
    assert torch.equal(folded["blocks.1.ffn.net.2.weight"], sd["blocks.1.ffn.net.2.weight"])  # FFN untouched


def test_flow_sigma_schedule():
    s = UniPCFlowRef(flow_shift=5.0)
    s.set_timesteps(50)
    assert len(s.timesteps) == 50 and len(s.sigmas) == 51 and float(s.sigmas[-1]) == 0.0
    assert int(s.timesteps[0]) == 999 and torch.all(s.timesteps[:-1] > s.timesteps[1:])
    sig = 1 - np.linspace(1, 1 / 1000, 51)[::-1][0]
    assert abs(float(s.sigmas[0]) - 5 * sig / (1 + 4 * sig)) < 1e-6


def test_unipc_coefficient_form_matches_tensor_form():
    from vist3a_b200.unipc import UniPCFlowSchedule

    for n, shift in ((50, 5.0), (10, 3.0), (4, 1.0)):
        ref = UniPCFlowRef(flow_shift=shift)
        ref.set_timesteps(n)
        sch = UniPCFlowSchedule(n, shift)
        assert np.array_equal(sch.timesteps, ref.timesteps.numpy())
        g = torch.Generator().manual_seed(n)
        x = torch.randn(3, 6, generator=g, dtype=torch.float64)
        xr = x.float()
        m1 = m2 = torch.zeros_like(x)
        last = None
        for i in range(n):
            eps = torch.randn(3, 6, generator=g, dtype=torch.float64)
            xr = ref.step(eps.float(), xr)
            x0 = x - sch.sigmas[i] * eps
            if i > 0:
                cl, c0, c1, ct = sch.corr[i]
                cur = cl * last + c0 * m1 + c1 * m2 + ct * x0
            else:
                cur = x
            cx, c0, c1 = sch.pred[i]
            last = cur
            x = cx * cur + c0 * x0 + c1 * m1
            m2, m1 = m1, x0
            assert float((x.float() - xr).abs().max()) < 2e-5
        # with sigma_N = 0 the last step returns the final x0 prediction
        assert torch.allclose(x, x0)


def test_unipc_exact_on_constant_velocity_field():
    # flow matching with a constant velocity v: x_t = x_0 + sigma * v, model output = v for every t.
    ref = UniPCFlowRef(flow_shift=5.0)
    ref.set_timesteps(20)
    x0 = torch.randn(2, 5)
    v = torch.randn(2, 5)
    x = x0 + ref.sigmas[0] * v
    for _ in range(20):
        x = ref.step(v, x)
    assert torch.allclose(x, x0, atol=1e-5)


def test_unipc_integrates_the_gaussian_flow():
    """Independent check of the restated solver (no diffusers here to compare with): for data ~ N(mu, s^2) the flow-matching
    velocity field is analytic, v(x, sigma) = (x - E[x0 | x]) / sigma, and its ODE has the closed-form solution
    x(sigma) = (1 - sigma) mu + sqrt((1 - sigma)^2 s^2 + sigma^2) z.  The order-2 predictor-corrector must follow it far more closely
    than its own order-1 form (DDIM), and converge as the step count grows."""
    import math

    mu, s = 0.7, 0.4

    def err(n, shift, order, stop_at=0.3):
        ref = UniPCFlowRef(flow_shift=shift, solver_order=order)
        ref.set_timesteps(n)
        z = torch.randn(2048, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
        sg = float(ref.sigmas[0])
        x = (1 - sg) * mu + math.sqrt((1 - sg) ** 2 * s * s + sg * sg) * z
        for i in range(n):
            sg = float(ref.sigmas[i])
            if sg < stop_at:          # the last steps towards sigma = 0 are dominated by the posterior-mean error of the final jump
                break
            x0 = mu + (1 - sg) * s * s / ((1 - sg) ** 2 * s * s + sg * sg) * (x - (1 - sg) * mu)
            x = ref.step((x - x0) / sg, x)
        exact = (1 - sg) * mu + math.sqrt((1 - sg) ** 2 * s * s + sg * sg) * z
        return float((x - exact).abs().max())

    for shift in (1.0, 5.0):
        e1 = [err(n, shift, 1) for n in (40, 80, 160)]
        e2 = [err(n, shift, 2) for n in (40, 80, 160)]
        assert e1[0] > e1[1] > e1[2] and e2[0] > e2[1] > e2[2]
        assert all(b < a / 10 for a, b in zip(e1, e2)), (e1, e2)
        assert e2[1] < 1e-4 and e2[2] < 2e-5, e2


@pytest.mark.slow
def test_config0_single_1p3b_block_on_cpu():
    """BASELINE.json configs[0]: single Wan-1.3B block, latent [1,16,4,64,64], 77-token text."""
    cfg = R.WAN_1_3B
    full = R.param_shapes(cfg)
    g = torch.Generator().manual_seed(0)
    sd = {}
    for k, shp in full.items():
        if k.startswith("blocks.") and not k.startswith("blocks.0."):
            continue
        if k.endswith("scale_shift_table"):
            sd[k] = torch.randn(shp, generator=g) / math.sqrt(cfg.inner_dim)
        elif "norm" in k and k.endswith("weight"):
            sd[k] = torch.ones(shp)
        elif k.endswith("bias"):
            sd[k] = torch.zeros(shp)
        else:
            sd[k] = (torch.randn(shp, generator=g) * 0.02).bfloat16().float()
    lat, txt = R.synthetic_inputs(cfg, frames=4, hw=64, text_len=77)
    out = R.single_block(sd, cfg, lat, torch.tensor([999.0]), txt)
    assert out.shape == (1, 4096, 1536) and torch.isfinite(out).all()
    assert 0.01 < float(out.std()) < 10
