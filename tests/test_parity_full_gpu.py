"""GPU parity at the BASELINE sizes: the configurations the metric is quoted on, compared with the oracles and
with golden vectors of the REAL reference -- not only with size-independent invariants.

  * 1.3B DiT, all 30 layers, L = 4096 latent tokens, 512 text tokens            vs oracle/wan_dit_ref.py (fp32, ~10 s of CPU)
  * 50-step CFG + UniPC trajectory (1.3B widths, reduced depth and grid)        vs oracle/unipc_ref.py's diffusers-style loop
  * stitched decoder, full widths, 13 views x 448x448 (2 609 152 Gaussians)     vs tests/golden/decoder_full_13v.pt (the real reference's
    fp32 forward, made by tests/golden/make_decoder_golden_full.py) and vs oracle/decoder_ref.py run in the test
  * attention at the decoder's global lengths 13 377 (104 x 128 + 65) and 21 609 vs fp32 SDPA on the device

Tolerances are per field, 3 x the error measured on the B200 (recorded next to each bound), never a blanket figure: the engine computes
in the reference's own mixed precision (bf16 tensor-core operands in the transformers, TF32 in the heads), the oracles in
fp32.  Fields that do not pass through the predicted cameras meet north_star's 1e-3; camera-conditioned fields are bounded by
FLOOR_MULT x the error of the reference's own GPU numerics on the same inputs where that is larger.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD_FULL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "decoder_full_13v.pt")
GAUSS = ("means", "covariances", "harmonics", "opacities", "scales", "rotations")


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


# --------------------------------------------------------------------------------------------------------------------------------
# DiT
# --------------------------------------------------------------------------------------------------------------------------------
def test_dit_1p3b_all_30_layers_baseline_shape():
    """BASELINE configs[1] forward: Wan-1.3B (30 layers, D = 1536, F = 8960), latent [1,16,4,64,64] (L = 4096), 512 text tokens."""
    from oracle import wan_dit_ref as R
    from vist3a_b200.wan_dit import WanTransformer3DModelB200

    cfg = R.WAN_1_3B
    sd = R.init_state_dict(cfg, seed=0, bias_std=0.02)
    lat, txt = R.synthetic_inputs(cfg, frames=4, hw=64, text_len=512, text_valid=200, seed=0)
    t = torch.tensor([875.0])
    ref = R.wan_forward(sd, cfg, lat, t, txt, cast_fp32=False)
    m = WanTransformer3DModelB200.from_state_dict(sd, cfg)
    out = m(lat.cuda(), t.cuda(), txt.cuda(), return_dict=False)[0]
    # the reference's own execution mode on this GPU: the same graph by torch under bf16 autocast (cuBLAS / SDPA)
    sdg = {k: (v.cuda().bfloat16() if v.dim() > 1 and "scale_shift_table" not in k else v.cuda()) for k, v in sd.items()}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        # (a) as the reference runs it: bf16 pipeline, the residual stream rounded to bf16 after every gated add (`.type_as(hidden_states)`)
        auto = R.wan_forward(sdg, cfg, lat.cuda(), t.cuda(), txt.cuda(), cast_fp32=False, stream_dtype=torch.bfloat16).float()
        # (b) the same graph with an fp32 residual stream (what type promotion gives when nothing casts back): a lower bound on bf16 error
        auto32 = R.wan_forward(sdg, cfg, lat.cuda(), t.cuda(), txt.cuda(), cast_fp32=False).float()
    e_ours, e_torch, e_torch32 = _rel(out, ref), _rel(auto, ref), _rel(auto32, ref)
    print(f"1.3B x 30 layers, L=4096, Lt=512: rel-L2 vs fp32 oracle: ours {e_ours:.3e}, torch-bf16 (bf16 stream, the reference's policy) {e_torch:.3e}, "
          f"torch-bf16 (fp32 stream) {e_torch32:.3e}")
    assert out.shape == (1, 16, 4, 64, 64) and bool(torch.isfinite(out).all())
    # bound: the engine keeps diffusers' dtype policy -- a bf16 residual stream, rounded after each of the 90 residual adds -- so its
    # error against the fp32 oracle is that policy's (measured on the B200: ours 1.37e-2; fp32-stream torch 4.7e-3).  Stay within 1.5 x
    # the error of torch executing the same policy, and below 3 x measured in absolute terms
    assert e_ours < 4e-2
    assert e_ours < 1.5 * e_torch + 1e-3


def test_50_step_trajectory_reduced_depth():
    """the whole sampling loop of BASELINE configs[1] -- 50 steps, CFG 6, UniPC flow (shift 5) -- at 1.3B widths, 2 layers, latent
    [1,16,2,32,32]: device loop (batched cond/uncond, CUDA-graphed DenoiseEngine) vs the oracle's diffusers-style loop"""
    import dataclasses

    from oracle import wan_dit_ref as R
    from oracle.unipc_ref import denoise_loop
    from vist3a_b200.pipeline import DenoiseEngine
    from vist3a_b200.wan_dit import WanTransformer3DModelB200

    cfg = dataclasses.replace(R.WAN_1_3B, num_layers=2)
    sd = R.init_state_dict(cfg, seed=5, bias_std=0.02)
    g = torch.Generator().manual_seed(6)
    noise = torch.randn(1, 16, 2, 32, 32, generator=g)
    _, tc = R.synthetic_inputs(cfg, text_len=64, text_valid=40, seed=7)
    _, tu = R.synthetic_inputs(cfg, text_len=64, text_valid=9, seed=8)
    sdf = {k: v.float() for k, v in sd.items()}
    ref = denoise_loop(lambda x, t, txt: R.wan_forward(sdf, cfg, x, t, txt, cast_fp32=False), noise, tc, tu, num_inference_steps=50,
                       guidance_scale=6.0, flow_shift=5.0)
    m = WanTransformer3DModelB200.from_state_dict(sd, cfg)
    eng = DenoiseEngine(m, noise.shape, 64, num_inference_steps=50, guidance_scale=6.0, flow_shift=5.0, use_graph=True)
    out = eng.run(noise.cuda(), tc.cuda(), tu.cuda())
    e = _rel(out, ref)
    print(f"50-step trajectory (2 layers, L=512): rel-L2 {e:.3e}")
    assert bool(torch.isfinite(out).all())
    assert e < 1.5e-2   # 3 x the 4.1e-3 measured on the B200 (guidance 6 amplifies the per-step bf16 error)


# --------------------------------------------------------------------------------------------------------------------------------
# attention at the decoder's global sequence lengths
# --------------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("L", [13377, 21609])
def test_fmha_global_lengths(L):
    """B1 H16 d64 at 13 x 1029 = 13 377 (= 104 x 128 + 65: partial last key block and query tile) and 21 x 1029 = 21 609 tokens,
    vs SDPA on the device in fp32 (math / mem-efficient backend, computed per head to bound memory)"""
    from vist3a_b200 import ops

    H, D = 16, 64
    g = torch.Generator(device="cuda").manual_seed(L)
    qkv = torch.randn(1, L, 3, H, D, device="cuda", generator=g).bfloat16()
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    out = ops.fmha(q, k, v)
    torch.cuda.synchronize()
    worst_abs, num, den = 0.0, 0.0, 0.0
    for h in range(H):   # fp32 reference one head at a time: softmax(q k^T / 8) v with the full 13377^2 / 21609^2 score matrix in fp32
        qh, kh, vh = q[0, :, h].float(), k[0, :, h].float(), v[0, :, h].float()
        ref = torch.empty(L, D, device="cuda")
        for r0 in range(0, L, 4096):
            s = (qh[r0:r0 + 4096] @ kh.t()) * (D ** -0.5)
            ref[r0:r0 + 4096] = torch.softmax(s, dim=-1) @ vh
        d = out[0, :, h].float() - ref
        worst_abs = max(worst_abs, float(d.abs().max()))
        num += float(d.double().pow(2).sum())
        den += float(ref.double().pow(2).sum())
    rel = (num / den) ** 0.5
    print(f"fmha L={L}: max abs err {worst_abs:.3e}, rel-L2 {rel:.3e}")
    # outputs are means of ~L unit-variance values: |O| ~ 1/sqrt(L) ~ 1e-2; P and O are rounded to bf16 (2^-9 relative)
    assert rel < 8e-3 and worst_abs < 2e-3


# --------------------------------------------------------------------------------------------------------------------------------
# decoder at 13 views x 448 x 448
# --------------------------------------------------------------------------------------------------------------------------------
# per-field bounds = 3 x the rel-L2 error measured on the B200 against the fp32 oracle at this size; the fields of
# the first group do not pass through the predicted cameras and meet north_star's 1e-3
# measured (B200, r2): scales 4.1e-4, opacities 2.3e-4, depth 7.9e-4, covariances 1.2e-3, rotations 1.9e-3, harmonics 1.9e-3, pose 3.0e-3,
# intrinsic 2.2e-3, extrinsic 6.1e-3, means 7.9e-3, scene scale 8.7e-4; the reference's own GPU numerics (floor): 3.0e-4, 1.7e-4, 5.3e-4,
# 1.5e-3, 2.8e-3, 1.2e-3, 4.2e-3, 2.4e-3, 8.6e-3, 1.2e-2, 3.8e-5
BOUNDS_13V = {"scales": 1e-3, "opacities": 1e-3, "depth": 1e-3, "covariances": 3.6e-3, "rotations": 6e-3, "harmonics": 6e-3,
              "last_pred_pose_enc": 9e-3, "intrinsic": 7e-3, "extrinsic": 1.8e-2, "means": 2.4e-2, "scene_scale": 3e-3}
FLOOR_MULT = 2.0


def _decoder_13v():
    from oracle import decoder_ref as D
    from vist3a_b200.stitched_decoder import DecoderConfig, StitchVAE3DB200

    sd = D.init_state_dict(D.FULL, seed=1)
    lat, img = D.synthetic_inputs(D.FULL, views_latent=4, latent_hw=64, image_hw=448, seed=3)
    m = StitchVAE3DB200.from_state_dict(sd, DecoderConfig(resolution=512), device="cuda:0")
    o = m.forward_with_latent(lat.cuda(), img.cuda())
    g = o.gaussians
    out = {k: getattr(g, k) for k in GAUSS}
    out.update(depth=o.depth_dict["depth"], extrinsic=o.pred_context_pose["extrinsic"], intrinsic=o.pred_context_pose["intrinsic"],
               last_pred_pose_enc=o.last_pred_pose_enc, scene_scale=o.infos["scene_scale"].reshape(1))
    return D, sd, lat, img, out


def test_decoder_13_views_matches_real_reference_golden():
    """engine vs the golden vectors the REAL reference produced at full width and full size (strided subsample + full-tensor checksums)"""
    if not os.path.exists(GOLD_FULL):
        pytest.fail("tests/golden/decoder_full_13v.pt is missing (python tests/golden/make_decoder_golden_full.py in the build container)")
    gold = torch.load(GOLD_FULL)
    assert gold["weight_seed"] == 1 and gold["input_seed"] == 3
    D, sd, lat, img, out = _decoder_13v()
    want, st = gold["outputs"], gold["stride"]
    errs = {k: _rel(out[k][:, ::st], want[k]) for k in GAUSS}
    errs["depth"] = _rel(out["depth"][:, :, ::gold["depth_stride"], ::gold["depth_stride"]], want["depth"])
    for k in ("extrinsic", "intrinsic", "last_pred_pose_enc"):
        errs[k] = _rel(out[k], want[k])
    # checksums over ALL 2 609 152 Gaussians (float64 sums of the reference's full tensors)
    sums = {k: _rel(out[k].double().abs().sum(dim=1), want["abs_checksum_" + k]) for k in ("scales", "opacities", "covariances", "harmonics")}
    sums["depth"] = _rel(out["depth"].double().sum(dim=(2, 3, 4)), want["checksum_depth"])
    print("13v golden: rel-L2 ", {k: f"{v:.2e}" for k, v in errs.items()})
    print("13v golden: checksums", {k: f"{v:.2e}" for k, v in sums.items()})
    # the reference's own GPU numerics on the same inputs (oracle graph under bf16 autocast, heads in fp32/TF32) as the floor for the
    # camera-conditioned fields
    with torch.device("cuda"):
        auto = D.decoder_forward({k: v.cuda() for k, v in sd.items()}, D.FULL, lat.cuda(), img.cuda(), resolution=512, gpu_autocast=True)
    floor = {k: _rel(auto[k][:, ::st], want[k]) for k in GAUSS}
    for k in ("extrinsic", "intrinsic", "last_pred_pose_enc"):
        floor[k] = _rel(auto[k], want[k])
    floor["depth"] = _rel(auto["depth"][:, :, ::gold["depth_stride"], ::gold["depth_stride"]], want["depth"])
    print("13v golden: floor  ", {k: f"{v:.2e}" for k, v in floor.items()})
    bad = {k: (v, BOUNDS_13V[k], floor[k]) for k, v in errs.items() if not v < max(BOUNDS_13V[k], FLOOR_MULT * floor[k])}
    assert not bad, bad
    assert all(v < 3e-3 for v in sums.values()), sums


def test_decoder_13_views_matches_oracle_in_test():
    """engine vs oracle/decoder_ref.py (pinned to the live reference) run here at the full size: every element of every output"""
    D, sd, lat, img, out = _decoder_13v()
    out = {k: v.cpu() for k, v in out.items()}
    torch.cuda.empty_cache()
    ref = D.decoder_forward(sd, D.FULL, lat, img, resolution=512)
    errs = {k: _rel(out[k], ref[k]) for k in BOUNDS_13V}
    print("13v oracle: rel-L2 ", {k: f"{v:.2e}" for k, v in errs.items()})
    with torch.device("cuda"):
        auto = D.decoder_forward({k: v.cuda() for k, v in sd.items()}, D.FULL, lat.cuda(), img.cuda(), resolution=512, gpu_autocast=True)
    floor = {k: _rel(auto[k], ref[k]) for k in BOUNDS_13V}
    print("13v oracle: floor  ", {k: f"{v:.2e}" for k, v in floor.items()})
    bad = {k: (v, BOUNDS_13V[k], floor[k]) for k, v in errs.items() if not v < max(BOUNDS_13V[k], FLOOR_MULT * floor[k])}
    assert not bad, bad
