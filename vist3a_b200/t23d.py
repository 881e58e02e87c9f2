"""Text -> 3D Gaussians on one GPU, and prompt sharding over the GPUs of one box.

Mirrors the per-prompt body of /root/reference/inference_t23d.py:84-137:
    latents = pipe(prompt..., num_inference_steps=50, guidance_scale=cfg, output_type="latent")   -> DenoiseEngine.run
    latents = latents / (1 / std) + mean                                                          -> de-normalise (:104-113)
    output  = stitched_decoder.forward_with_latent(latents, feedforward_image=..., train=False)   -> StitchVAE3DB200
and its data-parallel layout (:58-62): `prompt_list[rank::world_size]`, one full model replica per GPU, no
collective inside a prompt.  The one exchange step is the gather of the final Gaussian tensors
(`all_gather_gaussians`): means, scales, rotations, opacities, harmonics (+ covariances) of every rank.

Text encoding (UMT5) and the Wan VAE decode that produces `feedforward_image` are outside the hot path
(SURVEY §8f): the caller passes text embeddings and either the decoded views or its VAE module (`views_from_vae`).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import ops
from .pipeline import DenoiseEngine
from .stitched_decoder import EncoderOutput, Gaussians

# AutoencoderKLWan.config.latents_mean / latents_std (the reference's copy: utils/wan_utils.py:925-960)
WAN_LATENTS_MEAN = (-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632,
                    -0.1922, -0.9497, 0.2503, -0.2921)
WAN_LATENTS_STD = (2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382,
                   1.1253, 2.8251, 1.9160)

GAUSSIAN_FIELDS = ("means", "scales", "rotations", "opacities", "harmonics", "covariances")


def shard_prompts(prompts: Sequence, rank: int, world_size: int) -> list:
    """inference_t23d.py:62 -- `prompt_list[rank :: dist.get_world_size()]`"""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(prompts[rank::world_size])


class TextTo3DGS:
    """One prompt -> EncoderOutput; owns the CUDA-graphed denoise engine and the decoder."""

    def __init__(self, transformer, decoder, *, views: int = 13, resolution: int = 512, text_len: int = 512,
                 num_inference_steps: int = 50, guidance_scale: float = 6.0, flow_shift: float = 5.0, use_graph: bool = True,
                 decoder_graph: bool = True):
        if (views - 1) % 4:
            raise ValueError("views must be 4k+1 (Wan VAE temporal stride 4)")
        self.tr, self.dec = transformer, decoder
        # the stitched decode replayed from a CUDA graph as well (stitched_decoder.DecoderGraph: the eager forward needs 60 ms of host time
        # to queue 80 ms of kernels); outputs are copies, freshly allocated per prompt as in the eager call
        self.decoder_graph = decoder_graph and use_graph and hasattr(decoder, "forward_with_latent_graph")
        self.dev = transformer.device
        T = (views - 1) // 4 + 1
        self.latent_shape = (1, 16, T, resolution // 8, resolution // 8)
        self.engine = DenoiseEngine(transformer, self.latent_shape, text_len, num_inference_steps=num_inference_steps,
                                    guidance_scale=guidance_scale, flow_shift=flow_shift, use_graph=use_graph)
        # x * std + mean as one kernel:  out = std_c * x + mean_c * 1
        shp = (1, 16, 1, 1, 1)
        self._std = torch.tensor(WAN_LATENTS_STD, device=self.dev).view(shp).expand(self.latent_shape).contiguous()
        self._mean = torch.tensor(WAN_LATENTS_MEAN, device=self.dev).view(shp).expand(self.latent_shape).contiguous()
        self._lat = torch.empty(self.latent_shape, dtype=torch.float32, device=self.dev)

    @torch.no_grad()
    def denoise(self, noise: torch.Tensor, text_cond: torch.Tensor, text_uncond: torch.Tensor) -> torch.Tensor:
        """50-step CFG sampling -> de-normalised VAE latent [1, 16, T, h, w] fp32 (what the decoder is stitched onto)."""
        x = self.engine.run(noise, text_cond, text_uncond)
        ops.fma_rows(x.view(-1, x.shape[-1]), self._std.view(-1, x.shape[-1]), self._mean.view(-1, x.shape[-1]), out=self._lat.view(-1, x.shape[-1]))
        return self._lat

    @torch.no_grad()
    def generate(self, noise: torch.Tensor, text_cond: torch.Tensor, text_uncond: torch.Tensor,
                 feedforward_image: Optional[torch.Tensor] = None, *, vae=None) -> EncoderOutput:
        """One prompt.  `feedforward_image` [1, 3, V, 448, 448] in [-1, 1] are the decoded views the Gaussian head reads; when it is not
        given, `vae` (the caller's `pipe.vae`: any object with diffusers' `decode(latents, return_dict=False)`) decodes the de-normalised
        latent and the frames are resized as the reference does (inference_t23d.py:114-123) -- the VAE itself is outside this engine."""
        latent = self.denoise(noise, text_cond, text_uncond)
        if feedforward_image is None:
            if vae is None:
                raise ValueError("generate: pass feedforward_image, or vae= to decode the views from the latent")
            feedforward_image = views_from_vae(vae, latent)
        if self.decoder_graph:
            return self.dec.forward_with_latent_graph(latent, feedforward_image)
        return self.dec.forward_with_latent(latent, feedforward_image, train=False)


def views_from_vae(vae, latent: torch.Tensor, size: int = 448) -> torch.Tensor:
    """inference_t23d.py:114-123: `samples = pipe.vae.decode(latents, return_dict=False)[0]`, then a trilinear (align_corners=False)
    resize of the frames to size x size with the frame count kept.  With the engine's own decoder (`WanVAEDecoderB200`) both steps run on
    the sm_100a kernels; any other object is treated as the caller's VAE module (diffusers' `decode(latents, return_dict=False)`)."""
    import torch.nn.functional as F

    from .wan_vae import WanVAEDecoderB200

    if isinstance(vae, WanVAEDecoderB200):
        samples = vae.decode(latent, return_dict=False)[0]                # [B, 3, T, 8h, 8w] fp32 in [-1, 1]
        return ops.resize_planes(samples, size, size)
    p = next(iter(vae.parameters()), None) if hasattr(vae, "parameters") else None
    samples = vae.decode(latent if p is None else latent.to(p.dtype), return_dict=False)[0]
    return F.interpolate(samples.float(), (samples.shape[2], size, size), mode="trilinear", align_corners=False)


def pack_gaussians(g: Gaussians, with_covariances: bool = False) -> torch.Tensor:
    """[B, N, F] fp32 record per Gaussian: means 3 | scales 3 | rotations 4 | opacity 1 | harmonics 3*d_sh (| covariances 9)."""
    B, N = g.opacities.shape
    parts = [g.means, g.scales, g.rotations, g.opacities.unsqueeze(-1), g.harmonics.reshape(B, N, -1)]
    if with_covariances:
        parts.append(g.covariances.reshape(B, N, 9))
    return torch.cat(parts, dim=-1).contiguous()


def unpack_gaussians(rec: torch.Tensor, d_sh: int, with_covariances: bool = False) -> Dict[str, torch.Tensor]:
    B, N, _ = rec.shape
    sizes = [3, 3, 4, 1, 3 * d_sh] + ([9] if with_covariances else [])
    m, s, r, o, h, *c = rec.split(sizes, dim=-1)
    out = {"means": m, "scales": s, "rotations": r, "opacities": o.squeeze(-1), "harmonics": h.reshape(B, N, 3, d_sh)}
    if c:
        out["covariances"] = c[0].reshape(B, N, 3, 3)
    return out


def _field_views(flat: torch.Tensor, B: int, N: int, d_sh: int, with_covariances: bool) -> Dict[str, torch.Tensor]:
    """views into one rank's field-major buffer (layout of vist3a_b200.ops.alloc_gaussian_fields)"""
    P, out, off = B * N, {}, 0
    for k, w, shp in (("means", 3, (B, N, 3)), ("scales", 3, (B, N, 3)), ("rotations", 4, (B, N, 4)), ("opacities", 1, (B, N)),
                      ("harmonics", 3 * d_sh, (B, N, 3, d_sh))) + ((("covariances", 9, (B, N, 3, 3)),) if with_covariances else ()):
        out[k] = flat[off:off + P * w].view(shp)
        off += P * w
    return out


class GaussianGather:
    """An all-gather of Gaussians in flight (`all_gather_gaussians_async`).  `wait()` makes the CURRENT stream wait for the collective (the host is
    not blocked on NCCL) and returns one dict of tensors per rank; the handle keeps the send and receive buffers alive until then."""

    def __init__(self, work, finish, keep):
        self._work, self._finish, self._keep = work, finish, keep

    def wait(self) -> List[Dict[str, torch.Tensor]]:
        if self._work is not None:
            self._work.wait()
            self._work = None
        return self._finish()


def all_gather_gaussians_async(g: Gaussians, group=None, with_covariances: bool = False, fixed_count: bool = False) -> GaussianGather:
    """The single exchange step of the multi-GPU path: every rank contributes the Gaussians of its prompt and
    receives everybody's.  One `all_gather_into_tensor` of a fixed-size buffer (N = V*H*W per prompt when voxelisation is off), NCCL
    over NVLink on GPUs, gloo on CPU in the tests.  The collective is issued asynchronously (`async_op=True`: on NCCL it runs on the
    process group's own stream behind the work already queued on the current one), so the caller can queue the next prompt's denoising
    before calling `wait()` -- the gather then hides under it.

    The decoder writes its outputs field-major into ONE flat buffer (`Gaussians.packed`), so the collective sends that buffer as it is and
    the received segments are viewed, not copied (2.1 ms for 2 x 0.99 GB over NVLink; packing records with torch.cat and re-splitting them
    cost 35 ms).  Gaussians built elsewhere (no `packed`) take the record path.

    fixed_count: every rank holds the same number of Gaussians (voxelised fusion and the confidence branches are off: N = V*H*W).  Otherwise
    the counts are agreed on first -- one tiny collective and a host read, which synchronises the stream."""
    import torch.distributed as dist
    import torch.nn.functional as F

    world = dist.get_world_size(group)
    B, N = g.opacities.shape
    d_sh = g.harmonics.shape[-1]
    all_counts = None
    if not fixed_count:
        # voxelised fusion makes the Gaussian count data dependent: agree on the counts first (one tiny collective); ranks with fewer Gaussians
        # pad to the largest count with the reference's fill values (points -1e4, opacity 0: models/anysplat_stitched.py:448-455) and every
        # returned dict is cut back to its rank's own count
        counts = torch.tensor([N], dtype=torch.int64, device=g.opacities.device)
        gathered = torch.empty((world,), dtype=torch.int64, device=counts.device)
        dist.all_gather_into_tensor(gathered, counts, group=group)
        all_counts = [int(c) for c in gathered.tolist()]
        nmax = max(all_counts)
        if min(all_counts) != nmax:
            pad = nmax - N

            def padn(t, value=0.0):
                return F.pad(t, (0, 0) * (t.dim() - 2) + (0, pad), value=value) if pad else t

            g = Gaussians(means=padn(g.means, -1e4), covariances=padn(g.covariances), harmonics=padn(g.harmonics), opacities=padn(g.opacities),
                          scales=padn(g.scales), rotations=padn(g.rotations))
            N = nmax
    if g.packed is not None:
        per = B * N * (11 + 3 * d_sh + (9 if with_covariances else 0))   # covariances are the last field: leave them out by length
        src = g.packed[:per]
        out = torch.empty((world * per,), dtype=src.dtype, device=src.device)
        work = dist.all_gather_into_tensor(out, src, group=group, async_op=True)
        return GaussianGather(work, lambda: [_field_views(out[r * per:(r + 1) * per], B, N, d_sh, with_covariances) for r in range(world)], (g, src, out))
    rec = pack_gaussians(g, with_covariances)
    out = torch.empty((world * B,) + tuple(rec.shape[1:]), dtype=rec.dtype, device=rec.device)  # rank-major concatenation
    work = dist.all_gather_into_tensor(out, rec, group=group, async_op=True)

    def finish():
        res = [unpack_gaussians(out[r * B:(r + 1) * B], d_sh, with_covariances) for r in range(world)]
        return res if all_counts is None else [{k: v[:, :all_counts[r]] for k, v in d.items()} for r, d in enumerate(res)]

    return GaussianGather(work, finish, (rec, out))


def all_gather_gaussians(g: Gaussians, group=None, with_covariances: bool = False, fixed_count: bool = False) -> List[Dict[str, torch.Tensor]]:
    """`all_gather_gaussians_async(...).wait()`: the gather in line with the caller's stream."""
    return all_gather_gaussians_async(g, group, with_covariances, fixed_count).wait()


def generate_sharded(engine: "TextTo3DGS", items, *, group=None, with_covariances: bool = False):
    """The prompts of THIS rank (`shard_prompts`), each followed by the exchange of its Gaussians with the other ranks, pipelined: the
    all-gather of prompt p is in flight while prompt p + 1 is denoised and decoded.  `items` yields (noise, text_cond, text_uncond,
    feedforward_image or None [, vae]); yields (EncoderOutput of this rank, [dict of Gaussian tensors per rank]) per prompt.  Every rank must
    bring the same number of prompts (the exchange is a collective)."""
    cfg = engine.dec.cfg
    fixed = not (cfg.voxelize or cfg.render_conf or cfg.opacity_conf)
    pending = None
    for it in items:
        noise, tc, tu, img = it[:4]
        out = engine.generate(noise, tc, tu, img, vae=it[4] if len(it) > 4 else None)
        handle = all_gather_gaussians_async(out.gaussians, group, with_covariances, fixed_count=fixed)
        if pending is not None:
            yield pending[0], pending[1].wait()
        pending = (out, handle)
    if pending is not None:
        yield pending[0], pending[1].wait()
