"""Novel-view renderer on the sm_100a rasteriser -- drop-in for `DecoderSplattingCUDA.rendering_fn`
(/root/reference third_party_model/anysplat/src/model/decoder/decoder_splatting_cuda.py:43-125), the consumer of the decoder's
Gaussians in inference_t23d.py:146-155 (132 interpolated views -> mp4) and utils/reward.py.  The reference loops over batch and
views and calls gsplat.rasterization once per view with the full Gaussian set; so does this (vist3a_gs_project + vist3a_gs_rasterize)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import ops
from .stitched_decoder import Gaussians


@dataclass
class DecoderOutput:   # AS/model/decoder/decoder.py: color [B,V,3,H,W], depth [B,V,H,W], alpha [B,V,H,W]
    color: torch.Tensor
    depth: Optional[torch.Tensor]
    alpha: Optional[torch.Tensor]
    lod_rendering: Optional[dict] = None


class DecoderSplattingB200(torch.nn.Module):
    def __init__(self, background_color: Sequence[float] = (1.0, 1.0, 1.0)):
        super().__init__()
        self.background_color = tuple(float(c) for c in background_color)

    @torch.no_grad()
    def rendering_fn(self, gaussians: Gaussians, extrinsics: torch.Tensor, intrinsics: torch.Tensor, near=None, far=None,
                     image_shape=(448, 448), depth_mode=None, cam_rot_delta=None, cam_trans_delta=None, cov_ignore: bool = False) -> DecoderOutput:
        """extrinsics [B,V,4,4] camera-to-world, intrinsics [B,V,3,3] normalised by the image size (rows 0 / 1 divided by W / H), as the
        reference passes them (decoder_splatting_cuda.py:80-86).  near / far are accepted and ignored like in the reference call
        (near_plane=1e-10 is hard-coded there, :104-106)."""
        if cov_ignore:
            raise NotImplementedError("cov_ignore=True (covariances rebuilt from scales / rotations by gsplat) is not on the reference's path")
        if cam_rot_delta is not None or cam_trans_delta is not None:
            raise NotImplementedError("camera deltas are a training-time feature")
        B, V = intrinsics.shape[:2]
        H, W = image_shape
        w2c = torch.linalg.inv(extrinsics.float().cpu())
        K = intrinsics.float().cpu().clone()
        K[:, :, 0] *= W
        K[:, :, 1] *= H
        sh_degree = int(round(gaussians.harmonics.shape[-1] ** 0.5)) - 1
        color = torch.empty((B, V, 3, H, W), dtype=torch.float32, device=gaussians.means.device)
        depth = torch.empty((B, V, H, W), dtype=torch.float32, device=gaussians.means.device)
        alpha = torch.empty((B, V, H, W), dtype=torch.float32, device=gaussians.means.device)
        for b in range(B):
            m, c = gaussians.means[b].float().contiguous(), gaussians.covariances[b].float().contiguous()
            o, h = gaussians.opacities[b].float().contiguous(), gaussians.harmonics[b].float().contiguous()
            for v in range(V):
                r = ops.gs_render(m, c, o, h, w2c[b, v], K[b, v], W, H, sh_degree=sh_degree, background=self.background_color)
                color[b, v] = r["rgb"].clamp(0.0, 1.0).permute(2, 0, 1)     # :113-115
                depth[b, v], alpha[b, v] = r["depth"], r["alpha"]
        return DecoderOutput(color, depth, alpha)

    forward = rendering_fn

    @torch.no_grad()
    def render_context_views(self, output, image_shape=(448, 448)) -> DecoderOutput:
        """Re-render the scene from the cameras the decoder predicted for its own context views (`EncoderOutput.pred_context_pose`:
        camera-to-world extrinsics, normalised intrinsics) -- the first thing inference_t23d.py:139-155 does with the Gaussians."""
        pose = output.pred_context_pose
        return self.rendering_fn(output.gaussians, pose["extrinsic"], pose["intrinsic"], image_shape=image_shape)
