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


def interpolate_context_cameras(extrinsics: torch.Tensor, intrinsics: torch.Tensor, t: int = 10):
    """Camera path of the reference's result video (`save_interpolated_video`, AS/misc/image_io.py:111-183): between every pair of
    neighbouring context views, the view itself plus t interpolated ones -- translation and intrinsics linearly, rotation as the linear
    blend of the two matrices projected back onto SO(3) through its SVD (U V^T).  extrinsics [B,V,4,4] camera-to-world, intrinsics
    [B,V,3,3] -> [B,(V-1)(t+1),4,4], [B,(V-1)(t+1),3,3]  (13 views, t = 10: 132 frames; the last context view is not part of the path,
    as in the reference, which appends it only after the path has been concatenated, :176-181)."""
    B, V = extrinsics.shape[:2]
    ex, ix = [], []
    for i in range(V - 1):
        ex.append(extrinsics[:, i:i + 1])
        ix.append(intrinsics[:, i:i + 1])
        for j in range(1, t + 1):
            a = j / (t + 1)
            s, e = extrinsics[:, i], extrinsics[:, i + 1]
            rot = (1 - a) * s[:, :3, :3] + a * e[:, :3, :3]
            u, _, vh = torch.linalg.svd(rot)
            m = torch.eye(4, device=extrinsics.device, dtype=extrinsics.dtype).unsqueeze(0).repeat(B, 1, 1)
            m[:, :3, :3] = u @ vh
            m[:, :3, 3] = (1 - a) * s[:, :3, 3] + a * e[:, :3, 3]
            ex.append(m.unsqueeze(1))
            ix.append(((1 - a) * intrinsics[:, i] + a * intrinsics[:, i + 1]).unsqueeze(1))
    return torch.cat(ex, dim=1), torch.cat(ix, dim=1)


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

    @torch.no_grad()
    def render_interpolated_views(self, output, image_shape=(448, 448), t: int = 10) -> DecoderOutput:
        """The frames of the reference's gs.mp4 (`save_interpolated_video`, inference_t23d.py:146-155): the interpolated camera path through the
        predicted context views (132 frames for 13 views), rendered from all Gaussians.  Colour clipped to [0, 1] as in :193."""
        pose = output.pred_context_pose
        ex, ix = interpolate_context_cameras(pose["extrinsic"].float().cpu(), pose["intrinsic"].float().cpu(), t)
        return self.rendering_fn(output.gaussians, ex, ix, image_shape=image_shape)
