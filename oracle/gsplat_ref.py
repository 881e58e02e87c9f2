"""ORACLE (test infrastructure; never imported from vist3a_b200): CPU restatement of the 3D-Gaussian rasteriser the reference calls,
gsplat 1.4.0 `rasterization(..., covars=, sh_degree, render_mode="RGB+D", rasterize_mode="classic", radius_clip=0.1, near_plane=1e-10,
packed=False, tile_size=16, eps2d=0.3)` (call site: third_party_model/anysplat/src/model/decoder/decoder_splatting_cuda.py:92-112).

PARITY UNPINNED: gsplat is a third-party CUDA package pinned at 1.4.0 in requirements.txt:17 but absent from /root/reference and not
installable here, and the reference holds no golden renders.  The published algorithm is restated from the package's kernels
(fully_fused_projection_fwd, spherical_harmonics (sh_coeffs_to_color_fast), isect_tiles, rasterize_to_pixels_fwd):
  * camera-space mean / covariance, perspective projection with the clamped Jacobian (limits (W-cx)/fx + 0.3 tan_fovx, cx/fx + 0.3 tan_fovx),
    eps2d added to the diagonal of the 2-D covariance, conic = its inverse, radius = ceil(3 sqrt(b + sqrt(max(0.01, b^2 - det)))),
    culled when z outside [near, far], det <= 0, radius <= radius_clip, or the +-radius box misses the image;
  * colour = max(SH(degree, normalise(mean - camera centre)) + 0.5, 0);
  * a Gaussian contributes to the pixels of the 16x16 tiles its +-radius box touches; per pixel centre (+0.5) front to back by depth:
    sigma = 1/2 d^T conic d, alpha = min(0.999, opacity exp(-sigma)), skipped if sigma < 0 or alpha < 1/255, stop (without adding) when
    T (1 - alpha) <= 1e-4; colour and depth accumulate alpha T, background enters with the final T, alpha = 1 - T.
Analytic known answers pin the pieces that can be pinned (tests/test_oracle_render.py)."""
from __future__ import annotations

import math
from typing import Dict

import torch

TILE = 16


def sh_basis(deg: int, d: torch.Tensor) -> torch.Tensor:
    """real SH basis values [N, (deg+1)^2] at unit vectors d [N, 3] (Sloan's recurrences, as gsplat's sh_coeffs_to_color_fast)"""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    b = [torch.full_like(x, 0.2820947917738781)]
    if deg >= 1:
        b += [-0.48860251190292 * y, 0.48860251190292 * z, -0.48860251190292 * x]
    if deg >= 2:
        z2 = z * z
        f0b = -1.092548430592079 * z
        fC1, fS1 = x * x - y * y, 2 * x * y
        p6 = 0.9461746957575601 * z2 - 0.3153915652525201
        b += [0.5462742152960395 * fS1, f0b * y, p6, f0b * x, 0.5462742152960395 * fC1]
    if deg >= 3:
        f0c = -2.285228997322329 * z2 + 0.4570457994644658
        f1b = 1.445305721320277 * z
        fC2, fS2 = x * fC1 - y * fS1, x * fS1 + y * fC1
        p12 = z * (1.865881662950577 * z2 - 1.119528997770346)
        b += [-0.5900435899266435 * fS2, f1b * fS1, f0c * y, p12, f0c * x, f1b * fC1, -0.5900435899266435 * fC2]
    if deg >= 4:
        f0d = z * (-4.683325804901025 * z2 + 2.007139630671868)
        f1c = 3.31161143515146 * z2 - 0.47308734787878
        f2b = -1.770130769779931 * z
        fC3, fS3 = x * fC2 - y * fS2, x * fS2 + y * fC2
        p20 = 1.984313483298443 * z * p12 - 1.006230589874905 * p6
        b += [0.6258357354491763 * fS3, f2b * fS2, f1c * fS1, f0d * y, p20, f0d * x, f1c * fC1, f2b * fC2, 0.6258357354491763 * fC3]
    return torch.stack(b, dim=-1)


def project(means, covars, viewmat, K, W: int, H: int, near=1e-10, far=1e10, radius_clip=0.1, eps2d=0.3) -> Dict[str, torch.Tensor]:
    """fully_fused_projection + the tile box of isect_tiles.  means [N,3], covars [N,3,3], viewmat [4,4] world->camera, K [3,3] pixels."""
    R, t = viewmat[:3, :3], viewmat[:3, 3]
    mc = means @ R.T + t
    x, y, z = mc[:, 0], mc[:, 1], mc[:, 2]
    valid = (z >= near) & (z <= far)
    zs = torch.where(valid, z, torch.ones_like(z))
    Cc = R @ covars @ R.T
    fx, fy, cx, cy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    tfx, tfy = 0.5 * W / fx, 0.5 * H / fy
    rz = 1.0 / zs
    tx = zs * torch.clamp(x * rz, -(cx / fx + 0.3 * tfx), (W - cx) / fx + 0.3 * tfx)
    ty = zs * torch.clamp(y * rz, -(cy / fy + 0.3 * tfy), (H - cy) / fy + 0.3 * tfy)
    J = torch.zeros(means.shape[0], 2, 3, dtype=means.dtype)
    J[:, 0, 0], J[:, 0, 2] = fx * rz, -fx * tx * rz * rz
    J[:, 1, 1], J[:, 1, 2] = fy * rz, -fy * ty * rz * rz
    c2 = J @ Cc @ J.transpose(1, 2)
    m2 = torch.stack([fx * x * rz + cx, fy * y * rz + cy], dim=-1)
    c00, c01, c10, c11 = c2[:, 0, 0] + eps2d, c2[:, 0, 1], c2[:, 1, 0], c2[:, 1, 1] + eps2d
    det = c00 * c11 - c01 * c10
    valid &= det > 0
    dets = torch.where(valid, det, torch.ones_like(det))
    conic = torch.stack([c11 / dets, -c01 / dets, c00 / dets], dim=-1)
    b = 0.5 * (c00 + c11)
    radius = torch.ceil(3.0 * torch.sqrt(b + torch.sqrt(torch.clamp(b * b - det, min=0.01))))
    valid &= radius > radius_clip
    valid &= ~((m2[:, 0] + radius <= 0) | (m2[:, 0] - radius >= W) | (m2[:, 1] + radius <= 0) | (m2[:, 1] - radius >= H))
    tiles_x, tiles_y = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    tr = radius / TILE
    x0 = torch.clamp(torch.floor(m2[:, 0] / TILE - tr), 0, tiles_x)
    x1 = torch.clamp(torch.ceil(m2[:, 0] / TILE + tr), 0, tiles_x)
    y0 = torch.clamp(torch.floor(m2[:, 1] / TILE - tr), 0, tiles_y)
    y1 = torch.clamp(torch.ceil(m2[:, 1] / TILE + tr), 0, tiles_y)
    valid &= ((x1 - x0) * (y1 - y0)) > 0
    return dict(valid=valid, means2d=m2, depth=z, conic=conic, radius=radius, box=torch.stack([x0, x1, y0, y1], dim=-1))


def render(means, covars, opacities, harmonics, viewmat, K, W: int, H: int, sh_degree: int = 4, background=(0.0, 0.0, 0.0), near=1e-10, far=1e10,
           radius_clip=0.1, eps2d=0.3) -> Dict[str, torch.Tensor]:
    """One view.  harmonics [N, 3, d_sh] (the decoder's layout; the reference permutes to [N, d_sh, 3] for gsplat, same numbers).
    Returns rgb [H,W,3] (unclamped), depth [H,W] (accumulated alpha-weighted depth, render_mode "RGB+D"), alpha [H,W], n_isect."""
    means, covars, opacities, harmonics = means.float(), covars.float(), opacities.float(), harmonics.float()
    viewmat, K = viewmat.float(), K.float()
    p = project(means, covars, viewmat, K, W, H, near, far, radius_clip, eps2d)
    campos = -viewmat[:3, :3].T @ viewmat[:3, 3]
    d = means - campos
    d = d / d.norm(dim=-1, keepdim=True)
    nb = (sh_degree + 1) ** 2
    colors = torch.clamp((harmonics[:, :, :nb] * sh_basis(sh_degree, d)[:, None, :]).sum(-1) + 0.5, min=0.0)  # [N, 3]
    idx = torch.nonzero(p["valid"]).flatten()
    order = torch.argsort(p["depth"][idx], stable=True)      # front to back; equal depths keep the Gaussian order
    idx = idx[order]
    bg = torch.tensor(background, dtype=torch.float32)
    n_isect = int(((p["box"][idx, 1] - p["box"][idx, 0]) * (p["box"][idx, 3] - p["box"][idx, 2])).sum())
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    px, py = xs.flatten().float() + 0.5, ys.flatten().float() + 0.5
    tx, ty = (xs.flatten() // TILE).float(), (ys.flatten() // TILE).float()
    if idx.numel() == 0:
        return dict(rgb=bg.expand(H, W, 3).clone(), depth=torch.zeros(H, W), alpha=torch.zeros(H, W), n_isect=0)
    m2, con, box = p["means2d"][idx], p["conic"][idx], p["box"][idx]
    dx = m2[None, :, 0] - px[:, None]
    dy = m2[None, :, 1] - py[:, None]
    sigma = 0.5 * (con[None, :, 0] * dx * dx + con[None, :, 2] * dy * dy) + con[None, :, 1] * dx * dy
    alpha = torch.clamp(opacities[idx][None] * torch.exp(-sigma), max=0.999)
    in_tile = (tx[:, None] >= box[None, :, 0]) & (tx[:, None] < box[None, :, 1]) & (ty[:, None] >= box[None, :, 2]) & (ty[:, None] < box[None, :, 3])
    skip = (~in_tile) | (sigma < 0) | (alpha < 1.0 / 255.0)
    a_eff = torch.where(skip, torch.zeros_like(alpha), alpha)
    t_next = torch.cumprod(1.0 - a_eff, dim=1)
    t_before = torch.cat([torch.ones(t_next.shape[0], 1), t_next[:, :-1]], dim=1)
    stop_here = (~skip) & (t_next <= 1e-4)
    stopped = torch.cummax(stop_here.int(), dim=1)[0].bool()          # the stopping Gaussian and everything behind it
    use = (~skip) & (~stopped)
    vis = torch.where(use, alpha * t_before, torch.zeros_like(alpha))
    any_stop = stopped[:, -1]
    first_stop = torch.argmax(stop_here.int(), dim=1)
    t_final = torch.where(any_stop, t_before.gather(1, first_stop[:, None])[:, 0], t_next[:, -1])
    rgb = vis @ colors[idx] + t_final[:, None] * bg[None]
    depth = vis @ p["depth"][idx]
    return dict(rgb=rgb.view(H, W, 3), depth=depth.view(H, W), alpha=(1.0 - t_final).view(H, W), n_isect=n_isect)


def random_scene(n: int, seed: int = 0, spread: float = 1.0, scale=(0.02, 0.15), sh_degree: int = 4):
    """synthetic Gaussians in front of a camera at the origin looking down +z: means, covariances R S S^T R^T, opacities, SH [N,3,d_sh]"""
    g = torch.Generator().manual_seed(seed)
    means = torch.cat([(torch.rand(n, 2, generator=g) - 0.5) * 2 * spread, 1.5 + 2.5 * torch.rand(n, 1, generator=g)], dim=-1)
    q = torch.randn(n, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=-1).view(n, 3, 3)
    s = scale[0] + (scale[1] - scale[0]) * torch.rand(n, 3, generator=g)
    covars = R @ torch.diag_embed(s * s) @ R.transpose(1, 2)
    opac = 0.05 + 0.9 * torch.rand(n, generator=g)
    d_sh = (sh_degree + 1) ** 2
    harm = torch.randn(n, 3, d_sh, generator=g) * 0.3
    harm[:, :, 0] += 0.8
    return means, covars, opac, harm


def look_at_camera(W: int, H: int, fov_deg: float = 60.0, shift=(0.0, 0.0, 0.0), yaw_deg: float = 0.0):
    """world->camera matrix (camera near the origin looking down +z, optional yaw / shift) and pixel intrinsics"""
    f = 0.5 * W / math.tan(math.radians(fov_deg) / 2)
    K = torch.tensor([[f, 0.0, W / 2], [0.0, f, H / 2], [0.0, 0.0, 1.0]])
    c, s = math.cos(math.radians(yaw_deg)), math.sin(math.radians(yaw_deg))
    R = torch.tensor([[c, 0.0, -s], [0.0, 1.0, 0.0], [s, 0.0, c]])
    V = torch.eye(4)
    V[:3, :3] = R
    V[:3, 3] = -R @ torch.tensor(shift)
    return V, K
