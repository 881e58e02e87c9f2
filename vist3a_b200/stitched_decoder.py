"""Stitched latent -> 3D-Gaussian decoder on the sm_100a kernels -- drop-in for
`StitchVAE3D.forward_with_latent` (/root/reference/models/stitched_model.py:165-173), i.e. the
trilinear T-upsample + stitching Conv3d (models/stitching_layer_builder.py:21-42) feeding
`AnySplatStitched.forward` (models/anysplat_stitched.py:167-525): 22 DINOv2 blocks, 24 x (frame,
global) alternating-attention blocks, camera head, DPT depth head, DPT Gaussian head and the
per-pixel Gaussian adapter (`render_conf=False`, `opacity_conf=False`); `DecoderConfig.voxelize` selects the voxelised-fusion
branch (models/anysplat_stitched.py:419-455 -> AS/model/encoder/anysplat.py:298-335) the released AnySplat configs enable.

Weights come from a state dict with the REFERENCE's key names (`stitching_layer.*`,
`stitched_3d_model.encoder.*`), so `anysplat_stitched.pth` + the AnySplat checkpoint load unchanged.

Data layout in HBM
  tokens   fp32 [B*V*1029, C] residual stream (cls/camera + 4 register tokens first in every view);
           the stitching GEMM and the final DINO norm write patch tokens straight behind the special
           tokens through row maps, so no concatenation kernels exist;
  qkv      bf16 [rows, 3C] written by one GEMM, normalised / rotated in place, consumed by attention
           through strided views (frame: 13 sequences of 1029; global: one of 13377);
  inter    4 x fp32 [B*V*1029, 2C] (frame | global halves) read by the heads through row maps;
  heads    NHWC fp32 feature maps; every 3x3 convolution is an implicit GEMM whose A operand is
           fetched by 4-D TMA boxes (zero padding = out-of-bounds fill), TF32 tcgen05 MMAs
           (the reference runs these convs through cuDNN with TF32 allowed);
  output   means/scales/rotations/opacities/harmonics/covariances fp32, contiguous per field.
dtype policy: bf16 GEMM/attention operands and fp32 accumulation + fp32 residual stream in the
transformer; fp32 activations with TF32 MMAs in the heads; fp32 LayerNorm statistics everywhere.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import ops

E = "stitched_3d_model.encoder."
N_SPECIAL = 5  # cls/camera token + 4 register tokens


# ------------------------------------------------------------------------------------------------
# return types (AS/model/types.py:8-14, AS/model/encoder/encoder.py:16-23)
# ------------------------------------------------------------------------------------------------
@dataclass
class Gaussians:
    means: torch.Tensor          # [B, N, 3]
    covariances: torch.Tensor    # [B, N, 3, 3]
    harmonics: torch.Tensor      # [B, N, 3, d_sh]
    opacities: torch.Tensor      # [B, N]
    scales: torch.Tensor         # [B, N, 3]
    rotations: torch.Tensor      # [B, N, 4] (xyzw)
    # not a reference field: the flat fp32 buffer all the fields above are views of (field-major: means | scales | rotations | opacities |
    # harmonics | covariances, each over the B*N Gaussians) -- what the multi-GPU gather sends without packing (t23d.all_gather_gaussians)
    packed: Optional[torch.Tensor] = None


@dataclass
class EncoderOutput:
    gaussians: Gaussians
    pred_pose_enc_list: Optional[List[torch.Tensor]]
    pred_context_pose: dict
    depth_dict: dict
    infos: dict
    distill_infos: Optional[dict] = None
    last_pred_pose_enc: Optional[torch.Tensor] = None


@dataclass(frozen=True)
class DecoderConfig:
    embed_dim: int = 1024
    num_heads: int = 16
    dino_blocks: int = 22
    agg_depth: int = 24
    cam_heads: int = 16
    cam_trunk: int = 4
    dpt_features: int = 256
    dpt_out_channels: Tuple[int, ...] = (256, 512, 1024, 1024)
    pos_grid: int = 37
    patch: int = 14
    sh_degree: int = 4
    latent_channels: int = 16
    inter_layers: Tuple[int, ...] = (4, 11, 17, 23)
    resolution: int = 512  # video resolution; the VAE latent grid is resolution / 8
    patch_embed: bool = False  # un-stitched AnySplat encoder: DINOv2 patch embedding (image input, `forward_images`) instead of the stitching
                               # layer; dino_blocks then counts ALL DINO blocks (24 in the released model)
    voxelize: bool = False   # EncoderAnySplatCfg.voxelize (AS/model/encoder/anysplat.py:125; true in config/experiment/*.yaml)
    voxel_size: float = 0.002  # config/experiment/dl3dv.yaml:20
    # confidence-quantile branches (EncoderAnySplatCfg.render_conf / opacity_conf / conf_threshold, AS/model/encoder/anysplat.py:83-125;
    # models/anysplat_stitched.py:381-387, 443-467; off in every released config)
    render_conf: bool = False
    opacity_conf: bool = False
    conf_threshold: float = 0.1

    @property
    def d_sh(self):
        return (self.sh_degree + 1) ** 2

    @property
    def raw_gs_dim(self):
        return 1 + 7 + 3 * self.d_sh


def _conv_w(w: torch.Tensor) -> torch.Tensor:
    """[N, C, kh, kw] -> [N, kh*kw*C] with k = (dy*kw + dx)*C + c (the implicit-GEMM / im2col order)"""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)


def _sincos(p: torch.Tensor, d: int) -> torch.Tensor:
    om = torch.arange(d // 2, dtype=torch.double) / (d / 2.0)
    om = 1.0 / 100 ** om
    out = torch.einsum("m,d->md", p.reshape(-1).double(), om)
    return torch.cat([out.sin(), out.cos()], dim=1).float()


def _dpt_pos_tables(C: int, h: int, w: int, W_img: int, H_img: int, ratio: float = 0.1):
    """separable form of create_uv_grid + position_grid_to_embed (AS/.../heads/utils.py:11-108, dpt_head.py:267-277):
    channels [0, C/2) depend on x only, [C/2, C) on y only.  -> pos_x [w, C/2], pos_y [h, C/2]"""
    aspect = W_img / H_img
    diag = (aspect ** 2 + 1.0) ** 0.5
    sx, sy = aspect / diag, 1.0 / diag
    xs = torch.linspace(-sx * (w - 1) / w, sx * (w - 1) / w, steps=w)
    ys = torch.linspace(-sy * (h - 1) / h, sy * (h - 1) / h, steps=h)
    return (_sincos(xs, C // 2) * ratio).contiguous(), (_sincos(ys, C // 2) * ratio).contiguous()


def param_shapes(cfg: DecoderConfig) -> Dict[str, Tuple[int, ...]]:
    """Parameter manifest of the stitched decoder under the reference's state-dict keys (what a real
    `anysplat_stitched.pth` + AnySplat checkpoint provide; tests check it against the oracle's manifest)."""
    C, C2, Fd, oc = cfg.embed_dim, 2 * cfg.embed_dim, cfg.dpt_features, cfg.dpt_out_channels
    s: Dict[str, Tuple[int, ...]] = {}
    if cfg.patch_embed:   # AS/.../layers/patch_embed.py:65 (Conv2d k = s = patch)
        s[E + "aggregator.patch_embed.patch_embed.proj.weight"] = (C, 3, cfg.patch, cfg.patch)
        s[E + "aggregator.patch_embed.patch_embed.proj.bias"] = (C,)
    else:
        s["stitching_layer.weight"], s["stitching_layer.bias"] = (C, cfg.latent_channels, 5, 3, 3), (C,)

    def lin(name, n, k):
        s[name + ".weight"], s[name + ".bias"] = (n, k), (n,)

    def norm(name, n):
        s[name + ".weight"], s[name + ".bias"] = (n,), (n,)

    def block(p, dim, head_dim=None):
        norm(p + "norm1", dim)
        lin(p + "attn.qkv", 3 * dim, dim)
        if head_dim:
            norm(p + "attn.q_norm", head_dim)
            norm(p + "attn.k_norm", head_dim)
        lin(p + "attn.proj", dim, dim)
        s[p + "ls1.gamma"] = (dim,)
        norm(p + "norm2", dim)
        lin(p + "mlp.fc1", 4 * dim, dim)
        lin(p + "mlp.fc2", dim, 4 * dim)
        s[p + "ls2.gamma"] = (dim,)

    pe, ag, ch = E + "aggregator.patch_embed.", E + "aggregator.", E + "camera_head."
    s[pe + "cls_token"], s[pe + "pos_embed"] = (1, 1, C), (1, 1 + cfg.pos_grid ** 2, C)
    s[pe + "register_tokens"], s[pe + "mask_token"] = (1, 4, C), (1, C)
    for i in range(cfg.dino_blocks):
        block(pe + f"blocks.{i}.", C)
    norm(pe + "norm", C)
    s[ag + "camera_token"], s[ag + "register_token"] = (1, 2, 1, C), (1, 2, 4, C)
    for i in range(cfg.agg_depth):
        block(ag + f"frame_blocks.{i}.", C, C // cfg.num_heads)
        block(ag + f"global_blocks.{i}.", C, C // cfg.num_heads)
    s[ch + "empty_pose_tokens"] = (1, 1, 9)
    for i in range(cfg.cam_trunk):
        block(ch + f"trunk.{i}.", C2)
    norm(ch + "token_norm", C2)
    norm(ch + "trunk_norm", C2)
    lin(ch + "embed_pose", C2, 9)
    lin(ch + "poseLN_modulation.1", 3 * C2, C2)
    lin(ch + "pose_branch.fc1", C2 // 2, C2)
    lin(ch + "pose_branch.fc2", 9, C2 // 2)

    def conv(name, co, ci, k, bias=True):
        s[name + ".weight"] = (co, ci, k, k)
        if bias:
            s[name + ".bias"] = (co,)

    for head, out_dim in (("depth_head.", 2), ("gaussian_param_head.", cfg.raw_gs_dim + 1)):
        h = E + head
        hf2 = 128 if out_dim > 50 else 32
        norm(h + "norm", C2)
        for k in range(4):
            conv(h + f"projects.{k}", oc[k], C2, 1)
            conv(h + f"scratch.layer{k + 1}_rn", Fd, oc[k], 3, bias=False)
        conv(h + "resize_layers.0", oc[0], oc[0], 4)
        conv(h + "resize_layers.1", oc[1], oc[1], 2)
        conv(h + "resize_layers.3", oc[3], oc[3], 3)
        for r in (1, 2, 3, 4):
            p = h + f"scratch.refinenet{r}."
            conv(p + "out_conv", Fd, Fd, 1)
            for u in ((1, 2) if r != 4 else (2,)):
                conv(p + f"resConfUnit{u}.conv1", Fd, Fd, 3)
                conv(p + f"resConfUnit{u}.conv2", Fd, Fd, 3)
        conv(h + "scratch.output_conv1", Fd // 2, Fd, 3)
        conv(h + "scratch.output_conv2.0", hf2, Fd // 2, 3)
        conv(h + "scratch.output_conv2.2", out_dim, hf2, 1)
        if out_dim > 50:
            conv(h + "input_merger.0", hf2, 3, 7)
    return s


def random_state_dict(cfg: DecoderConfig, seed: int = 0, device="cuda") -> Dict[str, torch.Tensor]:
    """Random-init weights of the named architecture for benchmarking (no checkpoint is reachable offline), generated on
    `device`.  Scales keep activations O(1) through the 70 blocks: matrices U(-1,1) * 0.8 sqrt(3 / fan_in), biases
    0.02 U, norm weights 1 + 0.1 U, LayerScale 1.0 (DINOv2) / 0.01 (aggregator, camera trunk) as the reference
    constructors set them (AS/.../layers/block.py:44-77, aggregator.py:99-133)."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(cfg).items():
        u = torch.rand(shp, generator=g, device=device) * 2 - 1
        if k.endswith("gamma"):
            t = (1.0 if "patch_embed.blocks" in k else 0.01) * (1.0 + 0.2 * u)
        elif "norm" in k and k.endswith(".weight"):
            t = 1.0 + 0.1 * u
        elif k.endswith(".bias"):
            t = 0.02 * u
        elif k.endswith("token") or k.endswith("tokens") or k.endswith("pos_embed"):
            t = 0.02 * u
        else:
            fan_in = math.prod(shp[1:]) if len(shp) > 1 else shp[0]
            if "resize_layers.0" in k or "resize_layers.1" in k:
                fan_in = shp[0]
            t = u * math.sqrt(3.0 / fan_in) * 0.8
        if k.endswith("camera_head.pose_branch.fc2.bias"):  # quaternion w ~ 1, FoV ~ 1 rad after the 4 refinement iterations
            t = t + torch.tensor([0, 0, 0, 0, 0, 0, 0.25, 0.25, 0.25], device=device)
        sd[k] = t
    return sd


def _clone_tree(obj, packed_src=None, packed_dst=None):
    """Deep copy of the tensors of a decoder output (dataclasses, dicts, lists); tensors that are views of `packed_src` (the field-major
    Gaussian buffer) become the same views of `packed_dst`, so that the copy keeps the zero-copy layout the multi-GPU gather relies on."""
    if isinstance(obj, torch.Tensor):
        if packed_src is not None and obj.untyped_storage().data_ptr() == packed_src.untyped_storage().data_ptr() and obj.dtype == packed_src.dtype:
            return packed_dst.as_strided(obj.size(), obj.stride(), obj.storage_offset() - packed_src.storage_offset())
        return obj.clone()
    if isinstance(obj, dict):
        return {k: _clone_tree(v, packed_src, packed_dst) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_clone_tree(v, packed_src, packed_dst) for v in obj)
    if hasattr(obj, "__dataclass_fields__"):
        return type(obj)(**{k: _clone_tree(getattr(obj, k), packed_src, packed_dst) for k in obj.__dataclass_fields__})
    return obj


def clone_output(out: EncoderOutput) -> EncoderOutput:
    src = out.gaussians.packed
    dst = src.clone() if src is not None else None
    g = _clone_tree(out.gaussians, src, dst)
    if dst is not None:
        g.packed = dst
    rest = {k: _clone_tree(getattr(out, k)) for k in out.__dataclass_fields__ if k != "gaussians"}
    return EncoderOutput(gaussians=g, **rest)


class DecoderGraph:
    """One `StitchVAE3DB200.forward_with_latent` of fixed input shapes as a CUDA graph.  The eager forward queues ~900 kernels and needs
    ~60 ms of host time for 80 ms of device time (measured, tools/decoder_graph_try.py): a busy host makes it host-bound; the replay costs
    0.3 ms of host time and 78 ms on the device.  Inputs are copied into the graph's static buffers; `clone=True` (default) returns freshly
    allocated outputs like the eager call does (0.9 GB copied at 13 views: 0.3 ms), `clone=False` the graph's own output tensors, valid
    until the next replay.  Not available for the data-dependent branches (voxelised fusion, confidence quantiles): those read counts on
    the host."""

    def __init__(self, dec: "StitchVAE3DB200", latent: torch.Tensor, feedforward_image: torch.Tensor):
        cfg = dec.cfg
        if cfg.voxelize or cfg.render_conf or cfg.opacity_conf:
            raise NotImplementedError("DecoderGraph: the voxelised / confidence-filtered outputs have data-dependent sizes")
        self.dec = dec
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self.latent = latent.detach().to(dec.device).clone()
            self.image = feedforward_image.detach().to(dec.device).clone()
            for _ in range(2):   # warm-up outside the capture: function attributes, tensor-map caches, allocator pools
                dec.forward_with_latent(self.latent, self.image)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = dec.forward_with_latent(self.latent, self.image)

    @torch.no_grad()
    def __call__(self, latent: torch.Tensor, feedforward_image: torch.Tensor, clone: bool = True) -> EncoderOutput:
        if latent.shape != self.latent.shape or feedforward_image.shape != self.image.shape:
            raise ValueError(f"DecoderGraph captured for latent {tuple(self.latent.shape)} / views {tuple(self.image.shape)}, "
                             f"got {tuple(latent.shape)} / {tuple(feedforward_image.shape)}")
        self.latent.copy_(latent, non_blocking=True)
        self.image.copy_(feedforward_image, non_blocking=True)
        self.graph.replay()
        return clone_output(self.out) if clone else self.out


class StitchVAE3DB200(torch.nn.Module):
    """B200-native StitchVAE3D (inference: `forward_with_latent`)."""

    def __init__(self, config: DecoderConfig = DecoderConfig(), device="cuda"):
        super().__init__()
        self.cfg = config
        self.device = torch.device(device)
        self.w: Dict[str, torch.Tensor] = {}
        self._tables = {}
        self.keep_voxel_inputs = False
        self.voxel_inputs = None
        self.diffusion_vae = None        # optional: the caller's Wan VAE module, only used by forward(images, ...)
        if config.embed_dim // config.num_heads != 64:
            raise NotImplementedError("the QK-norm + 2-D RoPE kernel implements the aggregator's 64-wide heads")

    def eval(self):
        return self

    def _apply(self, fn):
        return self

    # ------------------------------------------------------------------ weights
    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor], config: DecoderConfig = DecoderConfig(), device="cuda"):
        m = cls(config, device)
        m.load_weights(sd)
        return m

    def load_weights(self, sd: Dict[str, torch.Tensor]):
        cfg, dev = self.cfg, self.device
        C = cfg.embed_dim
        w: Dict[str, torch.Tensor] = {}

        def f32(t):
            return t.detach().float().contiguous().to(dev)

        def bf(t):
            return t.detach().float().to(dev, torch.bfloat16).contiguous()

        pe = E + "aggregator.patch_embed."
        if not cfg.patch_embed and (pe + "patch_embed.proj.weight" in sd or pe + f"blocks.{cfg.dino_blocks}.norm1.weight" in sd):
            # an un-stitched AnySplat checkpoint numbers its DINO blocks 0..23; the stitched model's blocks.i are original blocks.(i + k)
            raise ValueError(f"load_weights: the state dict holds patch_embed.patch_embed.proj / blocks.{cfg.dino_blocks}: it carries the UN-stitched "
                             "block numbering; renumber it with vist3a_b200.checkpoint.renumber_stitched_blocks(sd, k) (k of enc_blocks_k) first")
        if cfg.patch_embed:
            pw = sd[pe + "patch_embed.proj.weight"].detach().float().reshape(C, -1)     # k = c*p*p + py*p + px
            self._pe_k = (pw.shape[1] + 7) // 8 * 8                                      # 588 -> 592 (16-byte TMA row stride)
            w["pe.w"] = bf(F.pad(pw, (0, self._pe_k - pw.shape[1])))
            w["pe.b"] = f32(sd[pe + "patch_embed.proj.bias"])
        else:
            self.stitching_layer = SimpleNamespace(weight=sd["stitching_layer.weight"], bias=sd["stitching_layer.bias"])
            w["stitch.w"] = bf(sd["stitching_layer.weight"].reshape(C, -1))  # k = c*45 + kt*9 + ky*3 + kx
            w["stitch.b"] = f32(sd["stitching_layer.bias"])
        self._pos_embed = sd[pe + "pos_embed"].detach().float().cpu()
        self._cls = sd[pe + "cls_token"].detach().float().cpu().reshape(1, C)
        self._reg = sd[pe + "register_tokens"].detach().float().cpu().reshape(4, C)

        def block(dst, src, conv=bf):
            for n in ("norm1", "norm2"):
                w[dst + n + ".w"], w[dst + n + ".b"] = f32(sd[src + n + ".weight"]), f32(sd[src + n + ".bias"])
            for n, m in (("qkv", "attn.qkv"), ("proj", "attn.proj"), ("fc1", "mlp.fc1"), ("fc2", "mlp.fc2")):
                w[dst + n + ".w"], w[dst + n + ".b"] = conv(sd[src + m + ".weight"]), f32(sd[src + m + ".bias"])
            w[dst + "ls1"], w[dst + "ls2"] = f32(sd[src + "ls1.gamma"]), f32(sd[src + "ls2.gamma"])
            if src + "attn.q_norm.weight" in sd:
                for n in ("q_norm", "k_norm"):
                    w[dst + n + ".w"], w[dst + n + ".b"] = f32(sd[src + f"attn.{n}.weight"]), f32(sd[src + f"attn.{n}.bias"])

        for i in range(cfg.dino_blocks):
            block(f"dino{i}.", pe + f"blocks.{i}.")
        w["dino.norm.w"], w["dino.norm.b"] = f32(sd[pe + "norm.weight"]), f32(sd[pe + "norm.bias"])
        ag = E + "aggregator."
        cam = sd[ag + "camera_token"].detach().float().reshape(2, 1, C)
        reg = sd[ag + "register_token"].detach().float().reshape(2, 4, C)
        w["agg.special"] = f32(torch.cat([cam, reg], dim=1))  # [2 slots (first view / other views), 5, C]
        for i in range(cfg.agg_depth):
            block(f"frame{i}.", ag + f"frame_blocks.{i}.")
            block(f"global{i}.", ag + f"global_blocks.{i}.")
        # camera head: fp32 weights streamed by the skinny-linear kernel
        ch = E + "camera_head."
        for i in range(cfg.cam_trunk):
            block(f"cam{i}.", ch + f"trunk.{i}.", conv=f32)
        for n in ("token_norm", "trunk_norm"):
            w["cam." + n + ".w"], w["cam." + n + ".b"] = f32(sd[ch + n + ".weight"]), f32(sd[ch + n + ".bias"])
        w["cam.empty"] = f32(F.pad(sd[ch + "empty_pose_tokens"].detach().float().reshape(1, 9), (0, 7)))
        w["cam.embed.w"] = f32(F.pad(sd[ch + "embed_pose.weight"].detach().float(), (0, 7)))  # K 9 -> 16
        w["cam.embed.b"] = f32(sd[ch + "embed_pose.bias"])
        w["cam.mod.w"], w["cam.mod.b"] = f32(sd[ch + "poseLN_modulation.1.weight"]), f32(sd[ch + "poseLN_modulation.1.bias"])
        w["cam.fc1.w"], w["cam.fc1.b"] = f32(sd[ch + "pose_branch.fc1.weight"]), f32(sd[ch + "pose_branch.fc1.bias"])
        w["cam.fc2.w"], w["cam.fc2.b"] = f32(sd[ch + "pose_branch.fc2.weight"]), f32(sd[ch + "pose_branch.fc2.bias"])
        # DPT heads (fp32 weights, TF32 MMAs)
        for tag, head in (("dh.", E + "depth_head."), ("gh.", E + "gaussian_param_head.")):
            w[tag + "norm.w"], w[tag + "norm.b"] = f32(sd[head + "norm.weight"]), f32(sd[head + "norm.bias"])
            for k in range(4):
                w[tag + f"proj{k}.w"] = f32(sd[head + f"projects.{k}.weight"].flatten(1))
                w[tag + f"proj{k}.b"] = f32(sd[head + f"projects.{k}.bias"])
                w[tag + f"rn{k}.w"] = f32(_conv_w(sd[head + f"scratch.layer{k + 1}_rn.weight"]))
            for k, ks in ((0, 4), (1, 2)):  # ConvTranspose2d(k = s): W'[(dy*k+dx)*Co + co, ci] = w[ci, co, dy, dx]
                wt = sd[head + f"resize_layers.{k}.weight"].detach().float()
                w[tag + f"up{k}.w"] = f32(wt.permute(2, 3, 1, 0).reshape(ks * ks * wt.shape[1], wt.shape[0]))
                w[tag + f"up{k}.b"] = f32(sd[head + f"resize_layers.{k}.bias"].detach().float().repeat(ks * ks))
            w[tag + "down3.w"] = f32(_conv_w(sd[head + "resize_layers.3.weight"]))
            w[tag + "down3.b"] = f32(sd[head + "resize_layers.3.bias"])
            for r in (1, 2, 3, 4):
                p = head + f"scratch.refinenet{r}."
                w[tag + f"rf{r}.out.w"], w[tag + f"rf{r}.out.b"] = f32(sd[p + "out_conv.weight"].flatten(1)), f32(sd[p + "out_conv.bias"])
                for u in ((1, 2) if r != 4 else (2,)):
                    for c in (1, 2):
                        w[tag + f"rf{r}.u{u}c{c}.w"] = f32(_conv_w(sd[p + f"resConfUnit{u}.conv{c}.weight"]))
                        w[tag + f"rf{r}.u{u}c{c}.b"] = f32(sd[p + f"resConfUnit{u}.conv{c}.bias"])
            w[tag + "oc1.w"], w[tag + "oc1.b"] = f32(_conv_w(sd[head + "scratch.output_conv1.weight"])), f32(sd[head + "scratch.output_conv1.bias"])
            w[tag + "oc2a.w"], w[tag + "oc2a.b"] = f32(_conv_w(sd[head + "scratch.output_conv2.0.weight"])), f32(sd[head + "scratch.output_conv2.0.bias"])
        w["dh.oc2b.w"] = f32(sd[E + "depth_head.scratch.output_conv2.2.weight"][0].flatten())  # depth channel
        self._depth_b = float(sd[E + "depth_head.scratch.output_conv2.2.bias"][0])
        w["dh.conf.w"] = f32(sd[E + "depth_head.scratch.output_conv2.2.weight"][1].flatten())  # confidence channel (render_conf / opacity_conf only)
        self._conf_b = float(sd[E + "depth_head.scratch.output_conv2.2.bias"][1])
        g = E + "gaussian_param_head."
        w["gh.oc2b.w"] = f32(sd[g + "scratch.output_conv2.2.weight"].flatten(1))
        w["gh.oc2b.b"] = f32(sd[g + "scratch.output_conv2.2.bias"])
        # 7x7 RGB input_merger as 7 k-blocks of an overlapping-window implicit GEMM: k = dy*32 + dx*4 + c (dx < 7, c < 3; rest zero)
        mw = sd[g + "input_merger.0.weight"].detach().float()            # [128, 3, 7, 7]
        mk = torch.zeros((mw.shape[0], 7, 8, 4), dtype=torch.float32)
        mk[:, :, :7, :3] = mw.permute(0, 2, 3, 1).cpu()
        w["gh.merger.w"] = f32(mk.reshape(mw.shape[0], 224))
        w["gh.merger.b"] = f32(sd[g + "input_merger.0.bias"])
        m = torch.ones(cfg.d_sh)
        for deg in range(1, cfg.sh_degree + 1):  # AS/model/encoder/common/gaussian_adapter.py:34-40
            m[deg ** 2:(deg + 1) ** 2] = 0.1 * 0.25 ** deg
        w["sh_mask"] = f32(m)
        # 2-D RoPE tables (AS/.../layers/rope.py:100-115): base 100, 16 frequencies per axis
        inv = 1.0 / (100.0 ** (torch.arange(0, 32, 2).float() / 32))
        ang = torch.arange(64).float()[:, None] * inv[None]
        w["rope.cos"], w["rope.sin"] = f32(ang.cos()), f32(ang.sin())
        self.w = w
        self._tables = {}
        # attributes the reference's drivers read off the model: `.stitched_3d_model.decoder` (the renderer handed to
        # save_interpolated_video, inference_t23d.py:154), `.stitching_layer.{weight,bias}` and the patch-embed tokens
        # (nvs_eval.py:51-61; model_stitching_training.py:33-72 saves exactly these)
        from .renderer import DecoderSplattingB200

        pe_ns = SimpleNamespace(cls_token=sd[pe + "cls_token"], register_tokens=sd[pe + "register_tokens"], mask_token=sd.get(pe + "mask_token"))
        self.stitched_3d_model = SimpleNamespace(decoder=DecoderSplattingB200((1.0, 1.0, 1.0)),
                                                 encoder=SimpleNamespace(aggregator=SimpleNamespace(patch_embed=pe_ns), cfg=cfg))

    def weight_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.w.values())

    # ------------------------------------------------------------------ shape-dependent constants
    def _shape_tables(self, gh: int, gw: int, H: int, W: int):
        key = (gh, gw, H, W)
        t = self._tables.get(key)
        if t is not None:
            return t
        cfg, dev = self.cfg, self.device
        C = cfg.embed_dim
        pe = self._pos_embed
        N = pe.shape[1] - 1
        M = int(math.sqrt(N))
        if gh * gw == N and gh == gw:
            patch_pe = pe[0, 1:]
        else:  # vision_transformer.py:184-216 (`size=` branch, bicubic + antialias): a per-shape constant
            patch_pe = F.interpolate(pe[:, 1:].reshape(1, M, M, C).permute(0, 3, 1, 2), size=(gw, gh), mode="bicubic",
                                     antialias=True).permute(0, 2, 3, 1).reshape(-1, C)
        t = SimpleNamespace()
        t.pos = patch_pe.contiguous().to(dev)                                         # [gh*gw, C]
        t.special = torch.cat([self._cls + pe[0, :1], self._reg], 0).contiguous().to(dev)  # [5, C]
        t.dpt = {}
        for k, oc in enumerate(cfg.dpt_out_channels):  # [gh*gw, oc] tables added by the projection GEMM epilogues of both heads
            px, py = _dpt_pos_tables(oc, gh, gw, W, H)
            t.dpt[k] = torch.cat([px[None].expand(gh, gw, -1), py[:, None].expand(gh, gw, -1)], -1).reshape(gh * gw, oc).contiguous().to(dev)
        px, py = _dpt_pos_tables(cfg.dpt_features // 2, H, W, W, H)
        t.full_px, t.full_py = px.to(dev), py.to(dev)
        self._tables[key] = t
        return t

    # ------------------------------------------------------------------ transformer
    def _block(self, p: str, x: torch.Tensor, *, seqs: int, seq_len: int, eps: float, rope: Optional[dict], ws) -> None:
        """AS/.../layers/block.py:81-107 on the fp32 stream x [rows, C] (in place)."""
        w, cfg = self.w, self.cfg
        C, Hn = cfg.embed_dim, cfg.num_heads
        ops.layernorm(x, mul=w[p + "norm1.w"], add=w[p + "norm1.b"], eps=eps, out=ws.h)
        ops.gemm(ws.h, w[p + "qkv.w"], w[p + "qkv.b"], out=ws.qkv)
        if rope is not None:
            ops.qknorm_rope2d_(ws.qkv, Hn, w[p + "q_norm.w"], w[p + "q_norm.b"], w[p + "k_norm.w"], w[p + "k_norm.b"], w["rope.cos"],
                               w["rope.sin"], tokens_per_view=rope["tpv"], n_special=N_SPECIAL, grid_w=rope["gw"], eps=1e-5)
        q5 = ws.qkv.view(seqs, seq_len, 3, Hn, C // Hn)
        ops.fmha(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], out=ws.att.view(seqs, seq_len, Hn, C // Hn))
        ops.gemm(ws.att, w[p + "proj.w"], w[p + "proj.b"], gate=w[p + "ls1"], residual=x, out=x)
        ops.layernorm(x, mul=w[p + "norm2.w"], add=w[p + "norm2.b"], eps=eps, out=ws.h)
        ops.gemm(ws.h, w[p + "fc1.w"], w[p + "fc1.b"], act="gelu_erf", out=ws.mlp)
        ops.gemm(ws.mlp, w[p + "fc2.w"], w[p + "fc2.b"], gate=w[p + "ls2"], residual=x, out=x)

    def _aggregate(self, latent: torch.Tensor, H: int, W: int) -> List[torch.Tensor]:
        cfg, w, dev = self.cfg, self.w, self.device
        C = cfg.embed_dim
        if cfg.patch_embed:   # `latent` is the image batch [B, V, 3, H, W] in [0, 1]
            B, V = latent.shape[:2]
            gh, gw = H // cfg.patch, W // cfg.patch
        else:
            B, _, T, lh, lw = latent.shape
            V = (T - 1) * 4 + 1
            gh, gw = lh // 2, lw // 2
            if (gh, gw) != (H // cfg.patch, W // cfg.patch):
                raise ValueError(f"latent grid {lh}x{lw} -> {gh}x{gw} tokens does not match the {H}x{W} image ({H // cfg.patch}x{W // cfg.patch} patches)")
        npatch = gh * gw
        P = npatch + N_SPECIAL
        BV = B * V
        rows = BV * P
        tb = self._shape_tables(gh, gw, H, W)
        ws = SimpleNamespace(h=torch.empty((rows, C), dtype=torch.bfloat16, device=dev),
                             qkv=torch.empty((rows, 3 * C), dtype=torch.bfloat16, device=dev),
                             att=torch.empty((rows, C), dtype=torch.bfloat16, device=dev),
                             mlp=torch.empty((rows, 4 * C), dtype=torch.bfloat16, device=dev))
        pmap = (npatch, P, N_SPECIAL)  # patch-token rows of the [BV*P] stream
        # --- stitching conv: upsample + replicate pad + im2col, then one GEMM (+bias +pos-embed) into the patch rows
        x = torch.empty((rows, C), dtype=torch.float32, device=dev)
        if cfg.patch_embed:
            # DINOv2 patch embedding: normalise + im2col (a permutation: the 14x14 patches do not overlap), then one GEMM (+bias +pos-embed)
            a = ops.patch_embed_im2col(latent.reshape(B * V, 3, H, W), cfg.patch, self._pe_k)
            ops.gemm(a, w["pe.w"], w["pe.b"], out=x, residual=tb.pos, rmap=(npatch, 0, 0), cmap=pmap)
        else:
            a = ops.im2col_stitch(latent)
            ops.gemm(a, w["stitch.w"], w["stitch.b"], out=x, residual=tb.pos, rmap=(npatch, 0, 0), cmap=pmap)
        del a
        x.view(BV, P, C)[:, :N_SPECIAL].copy_(tb.special)
        # --- DINOv2 blocks (eps 1e-6, no QK norm / RoPE), sequences = views
        for i in range(cfg.dino_blocks):
            self._block(f"dino{i}.", x, seqs=BV, seq_len=P, eps=1e-6, rope=None, ws=ws)
        # --- final DINO norm on the patch rows -> aggregator stream; camera/register tokens in front
        t = torch.empty((rows, C), dtype=torch.float32, device=dev)
        ops.layernorm(x, mul=w["dino.norm.w"], add=w["dino.norm.b"], eps=1e-6, out=t, in_map=pmap, out_map=pmap, rows=BV * npatch)
        del x
        t5 = t.view(B, V, P, C)
        t5[:, 0, :N_SPECIAL].copy_(w["agg.special"][0])
        if V > 1:
            t5[:, 1:, :N_SPECIAL].copy_(w["agg.special"][1])
        rope = {"tpv": P, "gw": gw}
        inters = []
        for i in range(cfg.agg_depth):
            keep = i in cfg.inter_layers
            self._block(f"frame{i}.", t, seqs=BV, seq_len=P, eps=1e-5, rope=rope, ws=ws)
            if keep:
                it = torch.empty((rows, 2 * C), dtype=torch.float32, device=dev)
                it[:, :C].copy_(t)
            self._block(f"global{i}.", t, seqs=B, seq_len=V * P, eps=1e-5, rope=rope, ws=ws)
            if keep:
                it[:, C:].copy_(t)
                inters.append(it)
        return inters

    # ------------------------------------------------------------------ camera head
    def _camera_head(self, inter_last: torch.Tensor, B: int, V: int, P: int, iters: int = 4) -> List[torch.Tensor]:
        """AS/.../heads/camera_head.py:87-170 -> list of activated pose encodings [B, V, 9]"""
        w, cfg = self.w, self.cfg
        C2 = 2 * cfg.embed_dim
        Hn = cfg.cam_heads
        if V > 32:
            raise NotImplementedError("camera head kernels handle up to 32 views per scene")
        TP = 16 if V <= 16 else 32  # padded token rows of the weight-streaming GEMMs' B operand
        per_batch = []
        dev = self.device

        def buf(n):  # padded token matrix (rows >= V stay zero): B operand of the weight-streaming GEMMs
            return torch.zeros((TP, n), dtype=torch.float32, device=dev)

        for b in range(B):
            cam_rows = inter_last.view(B, V, P, C2)[b, :, 0]  # [V, C2] view, row stride P*C2
            tok = ops.layernorm(cam_rows, mul=w["cam.token_norm.w"], add=w["cam.token_norm.b"], eps=1e-5, out_dtype=torch.float32)
            pred = None
            outs = []
            emb, h, x, att16 = buf(C2), buf(C2), buf(C2), buf(C2)
            for _ in range(iters):
                inp = w["cam.empty"].expand(V, -1).contiguous() if pred is None else pred
                # poseLN_modulation = Sequential(SiLU, Linear): the embedding is only consumed through the SiLU
                ops.skinny_linear(inp, w["cam.embed.w"], w["cam.embed.b"], act="silu", out=emb[:V])
                mod = ops.linear_tokens16(emb, V, w["cam.mod.w"], w["cam.mod.b"])  # shift | scale | gate
                mlo = ops.layernorm(tok, mul=mod[:V, C2:2 * C2], add=mod[:V, :C2], mul_bstride=3 * C2, add_bstride=3 * C2, rows_per_batch=1,
                                    eps=1e-6, mul_plus_one=True, out_dtype=torch.float32)
                ops.fma_rows(mod[:V, 2 * C2:], mlo, tok, out=x[:V])
                for i in range(cfg.cam_trunk):
                    p = f"cam{i}."
                    ops.layernorm(x[:V], mul=w[p + "norm1.w"], add=w[p + "norm1.b"], eps=1e-5, out=h[:V])
                    qkv = ops.linear_tokens16(h, V, w[p + "qkv.w"], w[p + "qkv.b"])
                    ops.attention_small(qkv[:V], 1, V, Hn, C2 // Hn, out=att16[:V])
                    ops.linear_tokens16(att16, V, w[p + "proj.w"], w[p + "proj.b"], gate=w[p + "ls1"], residual=x, out=x)
                    ops.layernorm(x[:V], mul=w[p + "norm2.w"], add=w[p + "norm2.b"], eps=1e-5, out=h[:V])
                    m = ops.linear_tokens16(h, V, w[p + "fc1.w"], w[p + "fc1.b"], act="gelu_erf")
                    ops.linear_tokens16(m, V, w[p + "fc2.w"], w[p + "fc2.b"], gate=w[p + "ls2"], residual=x, out=x)
                ops.layernorm(x[:V], mul=w["cam.trunk_norm.w"], add=w["cam.trunk_norm.b"], eps=1e-5, out=h[:V])
                m = ops.linear_tokens16(h, V, w["cam.fc1.w"], w["cam.fc1.b"], act="gelu_erf")
                new = torch.zeros((V, 16), dtype=torch.float32, device=dev)
                ops.skinny_linear(m[:V], w["cam.fc2.w"], w["cam.fc2.b"], out=new, residual=pred)  # pred + delta (cols 9..15 stay 0)
                pred = new
                outs.append(pred)
            per_batch.append(outs)
        return [torch.stack([per_batch[b][i][:, :9] for b in range(B)], 0) for i in range(iters)]  # raw (pre-activation) encodings

    # ------------------------------------------------------------------ DPT trunk
    def _conv3(self, x, wt, bias=None, **kw):
        n, h, wd, _ = x.shape
        out = torch.empty((n, h, wd, wt.shape[0]), dtype=torch.float32, device=x.device)
        ops.gemm(x, wt, bias, conv=dict(kh=3, kw=3, pad=1), out=out.view(-1, wt.shape[0]), **kw)
        return out

    def _fusion(self, tag: str, r: int, x: Optional[torch.Tensor], skip_relu: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
        """FeatureFusionBlock.forward (dpt_head.py:442-474).  `skip_relu` already holds relu(skip): ResidualConvUnit's
        in-place ReLU (dpt_head.py:362-404) rectifies its input tensor, so the skip connection adds relu(x)."""
        w = self.w
        p = tag + f"rf{r}."
        if x is not None:
            c1 = self._conv3(skip_relu, w[p + "u1c1.w"], w[p + "u1c1.b"], act="relu")
            C = x.shape[-1]
            # s = relu(x + conv2(c1) + b + relu(skip))  (the trailing relu is unit 2's in-place activation)
            s = self._conv3(c1, w[p + "u1c2.w"], w[p + "u1c2.b"], residual=skip_relu.view(-1, C), residual2=x.view(-1, C), post_act="relu")
        else:
            s = skip_relu
        c1 = self._conv3(s, w[p + "u2c1.w"], w[p + "u2c1.b"], act="relu")
        C = s.shape[-1]
        o = self._conv3(c1, w[p + "u2c2.w"], w[p + "u2c2.b"], residual=s.view(-1, C))
        o = ops.bilinear_nhwc(o, size[0], size[1])
        out = torch.empty_like(o)
        ops.gemm(o.view(-1, C), w[p + "out.w"], w[p + "out.b"], out=out.view(-1, C))
        return out

    def _dpt_trunk(self, tag: str, inters: List[torch.Tensor], BV: int, gh: int, gw: int, tb) -> torch.Tensor:
        """DPTHead._forward_impl up to scratch.output_conv1 (dpt_head.py:194-252,279-309) -> NHWC [BV, 8gh, 8gw, F/2]"""
        cfg, w, dev = self.cfg, self.w, self.device
        C2 = 2 * cfg.embed_dim
        npatch = gh * gw
        P = npatch + N_SPECIAL
        pmap = (npatch, P, N_SPECIAL)
        feats = []
        for k, it in enumerate(inters):
            oc = cfg.dpt_out_channels[k]
            xn = ops.layernorm(it, mul=w[tag + "norm.w"], add=w[tag + "norm.b"], eps=1e-5, in_map=pmap, rows=BV * npatch, out_dtype=torch.float32)
            x = ops.gemm(xn, w[tag + f"proj{k}.w"], w[tag + f"proj{k}.b"], residual=tb.dpt[k], rmap=(npatch, 0, 0))  # + pos embed
            del xn
            if k in (0, 1):
                ks = 4 if k == 0 else 2
                y = ops.gemm(x, w[tag + f"up{k}.w"], w[tag + f"up{k}.b"])
                x = ops.depth_to_space(y, BV, gh, gw, oc, ks)
                del y
            elif k == 2:
                x = x.view(BV, gh, gw, oc)
            else:
                a = ops.im2col_nhwc(x.view(BV, gh, gw, oc), 3, 3, 2, 1)
                ho, wo = (gh - 1) // 2 + 1, (gw - 1) // 2 + 1
                x = ops.gemm(a, w[tag + "down3.w"], w[tag + "down3.b"]).view(BV, ho, wo, oc)
                del a
            feats.append(self._conv3(x, w[tag + f"rn{k}.w"], None, act="relu"))  # relu: the consuming ResidualConvUnit's in-place activation
        l1, l2, l3, l4 = feats
        out = self._fusion(tag, 4, None, l4, l3.shape[1:3])
        out = self._fusion(tag, 3, out, l3, l2.shape[1:3])
        out = self._fusion(tag, 2, out, l2, l1.shape[1:3])
        out = self._fusion(tag, 1, out, l1, (2 * l1.shape[1], 2 * l1.shape[2]))
        return self._conv3(out, w[tag + "oc1.w"], w[tag + "oc1.b"])

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward_with_latent(self, latent: torch.Tensor, feedforward_image: torch.Tensor, train: bool = False) -> EncoderOutput:
        """latent [B, 16, T, res/8, res/8] (de-normalised VAE latent), feedforward_image [B, 3, V, H, W] in [-1, 1]."""
        if train:
            raise NotImplementedError("the B200 decoder is an inference engine (train=False)")
        if not self.w:
            raise RuntimeError("weights not loaded: use StitchVAE3DB200.from_state_dict(...)")
        if self.cfg.patch_embed:
            raise RuntimeError("this engine was built with DecoderConfig.patch_embed (un-stitched encoder): call forward_images(image)")
        cfg, dev = self.cfg, self.device
        B, ci, V, H, W = feedforward_image.shape
        if latent.dtype not in (torch.float32, torch.bfloat16):
            latent = latent.float()
        latent = latent.to(dev).contiguous()
        if latent.shape[0] != B or (latent.shape[2] - 1) * 4 + 1 != V:
            raise ValueError(f"latent {tuple(latent.shape)} does not match {V} views x batch {B}")
        lat_hw = cfg.resolution // 8
        if latent.shape[-2:] != (lat_hw, lat_hw):
            # upsampling_layer also resizes H, W to resolution/8 (stitched_model.py:92-107: one trilinear, align_corners=True interpolation
            # to [4(T-1)+1, res/8, res/8]).  Trilinear interpolation is separable: the spatial part runs here (bilinear kernel on the
            # [B*T, h, w, 16] view of the latent), the temporal part stays fused in the stitching conv's operand gather.
            Bl, Cl, Tl, hl, wl = latent.shape
            x = latent.float().permute(0, 2, 3, 4, 1).reshape(Bl * Tl, hl, wl, Cl).contiguous()
            if Cl % 8:
                raise NotImplementedError(f"spatial resampling of a {Cl}-channel latent (the bilinear kernel moves 8-channel groups)")
            x = ops.bilinear_nhwc(x, lat_hw, lat_hw)
            latent = x.view(Bl, Tl, lat_hw, lat_hw, Cl).permute(0, 4, 1, 2, 3).contiguous()
        return self._decode(latent, feedforward_image, B, V, H, W)

    @torch.no_grad()
    def forward_with_latent_graph(self, latent: torch.Tensor, feedforward_image: torch.Tensor, clone: bool = True) -> EncoderOutput:
        """`forward_with_latent` replayed from a CUDA graph (`DecoderGraph`, one per input shape, captured on first use); the eager call
        for the configurations a graph cannot hold."""
        cfg = self.cfg
        if cfg.voxelize or cfg.render_conf or cfg.opacity_conf or cfg.patch_embed or not latent.is_cuda or torch.cuda.is_current_stream_capturing():
            return self.forward_with_latent(latent, feedforward_image)
        graphs = self.__dict__.setdefault("_graphs", {})
        key = (tuple(latent.shape), latent.dtype, tuple(feedforward_image.shape), feedforward_image.dtype)
        g = graphs.get(key)
        if g is None:
            while len(graphs) >= 3:   # a graph pins the memory of a whole forward (8.8 GB at 13 views): keep the most recent shapes
                graphs.pop(next(iter(graphs)))
            g = graphs[key] = DecoderGraph(self, latent, feedforward_image)
        return g(latent, feedforward_image, clone=clone)

    @torch.no_grad()
    def forward_images(self, image: torch.Tensor) -> EncoderOutput:
        """Un-stitched AnySplat encoder, image -> 3D Gaussians (drop-in for `EncoderAnySplat.forward(image)`,
        AS/model/encoder/anysplat.py:337-610): image [B, V, 3, H, W] in [0, 1], H and W multiples of 14.  Needs an engine built
        with DecoderConfig(patch_embed=True, dino_blocks=<all DINO blocks>) from a state dict that holds the patch-embedding conv."""
        if not self.cfg.patch_embed:
            raise RuntimeError("this engine was built for the stitched path: call forward_with_latent(latent, feedforward_image)")
        if not self.w:
            raise RuntimeError("weights not loaded: use StitchVAE3DB200.from_state_dict(...)")
        B, V, ci, H, W = image.shape
        if ci != 3 or H % self.cfg.patch or W % self.cfg.patch:
            raise ValueError(f"image {tuple(image.shape)}: expected [B, V, 3, H, W] with H, W multiples of {self.cfg.patch}")
        if image.dtype not in (torch.float32, torch.bfloat16):
            image = image.float()
        image = image.to(self.device).contiguous()
        return self._decode(image, image, B, V, H, W)

    def _decode(self, latent: torch.Tensor, feedforward_image: torch.Tensor, B: int, V: int, H: int, W: int) -> EncoderOutput:
        cfg, w, dev = self.cfg, self.w, self.device
        gh, gw = H // cfg.patch, W // cfg.patch
        P = gh * gw + N_SPECIAL
        BV = B * V
        tb = self._shape_tables(gh, gw, H, W)
        inters = self._aggregate(latent, H, W)
        # --- cameras
        pose_raw = self._camera_head(inters[-1], B, V, P)
        cams = ops.pose_to_cameras(pose_raw[-1].reshape(BV, 9), H, W)
        pose_list = [ops.pose_to_cameras(p.reshape(BV, 9), H, W, cameras=False)["pose_act"].view(B, V, 9) for p in pose_raw[:-1]]
        pose_list.append(cams["pose_act"].view(B, V, 9))
        # --- depth head
        d = self._dpt_trunk("dh.", inters, BV, gh, gw, tb)
        d = ops.bilinear_nhwc(d, H, W, pos_x=tb.full_px, pos_y=tb.full_py)
        dfeat = self._conv3(d, w["dh.oc2a.w"], w["dh.oc2a.b"], act="relu")
        del d
        # --- Gaussian head
        if feedforward_image.dtype not in (torch.float32, torch.bfloat16):
            feedforward_image = feedforward_image.float()
        if cfg.patch_embed:
            rgb = ops.rgb01_views_to_nhwc4pad(feedforward_image)          # image already in [0, 1], [B, V, 3, H, W]
        else:
            rgb = ops.rgb_to_nhwc4pad(feedforward_image.to(dev))          # [BV, H, W+8, 4] in [0, 1], rows zero-padded
        merged = torch.empty((BV * H * W, w["gh.merger.w"].shape[0]), dtype=torch.float32, device=dev)
        ops.gemm(rgb, w["gh.merger.w"], w["gh.merger.b"], act="relu", out=merged,
                 conv=dict(kh=7, kw=1, pad=3, pad_x=0, geom=(BV, H, W, 32), strides=(4, (W + 8) * 4, H * (W + 8) * 4)))
        del rgb
        g = self._dpt_trunk("gh.", inters, BV, gh, gw, tb)
        del inters
        g = ops.bilinear_nhwc(g, H, W, add=merged, pos_x=tb.full_px, pos_y=tb.full_py)
        del merged
        gfeat = self._conv3(g, w["gh.oc2a.w"], w["gh.oc2a.b"], act="relu")
        del g
        raw = ops.gemm(gfeat.view(BV * H * W, -1), w["gh.oc2b.w"], w["gh.oc2b.b"])
        del gfeat
        # --- fused depth activation + unprojection + Gaussian adapter
        o = ops.gaussian_epilogue(dfeat.view(BV * H * W, -1), w["dh.oc2b.w"], self._depth_b, raw, cams["extr"], cams["intr"], w["sh_mask"], BV, H, W)
        N = V * H * W
        conf_mask, valid_counts = None, None
        if (cfg.render_conf or cfg.opacity_conf) and cfg.voxelize:
            raise NotImplementedError("render_conf / opacity_conf together with voxelize: the reference indexes the damping factor with the pixel mask "
                                      "against voxel-sized tensors (anysplat_stitched.py:463-467), which only works with voxelize off")
        if cfg.render_conf or cfg.opacity_conf:
            # confidence-quantile branches (anysplat_stitched.py:381-387, 442-455, 463-467): the quantile is taken over ALL pixels of the call;
            # the kept pixels of every batch element are compacted in (view, row, column) order and padded to the largest count
            if cfg.opacity_conf and B != 1:
                raise ValueError("opacity_conf: the reference's broadcast (anysplat_stitched.py:465-467) only works for one batch element")
            Cr = cfg.raw_gs_dim
            conf = ops.depth_conf(dfeat.view(BV * H * W, -1), w["dh.conf.w"], self._conf_b)
            qv = ops.quantile(conf, cfg.conf_threshold)
            rawb, ptsb, confb = raw.view(B, N, -1), o["means"].view(B, N, 3), conf.view(B, N)
            kept = [ops.compact_rows(confb[b], qv, rawb[b], ptsb[b], feat_dim=Cr, use_threshold=cfg.render_conf, want_damp=cfg.opacity_conf) for b in range(B)]
            valid_counts = [k["count"] for k in kept]
            conf_mask = (conf.view(B, V, H, W) > qv) if cfg.render_conf else None
            N = max(valid_counts)
            if B == 1:
                vp, vf = kept[0]["pts"], kept[0]["feats"]
            else:
                vp = torch.full((B, N, 3), -1e4, dtype=torch.float32, device=dev)
                vf = torch.full((B, N, Cr), -1e10, dtype=torch.float32, device=dev)
                for b, k in enumerate(kept):
                    vp[b, :k["count"]] = k["pts"]
                    vf[b, :k["count"]] = k["feats"]
            depth_out = o["depth"]
            o = ops.gaussian_adapter(vp.reshape(B * N, 3), vf.reshape(B * N, Cr), w["sh_mask"]) | dict(depth=depth_out, scene_sum=o["scene_sum"])
            if cfg.opacity_conf:   # opacity * sigmoid(depth_conf - quantile) of the kept pixels
                op = o["opacities"].view(1, -1)
                ops.fma_rows(op, kept[0]["damp"].view(1, -1), torch.zeros_like(op), out=op)
            del kept, vp, vf
        elif cfg.voxelize:
            # voxelised fusion per batch element, padded to the largest voxel count (anysplat_stitched.py:419-455), then the adapter
            Cr = cfg.raw_gs_dim
            rawb, ptsb = raw.view(B, N, -1), o["means"].view(B, N, 3)
            fused = [ops.voxel_fusion(ptsb[b], rawb[b], rawb[b][:, Cr], cfg.voxel_size, feat_dim=Cr) for b in range(B)]
            if self.keep_voxel_inputs:  # tests: the fusion stage is checked against the oracle on exactly these inputs
                self.voxel_inputs = dict(pts=ptsb.clone(), raw=rawb.clone(), counts=[f["n_voxels"] for f in fused])
            N = max(f["n_voxels"] for f in fused)
            if B == 1:
                vp, vf = fused[0]["pts"], fused[0]["feats"]
            else:
                vp = torch.full((B, N, 3), -1e4, dtype=torch.float32, device=dev)
                vf = torch.full((B, N, Cr), -1e10, dtype=torch.float32, device=dev)
                for b, f in enumerate(fused):
                    vp[b, :f["n_voxels"]] = f["pts"]
                    vf[b, :f["n_voxels"]] = f["feats"]
            depth_out = o["depth"]
            o = ops.gaussian_adapter(vp.reshape(B * N, 3), vf.reshape(B * N, Cr), w["sh_mask"]) | dict(depth=depth_out, scene_sum=o["scene_sum"])
            del fused, vp, vf
        gauss = Gaussians(means=o["means"].view(B, N, 3), covariances=o["covariances"].view(B, N, 3, 3),
                          harmonics=o["harmonics"].view(B, N, 3, cfg.d_sh), opacities=o["opacities"].view(B, N),
                          scales=o["scales"].view(B, N, 3), rotations=o["rotations"].view(B, N, 4), packed=o.get("packed"))
        scene_scale = (o["scene_sum"] / float(BV * H * W)).clamp_min(1e-8).reshape(())
        return EncoderOutput(
            gaussians=gauss,
            pred_pose_enc_list=pose_list,
            pred_context_pose=dict(extrinsic=cams["c2w"].view(B, V, 4, 4), intrinsic=cams["intr_norm"].view(B, V, 3, 3)),
            depth_dict=dict(depth=o["depth"].view(B, V, H, W, 1),
                            conf_valid_mask=conf_mask if conf_mask is not None else torch.ones((B, V, H, W), dtype=torch.bool, device=dev)),
            infos=dict(scene_scale=scene_scale, voxelize_ratio=float(N) / (H * W * V), **({"valid_counts": valid_counts} if valid_counts is not None else {})),
            distill_infos=None,
            last_pred_pose_enc=pose_list[-1],
        )

    def forward(self, images: torch.Tensor, feedforward_image: torch.Tensor, train: bool = False) -> EncoderOutput:
        """StitchVAE3D.forward (models/stitched_model.py:140-163): `diffusion_vae.encode(images).latent_dist.sample()` (images [B, 3, T, H, W]
        in [-1, 1]) and then the stitched path on that latent.  The VAE is the caller's module (`self.diffusion_vae = pipe.vae`; the Wan VAE
        is not part of this engine yet); without one the call is refused."""
        vae = getattr(self, "diffusion_vae", None)
        if vae is None:
            raise NotImplementedError("StitchVAE3D.forward needs the Wan VAE encoder, which is outside this engine: set `.diffusion_vae` to the "
                                      "caller's VAE module, or call forward_with_latent(latent, feedforward_image)")
        with torch.no_grad():
            latent = vae.encode(images).latent_dist.sample()
        return self.forward_with_latent(latent, feedforward_image, train=train)
