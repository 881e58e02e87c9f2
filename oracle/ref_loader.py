"""ORACLE support (test infrastructure; only usable where /root/reference is mounted, i.e. in the
build container -- never on the GPU box and never from vist3a_b200): imports the REAL reference
decoder (`models/stitched_model.py:StitchVAE3D` -> `models/anysplat_stitched.py:AnySplatStitched`)
on CPU so that oracle/decoder_ref.py can be validated against it and golden vectors generated.

What is stubbed: third-party roots that the reference imports at module load but never
executes on the `forward_with_latent(train=False, voxelize=False)` path, `VGGT.from_pretrained`
(network) and the `AutoencoderKLWan` isinstance check.  The reference's forward code runs unmodified.
The voxelize=True branch calls torch_scatter (absent here; rusty1s/pytorch_scatter, unpinned in requirements.txt):
`scatter_add` / `scatter_max` are provided with that package's documented semantics (out[index[i]] += src[i] along dim 0;
max with -inf-free reduction over present indices) on top of torch.index_add_ / index_reduce_.

`width="tiny"` shrinks constructor hyper-parameters only (embed dim, head counts, DPT feature
widths) through the reference classes' own keyword arguments, so that a full weight set is a few MB
and golden vectors can be committed; depth (24 alternating blocks, stitch after enc block 2) and
every code path are those of the full model.
"""
from __future__ import annotations

import functools
import os
import sys
import types
from dataclasses import dataclass

import torch

REFERENCE_ROOT = "/root/reference"

_STUBS = [
    "diffusers", "diffusers.configuration_utils", "diffusers.loaders", "diffusers.loaders.single_file_model",
    "diffusers.models", "diffusers.models.activations", "diffusers.models.autoencoders", "diffusers.models.autoencoders.vae",
    "diffusers.models.modeling_outputs", "diffusers.models.modeling_utils", "diffusers.pipelines", "diffusers.pipelines.wan",
    "diffusers.pipelines.wan.pipeline_wan", "diffusers.utils", "diffusers.utils.accelerate_utils",
    "dacite", "lightning", "lightning.pytorch", "lightning.pytorch.utilities", "skvideo", "skvideo.io", "matplotlib", "matplotlib.figure",
    "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "omegaconf", "torch_scatter", "xformers", "xformers.ops", "e3nn", "e3nn.o3", "gsplat",
    "colorspacious", "plyfile", "moviepy", "moviepy.editor", "lpips", "wandb", "imageio", "cv2", "open3d", "trimesh", "roma", "kornia", "hydra",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def _install_torch_scatter():
    """functional stand-in for the two torch_scatter calls of voxelizaton_with_fusion (AS/model/encoder/anysplat.py:314-333)"""
    if "torch_scatter" in sys.modules and hasattr(sys.modules["torch_scatter"], "_vist3a_oracle"):
        return
    try:
        import torch_scatter  # noqa: F401  (the real package, if it is ever installed)
        return
    except Exception:
        pass
    m = types.ModuleType("torch_scatter")

    def scatter_add(src, index, dim=0, out=None, dim_size=None):
        assert dim == 0 and out is None
        n = int(index.max()) + 1 if dim_size is None else dim_size
        return torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device).index_add_(0, index, src)

    def scatter_max(src, index, dim=0, out=None, dim_size=None):
        assert dim == 0 and out is None and src.dim() == 1
        n = int(index.max()) + 1 if dim_size is None else dim_size
        mx = torch.full((n,), -float("inf"), dtype=src.dtype, device=src.device).index_reduce_(0, index, src, "amax", include_self=True)
        return mx, None  # the reference discards argmax (:314)

    m.scatter_add, m.scatter_max, m._vist3a_oracle = scatter_add, scatter_max, True
    sys.modules["torch_scatter"] = m


def _make_stub(name: str):
    if name in sys.modules:
        return
    try:
        __import__(name)
        return
    except Exception:
        pass
    m = types.ModuleType(name)
    m.__path__ = []

    def ga(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)

        def _decorator_or_obj(*a, **k):
            if len(a) == 1 and callable(a[0]) and not k:
                return a[0]
            return None

        cls = type(attr, (), {"__init__": lambda self, *a, **k: None, "__class_getitem__": classmethod(lambda c, x: c),
                              "__call__": lambda self, *a, **k: None})
        return cls

    m.__getattr__ = ga
    sys.modules[name] = m


@dataclass(frozen=True)
class RefWidth:
    """constructor hyper-parameters of the reference model (defaults = the released model)"""
    embed_dim: int = 1024
    num_heads: int = 16            # aggregator + DINO heads (head_dim 64)
    dino_depth: int = 24
    cam_heads: int = 16
    dpt_features: int = 256
    dpt_out_channels: tuple = (256, 512, 1024, 1024)
    pos_grid: int = 37             # DINO pos-embed grid (img_size 518 / 14)


FULL = RefWidth()
# dpt_features stays 256: the GS head hard-codes 128 merger channels = features // 2 (vggt_dpt_gs_head.py:69-76)
TINY = RefWidth(embed_dim=64, num_heads=1, dino_depth=4, cam_heads=2, dpt_features=256, dpt_out_channels=(32, 32, 64, 64), pos_grid=37)


def load_teacher(width: RefWidth = FULL, seed: int = 0, sh_degree: int = 4, voxelize: bool = False, voxel_size: float = 0.002):
    """Returns the reference's UN-STITCHED AnySplat encoder (`EncoderAnySplat`, AS/model/encoder/anysplat.py: DINOv2 patch embedding +
    all DINO blocks + aggregator + heads; image [B, V, 3, H, W] in [0, 1] -> EncoderOutput), fp32, eval, seeded random init."""
    return load_reference(width, seed=seed, sh_degree=sh_degree, voxelize=voxelize, voxel_size=voxel_size, _teacher=True)


def load_reference(width: RefWidth = FULL, resolution: int = 512, seed: int = 0, sh_degree: int = 4, voxelize: bool = False,
                   voxel_size: float = 0.002, _teacher: bool = False, render_conf: bool = False, opacity_conf: bool = False,
                   conf_threshold: float = 0.1):
    """Returns the reference StitchVAE3D (fp32, eval, checkpointing off) with seeded random init."""
    if not available():
        raise RuntimeError(f"{REFERENCE_ROOT} is not mounted here; the real reference can only be imported in the build container")
    _install_torch_scatter()
    for n in _STUBS:
        _make_stub(n)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    P = "third_party_model.anysplat.src.model"
    import importlib

    vggt_mod = importlib.import_module(P + ".encoder.vggt.models.vggt")
    agg_mod = importlib.import_module(P + ".encoder.vggt.models.aggregator")
    vit_mod = importlib.import_module(P + ".encoder.vggt.layers.vision_transformer")
    cam_mod = importlib.import_module(P + ".encoder.vggt.heads.camera_head")
    dpt_mod = importlib.import_module(P + ".encoder.vggt.heads.dpt_head")
    enc_mod = importlib.import_module(P + ".encoder.anysplat")
    gsh_mod = importlib.import_module(P + ".encoder.heads.vggt_dpt_gs_head")
    import models.stitched_model as sm
    from models.stitching_layer_builder import parse_conv_spec

    W = width
    VGGT = vggt_mod.VGGT

    def vit_small_cfg(patch_size=16, num_register_tokens=0, **kw):
        return vit_mod.DinoVisionTransformer(patch_size=patch_size, embed_dim=W.embed_dim, depth=W.dino_depth, num_heads=W.num_heads,
                                             mlp_ratio=4, block_fn=functools.partial(vit_mod.Block, attn_class=vit_mod.MemEffAttention),
                                             num_register_tokens=num_register_tokens, **kw)

    saved = {}

    def patch(obj, name, val):
        saved[(obj, name)] = getattr(obj, name)
        setattr(obj, name, val)

    try:
        if W != FULL:
            patch(agg_mod, "vit_large", vit_small_cfg)
            patch(vggt_mod, "Aggregator", functools.partial(agg_mod.Aggregator, num_heads=W.num_heads))
            patch(vggt_mod, "CameraHead", functools.partial(cam_mod.CameraHead, num_heads=W.cam_heads))
            patch(vggt_mod, "DPTHead", functools.partial(dpt_mod.DPTHead, features=W.dpt_features, out_channels=list(W.dpt_out_channels)))

            class _GS(gsh_mod.VGGT_DPT_GS_Head):
                def __init__(self, dim_in, patch_size, output_dim, activation, conf_activation, features):
                    super().__init__(dim_in=2 * W.embed_dim, patch_size=patch_size, output_dim=output_dim, activation=activation,
                                     conf_activation=conf_activation, features=W.dpt_features, out_channels=list(W.dpt_out_channels))

            patch(enc_mod, "VGGT_DPT_GS_Head", _GS)
        emb = W.embed_dim
        patch(VGGT, "from_pretrained", classmethod(lambda cls, *a, **k: cls(embed_dim=emb)))

        class FakeVAE(torch.nn.Module):
            pass

        patch(sm, "AutoencoderKLWan", FakeVAE)
        patch(sm, "AutoencoderKLWan_wan", FakeVAE)

        anysplat_mod = importlib.import_module(P + ".model.anysplat")
        ga_mod = importlib.import_module(P + ".encoder.common.gaussian_adapter")
        dec_mod = importlib.import_module(P + ".decoder.decoder_splatting_cuda")
        cfg = enc_mod.EncoderAnySplatCfg(
            name="anysplat", anchor_feat_dim=83, voxel_size=voxel_size, n_offsets=2, d_feature=32, add_view=False,
            num_monocular_samples=32, backbone=None, visualizer=None,
            gaussian_adapter=ga_mod.GaussianAdapterCfg(0.5, 15.0, sh_degree), apply_bounds_shim=True,
            opacity_mapping=enc_mod.OpacityMappingCfg(0.0, 0.0, 1), gaussians_per_pixel=1, num_surfaces=1,
            gs_params_head_type="dpt_gs", pred_head_type="depth", voxelize=voxelize, intermediate_layer_idx=[4, 11, 17, 23],
            render_conf=render_conf, opacity_conf=opacity_conf, conf_threshold=conf_threshold)
        torch.manual_seed(seed)
        ff = anysplat_mod.AnySplat(cfg, dec_mod.DecoderSplattingCUDACfg("splatting_cuda", [1.0, 1.0, 1.0], False))
        if _teacher:
            enc = ff.encoder.float().eval()
            enc.aggregator.use_checkpoint = False
            return enc
        model = sm.StitchVAE3D(FakeVAE(), ff, torch.device("cpu"), "enc_blocks_2",
                               parse_conv_spec(f"conv3d_k5x3x3_o{W.embed_dim}_s1x2x2_p2x1x1"), resolution)
    finally:
        for (obj, name), val in saved.items():
            setattr(obj, name, val)
    model = model.float().eval()
    model.stitched_3d_model.grad_checkpointing = False
    model.stitched_3d_model.encoder.aggregator.use_checkpoint = False
    return model


def decoder_state_dict(model) -> dict:
    """the tensors the decoder path reads, with the reference's own state-dict key names (diffusion_vae excluded)"""
    sd = {k: v.detach().clone() for k, v in model.state_dict().items() if not k.startswith("diffusion_vae")}
    return sd


def outputs_to_dict(out) -> dict:
    """flatten the reference's EncoderOutput into plain tensors"""
    g = out.gaussians
    d = {"means": g.means, "covariances": g.covariances, "harmonics": g.harmonics, "opacities": g.opacities, "scales": g.scales,
         "rotations": g.rotations, "extrinsic": out.pred_context_pose["extrinsic"], "intrinsic": out.pred_context_pose["intrinsic"],
         "depth": out.depth_dict["depth"],
         "last_pred_pose_enc": out.last_pred_pose_enc if getattr(out, "last_pred_pose_enc", None) is not None else out.pred_pose_enc_list[-1],
         "scene_scale": out.infos["scene_scale"].reshape(1)}
    for i, p in enumerate(out.pred_pose_enc_list):
        d[f"pred_pose_enc_{i}"] = p
    return {k: v.detach().float().contiguous() for k, v in d.items()}


class LiveWanVAE:
    """The reference's vendored Wan-2.1 VAE (`utils/wan_utils.py:96-1180`) on CPU: its own `WanEncoder3d` / `WanDecoder3d` / `WanCausalConv3d`
    modules and its own chunked, cache-carrying `_encode` / `_decode` loops (:1021-1048, :1078-1117), run unmodified.  Only the diffusers
    base classes of `AutoencoderKLWan` (ModelMixin / ConfigMixin: config + checkpoint I/O, no arithmetic) are absent, so the two loops are
    called as plain functions on this holder; `get_activation("silu")` is served as `nn.SiLU()` (what diffusers returns)."""

    def __init__(self, base_dim=96, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temperal_downsample=(False, True, True), seed=0):
        if not available():
            raise RuntimeError(f"{REFERENCE_ROOT} is not mounted here; the real reference can only be imported in the build container")
        for n in _STUBS:
            _make_stub(n)
        act = sys.modules["diffusers.models.activations"]
        if "_vist3a_oracle" not in act.__dict__:   # (the stub modules answer every attribute, so no hasattr)
            def get_activation(name):
                assert name == "silu", name
                return torch.nn.SiLU()
            act.get_activation, act._vist3a_oracle = get_activation, True
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        import importlib

        wu = importlib.import_module("utils.wan_utils")
        if wu.get_activation is not act.__dict__["get_activation"]:   # module imported earlier with the inert stub
            wu.get_activation = act.__dict__["get_activation"]
        self._wu = wu
        self.z_dim = z_dim
        self.temperal_downsample = list(temperal_downsample)
        self.temperal_upsample = self.temperal_downsample[::-1]
        torch.manual_seed(seed)
        self.encoder = wu.WanEncoder3d(base_dim, z_dim * 2, list(dim_mult), num_res_blocks, [], self.temperal_downsample, 0.0).float().eval()
        self.quant_conv = wu.WanCausalConv3d(z_dim * 2, z_dim * 2, 1).float().eval()
        self.post_quant_conv = wu.WanCausalConv3d(z_dim, z_dim, 1).float().eval()
        self.decoder = wu.WanDecoder3d(base_dim, z_dim, list(dim_mult), num_res_blocks, [], self.temperal_upsample, 0.0).float().eval()
        self._parts = {"encoder": self.encoder, "quant_conv": self.quant_conv, "post_quant_conv": self.post_quant_conv, "decoder": self.decoder}

    # the two helpers `_encode` / `_decode` call on self
    def _count_conv3d(self, model):
        return self._wu.AutoencoderKLWan._count_conv3d(self, model)

    def _create_cache_state(self, model):
        return self._wu.AutoencoderKLWan._create_cache_state(self, model)

    @torch.no_grad()
    def encode_moments(self, x: torch.Tensor) -> torch.Tensor:
        """[B, 3, T, H, W] in [-1, 1] -> [B, 2 z_dim, 1 + (T - 1) / 4, H / 8, W / 8] (mean | logvar), the tensor `encode` wraps in
        DiagonalGaussianDistribution (:1068-1072)"""
        return self._wu.AutoencoderKLWan._encode(self, x)

    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """[B, z_dim, T', h, w] -> [B, 3, 1 + 4 (T' - 1), 8 h, 8 w], clamped to [-1, 1] (:1078-1117)"""
        return self._wu.AutoencoderKLWan._decode(self, z, return_dict=False)[0]

    def state_dict(self) -> dict:
        return {f"{p}.{k}": v.detach().clone() for p, m in self._parts.items() for k, v in m.state_dict().items()}

    def load_state_dict(self, sd: dict):
        for p, m in self._parts.items():
            m.load_state_dict({k[len(p) + 1:]: v for k, v in sd.items() if k.startswith(p + ".")}, strict=True)
