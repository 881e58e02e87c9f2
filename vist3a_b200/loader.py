"""`load_stitching_model(args)` for the B200 engine -- the constructor the reference's drivers call
(/root/reference/evaluation/novel_view_synthesis_bench/nvs_eval.py:21-63, used at inference_t23d.py:83 and train_vdm.py:401-403),
with the same argument names and the same error behaviour, returning a `StitchVAE3DB200`.

What the reference does there and what happens here instead:
  * `load_feedforward_model(args, device)` downloads AnySplat from the Hugging Face hub (utils/utils_for_thirdparty.py:14-29).  There is
    no network behind this engine: the AnySplat weights come from `args.feedforward_weights` (a local .safetensors / .pt state dict of
    the AnySplat module, keys `encoder.*`) or from the `feedforward_state_dict=` argument.
  * `load_vae` + `StitchVAE3D(...)`: the VAE only supplies the latent geometry (16 channels, resolution // 8; stitched_model.py:50-63);
    `convert_model_to_stitched_model` drops the first k DINO blocks and the patch-embedding conv (anysplat_stitched.py:158-165) --
    `renumber_stitched_blocks` does the same on the state dict.
  * `add_lora` + `load_state_dict(state_dict["lora"])` + `.eval()` merge (utils/lora_util/layers.py:149-165): the LoRA factors of
    `args.checkpoint_path` are folded into the weights at load (`apply_stitched_checkpoint`), alpha / r from `args.lora_config`.
  * `cast_to_bfloat16` (utils_for_thirdparty.py:53-69): the engine's own dtype policy (bf16 transformer weights, fp32 heads) is that rule.
The two mini-grammars of the argument parser are mirrored so that the same command lines parse:
`parse_conv_spec` (models/stitching_layer_builder.py:48-89) and `parse_lora_mode` (utils/lora_util/utils.py:68-117).
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass
from typing import Dict, Optional, Tuple, Union

import torch

from .checkpoint import apply_stitched_checkpoint, renumber_stitched_blocks
from .stitched_decoder import DecoderConfig, StitchVAE3DB200

IntOrTuple = Union[int, Tuple[int, ...]]


@dataclass(frozen=True)
class ConvSpec:
    """models/stitching_layer_builder.py:12-19 (fields only: the engine never builds an nn.Module from it)."""
    dim: int
    out_channels: int
    kernel_size: IntOrTuple
    stride: IntOrTuple = 1
    padding: IntOrTuple = 0
    dilation: IntOrTuple = 1


_CONV_RE = re.compile(r"^conv(?P<dim>[123])d_k(?P<k>[0-9x]+)_o(?P<o>[0-9]+)(?:_s(?P<s>[0-9x]+))?(?:_p(?P<p>[0-9x]+))?(?:_d(?P<d>[0-9x]+))?$", re.IGNORECASE)


def _ints(txt: str) -> IntOrTuple:
    return tuple(int(n) for n in txt.split("x")) if "x" in txt else int(txt)


def parse_conv_spec(spec: str) -> ConvSpec:
    """'conv3d_k5x3x3_o1024_s1x2x2_p2x1x1' -> ConvSpec; ValueError when the string is outside the grammar (as the reference)."""
    m = _CONV_RE.fullmatch(spec)
    if not m:
        raise ValueError(f"Bad CONV_SPEC {spec!r}. Expected something like 'conv2d_k3_o64', 'conv3d_k3x3x3_o32_s2_p1', ...")
    g = m.groupdict()
    return ConvSpec(dim=int(g["dim"]), out_channels=int(g["o"]), kernel_size=_ints(g["k"]), stride=_ints(g["s"]) if g["s"] else 1,
                    padding=_ints(g["p"]) if g["p"] else 0, dilation=_ints(g["d"]) if g["d"] else 1)


@dataclass
class LoraConfig:
    """utils/lora_util/utils.py:52-65 defaults"""
    r: int = 4
    alpha: int = 1
    dropout: float = 0.0
    bias: str = "lora_only"
    target_modules: Optional[Tuple[str, ...]] = None
    fan_in_fan_out: bool = False
    finetune_encoder: bool = False
    freeze_head: bool = False


def parse_lora_mode(spec: str) -> LoraConfig:
    """'r8,a16,d0.05,f0' (+ b<none|all|lora_only>, t<a|b|c>, enc, fix_head) -> LoraConfig; ValueError on a chunk outside the grammar."""
    cfg = LoraConfig()
    for chunk in spec.split(","):
        c = chunk.strip().lower()
        if c == "enc":
            cfg.finetune_encoder = True
        elif c in ("fix_head", "fixhead"):
            cfg.freeze_head = True
        elif re.fullmatch(r"[radf][\d.]+", c):
            key, num = c[0], c[1:]
            if key == "r":
                cfg.r = int(num)
            elif key == "a":
                cfg.alpha = int(num)
            elif key == "d":
                cfg.dropout = float(num)
            else:
                cfg.fan_in_fan_out = bool(int(num))
        elif re.fullmatch(r"b[^,]+", c):
            if c[1:] not in ("none", "all", "lora_only"):
                raise ValueError("b chunk must be none|all|lora_only")
            cfg.bias = c[1:]
        elif re.fullmatch(r"t[^,]+", c):
            cfg.target_modules = tuple(c[1:].split("|"))
        else:
            raise ValueError(f"Bad LoRA chunk: {c!r}")
    return cfg


def _read_state_dict(path: str) -> Dict[str, torch.Tensor]:
    if os.path.isdir(path):
        for name in ("model.safetensors", "pytorch_model.bin"):
            if os.path.exists(os.path.join(path, name)):
                path = os.path.join(path, name)
                break
        else:
            raise FileNotFoundError(f"no model.safetensors / pytorch_model.bin under {path}")
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file

        return load_file(path)
    sd = torch.load(path, map_location="cpu", weights_only=True)
    return sd["state_dict"] if "state_dict" in sd and "stitching_layer" not in sd else sd


def stitched_layer_index(location: str) -> int:
    """'enc_blocks_k' -> k (models/anysplat_stitched.py:150-157)"""
    m = re.fullmatch(r"enc_blocks_(\d+)", location)
    if not m:
        raise NotImplementedError(f"stitching_layer_location {location!r}: the stitched AnySplat model takes 'enc_blocks_<k>'")
    return int(m.group(1))


def load_stitching_model(args, *, feedforward_state_dict: Optional[Dict[str, torch.Tensor]] = None, device=None, voxelize: bool = False,
                         render_conf: bool = False, opacity_conf: bool = False, conf_threshold: float = 0.1,
                         config_overrides: Optional[dict] = None) -> StitchVAE3DB200:
    """Same argument names as the reference: feedforward_model, video_model, stitching_layer_location, stitching_layer_config,
    resolution, initialization_weight_path, lora_config, checkpoint_path (argparse.Namespace or any object with these attributes).
    The keyword-only extras select the branches the AnySplat HF config selects in the reference (EncoderAnySplatCfg);
    `config_overrides` sets DecoderConfig fields of non-released widths (head counts, DPT channels: the tests' tiny model)."""
    if getattr(args, "feedforward_model", "anysplat") != "anysplat" and not os.path.exists(str(args.feedforward_model)):
        raise NotImplementedError(f"Feedforward model {args.feedforward_model} is not implemented.")
    if getattr(args, "video_model", "wan") != "wan":
        raise NotImplementedError(f"Video diffusion model {args.video_model} is not implemented.")
    dev = torch.device(device if device is not None else "cuda")
    spec = args.stitching_layer_config
    if isinstance(spec, str):
        spec = parse_conv_spec(spec)
    k = stitched_layer_index(args.stitching_layer_location)
    if (spec.dim, spec.kernel_size, spec.stride, spec.padding, spec.dilation) != (3, (5, 3, 3), (1, 2, 2), (2, 1, 1), 1):
        raise NotImplementedError(f"stitching layer {spec}: the engine implements conv3d_k5x3x3_s1x2x2_p2x1x1 (the released configuration)")
    # AnySplat weights
    if feedforward_state_dict is None:
        src = getattr(args, "feedforward_weights", None) or (args.feedforward_model if os.path.exists(str(args.feedforward_model)) else None)
        if src is None:
            raise FileNotFoundError("load_stitching_model: AnySplat weights are downloaded from the Hugging Face hub by the reference; this engine "
                                    "has no network path: pass args.feedforward_weights=<local state dict> or feedforward_state_dict=")
        feedforward_state_dict = _read_state_dict(str(src))
    sd = {("stitched_3d_model." + n if not n.startswith("stitched_3d_model.") else n): v for n, v in feedforward_state_dict.items()
          if n.startswith(("encoder.", "stitched_3d_model.encoder."))}
    pe = "stitched_3d_model.encoder.aggregator.patch_embed."
    n_dino = 1 + max((int(n[len(pe + "blocks."):].split(".", 1)[0]) for n in sd if n.startswith(pe + "blocks.")), default=-1)
    if pe + "patch_embed.proj.weight" in sd:        # un-stitched checkpoint: drop the first k blocks, renumber
        sd = renumber_stitched_blocks(sd, k)
        n_dino -= k
    # stitching layer: explicit init file first (stitched_model.py:109-121), then the trained checkpoint overrides it (nvs_eval.py:51-52)
    init = getattr(args, "initialization_weight_path", None)
    if init:
        st = _read_state_dict(init)
        sd["stitching_layer.weight"], sd["stitching_layer.bias"] = st["weight"], st["bias"]
    lora = parse_lora_mode(args.lora_config) if isinstance(getattr(args, "lora_config", None), str) else getattr(args, "lora_config", None) or LoraConfig()
    ckpt_path = getattr(args, "checkpoint_path", None)
    if ckpt_path:
        ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=True)
        sd = apply_stitched_checkpoint(sd, ckpt, lora_alpha=float(lora.alpha), lora_r=int(lora.r))
    if "stitching_layer.weight" not in sd:
        raise KeyError("load_stitching_model: no stitching layer weights (neither initialization_weight_path nor checkpoint_path provides them)")
    C = sd["stitching_layer.weight"].shape[0]
    if spec.out_channels != C:
        raise ValueError(f"stitching_layer_config says {spec.out_channels} output channels, the weights have {C}")
    cfg = DecoderConfig(embed_dim=C, dino_blocks=n_dino, resolution=int(args.resolution), latent_channels=sd["stitching_layer.weight"].shape[1],
                        voxelize=voxelize, render_conf=render_conf, opacity_conf=opacity_conf, conf_threshold=conf_threshold,
                        **(config_overrides or {}))
    model = StitchVAE3DB200.from_state_dict(sd, cfg, device=dev)
    model.lora_config = lora
    return model
