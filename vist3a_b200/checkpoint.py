"""On-disk formats either side of the hot path (host code): the reference's trainable-parameter checkpoint of the
stitched decoder and the PEFT adapter of the DiT, both folded into plain state dicts that `StitchVAE3DB200.from_state_dict` /
`WanTransformer3DModelB200.from_state_dict` ingest.

  * `anysplat_stitched.pth` (written by train_stitching.py, read by evaluation/novel_view_synthesis_bench/nvs_eval.py:21-63):
        {"lora": {<module>.lora_A [r, in], <module>.lora_B [out, r], ...},      # loralib-style, keys relative to stitched_3d_model
         "stitching_layer": {"weight" [1024,16,5,3,3], "bias" [1024]}, "mask_token", "cls_token", "register_tokens"}
    LoRA semantics (utils/lora_util/layers.py:106-182, 289-...): W_eff = W + (B @ A).view(W.shape) * alpha / r; in eval mode
    the reference merges exactly this into the weight, so folding at load is bit-equivalent to its inference path.
  * PEFT adapter directory of the DiT (train_vdm.py:370-388; inference_t23d.py:75-78): adapter_config.json (r, lora_alpha,
    target_modules) + adapter_model.safetensors with keys base_model.model.<module>.lora_{A,B}.weight.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional, Tuple

import torch


def fold_loralib(sd: Dict[str, torch.Tensor], lora: Dict[str, torch.Tensor], *, alpha: float, r: Optional[int] = None,
                 prefix: str = "") -> Dict[str, torch.Tensor]:
    """Returns a copy of `sd` with every `<m>.lora_A` / `<m>.lora_B` pair of `lora` merged into `prefix + <m>.weight`."""
    out = dict(sd)
    for ka, A in lora.items():
        if not ka.endswith(".lora_A"):
            continue
        mod = ka[: -len(".lora_A")]
        kb = mod + ".lora_B"
        kw = prefix + mod + ".weight"
        if kb not in lora:
            raise KeyError(f"LoRA factor {kb} missing")
        if kw not in out:
            raise KeyError(f"LoRA target {kw} is not a parameter of the model")
        W = out[kw].float()
        rank = r if r is not None else (A.shape[0] if W.dim() == 2 else A.shape[0] // W.shape[-1])
        delta = (lora[kb].float() @ A.float()).reshape(W.shape) * (float(alpha) / rank)
        out[kw] = (W + delta).to(out[kw].dtype)
    return out


_PE = "stitched_3d_model.encoder.aggregator.patch_embed."


def renumber_stitched_blocks(sd: Dict[str, torch.Tensor], stitched_layer_index: int) -> Dict[str, torch.Tensor]:
    """State dict of the UN-stitched model (AnySplat checkpoint: 24 DINO blocks `patch_embed.blocks.0..23` + the 14x14
    `patch_embed.patch_embed.proj`) -> the numbering of the stitched model.  `convert_model_to_stitched_model`
    (models/anysplat_stitched.py:158-165) deletes `patch_embed.blocks[0]` `stitched_layer_index` times and the patch-embedding conv;
    nn.ModuleList renumbers what is left, so stitched `blocks.i` is original `blocks.(i + stitched_layer_index)`.  The `lora` keys of
    anysplat_stitched.pth and `StitchVAE3D.state_dict()` use the stitched numbering."""
    k = int(stitched_layer_index)
    if k < 0:
        raise ValueError("stitched_layer_index must be >= 0")
    out = {}
    for name, v in sd.items():
        if name.startswith(_PE + "patch_embed."):
            continue                                   # the 14x14 patch-embedding conv does not exist in the stitched model
        if name.startswith(_PE + "blocks."):
            rest = name[len(_PE + "blocks."):]
            idx, tail = rest.split(".", 1)
            if int(idx) < k:
                continue
            name = f"{_PE}blocks.{int(idx) - k}.{tail}"
        out[name] = v
    return out


def apply_stitched_checkpoint(sd: Dict[str, torch.Tensor], ckpt: Dict, *, lora_alpha: float = 32.0,
                              lora_r: Optional[int] = None, stitched_layer_index: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Base state dict (reference key names: `stitching_layer.*`, `stitched_3d_model.encoder.*`) + the dict stored in
    anysplat_stitched.pth -> the state dict the inference model uses (nvs_eval.py:45-62).

    The base dict must carry the STITCHED block numbering (what `StitchVAE3D.state_dict()` gives).  A plain AnySplat checkpoint numbers
    its DINO blocks 0..23 and still holds the patch-embedding conv: pass `stitched_layer_index` (k of "enc_blocks_k",
    --stitching_layer_location) and it is renumbered first (`renumber_stitched_blocks`); without it such a dict is refused, because the
    LoRA deltas would fold into the wrong blocks."""
    pe = _PE
    if stitched_layer_index is not None:
        sd = renumber_stitched_blocks(sd, stitched_layer_index)
    elif pe + "patch_embed.proj.weight" in sd:
        raise ValueError("apply_stitched_checkpoint: the base state dict holds patch_embed.patch_embed.proj, i.e. it is an UN-stitched "
                         "AnySplat checkpoint (DINO blocks numbered 0..23); pass stitched_layer_index=k (enc_blocks_k) to renumber it")
    out = fold_loralib(sd, ckpt.get("lora", {}), alpha=lora_alpha, r=lora_r, prefix="stitched_3d_model.")
    # non-LoRA entries of the "lora" dict (bias = "lora_only" / "all" checkpoints carry biases too)
    for k, v in ckpt.get("lora", {}).items():
        if "lora_" not in k:
            out["stitched_3d_model." + k] = v
    out["stitching_layer.weight"] = ckpt["stitching_layer"]["weight"]
    out["stitching_layer.bias"] = ckpt["stitching_layer"]["bias"]
    for name in ("mask_token", "cls_token", "register_tokens"):
        if name in ckpt:
            out[pe + name] = ckpt[name]
    return out


def load_stitched_checkpoint(sd: Dict[str, torch.Tensor], path: str, **kw) -> Dict[str, torch.Tensor]:
    ckpt = torch.load(path, map_location="cpu", weights_only=True)   # dict of tensors / dicts of tensors: no pickled code is needed
    if "state_dict" in ckpt and "stitching_layer" not in ckpt:
        ckpt = ckpt["state_dict"]
    return apply_stitched_checkpoint(sd, ckpt, **kw)


def load_peft_adapter(path: str) -> Tuple[Dict[str, torch.Tensor], float, int]:
    """PEFT adapter directory -> (lora tensors, lora_alpha, r) for `WanTransformer3DModelB200.from_state_dict(lora=...)`."""
    from safetensors.torch import load_file

    with open(os.path.join(path, "adapter_config.json")) as f:
        cfg = json.load(f)
    st = os.path.join(path, "adapter_model.safetensors")
    tensors = load_file(st) if os.path.exists(st) else torch.load(os.path.join(path, "adapter_model.bin"), map_location="cpu", weights_only=True)
    return tensors, float(cfg.get("lora_alpha", 16)), int(cfg.get("r", 8))
