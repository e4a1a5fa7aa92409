"""Adapter checkpoint I/O in the on-disk layout the reference reads and writes.

The reference stores its trained weights through diffusers' ``ModelMixin``:

* ``I2VAdapterModule.from_pretrained(path)`` / ``.save_pretrained(dir)``  (``src/pipelines/pipeline_i2v_adapter.py:740``,
  ``src/models/unet_motion_cross_frame_attn.py:1080-1097``): a directory holding ``config.json`` (the constructor
  arguments plus ``_class_name``) and ``diffusion_pytorch_model.safetensors`` (or ``.bin``), keys
  ``{down,up}_blocks.N.attentions.M.transformer_blocks.K.i2v_adapter.{to_q,to_k,to_v,to_out.0}.*`` and
  ``mid_block.attentions...``;
* ``MotionAdapter.from_pretrained(path)`` (``src/pipelines/pipeline_i2v_adapter.py:733,745``): same container, keys
  ``{down,up}_blocks.N.motion_modules.M.*`` and ``mid_block.motion_modules.M.*``;
* the IP-Adapter file ``ip-adapter_sd15.bin`` / ``.safetensors`` consumed by ``_load_ip_adapter_weights``
  (``src/models/unet_motion_cross_frame_attn.py:1230-1287``): ``{"image_proj": {...}, "ip_adapter": {"1.to_k_ip.weight": ...}}``
  (flat ``image_proj.*`` / ``ip_adapter.*`` keys in the safetensors flavour).

Host-side plumbing only: tensors are read on the CPU and copied into the modules' parameters in place, so the packed
weight caches of the fast path (keyed on ``data_ptr`` / ``_version``) see the change.
"""
from __future__ import annotations

import json
import os
from typing import Any, Dict, Optional

import torch

WEIGHTS_NAME = "diffusion_pytorch_model.bin"
SAFETENSORS_WEIGHTS_NAME = "diffusion_pytorch_model.safetensors"
CONFIG_NAME = "config.json"


def _variant_name(name: str, variant: Optional[str]) -> str:
    if not variant:
        return name
    stem, ext = name.rsplit(".", 1)
    return f"{stem}.{variant}.{ext}"


def save_model_directory(module: torch.nn.Module, config: Dict[str, Any], class_name: str, save_directory: str,
                         safe_serialization: bool = True, variant: Optional[str] = None) -> str:
    """``ModelMixin.save_pretrained``: ``config.json`` + one weights file.  Returns the weights path."""
    if os.path.isfile(save_directory):
        raise ValueError(f"Provided path ({save_directory}) should be a directory, not a file")
    os.makedirs(save_directory, exist_ok=True)
    cfg = {"_class_name": class_name, "_diffusers_version": "0.25.0"}
    cfg.update({k: (list(v) if isinstance(v, tuple) else v) for k, v in config.items()})
    with open(os.path.join(save_directory, CONFIG_NAME), "w", encoding="utf-8") as f:
        f.write(json.dumps(cfg, indent=2, sort_keys=True) + "\n")
    state = {k: v.detach().cpu().contiguous() for k, v in module.state_dict().items()}
    if safe_serialization:
        from safetensors.torch import save_file

        path = os.path.join(save_directory, _variant_name(SAFETENSORS_WEIGHTS_NAME, variant))
        save_file(state, path, metadata={"format": "pt"})
    else:
        path = os.path.join(save_directory, _variant_name(WEIGHTS_NAME, variant))
        torch.save(state, path)
    return path


def load_state_dict_file(path: str) -> Dict[str, torch.Tensor]:
    """One weights file: ``.safetensors`` through the safetensors reader, anything else as a torch pickle."""
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file

        return load_file(path, device="cpu")
    return torch.load(path, map_location="cpu", weights_only=True)


def load_model_directory(pretrained_path: str, variant: Optional[str] = None):
    """``ModelMixin.from_pretrained`` for a local directory: returns ``(config, state_dict)``.  The safetensors file is
    preferred when both flavours are present, as diffusers does."""
    if not os.path.isdir(pretrained_path):
        raise EnvironmentError(f"{pretrained_path} is not a directory (this sandbox has no hub access)")
    cfg_path = os.path.join(pretrained_path, CONFIG_NAME)
    if not os.path.isfile(cfg_path):
        raise EnvironmentError(f"Error no file named {CONFIG_NAME} found in directory {pretrained_path}.")
    with open(cfg_path, encoding="utf-8") as f:
        config = json.load(f)
    for name in (_variant_name(SAFETENSORS_WEIGHTS_NAME, variant), _variant_name(WEIGHTS_NAME, variant)):
        path = os.path.join(pretrained_path, name)
        if os.path.isfile(path):
            return config, load_state_dict_file(path)
    raise EnvironmentError(
        f"Error no file named {_variant_name(SAFETENSORS_WEIGHTS_NAME, variant)} or "
        f"{_variant_name(WEIGHTS_NAME, variant)} found in directory {pretrained_path}.")


def constructor_kwargs(config: Dict[str, Any], names) -> Dict[str, Any]:
    """The subset of a ``config.json`` that a mirror class's constructor takes (``_class_name`` etc. dropped)."""
    return {k: (tuple(v) if isinstance(v, list) else v) for k, v in config.items() if k in names}


def load_ip_adapter_file(path: str) -> Dict[str, Dict[str, torch.Tensor]]:
    """``ip-adapter_sd15.bin`` / ``.safetensors`` -> ``{"image_proj": {...}, "ip_adapter": {...}}`` as
    ``_load_ip_adapter_weights`` expects (diffusers' ``load_ip_adapter`` does the same regrouping for safetensors)."""
    flat = load_state_dict_file(path)
    if "image_proj" in flat and "ip_adapter" in flat:
        return {"image_proj": dict(flat["image_proj"]), "ip_adapter": dict(flat["ip_adapter"])}
    out: Dict[str, Dict[str, torch.Tensor]] = {"image_proj": {}, "ip_adapter": {}}
    for key, value in flat.items():
        if key.startswith("image_proj."):
            out["image_proj"][key[len("image_proj."):]] = value
        elif key.startswith("ip_adapter."):
            out["ip_adapter"][key[len("ip_adapter."):]] = value
    if not out["image_proj"] or not out["ip_adapter"]:
        raise ValueError(f"{path} holds neither an 'image_proj' nor an 'ip_adapter' group")
    return out


def save_ip_adapter_file(state: Dict[str, Dict[str, torch.Tensor]], path: str) -> None:
    """Inverse of ``load_ip_adapter_file`` (test fixtures; the reference only reads this format)."""
    if path.endswith(".safetensors"):
        from safetensors.torch import save_file

        flat = {f"{group}.{k}": v.detach().cpu().contiguous() for group in ("image_proj", "ip_adapter")
                for k, v in state[group].items()}
        save_file(flat, path, metadata={"format": "pt"})
    else:
        torch.save({g: {k: v.detach().cpu() for k, v in state[g].items()} for g in ("image_proj", "ip_adapter")}, path)
