"""Shared fixtures for the test-suite: tiny model configs with the SD1.5 structure, seeded inputs."""
from __future__ import annotations

import torch

from i2v_adapter_unofficial_b200.hostmodel import (
    IPAdapterAttnProcessor2_0,
    UNetMotionCrossFrameAttnModel,
)

# SD1.5 structure (3 cross-frame down blocks + 1 motion-only, mirrored up path, mid block) at toy widths.
TINY_CFG = dict(block_out_channels=(32, 64, 64, 64), cross_attention_dim=48, num_attention_heads=4,
                motion_num_attention_heads=4, norm_num_groups=8, layers_per_block=1)
# head dims 40 / 80 / 160 / 160 like SD1.5 (8 heads), one layer per block to keep it small
SD15_HEADDIM_CFG = dict(block_out_channels=(320, 640, 1280, 1280), cross_attention_dim=768, num_attention_heads=8,
                        motion_num_attention_heads=8, norm_num_groups=32, layers_per_block=1)


def make_unet(cfg=None, seed=0, ip_adapter=False, dtype=torch.float32, device="cpu"):
    torch.manual_seed(seed)
    cfg = dict(TINY_CFG if cfg is None else cfg)
    unet = UNetMotionCrossFrameAttnModel(**cfg).eval()
    if ip_adapter:
        unet._load_ip_adapter_weights(fake_ip_adapter_state_dict(unet, seed + 1))
    return unet.to(device=device, dtype=dtype)


def fake_ip_adapter_state_dict(unet, seed=1, image_embed_dim=64):
    """Random weights in the layout of ``ip-adapter_sd15.bin`` (keys ``ip_adapter.{1,3,..}.to_{k,v}_ip.weight``,
    ``image_proj.proj/norm``) for the given UNet."""
    g = torch.Generator().manual_seed(seed)
    cross = unet.config.cross_attention_dim
    sd = {"image_proj": {
        "proj.weight": torch.randn(4 * cross, image_embed_dim, generator=g) * 0.05,
        "proj.bias": torch.randn(4 * cross, generator=g) * 0.05,
        "norm.weight": torch.ones(cross) + 0.1 * torch.randn(cross, generator=g),
        "norm.bias": 0.1 * torch.randn(cross, generator=g)}, "ip_adapter": {}}
    key_id = 1
    for name in unet.attn_processors.keys():
        if name.endswith("attn2.processor") and "motion_modules" not in name:
            if name.startswith("mid_block"):
                hidden = unet.config.block_out_channels[-1]
            elif name.startswith("up_blocks"):
                hidden = list(reversed(unet.config.block_out_channels))[int(name[len("up_blocks.")])]
            else:
                hidden = unet.config.block_out_channels[int(name[len("down_blocks.")])]
            sd["ip_adapter"][f"{key_id}.to_k_ip.weight"] = torch.randn(hidden, cross, generator=g) * cross**-0.5
            sd["ip_adapter"][f"{key_id}.to_v_ip.weight"] = torch.randn(hidden, cross, generator=g) * cross**-0.5
            key_id += 2
    return sd


def randomize_zero_init(unet, seed=7):
    """Give every parameter a non-degenerate value (LayerNorm/GroupNorm affine away from 1/0) so tests exercise
    all terms; deterministic."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in unet.named_parameters():
            if name.endswith("norm.weight") or ".norm1.weight" in name or ".norm2.weight" in name or \
                    ".norm3.weight" in name or name.endswith("conv_norm_out.weight"):
                p.add_(0.1 * torch.randn(p.shape, generator=g).to(p))
            elif name.endswith(".bias"):
                p.add_(0.05 * torch.randn(p.shape, generator=g).to(p))
    return unet


def unet_inputs(unet, videos=1, frames=4, size=16, tokens=7, seed=1, image_embed_dim=None):
    g = torch.Generator().manual_seed(seed)
    cross = unet.config.cross_attention_dim
    sample = torch.randn(videos, frames, 4, size, size, generator=g)
    ctx = torch.randn(videos, tokens, cross, generator=g)
    img = None if image_embed_dim is None else torch.randn(videos, image_embed_dim, generator=g)
    return sample, ctx, img


def unet_cfg_dict(unet):
    return dict(unet.config)
