"""UNetMotion with cross-frame attention — host-side mirror of
``/root/reference/src/models/unet_motion_cross_frame_attn.py``.

Same class names, block layout, ``forward`` signatures, processor plumbing (``attn_processors`` /
``set_attn_processor``, reference :1118-1161) and IP-Adapter installation order (reference :1230-1287) as the
reference.  Convolutions, resnets, time embedding stay on PyTorch; the attention arithmetic is whatever processors are
installed (stock SDPA ones by default, the B200 ones after ``i2v_adapter_unofficial_b200.install``).
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple, Union

import torch
from torch import nn

from . import checkpoint
from .attention import AttnProcessor2_0, IPAdapterAttnProcessor2_0
from .i2v_adapter import I2VAdapterModule, I2VAdapterTransformer2DModel, _Sample
from .layers import Downsample2D, ImageProjection, ResnetBlock2D, TimestepEmbedding, Timesteps, Upsample2D
from .temporal import DownBlockMotion, MotionAdapter, UpBlockMotion, _motion_module


def _spatial_transformer(channels, heads, cross_dim, groups, use_linear_projection=False, only_cross_attention=False,
                         upcast_attention=False, layers=1):
    return I2VAdapterTransformer2DModel(
        heads, channels // heads, in_channels=channels, num_layers=layers, cross_attention_dim=cross_dim,
        norm_num_groups=groups, use_linear_projection=use_linear_projection,
        only_cross_attention=only_cross_attention, upcast_attention=upcast_attention)


def _attend(attn, hidden_states, enable_cross_frame_attn, num_frames, encoder_hidden_states, cross_attention_kwargs,
            attention_mask, encoder_attention_mask):
    return attn(hidden_states, enable_cross_frame_attn=enable_cross_frame_attn, num_frames=num_frames,
                encoder_hidden_states=encoder_hidden_states, cross_attention_kwargs=cross_attention_kwargs,
                attention_mask=attention_mask, encoder_attention_mask=encoder_attention_mask, return_dict=False)[0]


class CrossFrameAttnDownBlockMotion(nn.Module):
    """resnet -> I2V spatial transformer -> motion module per layer, then stride-2 conv (reference :164-340)."""

    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, dropout: float = 0.0,
                 num_layers: int = 1, transformer_layers_per_block: int = 1, resnet_eps: float = 1e-6,
                 resnet_groups: int = 32, num_attention_heads: int = 1, cross_attention_dim: int = 1280,
                 output_scale_factor: float = 1.0, downsample_padding: int = 1, add_downsample: bool = True,
                 use_linear_projection: bool = False, only_cross_attention: bool = False,
                 upcast_attention: bool = False, temporal_cross_attention_dim: Optional[int] = None,
                 temporal_num_attention_heads: int = 8, temporal_max_seq_length: int = 32, **_unused):
        super().__init__()
        self.has_cross_attention = True
        self.num_attention_heads = num_attention_heads
        resnets, attentions, motion_modules = [], [], []
        for i in range(num_layers):
            resnets.append(ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels,
                                         eps=resnet_eps, groups=resnet_groups, dropout=dropout,
                                         output_scale_factor=output_scale_factor))
            attentions.append(_spatial_transformer(out_channels, num_attention_heads, cross_attention_dim,
                                                   resnet_groups, use_linear_projection, only_cross_attention,
                                                   upcast_attention, transformer_layers_per_block))
            motion_modules.append(_motion_module(out_channels, temporal_num_attention_heads, resnet_groups,
                                                 temporal_max_seq_length, temporal_cross_attention_dim))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.motion_modules = nn.ModuleList(motion_modules)
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels,
                                                         padding=downsample_padding, name="op")])
                             if add_downsample else None)

    def forward(self, hidden_states, temb=None, enable_cross_frame_attn: bool = False, encoder_hidden_states=None,
                attention_mask=None, num_frames: int = 1, encoder_attention_mask=None,
                cross_attention_kwargs: Optional[Dict[str, Any]] = None, additional_residuals=None):
        output_states = ()
        n = len(self.resnets)
        for i, (resnet, attn, motion_module) in enumerate(zip(self.resnets, self.attentions, self.motion_modules)):
            hidden_states = resnet(hidden_states, temb)
            hidden_states = _attend(attn, hidden_states, enable_cross_frame_attn, num_frames, encoder_hidden_states,
                                    cross_attention_kwargs, attention_mask, encoder_attention_mask)
            hidden_states = motion_module(hidden_states, num_frames=num_frames)[0]
            if i == n - 1 and additional_residuals is not None:
                hidden_states = hidden_states + additional_residuals
            output_states = output_states + (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            output_states = output_states + (hidden_states,)
        return hidden_states, output_states


class CrossFrameAttnUpBlockMotion(nn.Module):
    """skip-concat -> resnet -> I2V spatial transformer -> motion module per layer, then upsample
    (reference :342-529)."""

    def __init__(self, in_channels: int, out_channels: int, prev_output_channel: int, temb_channels: int,
                 resolution_idx: Optional[int] = None, dropout: float = 0.0, num_layers: int = 1,
                 transformer_layers_per_block: int = 1, resnet_eps: float = 1e-6, resnet_groups: int = 32,
                 num_attention_heads: int = 1, cross_attention_dim: int = 1280, output_scale_factor: float = 1.0,
                 add_upsample: bool = True, use_linear_projection: bool = False, only_cross_attention: bool = False,
                 upcast_attention: bool = False, temporal_cross_attention_dim: Optional[int] = None,
                 temporal_num_attention_heads: int = 8, temporal_max_seq_length: int = 32, **_unused):
        super().__init__()
        self.has_cross_attention = True
        self.num_attention_heads = num_attention_heads
        resnets, attentions, motion_modules = [], [], []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            rin = prev_output_channel if i == 0 else out_channels
            resnets.append(ResnetBlock2D(rin + skip, out_channels, temb_channels, eps=resnet_eps,
                                         groups=resnet_groups, dropout=dropout,
                                         output_scale_factor=output_scale_factor))
            attentions.append(_spatial_transformer(out_channels, num_attention_heads, cross_attention_dim,
                                                   resnet_groups, use_linear_projection, only_cross_attention,
                                                   upcast_attention, transformer_layers_per_block))
            motion_modules.append(_motion_module(out_channels, temporal_num_attention_heads, resnet_groups,
                                                 temporal_max_seq_length, temporal_cross_attention_dim))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.motion_modules = nn.ModuleList(motion_modules)
        self.upsamplers = (nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)])
                           if add_upsample else None)
        self.resolution_idx = resolution_idx

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, enable_cross_frame_attn: bool = False,
                encoder_hidden_states=None, cross_attention_kwargs=None, upsample_size=None, attention_mask=None,
                encoder_attention_mask=None, num_frames: int = 1):
        for resnet, attn, motion_module in zip(self.resnets, self.attentions, self.motion_modules):
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res], dim=1)
            hidden_states = resnet(hidden_states, temb)
            hidden_states = _attend(attn, hidden_states, enable_cross_frame_attn, num_frames, encoder_hidden_states,
                                    cross_attention_kwargs, attention_mask, encoder_attention_mask)
            hidden_states = motion_module(hidden_states, num_frames=num_frames)[0]
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


class UNetMidBlockCrossFrameAttnMotion(nn.Module):
    """resnet -> (I2V spatial transformer -> motion module -> resnet)* (reference :531-694)."""

    def __init__(self, in_channels: int, temb_channels: int, dropout: float = 0.0, num_layers: int = 1,
                 transformer_layers_per_block: int = 1, resnet_eps: float = 1e-6, resnet_groups: int = 32,
                 num_attention_heads: int = 1, output_scale_factor: float = 1.0, cross_attention_dim: int = 1280,
                 use_linear_projection: bool = False, upcast_attention: bool = False,
                 temporal_num_attention_heads: int = 1, temporal_cross_attention_dim: Optional[int] = None,
                 temporal_max_seq_length: int = 32, **_unused):
        super().__init__()
        self.has_cross_attention = True
        self.num_attention_heads = num_attention_heads
        resnet_groups = resnet_groups if resnet_groups is not None else min(in_channels // 4, 32)

        def resnet():
            return ResnetBlock2D(in_channels, in_channels, temb_channels, eps=resnet_eps, groups=resnet_groups,
                                 dropout=dropout, output_scale_factor=output_scale_factor)

        resnets, attentions, motion_modules = [resnet()], [], []
        for _ in range(num_layers):
            attentions.append(_spatial_transformer(in_channels, num_attention_heads, cross_attention_dim,
                                                   resnet_groups, use_linear_projection, False, upcast_attention,
                                                   transformer_layers_per_block))
            resnets.append(resnet())
            motion_modules.append(_motion_module(in_channels, temporal_num_attention_heads, resnet_groups,
                                                 temporal_max_seq_length, temporal_cross_attention_dim))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.motion_modules = nn.ModuleList(motion_modules)

    def forward(self, hidden_states, temb=None, enable_cross_frame_attn: bool = False, encoder_hidden_states=None,
                attention_mask=None, cross_attention_kwargs=None, encoder_attention_mask=None, num_frames: int = 1):
        hidden_states = self.resnets[0](hidden_states, temb)
        for attn, resnet, motion_module in zip(self.attentions, self.resnets[1:], self.motion_modules):
            hidden_states = _attend(attn, hidden_states, enable_cross_frame_attn, num_frames, encoder_hidden_states,
                                    cross_attention_kwargs, attention_mask, encoder_attention_mask)
            hidden_states = motion_module(hidden_states, num_frames=num_frames)[0]
            hidden_states = resnet(hidden_states, temb)
        return hidden_states


class _Config(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class UNetMotionCrossFrameAttnModel(nn.Module):
    def __init__(
        self,
        sample_size: Optional[int] = None,
        in_channels: int = 4,
        out_channels: int = 4,
        down_block_types: Tuple[str, ...] = ("CrossFrameAttnDownBlockMotion", "CrossFrameAttnDownBlockMotion",
                                             "CrossFrameAttnDownBlockMotion", "DownBlockMotion"),
        up_block_types: Tuple[str, ...] = ("UpBlockMotion", "CrossFrameAttnUpBlockMotion",
                                           "CrossFrameAttnUpBlockMotion", "CrossFrameAttnUpBlockMotion"),
        block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280),
        layers_per_block: int = 2,
        downsample_padding: int = 1,
        mid_block_scale_factor: float = 1,
        act_fn: str = "silu",
        norm_num_groups: int = 32,
        norm_eps: float = 1e-5,
        cross_attention_dim: int = 1280,
        use_linear_projection: bool = False,
        num_attention_heads: Union[int, Tuple[int, ...]] = 8,
        motion_max_seq_length: int = 32,
        motion_num_attention_heads: int = 8,
        use_motion_mid_block: int = True,
        encoder_hid_dim: Optional[int] = None,
        encoder_hid_dim_type: Optional[str] = None,
    ):
        super().__init__()
        self.config = _Config(
            sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
            down_block_types=tuple(down_block_types), up_block_types=tuple(up_block_types),
            block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
            downsample_padding=downsample_padding, mid_block_scale_factor=mid_block_scale_factor, act_fn=act_fn,
            norm_num_groups=norm_num_groups, norm_eps=norm_eps, cross_attention_dim=cross_attention_dim,
            use_linear_projection=use_linear_projection, num_attention_heads=num_attention_heads,
            motion_max_seq_length=motion_max_seq_length, motion_num_attention_heads=motion_num_attention_heads,
            use_motion_mid_block=use_motion_mid_block, encoder_hid_dim=encoder_hid_dim,
            encoder_hid_dim_type=encoder_hid_dim_type)
        self.sample_size = sample_size
        self.layers_per_block = layers_per_block
        self.num_attention_heads = num_attention_heads

        if len(down_block_types) != len(up_block_types):
            raise ValueError(
                f"Must provide the same number of `down_block_types` as `up_block_types`. `down_block_types`: "
                f"{down_block_types}. `up_block_types`: {up_block_types}.")
        if len(block_out_channels) != len(down_block_types):
            raise ValueError(
                f"Must provide the same number of `block_out_channels` as `down_block_types`. `block_out_channels`: "
                f"{block_out_channels}. `down_block_types`: {down_block_types}.")
        if not isinstance(num_attention_heads, int) and len(num_attention_heads) != len(down_block_types):
            raise ValueError(
                f"Must provide the same number of `num_attention_heads` as `down_block_types`. "
                f"`num_attention_heads`: {num_attention_heads}. `down_block_types`: {down_block_types}.")

        self.conv_in = nn.Conv2d(in_channels, block_out_channels[0], kernel_size=3, padding=1)
        time_embed_dim = block_out_channels[0] * 4
        self.time_proj = Timesteps(block_out_channels[0], True, 0)
        self.time_embedding = TimestepEmbedding(block_out_channels[0], time_embed_dim, act_fn=act_fn)
        self.encoder_hid_proj = None

        # created before mid_block so that processor enumeration runs down -> up -> mid
        self.down_blocks = nn.ModuleList([])
        self.up_blocks = nn.ModuleList([])

        heads = (num_attention_heads,) * len(down_block_types) if isinstance(num_attention_heads, int) \
            else tuple(num_attention_heads)
        common = dict(temb_channels=time_embed_dim, resnet_eps=norm_eps, resnet_groups=norm_num_groups,
                      temporal_num_attention_heads=motion_num_attention_heads,
                      temporal_max_seq_length=motion_max_seq_length)

        output_channel = block_out_channels[0]
        for i, kind in enumerate(down_block_types):
            input_channel, output_channel = output_channel, block_out_channels[i]
            final = i == len(block_out_channels) - 1
            if kind == "DownBlockMotion":
                blk = DownBlockMotion(in_channels=input_channel, out_channels=output_channel,
                                      num_layers=layers_per_block, add_downsample=not final,
                                      downsample_padding=downsample_padding, **common)
            elif kind == "CrossFrameAttnDownBlockMotion":
                blk = CrossFrameAttnDownBlockMotion(
                    in_channels=input_channel, out_channels=output_channel, num_layers=layers_per_block,
                    add_downsample=not final, downsample_padding=downsample_padding,
                    cross_attention_dim=cross_attention_dim, num_attention_heads=heads[i],
                    use_linear_projection=use_linear_projection, **common)
            else:
                raise ValueError(f"{kind} does not exist.")
            self.down_blocks.append(blk)

        self.mid_block = UNetMidBlockCrossFrameAttnMotion(
            in_channels=block_out_channels[-1], output_scale_factor=mid_block_scale_factor,
            cross_attention_dim=cross_attention_dim, num_attention_heads=heads[-1], **common)

        self.num_upsamplers = 0
        rev_channels = list(reversed(block_out_channels))
        rev_heads = list(reversed(heads))
        output_channel = rev_channels[0]
        for i, kind in enumerate(up_block_types):
            final = i == len(block_out_channels) - 1
            prev_output_channel, output_channel = output_channel, rev_channels[i]
            input_channel = rev_channels[min(i + 1, len(block_out_channels) - 1)]
            if not final:
                self.num_upsamplers += 1
            if kind == "UpBlockMotion":
                blk = UpBlockMotion(in_channels=input_channel, prev_output_channel=prev_output_channel,
                                    out_channels=output_channel, num_layers=layers_per_block + 1,
                                    add_upsample=not final, resolution_idx=i, **common)
            elif kind == "CrossFrameAttnUpBlockMotion":
                blk = CrossFrameAttnUpBlockMotion(
                    in_channels=input_channel, out_channels=output_channel, prev_output_channel=prev_output_channel,
                    num_layers=layers_per_block + 1, add_upsample=not final,
                    cross_attention_dim=cross_attention_dim, num_attention_heads=rev_heads[i], resolution_idx=i,
                    use_linear_projection=use_linear_projection, **common)
            else:
                raise ValueError(f"{kind} does not exist.")
            self.up_blocks.append(blk)

        if norm_num_groups is not None:
            self.conv_norm_out = nn.GroupNorm(num_channels=block_out_channels[0], num_groups=norm_num_groups,
                                              eps=norm_eps)
            self.conv_act = nn.SiLU()
        else:
            self.conv_norm_out = None
            self.conv_act = None
        self.conv_out = nn.Conv2d(block_out_channels[0], out_channels, kernel_size=3, padding=1)

    # ------------------------------------------------------------------ dtype / device helpers
    @property
    def dtype(self) -> torch.dtype:
        return next(self.parameters()).dtype

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    # ------------------------------------------------------------------ adapter weight plumbing
    def load_i2v_adapter(self, i2v_adapter: I2VAdapterModule) -> None:
        self.down_blocks.load_state_dict(i2v_adapter.down_blocks.state_dict(), strict=False)
        self.up_blocks.load_state_dict(i2v_adapter.up_blocks.state_dict(), strict=False)
        self.mid_block.load_state_dict(i2v_adapter.mid_block.state_dict(), strict=False)

    def obtain_i2v_adapter_modules(self) -> I2VAdapterModule:
        sd = {k: v for k, v in self.state_dict().items() if "i2v_adapter" in k}
        heads = self.num_attention_heads if isinstance(self.num_attention_heads, int) else self.num_attention_heads[0]
        module = I2VAdapterModule(self.layers_per_block, self.config.block_out_channels, heads)
        module.load_state_dict(sd)
        return module

    def load_motion_modules(self, motion_adapter: Optional[MotionAdapter]) -> None:
        """Reference :1028-1036."""
        for i, down_block in enumerate(motion_adapter.down_blocks):
            self.down_blocks[i].motion_modules.load_state_dict(down_block.motion_modules.state_dict())
        for i, up_block in enumerate(motion_adapter.up_blocks):
            self.up_blocks[i].motion_modules.load_state_dict(up_block.motion_modules.state_dict())
        # to support older motion modules that don't have a mid_block
        if hasattr(self.mid_block, "motion_modules") and motion_adapter.mid_block is not None:
            self.mid_block.motion_modules.load_state_dict(motion_adapter.mid_block.motion_modules.state_dict())

    def obtain_motion_modules(self) -> MotionAdapter:
        """Reference :1060-1078."""
        sd = {k: v for k, v in self.state_dict().items() if "motion_modules" in k}
        adapter = MotionAdapter(
            block_out_channels=self.config["block_out_channels"],
            motion_layers_per_block=self.config["layers_per_block"],
            motion_norm_num_groups=self.config["norm_num_groups"],
            motion_num_attention_heads=self.config["motion_num_attention_heads"],
            motion_max_seq_length=self.config["motion_max_seq_length"],
            use_motion_mid_block=bool(self.config["use_motion_mid_block"]))
        adapter.load_state_dict(sd)
        return adapter

    def save_i2v_adapter_modules(self, save_directory: str, is_main_process: bool = True,
                                 safe_serialization: bool = True, variant: Optional[str] = None,
                                 push_to_hub: bool = False, **kwargs) -> None:
        """Reference :1080-1097."""
        self.obtain_i2v_adapter_modules().save_pretrained(
            save_directory=save_directory, is_main_process=is_main_process, safe_serialization=safe_serialization,
            variant=variant, push_to_hub=push_to_hub, **kwargs)

    def save_motion_modules(self, save_directory: str, is_main_process: bool = True, safe_serialization: bool = True,
                            variant: Optional[str] = None, push_to_hub: bool = False, **kwargs) -> None:
        """Reference :1099-1116."""
        self.obtain_motion_modules().save_pretrained(
            save_directory=save_directory, is_main_process=is_main_process, safe_serialization=safe_serialization,
            variant=variant, push_to_hub=push_to_hub, **kwargs)

    def load_ip_adapter(self, weights_path: str) -> None:
        """``ip-adapter_sd15.bin`` / ``.safetensors`` -> IP-Adapter processors + image projection (what the pipeline's
        ``load_ip_adapter`` ends up calling: ``_load_ip_adapter_weights``, reference :1230-1287)."""
        self._load_ip_adapter_weights(checkpoint.load_ip_adapter_file(weights_path))

    def freeze_unet_params(self, freeze_animatediff: bool = True) -> None:
        for p in self.parameters():
            p.requires_grad = False
        for name, p in self.named_parameters():
            if ".i2v_adapter.to_q." in name or ".i2v_adapter.to_out." in name:
                p.requires_grad = True
            elif not freeze_animatediff and ".motion_modules." in name:
                p.requires_grad = True

    # ------------------------------------------------------------------ processor plumbing (drop-in boundary)
    @property
    def attn_processors(self) -> Dict[str, Any]:
        processors: Dict[str, Any] = {}

        def walk(name: str, module: nn.Module):
            if hasattr(module, "get_processor"):
                processors[f"{name}.processor"] = module.get_processor(return_deprecated_lora=True)
            for sub_name, child in module.named_children():
                walk(f"{name}.{sub_name}", child)

        for name, module in self.named_children():
            walk(name, module)
        return processors

    def set_attn_processor(self, processor, _remove_lora: bool = False) -> None:
        count = len(self.attn_processors.keys())
        if isinstance(processor, dict) and len(processor) != count:
            raise ValueError(
                f"A dict of processors was passed, but the number of processors {len(processor)} does not match the"
                f" number of attention layers: {count}. Please make sure to pass {count} processor classes.")

        def walk(name: str, module: nn.Module):
            if hasattr(module, "set_processor"):
                if not isinstance(processor, dict):
                    module.set_processor(processor, _remove_lora=_remove_lora)
                else:
                    module.set_processor(processor.pop(f"{name}.processor"), _remove_lora=_remove_lora)
            for sub_name, child in module.named_children():
                if sub_name == "processor":
                    continue  # a module-type processor is not an attention layer
                walk(f"{name}.{sub_name}", child)

        for name, module in self.named_children():
            walk(name, module)

    def _load_ip_adapter_weights(self, state_dict) -> None:
        """Install IP-Adapter processors on every spatial attn2 and the image projection (reference :1230-1287).
        ``state_dict`` = {"image_proj": {...}, "ip_adapter": {"<key_id>.to_k_ip.weight": ..., ...}}."""
        image_proj = state_dict["image_proj"]
        if "proj.weight" not in image_proj:
            raise ValueError("only the plain IP-Adapter image projection (`proj.weight`) is supported")
        num_image_text_embeds = 4
        self.encoder_hid_proj = None
        attn_procs = {}
        key_id = 1
        for name in self.attn_processors.keys():
            cross_dim = None if not name.endswith("attn2.processor") else self.config.cross_attention_dim
            if name.startswith("mid_block"):
                hidden_size = self.config.block_out_channels[-1]
            elif name.startswith("up_blocks"):
                hidden_size = list(reversed(self.config.block_out_channels))[int(name[len("up_blocks.")])]
            else:
                hidden_size = self.config.block_out_channels[int(name[len("down_blocks.")])]
            if cross_dim is None or "motion_modules" in name:
                attn_procs[name] = AttnProcessor2_0()
            else:
                proc = IPAdapterAttnProcessor2_0(hidden_size=hidden_size, cross_attention_dim=cross_dim, scale=1.0,
                                                 num_tokens=num_image_text_embeds).to(dtype=self.dtype,
                                                                                      device=self.device)
                proc.load_state_dict({k: state_dict["ip_adapter"][f"{key_id}.{k}"] for k in proc.state_dict()})
                attn_procs[name] = proc
                key_id += 2
        self.set_attn_processor(attn_procs)

        cross = image_proj["proj.weight"].shape[0] // 4
        proj = ImageProjection(image_embed_dim=image_proj["proj.weight"].shape[-1], cross_attention_dim=cross,
                               num_image_text_embeds=4)
        proj.load_state_dict({"image_embeds.weight": image_proj["proj.weight"],
                              "image_embeds.bias": image_proj["proj.bias"],
                              "norm.weight": image_proj["norm.weight"], "norm.bias": image_proj["norm.bias"]})
        self.encoder_hid_proj = proj.to(device=self.device, dtype=self.dtype)
        self.config.encoder_hid_dim_type = "ip_image_proj"

    # ------------------------------------------------------------------ forward (reference :1289-1451)
    def forward(self, sample, timestep, enable_cross_frame_attn: bool, encoder_hidden_states, timestep_cond=None,
                attention_mask=None, cross_attention_kwargs=None, added_cond_kwargs=None,
                down_block_additional_residuals=None, mid_block_additional_residual=None, return_dict: bool = True):
        up_factor = 2 ** self.num_upsamplers
        forward_upsample_size = any(s % up_factor != 0 for s in sample.shape[-2:])
        upsample_size = None
        if attention_mask is not None:
            attention_mask = ((1 - attention_mask.to(sample.dtype)) * -10000.0).unsqueeze(1)

        timesteps = timestep
        if not torch.is_tensor(timesteps):
            dtype = torch.float64 if isinstance(timestep, float) else torch.int64
            timesteps = torch.tensor([timesteps], dtype=dtype, device=sample.device)
        elif timesteps.dim() == 0:
            timesteps = timesteps[None].to(sample.device)

        num_frames = sample.shape[1]
        timesteps = timesteps.expand(sample.shape[0])
        emb = self.time_embedding(self.time_proj(timesteps).to(dtype=self.dtype), timestep_cond)
        emb = emb.repeat_interleave(repeats=num_frames, dim=0)

        if self.encoder_hid_proj is not None and self.config.encoder_hid_dim_type == "ip_image_proj":
            if added_cond_kwargs is None or "image_embeds" not in added_cond_kwargs:
                raise ValueError(
                    f"{self.__class__} has the config param `encoder_hid_dim_type` set to 'ip_image_proj' which "
                    f"requires the keyword argument `image_embeds` to be passed in  `added_conditions`")
            image_embeds = self.encoder_hid_proj(added_cond_kwargs.get("image_embeds")).to(encoder_hidden_states.dtype)
            encoder_hidden_states = torch.cat([encoder_hidden_states, image_embeds], dim=1)
        encoder_hidden_states = encoder_hidden_states.repeat_interleave(repeats=num_frames, dim=0)

        # frames fold into the batch: row = video * num_frames + frame
        sample = sample.reshape((sample.shape[0] * num_frames, -1) + sample.shape[3:])
        sample = self.conv_in(sample)

        down_res = (sample,)
        for blk in self.down_blocks:
            if getattr(blk, "has_cross_attention", False):
                sample, res = blk(hidden_states=sample, temb=emb, enable_cross_frame_attn=enable_cross_frame_attn,
                                  encoder_hidden_states=encoder_hidden_states, attention_mask=attention_mask,
                                  num_frames=num_frames, cross_attention_kwargs=cross_attention_kwargs)
            else:
                sample, res = blk(hidden_states=sample, temb=emb, num_frames=num_frames)
            down_res += res
        if down_block_additional_residuals is not None:
            down_res = tuple(r + a for r, a in zip(down_res, down_block_additional_residuals))

        if self.mid_block is not None:
            sample = self.mid_block(sample, emb, enable_cross_frame_attn=enable_cross_frame_attn,
                                    encoder_hidden_states=encoder_hidden_states, attention_mask=attention_mask,
                                    num_frames=num_frames, cross_attention_kwargs=cross_attention_kwargs)
        if mid_block_additional_residual is not None:
            sample = sample + mid_block_additional_residual

        for i, blk in enumerate(self.up_blocks):
            final = i == len(self.up_blocks) - 1
            res = down_res[-len(blk.resnets):]
            down_res = down_res[: -len(blk.resnets)]
            if not final and forward_upsample_size:
                upsample_size = down_res[-1].shape[2:]
            if getattr(blk, "has_cross_attention", False):
                sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=res,
                             enable_cross_frame_attn=enable_cross_frame_attn,
                             encoder_hidden_states=encoder_hidden_states, upsample_size=upsample_size,
                             attention_mask=attention_mask, num_frames=num_frames,
                             cross_attention_kwargs=cross_attention_kwargs)
            else:
                sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=res,
                             upsample_size=upsample_size, num_frames=num_frames)

        if self.conv_norm_out is not None:
            sample = self.conv_act(self.conv_norm_out(sample))
        sample = self.conv_out(sample)
        sample = sample[None, :].reshape((-1, num_frames) + sample.shape[1:])
        if not return_dict:
            return (sample,)
        return _Sample(sample)
