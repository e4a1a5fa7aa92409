"""AnimateDiff motion module and the attention-free motion blocks (diffusers-0.25 ``TransformerTemporalModel``,
``DownBlockMotion``, ``UpBlockMotion``).  The reference constructs them at
``/root/reference/src/models/unet_motion_cross_frame_attn.py:55-68, 123-137, 232-244``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from .i2v_adapter import _Sample
from .layers import BasicTransformerBlock, Downsample2D, ResnetBlock2D, Upsample2D


class TransformerTemporalModel(nn.Module):
    def __init__(self, num_attention_heads: int = 16, attention_head_dim: int = 88, in_channels: Optional[int] = None,
                 out_channels: Optional[int] = None, num_layers: int = 1, dropout: float = 0.0,
                 norm_num_groups: int = 32, cross_attention_dim: Optional[int] = None, attention_bias: bool = False,
                 sample_size: Optional[int] = None, activation_fn: str = "geglu", norm_elementwise_affine: bool = True,
                 double_self_attention: bool = True, positional_embeddings: Optional[str] = None,
                 num_positional_embeddings: Optional[int] = None):
        super().__init__()
        inner_dim = num_attention_heads * attention_head_dim
        self.in_channels = in_channels
        self.norm = nn.GroupNorm(num_groups=norm_num_groups, num_channels=in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner_dim)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner_dim, num_attention_heads, attention_head_dim, dropout=dropout,
                                  cross_attention_dim=cross_attention_dim, attention_bias=attention_bias,
                                  double_self_attention=double_self_attention,
                                  positional_embeddings=positional_embeddings,
                                  num_positional_embeddings=num_positional_embeddings)
            for _ in range(num_layers)
        ])
        self.proj_out = nn.Linear(inner_dim, in_channels)

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, class_labels=None,
                num_frames: int = 1, cross_attention_kwargs=None, return_dict: bool = True):
        batch_frames, channel, height, width = hidden_states.shape
        batch = batch_frames // num_frames
        residual = hidden_states

        x = hidden_states[None, :].reshape(batch, num_frames, channel, height, width).permute(0, 2, 1, 3, 4)
        x = self.norm(x)  # statistics over (C/groups, F, h, w): couples the frames of a video
        x = x.permute(0, 3, 4, 2, 1).reshape(batch * height * width, num_frames, channel)
        x = self.proj_in(x)
        for block in self.transformer_blocks:
            x = block(x, encoder_hidden_states=encoder_hidden_states, timestep=timestep,
                      cross_attention_kwargs=cross_attention_kwargs, class_labels=class_labels)
        x = self.proj_out(x)
        x = (x[None, None, :].reshape(batch, height, width, num_frames, channel)
             .permute(0, 3, 4, 1, 2).contiguous().reshape(batch_frames, channel, height, width))
        output = x + residual
        if not return_dict:
            return (output,)
        return _Sample(output)


def _motion_module(channels: int, heads: int, groups: int, max_seq: int, cross_dim=None) -> TransformerTemporalModel:
    return TransformerTemporalModel(
        num_attention_heads=heads, in_channels=channels, norm_num_groups=groups, cross_attention_dim=cross_dim,
        activation_fn="geglu", positional_embeddings="sinusoidal", num_positional_embeddings=max_seq,
        attention_head_dim=channels // heads, attention_bias=False)


class DownBlockMotion(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, dropout: float = 0.0,
                 num_layers: int = 1, resnet_eps: float = 1e-6, resnet_groups: int = 32,
                 output_scale_factor: float = 1.0, add_downsample: bool = True, downsample_padding: int = 1,
                 temporal_num_attention_heads: int = 1, temporal_cross_attention_dim: Optional[int] = None,
                 temporal_max_seq_length: int = 32, **_unused):
        super().__init__()
        resnets, motion_modules = [], []
        for i in range(num_layers):
            resnets.append(ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels,
                                         eps=resnet_eps, groups=resnet_groups, dropout=dropout,
                                         output_scale_factor=output_scale_factor))
            motion_modules.append(_motion_module(out_channels, temporal_num_attention_heads, resnet_groups,
                                                 temporal_max_seq_length, temporal_cross_attention_dim))
        self.resnets = nn.ModuleList(resnets)
        self.motion_modules = nn.ModuleList(motion_modules)
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels,
                                                         padding=downsample_padding, name="op")])
                             if add_downsample else None)

    def forward(self, hidden_states, temb=None, scale: float = 1.0, num_frames: int = 1):
        output_states: Tuple[torch.Tensor, ...] = ()
        for resnet, motion_module in zip(self.resnets, self.motion_modules):
            hidden_states = resnet(hidden_states, temb)
            hidden_states = motion_module(hidden_states, num_frames=num_frames)[0]
            output_states = output_states + (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            output_states = output_states + (hidden_states,)
        return hidden_states, output_states


class UpBlockMotion(nn.Module):
    def __init__(self, in_channels: int, prev_output_channel: int, out_channels: int, temb_channels: int,
                 resolution_idx: Optional[int] = None, dropout: float = 0.0, num_layers: int = 1,
                 resnet_eps: float = 1e-6, resnet_groups: int = 32, output_scale_factor: float = 1.0,
                 add_upsample: bool = True, temporal_num_attention_heads: int = 8,
                 temporal_cross_attention_dim: Optional[int] = None, temporal_max_seq_length: int = 32, **_unused):
        super().__init__()
        resnets, motion_modules = [], []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            rin = prev_output_channel if i == 0 else out_channels
            resnets.append(ResnetBlock2D(rin + skip, out_channels, temb_channels, eps=resnet_eps,
                                         groups=resnet_groups, dropout=dropout,
                                         output_scale_factor=output_scale_factor))
            motion_modules.append(_motion_module(out_channels, temporal_num_attention_heads, resnet_groups,
                                                 temporal_max_seq_length, temporal_cross_attention_dim))
        self.resnets = nn.ModuleList(resnets)
        self.motion_modules = nn.ModuleList(motion_modules)
        self.upsamplers = (nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)])
                           if add_upsample else None)
        self.resolution_idx = resolution_idx

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, upsample_size=None, scale: float = 1.0,
                num_frames: int = 1):
        for resnet, motion_module in zip(self.resnets, self.motion_modules):
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res], dim=1)
            hidden_states = resnet(hidden_states, temb)
            hidden_states = motion_module(hidden_states, num_frames=num_frames)[0]
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


class MotionModules(nn.Module):
    """``layers_per_block`` temporal transformers of one resolution (diffusers ``unet_motion_model.MotionModules``)."""

    def __init__(self, in_channels: int, layers_per_block: int = 2, num_attention_heads: int = 8,
                 norm_num_groups: int = 32, max_seq_length: int = 32):
        super().__init__()
        self.motion_modules = nn.ModuleList(
            [_motion_module(in_channels, num_attention_heads, norm_num_groups, max_seq_length)
             for _ in range(layers_per_block)])


class MotionAdapter(nn.Module):
    """Weights-only container of the AnimateDiff motion modules with the UNet's key layout
    (``{down,up}_blocks.N.motion_modules.M.*``, ``mid_block.motion_modules.M.*``): the diffusers ``MotionAdapter`` the
    reference loads with ``from_pretrained`` (``src/pipelines/pipeline_i2v_adapter.py:733,745``) and rebuilds in
    ``obtain_motion_modules`` (``src/models/unet_motion_cross_frame_attn.py:1060-1078``)."""

    def __init__(self, block_out_channels=(320, 640, 1280, 1280), motion_layers_per_block: int = 2,
                 motion_mid_block_layers_per_block: int = 1, motion_num_attention_heads: int = 8,
                 motion_norm_num_groups: int = 32, motion_max_seq_length: int = 32, use_motion_mid_block: bool = True,
                 conv_in_channels: Optional[int] = None):
        super().__init__()
        self.config = dict(block_out_channels=tuple(block_out_channels), motion_layers_per_block=motion_layers_per_block,
                           motion_mid_block_layers_per_block=motion_mid_block_layers_per_block,
                           motion_num_attention_heads=motion_num_attention_heads,
                           motion_norm_num_groups=motion_norm_num_groups, motion_max_seq_length=motion_max_seq_length,
                           use_motion_mid_block=use_motion_mid_block, conv_in_channels=conv_in_channels)
        if conv_in_channels is not None:
            raise ValueError("conv_in_channels (PIA) is not part of the SD1.5 motion adapter this mirror covers")

        def mods(ch, layers):
            return MotionModules(ch, layers, motion_num_attention_heads, motion_norm_num_groups, motion_max_seq_length)

        self.down_blocks = nn.ModuleList([mods(c, motion_layers_per_block) for c in block_out_channels])
        self.mid_block = (mods(block_out_channels[-1], motion_mid_block_layers_per_block)
                          if use_motion_mid_block else None)
        self.up_blocks = nn.ModuleList([mods(c, motion_layers_per_block + 1) for c in reversed(block_out_channels)])

    def forward(self, sample):  # pragma: no cover - container only
        pass

    def save_pretrained(self, save_directory: str, is_main_process: bool = True, safe_serialization: bool = True,
                        variant: Optional[str] = None, push_to_hub: bool = False, **_unused) -> None:
        from . import checkpoint

        if push_to_hub:
            raise ValueError("push_to_hub is not available: this environment has no network")
        if is_main_process:
            checkpoint.save_model_directory(self, self.config, "MotionAdapter", save_directory, safe_serialization,
                                            variant)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, torch_dtype: Optional[torch.dtype] = None,
                        variant: Optional[str] = None, **_unused) -> "MotionAdapter":
        from . import checkpoint

        config, state = checkpoint.load_model_directory(pretrained_model_name_or_path, variant)
        module = cls(**checkpoint.constructor_kwargs(
            config, ("block_out_channels", "motion_layers_per_block", "motion_mid_block_layers_per_block",
                     "motion_num_attention_heads", "motion_norm_num_groups", "motion_max_seq_length",
                     "use_motion_mid_block", "conv_in_channels")))
        module.load_state_dict(state)
        if torch_dtype is not None:
            module.to(torch_dtype)
        return module.eval()
