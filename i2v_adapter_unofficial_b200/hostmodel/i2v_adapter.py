"""I2V-Adapter transformer modules — host-side mirror of ``/root/reference/src/modules/i2v_adapter.py``.

Same class names, constructor arguments (the subset SD1.5 uses), ``forward`` signatures, parameter names and error
behaviour as the reference, so the reference's own tests read the same against these classes:

* ``I2VAdapterTransformerBlock``   reference :356-565  (spatial self-attn + cross-frame attn to frame 0, text/IP
                                   cross-attn, GEGLU feed-forward)
* ``I2VAdapterTransformer2DModel`` reference :95-354   (GroupNorm + 1x1 conv in, blocks, 1x1 conv out + residual)
* ``I2VAdapterModule``             reference :17-93    (weight-only container with the UNet's key layout)

All arithmetic is delegated to the ``Attention`` modules' processors — the drop-in boundary.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch
from torch import nn

from . import checkpoint
from .attention import Attention
from .layers import BasicTransformerBlock


class I2VAdapterTransformerBlock(BasicTransformerBlock):
    def __init__(self, dim: int, num_attention_heads: int, attention_head_dim: int, dropout: float = 0.0,
                 cross_attention_dim: Optional[int] = None, attention_bias: bool = False,
                 only_cross_attention: bool = False, double_self_attention: bool = False,
                 upcast_attention: bool = False, norm_eps: float = 1e-5, attention_out_bias: bool = True, **_unused):
        super().__init__(dim, num_attention_heads, attention_head_dim, dropout=dropout,
                         cross_attention_dim=cross_attention_dim, attention_bias=attention_bias,
                         only_cross_attention=only_cross_attention, double_self_attention=double_self_attention,
                         upcast_attention=upcast_attention, norm_eps=norm_eps, attention_out_bias=attention_out_bias)
        # registered after attn1/attn2 -> enumeration order attn1, attn2, i2v_adapter
        self.i2v_adapter = Attention(query_dim=dim, heads=num_attention_heads, dim_head=attention_head_dim,
                                     dropout=dropout, bias=attention_bias, cross_attention_dim=dim,
                                     upcast_attention=upcast_attention, out_bias=attention_out_bias)

    def forward(self, hidden_states, enable_cross_frame_attn: bool = False, num_frames: Optional[int] = None,
                attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None, timestep=None,
                cross_attention_kwargs: Optional[Dict[str, Any]] = None, class_labels=None, added_cond_kwargs=None):
        batch_size = hidden_states.shape[0]
        kw = dict(cross_attention_kwargs) if cross_attention_kwargs is not None else {}
        kw.pop("gligen", None)

        norm_h = self.norm1(hidden_states)
        if self.pos_embed is not None:
            norm_h = self.pos_embed(norm_h)

        attn_output = self.attn1(
            norm_h, encoder_hidden_states=encoder_hidden_states if self.only_cross_attention else None,
            attention_mask=attention_mask, **kw)

        if enable_cross_frame_attn:
            if num_frames is None:
                raise ValueError("`num_frames` must be provided when `enable_cross_frame_attn` is True.")
            if batch_size % num_frames != 0:
                raise ValueError(
                    f"Batch size {batch_size} must be divisible by the number of frames {num_frames}.")
            # every frame attends to the first frame of its clip: rows b*F .. b*F+F-1 all see row b*F
            first = norm_h[0:batch_size:num_frames]
            first = first.repeat_interleave(num_frames, dim=0)
            cross = self.i2v_adapter(norm_h, encoder_hidden_states=first, attention_mask=None, **kw)
            attn_output = attn_output + cross

        hidden_states = attn_output + hidden_states

        if self.attn2 is not None:
            norm_h = self.norm2(hidden_states)
            if self.pos_embed is not None:
                norm_h = self.pos_embed(norm_h)
            attn_output = self.attn2(norm_h, encoder_hidden_states=encoder_hidden_states,
                                     attention_mask=encoder_attention_mask, **kw)
            hidden_states = attn_output + hidden_states

        hidden_states = self.ff(self.norm3(hidden_states)) + hidden_states
        return hidden_states


class I2VAdapterTransformer2DModel(nn.Module):
    """Continuous-input Transformer2DModel (SD1.5: conv projections) whose blocks are I2VAdapterTransformerBlocks."""

    def __init__(self, num_attention_heads: int = 16, attention_head_dim: int = 88, in_channels: Optional[int] = None,
                 out_channels: Optional[int] = None, num_layers: int = 1, dropout: float = 0.0,
                 norm_num_groups: int = 32, cross_attention_dim: Optional[int] = None, attention_bias: bool = False,
                 use_linear_projection: bool = False, only_cross_attention: bool = False,
                 double_self_attention: bool = False, upcast_attention: bool = False, norm_eps: float = 1e-5,
                 **_unused):
        super().__init__()
        inner_dim = num_attention_heads * attention_head_dim
        self.in_channels = in_channels
        self.out_channels = in_channels if out_channels is None else out_channels
        self.use_linear_projection = use_linear_projection
        self.norm = nn.GroupNorm(num_groups=norm_num_groups, num_channels=in_channels, eps=1e-6, affine=True)
        if use_linear_projection:
            self.proj_in = nn.Linear(in_channels, inner_dim)
            self.proj_out = nn.Linear(inner_dim, in_channels)
        else:
            self.proj_in = nn.Conv2d(in_channels, inner_dim, kernel_size=1)
            self.proj_out = nn.Conv2d(inner_dim, in_channels, kernel_size=1)
        self.transformer_blocks = nn.ModuleList([
            I2VAdapterTransformerBlock(inner_dim, num_attention_heads, attention_head_dim, dropout=dropout,
                                       cross_attention_dim=cross_attention_dim, attention_bias=attention_bias,
                                       only_cross_attention=only_cross_attention,
                                       double_self_attention=double_self_attention,
                                       upcast_attention=upcast_attention, norm_eps=norm_eps)
            for _ in range(num_layers)
        ])

    def from_transformer2d_model(self, transformer2d_model: nn.Module) -> None:
        """Initialise from a plain 2-D transformer: i2v_adapter <- attn1 with a zero output projection
        (reference :171-182)."""
        self.load_state_dict(transformer2d_model.state_dict(), strict=False)
        for mine, theirs in zip(self.transformer_blocks, transformer2d_model.transformer_blocks):
            mine.i2v_adapter.load_state_dict(theirs.attn1.state_dict())
            mine.i2v_adapter.to_out[0].weight.data.zero_()
            mine.i2v_adapter.to_out[0].bias.data.zero_()

    def forward(self, hidden_states, enable_cross_frame_attn: bool = False, encoder_hidden_states=None,
                num_frames: Optional[int] = None, timestep=None, added_cond_kwargs=None, class_labels=None,
                cross_attention_kwargs=None, attention_mask=None, encoder_attention_mask=None,
                return_dict: bool = True):
        if attention_mask is not None and attention_mask.ndim == 2:
            attention_mask = ((1 - attention_mask.to(hidden_states.dtype)) * -10000.0).unsqueeze(1)
        if encoder_attention_mask is not None and encoder_attention_mask.ndim == 2:
            encoder_attention_mask = ((1 - encoder_attention_mask.to(hidden_states.dtype)) * -10000.0).unsqueeze(1)

        batch, _, height, width = hidden_states.shape
        residual = hidden_states
        hidden_states = self.norm(hidden_states)
        if not self.use_linear_projection:
            hidden_states = self.proj_in(hidden_states)
            inner_dim = hidden_states.shape[1]
            hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(batch, height * width, inner_dim)
        else:
            inner_dim = hidden_states.shape[1]
            hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(batch, height * width, inner_dim)
            hidden_states = self.proj_in(hidden_states)

        for block in self.transformer_blocks:
            hidden_states = block(
                hidden_states, enable_cross_frame_attn=enable_cross_frame_attn, num_frames=num_frames,
                attention_mask=attention_mask, encoder_hidden_states=encoder_hidden_states,
                encoder_attention_mask=encoder_attention_mask, timestep=timestep,
                cross_attention_kwargs=cross_attention_kwargs, class_labels=class_labels)

        if not self.use_linear_projection:
            hidden_states = hidden_states.reshape(batch, height, width, inner_dim).permute(0, 3, 1, 2).contiguous()
            hidden_states = self.proj_out(hidden_states)
        else:
            hidden_states = self.proj_out(hidden_states)
            hidden_states = hidden_states.reshape(batch, height, width, inner_dim).permute(0, 3, 1, 2).contiguous()
        output = hidden_states + residual
        if not return_dict:
            return (output,)
        return _Sample(output)


class _Sample:
    """Stand-in for diffusers' output dataclasses: ``.sample`` attribute and tuple-style ``[0]``."""

    def __init__(self, sample):
        self.sample = sample

    def __getitem__(self, idx):
        return (self.sample,)[idx]


def _adapter_block(out_channel: int, depth: int, heads: int, layers: int) -> nn.Module:
    block = nn.Module()
    attentions = []
    for _ in range(depth):
        attn_block = nn.Module()
        tblocks = []
        for _ in range(layers):
            tb = nn.Module()
            tb.i2v_adapter = Attention(query_dim=out_channel, heads=heads, dim_head=out_channel // heads,
                                       cross_attention_dim=out_channel)
            tblocks.append(tb)
        attn_block.transformer_blocks = nn.ModuleList(tblocks)
        attentions.append(attn_block)
    block.attentions = nn.ModuleList(attentions)
    return block


class I2VAdapterModule(nn.Module):
    """Weights-only container ``{down,up,mid}_blocks[].attentions[].transformer_blocks[].i2v_adapter`` used to save
    and load the trained adapter separately from the UNet (reference :49-93)."""

    def __init__(self, block_depth: int, block_out_channels, num_attention_heads: int,
                 transformer_layers_per_block: int = 1, mid_block_depth: int = 1):
        super().__init__()
        self.config = dict(block_depth=block_depth, block_out_channels=tuple(block_out_channels),
                           num_attention_heads=num_attention_heads,
                           transformer_layers_per_block=transformer_layers_per_block, mid_block_depth=mid_block_depth)
        chans = list(block_out_channels[:-1])
        self.down_blocks = nn.ModuleList(
            [_adapter_block(c, block_depth, num_attention_heads, transformer_layers_per_block) for c in chans])
        ups = [_adapter_block(c, block_depth + 1, num_attention_heads, transformer_layers_per_block)
               for c in reversed(chans)]
        # dummy first entry keeps the indices aligned with the UNet's up_blocks (the first one has no attention)
        self.up_blocks = nn.ModuleList([nn.Identity()] + ups)
        self.mid_block = _adapter_block(block_out_channels[-1], mid_block_depth, num_attention_heads,
                                        transformer_layers_per_block)

    def forward(self):  # pragma: no cover - container only
        pass

    # ---- on-disk format of the trained adapter (diffusers ModelMixin layout; reference call sites
    # src/pipelines/pipeline_i2v_adapter.py:740, src/models/unet_motion_cross_frame_attn.py:1080-1097) ----
    def save_pretrained(self, save_directory: str, is_main_process: bool = True, safe_serialization: bool = True,
                        variant: Optional[str] = None, push_to_hub: bool = False, **_unused) -> None:
        if push_to_hub:
            raise ValueError("push_to_hub is not available: this environment has no network")
        if is_main_process:
            checkpoint.save_model_directory(self, self.config, "I2VAdapterModule", save_directory, safe_serialization,
                                            variant)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, torch_dtype: Optional[torch.dtype] = None,
                        variant: Optional[str] = None, **_unused) -> "I2VAdapterModule":
        config, state = checkpoint.load_model_directory(pretrained_model_name_or_path, variant)
        module = cls(**checkpoint.constructor_kwargs(
            config, ("block_depth", "block_out_channels", "num_attention_heads", "transformer_layers_per_block",
                     "mid_block_depth")))
        module.load_state_dict(state)   # strict, as ModelMixin: missing / unexpected keys are an error
        if torch_dtype is not None:
            module.to(torch_dtype)
        return module.eval()
