"""Building blocks of the SD1.5 / AnimateDiff UNet around the attention hot path (diffusers-0.25 semantics,
SURVEY.md Appendix A2, A5-A9).  These stay on the PyTorch path (cuBLAS / cuDNN / ATen), as BASELINE.json's north_star
prescribes; parameter names match diffusers so reference checkpoints load with ``load_state_dict``.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from .attention import Attention


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x, scale: float = 1.0):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    """``net.0`` GEGLU(dim -> 4 dim), ``net.1`` dropout, ``net.2`` Linear(4 dim -> dim)."""

    def __init__(self, dim: int, mult: int = 4, dropout: float = 0.0):
        super().__init__()
        inner = dim * mult
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(dropout), nn.Linear(inner, dim)])

    def forward(self, x, scale: float = 1.0):
        for m in self.net:
            x = m(x)
        return x


class SinusoidalPositionalEmbedding(nn.Module):
    def __init__(self, embed_dim: int, max_seq_length: int = 32):
        super().__init__()
        position = torch.arange(max_seq_length).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, embed_dim, 2) * (-math.log(10000.0) / embed_dim))
        pe = torch.zeros(1, max_seq_length, embed_dim)
        pe[0, :, 0::2] = torch.sin(position * div_term)
        pe[0, :, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)

    def forward(self, x):
        return x + self.pe[:, : x.shape[1]]


class BasicTransformerBlock(nn.Module):
    """layer_norm flavour of diffusers' BasicTransformerBlock (the only one on the SD1.5 / motion-module path)."""

    def __init__(
        self,
        dim: int,
        num_attention_heads: int,
        attention_head_dim: int,
        dropout: float = 0.0,
        cross_attention_dim: Optional[int] = None,
        attention_bias: bool = False,
        only_cross_attention: bool = False,
        double_self_attention: bool = False,
        upcast_attention: bool = False,
        norm_eps: float = 1e-5,
        positional_embeddings: Optional[str] = None,
        num_positional_embeddings: Optional[int] = None,
        attention_out_bias: bool = True,
    ):
        super().__init__()
        self.only_cross_attention = only_cross_attention
        if positional_embeddings == "sinusoidal":
            self.pos_embed = SinusoidalPositionalEmbedding(dim, max_seq_length=num_positional_embeddings)
        else:
            self.pos_embed = None
        self.norm1 = nn.LayerNorm(dim, eps=norm_eps)
        self.attn1 = Attention(
            query_dim=dim, heads=num_attention_heads, dim_head=attention_head_dim, dropout=dropout,
            bias=attention_bias, cross_attention_dim=cross_attention_dim if only_cross_attention else None,
            upcast_attention=upcast_attention, out_bias=attention_out_bias,
        )
        if cross_attention_dim is not None or double_self_attention:
            self.norm2 = nn.LayerNorm(dim, eps=norm_eps)
            self.attn2 = Attention(
                query_dim=dim, cross_attention_dim=cross_attention_dim if not double_self_attention else None,
                heads=num_attention_heads, dim_head=attention_head_dim, dropout=dropout, bias=attention_bias,
                upcast_attention=upcast_attention, out_bias=attention_out_bias,
            )
        else:
            self.norm2 = None
            self.attn2 = None
        self.norm3 = nn.LayerNorm(dim, eps=norm_eps)
        self.ff = FeedForward(dim, dropout=dropout)

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                timestep=None, cross_attention_kwargs=None, class_labels=None):
        kw = dict(cross_attention_kwargs) if cross_attention_kwargs is not None else {}
        norm_h = self.norm1(hidden_states)
        if self.pos_embed is not None:
            norm_h = self.pos_embed(norm_h)
        attn_output = self.attn1(
            norm_h, encoder_hidden_states=encoder_hidden_states if self.only_cross_attention else None,
            attention_mask=attention_mask, **kw)
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


class Timesteps(nn.Module):
    def __init__(self, num_channels: int, flip_sin_to_cos: bool, downscale_freq_shift: float):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift

    def forward(self, timesteps):
        half = self.num_channels // 2
        exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device)
        exponent = exponent / (half - self.downscale_freq_shift)
        emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
        emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
        if self.flip_sin_to_cos:
            emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
        return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int, act_fn: str = "silu"):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample, condition=None):
        return self.linear_2(self.act(self.linear_1(sample)))


class ImageProjection(nn.Module):
    def __init__(self, image_embed_dim: int = 1024, cross_attention_dim: int = 768, num_image_text_embeds: int = 4):
        super().__init__()
        self.num_image_text_embeds = num_image_text_embeds
        self.image_embeds = nn.Linear(image_embed_dim, num_image_text_embeds * cross_attention_dim)
        self.norm = nn.LayerNorm(cross_attention_dim)

    def forward(self, image_embeds):
        b = image_embeds.shape[0]
        x = self.image_embeds(image_embeds).reshape(b, self.num_image_text_embeds, -1)
        return self.norm(x)


class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, eps: float = 1e-6, groups: int = 32,
                 dropout: float = 0.0, output_scale_factor: float = 1.0, **_unused):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None
        self.output_scale_factor = output_scale_factor

    def forward(self, x, temb, scale: float = 1.0):
        h = self.conv1(self.nonlinearity(self.norm1(x)))
        h = h + self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return (x + h) / self.output_scale_factor


class Downsample2D(nn.Module):
    def __init__(self, channels: int, use_conv: bool = True, out_channels: Optional[int] = None, padding: int = 1,
                 name: str = "conv"):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(channels, out_channels or channels, 3, stride=2, padding=padding)

    def forward(self, x, scale: float = 1.0):
        if self.padding == 0:   # diffusers: asymmetric (right / bottom) zero padding in front of the unpadded conv
            x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels: int, use_conv: bool = True, out_channels: Optional[int] = None):
        super().__init__()
        self.conv = nn.Conv2d(channels, out_channels or channels, 3, padding=1)

    def forward(self, x, output_size=None, scale: float = 1.0):
        if output_size is None:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        else:
            x = F.interpolate(x, size=output_size, mode="nearest")
        return self.conv(x)
