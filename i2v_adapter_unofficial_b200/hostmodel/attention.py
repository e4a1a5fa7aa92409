"""``Attention`` module and the stock processors, diffusers-0.25 semantics.

The reference builds every attention of the hot path from ``diffusers.models.attention_processor.Attention``
(``/root/reference/src/modules/i2v_adapter.py:32-37, 409-418``) and dispatches through a pluggable processor.
``diffusers`` is not installable in this image, so this file provides the same interface: parameter names
(``to_q``, ``to_k``, ``to_v``, ``to_out.0``), ``set_processor`` / ``get_processor`` and the processor calling
convention ``processor(attn, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw)``.

``AttnProcessor2_0`` / ``IPAdapterAttnProcessor2_0`` below are the *stock* PyTorch-SDPA processors — the path the
reference runs — and are what the B200 processors in ``..processors`` replace.  They are host/plumbing code: the
B200 processors never call them.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn


class Attention(nn.Module):
    def __init__(
        self,
        query_dim: int,
        cross_attention_dim: Optional[int] = None,
        heads: int = 8,
        dim_head: int = 64,
        dropout: float = 0.0,
        bias: bool = False,
        upcast_attention: bool = False,
        out_bias: bool = True,
        processor=None,
    ):
        super().__init__()
        self.inner_dim = dim_head * heads
        self.query_dim = query_dim
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.is_cross_attention = cross_attention_dim is not None
        self.heads = heads
        self.scale = dim_head**-0.5
        self.upcast_attention = upcast_attention
        # knobs diffusers exposes and the SD1.5 path leaves at their defaults
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.group_norm = None
        self.spatial_norm = None
        self.norm_cross = None

        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(dropout)])
        self.set_processor(processor if processor is not None else AttnProcessor2_0())

    def set_processor(self, processor, _remove_lora: bool = False) -> None:
        # a processor that is an nn.Module (IP-Adapter) must be (un)registered as a sub-module
        if hasattr(self, "processor") and isinstance(self.processor, nn.Module) and not isinstance(processor, nn.Module):
            self._modules.pop("processor")
        self.processor = processor

    def get_processor(self, return_deprecated_lora: bool = False):
        return self.processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(
            self,
            hidden_states,
            encoder_hidden_states=encoder_hidden_states,
            attention_mask=attention_mask,
            **cross_attention_kwargs,
        )


def _split_heads(t: torch.Tensor, heads: int) -> torch.Tensor:
    b, n, c = t.shape
    return t.view(b, n, heads, c // heads).transpose(1, 2)


def prepare_attention_mask(mask: torch.Tensor, target_length: int, batch: int, heads: int) -> torch.Tensor:
    """diffusers 0.25 ``Attention.prepare_attention_mask`` followed by the view ``AttnProcessor2_0`` takes of it:
    a (B, 1, L) or (B, L) additive mask, padded to ``target_length`` keys, repeated per head ->
    (B, heads, 1 or Sq, L) for ``F.scaled_dot_product_attention``."""
    if mask.dim() == 2:
        mask = mask[:, None, :]
    if mask.shape[-1] < target_length:   # (diffusers pads by target_length itself, which never matches; pad the rest)
        mask = F.pad(mask, (0, target_length - mask.shape[-1]), value=0.0)
    if mask.shape[0] == batch:
        mask = mask.repeat_interleave(heads, dim=0)   # (B * heads, q, L), head index fastest as in diffusers
    return mask.view(batch, heads, -1, mask.shape[-1])


class AttnProcessor2_0:
    """softmax(q k^T / sqrt(d)) v through ``F.scaled_dot_product_attention`` followed by ``to_out``."""

    def __call__(self, attn: Attention, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 scale: float = 1.0, **kwargs):
        batch = hidden_states.shape[0]
        context = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        if attention_mask is not None:
            attention_mask = prepare_attention_mask(attention_mask, context.shape[1], batch, attn.heads)
        q = _split_heads(attn.to_q(hidden_states), attn.heads)
        k = _split_heads(attn.to_k(context), attn.heads)
        v = _split_heads(attn.to_v(context), attn.heads)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(batch, -1, attn.inner_dim).to(q.dtype)
        o = attn.to_out[0](o)
        o = attn.to_out[1](o)
        if attn.residual_connection:
            o = o + hidden_states
        return o / attn.rescale_output_factor


class IPAdapterAttnProcessor2_0(nn.Module):
    """Decoupled text + image-prompt cross-attention (two softmaxes, scaled add), then ``to_out``."""

    def __init__(self, hidden_size: int, cross_attention_dim: Optional[int] = None, num_tokens: int = 4,
                 scale: float = 1.0):
        super().__init__()
        self.hidden_size = hidden_size
        self.cross_attention_dim = cross_attention_dim
        self.num_tokens = num_tokens
        self.scale = scale
        self.to_k_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)
        self.to_v_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)

    def forward(self, attn: Attention, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                scale: float = 1.0, **kwargs):
        batch = hidden_states.shape[0]
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states
        end = encoder_hidden_states.shape[1] - self.num_tokens
        text, image = encoder_hidden_states[:, :end, :], encoder_hidden_states[:, end:, :]
        q = _split_heads(attn.to_q(hidden_states), attn.heads)
        k = _split_heads(attn.to_k(text), attn.heads)
        v = _split_heads(attn.to_v(text), attn.heads)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(batch, -1, attn.inner_dim).to(q.dtype)
        k_ip = _split_heads(self.to_k_ip(image), attn.heads)
        v_ip = _split_heads(self.to_v_ip(image), attn.heads)
        o_ip = F.scaled_dot_product_attention(q, k_ip, v_ip, attn_mask=None, dropout_p=0.0, is_causal=False)
        o_ip = o_ip.transpose(1, 2).reshape(batch, -1, attn.inner_dim).to(q.dtype)
        o = o + self.scale * o_ip
        o = attn.to_out[0](o)
        o = attn.to_out[1](o)
        if attn.residual_connection:
            o = o + hidden_states
        return o / attn.rescale_output_factor
