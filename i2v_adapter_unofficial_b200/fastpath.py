"""Module-level fast path: the normalisation prologues, layout changes and residual epilogues *around* the attention
operators, which the AttnProcessor boundary cannot reach ("the only way to absorb LN prologue +
residual epilogue").  ``install(unet, fast_path=True)`` swaps the ``forward`` of four module types — parameters,
state-dict keys and call signatures untouched — for versions that follow the same reference lines but run the
bandwidth-bound steps in ``libi2v_attn_b200.so``:

===============================  ===================================================  ================================
module (reference)               replaced passes                                      kernels
===============================  ===================================================  ================================
I2VAdapterTransformer2DModel     GroupNorm, 1x1-conv proj_in + (BF,C,h,w)->(BF,S,C)   i2v_gn_stats,
 src/modules/i2v_adapter.py      copy, (BF,S,C)->(BF,C,h,w) copy + proj_out +         i2v_gn_apply_transpose,
 :214-234, :298-314              residual                                             i2v_untranspose_residual
I2VAdapterTransformerBlock       norm1/2/3, GEGLU                                     i2v_layernorm_fwd, i2v_geglu_fwd
 :420-565
TransformerTemporalModel         GroupNorm over (C/G,F,h,w), both permute copies,     same three layout kernels with
 (diffusers; Appendix A4)        residual add                                         fg = num_frames
temporal BasicTransformerBlock   norm1/2/3 with ``+ pos_embed`` folded in, GEGLU      i2v_layernorm_fwd (pe), geglu
===============================  ===================================================  ================================

The 1x1 convolutions ``proj_in`` / ``proj_out`` of the spatial transformer are evaluated as token-major GEMMs on the
same weights (``weight.view(out, in)``): a 1x1 convolution *is* that GEMM, and token-major output is what the blocks
consume, so both layout copies disappear.  The attention calls still go through ``self.attn1(...)`` etc., i.e. through
whatever processors are installed.

Every fast forward checks its preconditions (CUDA, bf16, inference, supported shape, no mask / ada-norm options) and
otherwise calls the module's original forward — the stock PyTorch GPU path of the host model, not a CPU fallback.
"""
from __future__ import annotations

import os
import functools
from typing import Any, Callable, Dict, List, Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .processors import _CACHE_EPOCH


#: how often each swapped forward handed the call back to the module's original (stock PyTorch) forward because a
#: precondition failed -- `bench.py` asserts this stays empty for the measured configuration
# Transformer2DModel tail (proj_out -> + residual, src/modules/i2v_adapter.py:298-314) on the library's token GEMM with the
# residual in its epilogue; False keeps cuBLAS + a separate elementwise add (developer A / B switch)
FUSE_PROJ_OUT_RESIDUAL = os.environ.get("I2V_FUSE_PROJ_OUT", "1") != "0"
FALLBACKS: Dict[str, int] = {}


def _fallback(kind: str, original: Callable, *args, **kwargs):
    FALLBACKS[kind] = FALLBACKS.get(kind, 0) + 1
    return original(*args, **kwargs)


def fallback_counts() -> Dict[str, int]:
    return dict(FALLBACKS)


def reset_fallback_counts() -> None:
    FALLBACKS.clear()


def _fast_ok(x: torch.Tensor, module: nn.Module) -> bool:
    return x.is_cuda and x.dtype == torch.bfloat16 and not module.training and not torch.is_grad_enabled()


def _norm_dtype_ok(norm: nn.Module, x: torch.Tensor) -> bool:
    """The norm kernels read bf16 parameters (autocast setups keep them in fp32: those go to the stock forward)."""
    w = getattr(norm, "weight", None)
    return w is None or w.dtype == x.dtype


def _ln(x: torch.Tensor, norm: nn.LayerNorm, pe: Optional[torch.Tensor] = None,
        pre: Optional[torch.Tensor] = None) -> torch.Tensor:
    if norm.weight is None or norm.bias is None or norm.weight.dtype != x.dtype:
        if pre is not None:
            x = x + pre
        y = F.layer_norm(x, norm.normalized_shape, norm.weight, norm.bias, norm.eps)
        return y if pe is None else y + pe
    return ops.layernorm(x.contiguous(), norm.weight, norm.bias, norm.eps, pe, pre)


FUSED_FF_GEGLU = True   # feed-forward input projection + GEGLU in one tcgen05 GEMM (i2v_ff_geglu_fwd) where its shapes fit


def _geglu_proj(x: torch.Tensor, proj: nn.Linear, ones_column: bool) -> torch.Tensor:
    """``hidden * gelu(gate)`` of ``proj(x).chunk(2, -1)`` (diffusers GEGLU; src/modules/i2v_adapter.py:535-561)."""
    if FUSED_FF_GEGLU and ops.ff_geglu_supported(x, proj.weight) and x.is_contiguous() and \
            (proj.bias is None or proj.bias.is_contiguous()):
        return ops.ff_geglu(x, proj.weight, proj.bias, ones_column=ones_column)
    return ops.geglu(F.linear(x, proj.weight, proj.bias), ones_column=ones_column)


def _ff(ff: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """FeedForward with GEGLU: net.0.proj (Linear dim -> 8 dim), fused GEGLU, net.2 (Linear 4 dim -> dim)."""
    proj = ff.net[0].proj
    h = _geglu_proj(x, proj, False)
    out = ff.net[2]
    return F.linear(h, out.weight, out.bias)


class _Deferred:
    """Per-block cache of the deferred-bias vectors and of the feed-forward's augmented output weight.

    With the residual add of every sub-layer riding in its output GEMM (``processors._out_proj``), the output biases
    b1 (attn1 [+ i2v_adapter]), b2 (attn2), b3 (ff) are constants that still have to reach the residual stream:
    LayerNorm 2 reads ``hidden + b1``, LayerNorm 3 reads ``hidden + b1 + b2`` (``pre`` of the LayerNorm kernel) and the
    last GEMM of the block gets ``[W_ff2 | b1 + b2 + b3 | 0]`` against a GEGLU output with a ones column, so the block
    returns the exact residual stream with no separate add or bias pass."""

    def __init__(self):
        self.key = None
        self.val = None

    @staticmethod
    def _sig(t):
        return None if t is None else (t.data_ptr(), t._version)

    def get1(self, b1, like: torch.Tensor):
        key = (_CACHE_EPOCH[0], self._sig(b1), like.dtype, like.device)
        if key != getattr(self, "key1", None):
            self.val1 = b1.to(like.dtype).contiguous()
            self.key1 = key
        return self.val1

    def get(self, b1, b2, ff_out: nn.Linear, like: torch.Tensor):
        key = (_CACHE_EPOCH[0], self._sig(b1), self._sig(b2), self._sig(ff_out.weight), self._sig(ff_out.bias),
               like.dtype, like.device)
        if key != self.key:
            C = ff_out.weight.shape[0]
            zero = torch.zeros(C, dtype=torch.float32, device=like.device)
            p1 = zero if b1 is None else b1.float()
            p2 = p1 + (zero if b2 is None else b2.float())
            p3 = p2 + (zero if ff_out.bias is None else ff_out.bias.float())
            w_aug = torch.cat([ff_out.weight.float(), p3[:, None], ff_out.weight.new_zeros(C, 7).float()], dim=1)
            self.val = (p1.to(like.dtype).contiguous(), p2.to(like.dtype).contiguous(),
                        w_aug.to(like.dtype).contiguous())
            self.key = key
        return self.val


def _is_geglu_ff(ff: nn.Module) -> bool:
    net = getattr(ff, "net", None)
    return (net is not None and len(net) == 3 and type(net[0]).__name__ == "GEGLU" and isinstance(net[2], nn.Linear)
            and getattr(net[1], "p", 0.0) == 0.0)


class _PE:
    """bf16 copy of the sinusoidal table rows [0, F) (cached per block; rebuilt if the buffer or F changes)."""

    def __init__(self):
        self.key = None
        self.val = None

    def get(self, pos_embed: nn.Module, frames: int, like: torch.Tensor) -> torch.Tensor:
        pe = pos_embed.pe
        key = (_CACHE_EPOCH[0], pe.data_ptr(), pe._version, frames, like.device, like.dtype)
        if key != self.key:
            self.val = pe[0, :frames].to(device=like.device, dtype=like.dtype).contiguous()
            self.key = key
        return self.val


# ----------------------------------------------------------------------------------------------------------------
# transformer blocks
# ----------------------------------------------------------------------------------------------------------------
def _fused_residual_ok(self) -> bool:
    """All sub-layers of the block can take their residual inside the output GEMM (our processors, plain GEGLU ff)."""
    from .processors import B200AttnProcessor, B200IPAdapterAttnProcessor, B200TemporalAttnProcessor

    def ours(attn):
        proc = attn.get_processor() if hasattr(attn, "get_processor") else None
        return isinstance(proc, (B200AttnProcessor, B200IPAdapterAttnProcessor, B200TemporalAttnProcessor))

    return (self.attn2 is not None and ours(self.attn1) and ours(self.attn2)
            and self.ff.net[2].weight.shape[1] % 8 == 0 and not self.only_cross_attention)


OWNED_INPUT_FLAG = "_b200_owned_input"   # set by our wrapper forwards right before calling a block: the tokens tensor
#                                          is theirs alone, the block may accumulate its first residual into it in place


def _block_forward_fast(self, hidden_states, encoder_hidden_states, enable_cross_frame_attn, num_frames, kw, pe_cache,
                        deferred=None, owned_input=False):
    """Shared body of I2VAdapterTransformerBlock.forward (src/modules/i2v_adapter.py:420-565) and the diffusers
    BasicTransformerBlock.forward it extends (layer_norm flavour).

    With ``deferred`` (a ``_Deferred`` cache; our processors installed) the three residual adds (:501, :533, :561)
    ride in the output GEMMs of attn1 / attn2 / ff and their biases are deferred as described in ``_Deferred``."""
    from .processors import RESIDUAL_INPLACE_KW, RESIDUAL_KW

    pe = None
    if self.pos_embed is not None:
        pe = pe_cache.get(self.pos_embed, hidden_states.shape[1], hidden_states)
    fuse = deferred is not None and hidden_states.is_contiguous()
    norm_h = _ln(hidden_states, self.norm1, pe)                                                   # :445, :458-459
    kw1 = dict(kw, **{RESIDUAL_KW: hidden_states, RESIDUAL_INPLACE_KW: bool(owned_input)}) if fuse else kw
    attn_output = self.attn1(norm_h, encoder_hidden_states=encoder_hidden_states if self.only_cross_attention else None,
                             attention_mask=None, **kw1)                                          # :468-473
    proc1 = self.attn1.get_processor() if hasattr(self.attn1, "get_processor") else None
    fused1 = fuse and getattr(proc1, "fused_residual", False)
    b1 = getattr(proc1, "deferred_bias", None) if fused1 else None
    if enable_cross_frame_attn:
        batch_size = hidden_states.shape[0]
        if num_frames is None:
            raise ValueError("`num_frames` must be provided when `enable_cross_frame_attn` is True.")
        if batch_size % num_frames != 0:
            raise ValueError(f"Batch size {batch_size} must be divisible by the number of frames {num_frames}.")
        state = getattr(proc1, "state", None)
        if state is not None and state.cross_done:
            state.cross_done = False          # attn1's processor already returned self + cross-frame (one launch)
        else:
            first = norm_h[0:batch_size:num_frames].repeat_interleave(num_frames, dim=0)          # :484-485
            attn_output = attn_output + self.i2v_adapter(norm_h, encoder_hidden_states=first, attention_mask=None,
                                                         **kw)                                    # :487-494
    hidden_states = attn_output if fused1 else attn_output + hidden_states                        # :501
    b2 = None
    if self.attn2 is not None:
        pre1 = None
        if fused1 and b1 is not None:
            pre1 = deferred.get1(b1, hidden_states)
        norm_h = _ln(hidden_states, self.norm2, pe, pre1)                                         # :514, :524-525
        # hidden_states is an intermediate of this block from here on: the GEMM may accumulate into it in place
        kw2 = dict(kw, **{RESIDUAL_KW: hidden_states, RESIDUAL_INPLACE_KW: True}) if fuse else kw
        attn_output = self.attn2(norm_h, encoder_hidden_states=encoder_hidden_states, attention_mask=None, **kw2)
        proc2 = self.attn2.get_processor() if hasattr(self.attn2, "get_processor") else None
        fused2 = fuse and getattr(proc2, "fused_residual", False)
        b2 = getattr(proc2, "deferred_bias", None) if fused2 else None
        hidden_states = attn_output if fused2 else attn_output + hidden_states                    # :533
    if fuse:
        p1, p2, w_aug = deferred.get(b1, b2, self.ff.net[2], hidden_states)
        pending = b1 is not None or b2 is not None
        norm_h = _ln(hidden_states, self.norm3, None, p2 if pending else None)                    # :539
        proj = self.ff.net[0].proj
        h = _geglu_proj(norm_h, proj, True)                                                       # [.., 4C + 8]
        C = hidden_states.shape[-1]
        if self.attn2 is not None:   # hidden_states is this block's own tensor (see above): accumulate in place
            hidden_states.view(-1, C).addmm_(h.view(-1, h.shape[-1]), w_aug.t())
            return hidden_states
        return torch.addmm(hidden_states.view(-1, C), h.view(-1, h.shape[-1]), w_aug.t()).view(hidden_states.shape)
    hidden_states = _ff(self.ff, _ln(hidden_states, self.norm3)) + hidden_states                  # :539-561
    return hidden_states


def _make_block_forward(module: nn.Module, original: Callable, is_i2v: bool):
    pe_cache = _PE()
    deferred = _Deferred()

    @functools.wraps(original)  # keeps the reference signature visible to inspect.signature (install()'s hooks bind it)
    def forward(hidden_states, *args, **kwargs):
        names = (("enable_cross_frame_attn", "num_frames", "attention_mask", "encoder_hidden_states",
                  "encoder_attention_mask", "timestep", "cross_attention_kwargs", "class_labels", "added_cond_kwargs")
                 if is_i2v else
                 ("attention_mask", "encoder_hidden_states", "encoder_attention_mask", "timestep",
                  "cross_attention_kwargs", "class_labels"))
        bound: Dict[str, Any] = dict(zip(names, args))
        bound.update(kwargs)
        owned = module.__dict__.pop(OWNED_INPUT_FLAG, False)
        simple = (_fast_ok(hidden_states, module) and hidden_states.dim() == 3 and hidden_states.shape[-1] % 8 == 0
                  and bound.get("attention_mask") is None and bound.get("encoder_attention_mask") is None
                  and len(args) <= len(names) and set(kwargs) <= set(names)
                  and isinstance(module.norm1, nn.LayerNorm) and isinstance(module.norm3, nn.LayerNorm)
                  and _is_geglu_ff(module.ff) and getattr(module, "_chunk_size", None) is None)
        if not simple:
            return _fallback("transformer_block", original, hidden_states, *args, **kwargs)
        kw = dict(bound.get("cross_attention_kwargs") or {})
        kw.pop("gligen", None)
        return _block_forward_fast(module, hidden_states, bound.get("encoder_hidden_states"),
                                   bool(bound.get("enable_cross_frame_attn", False)) if is_i2v else False,
                                   bound.get("num_frames") if is_i2v else None, kw, pe_cache,
                                   deferred if _fused_residual_ok(module) else None, owned)

    return forward


# ----------------------------------------------------------------------------------------------------------------
# spatial transformer wrapper
# ----------------------------------------------------------------------------------------------------------------
def _layout_ok(x: torch.Tensor, norm: nn.GroupNorm) -> bool:
    if x.dim() != 4 or not x.is_contiguous() or not norm.affine:
        return False
    _, C, h, w = x.shape
    G = norm.num_groups
    return (C % 64 == 0 and (h * w) % 8 == 0 and C % G == 0 and (C // G) % 2 == 0 and x.shape[0] <= 65535
            and _norm_dtype_ok(norm, x))


def _nhwc_ok(x: torch.Tensor, norm: nn.GroupNorm) -> bool:
    """Channels-last activation the NHWC GroupNorm kernels take."""
    if not (ops.is_channels_last(x) and isinstance(norm, nn.GroupNorm) and norm.affine):
        return False
    C, G = x.shape[1], norm.num_groups
    return (C % 8 == 0 and C % G == 0 and C <= 4096 and x.shape[0] <= 65535 and not x.is_contiguous()
            and _norm_dtype_ok(norm, x))


def _make_transformer2d_forward(module: nn.Module, original: Callable):
    @functools.wraps(original)
    def forward(hidden_states, *args, **kwargs):
        names = ("enable_cross_frame_attn", "encoder_hidden_states", "num_frames", "timestep", "added_cond_kwargs",
                 "class_labels", "cross_attention_kwargs", "attention_mask", "encoder_attention_mask", "return_dict")
        bound: Dict[str, Any] = dict(zip(names, args))
        bound.update(kwargs)
        conv_proj = isinstance(module.proj_in, nn.Conv2d) and module.proj_in.kernel_size == (1, 1)
        lin_proj = isinstance(module.proj_in, nn.Linear)
        simple = (_fast_ok(hidden_states, module) and _layout_ok(hidden_states, module.norm)
                  and bound.get("attention_mask") is None and bound.get("encoder_attention_mask") is None
                  and len(args) <= len(names) and set(kwargs) <= set(names) and (conv_proj or lin_proj)
                  and module.proj_out.weight.shape[0] == hidden_states.shape[1])
        nhwc = (_fast_ok(hidden_states, module) and _nhwc_ok(hidden_states, module.norm)
                and bound.get("attention_mask") is None and bound.get("encoder_attention_mask") is None
                and len(args) <= len(names) and set(kwargs) <= set(names) and (conv_proj or lin_proj)
                and module.proj_out.weight.shape[0] == hidden_states.shape[1])
        if not simple and not nhwc:
            return _fallback("transformer_2d", original, hidden_states, *args, **kwargs)
        N, C, h, w = hidden_states.shape
        norm = module.norm
        # :218-234  GroupNorm(32, eps 1e-6) -> proj_in -> (BF, S, inner), without the NCHW intermediate
        if nhwc:   # channels-last activation == token-major already: no transpose either way
            tokens = ops.group_norm_nhwc(hidden_states, norm.weight, norm.bias, norm.num_groups, norm.eps, 1)
            tokens = tokens.permute(0, 2, 3, 1).reshape(N, h * w, C)
        else:
            tokens = ops.group_norm_tokens(hidden_states, norm.weight, norm.bias, norm.num_groups, norm.eps, 1)
        inner = module.proj_in.weight.shape[0]
        tokens = F.linear(tokens, module.proj_in.weight.reshape(inner, C), module.proj_in.bias)
        for block in module.transformer_blocks:                                                  # :244-296
            block.__dict__[OWNED_INPUT_FLAG] = True   # `tokens` is a fresh GEMM output / the previous block's result
            tokens = block(tokens, enable_cross_frame_attn=bound.get("enable_cross_frame_attn", False),
                           num_frames=bound.get("num_frames"), attention_mask=None,
                           encoder_hidden_states=bound.get("encoder_hidden_states"), encoder_attention_mask=None,
                           timestep=bound.get("timestep"), cross_attention_kwargs=bound.get("cross_attention_kwargs"),
                           class_labels=bound.get("class_labels"))
        # :298-314  proj_out -> (BF, C, h, w) -> + residual
        w_po = module.proj_out.weight.reshape(C, inner)
        if nhwc and FUSE_PROJ_OUT_RESIDUAL and ops.linear_supported(tokens, w_po):
            # proj_out + bias + residual in one launch of the library's token GEMM (the residual tile rides in through TMA and
            # is added after the product is rounded to bf16, as the reference's separate add): one full read-read-write
            # pass less than `linear` followed by `+=`
            tokens = ops.linear(tokens, w_po, module.proj_out.bias,
                                residual=hidden_states.permute(0, 2, 3, 1).reshape(N, h * w, C))
            output = tokens.view(N, h, w, C).permute(0, 3, 1, 2)   # a channels-last (N, C, h, w) tensor, no copy
        elif nhwc:
            tokens = F.linear(tokens, w_po, module.proj_out.bias)
            tokens += hidden_states.permute(0, 2, 3, 1).reshape(N, h * w, C)
            output = tokens.view(N, h, w, C).permute(0, 3, 1, 2)
        else:
            tokens = F.linear(tokens, w_po, module.proj_out.bias)
            output = ops.tokens_to_nchw_residual(tokens.contiguous(), hidden_states, 1)
        if not bound.get("return_dict", True):
            return (output,)
        return _Sample(output)

    return forward


# ----------------------------------------------------------------------------------------------------------------
# motion module
# ----------------------------------------------------------------------------------------------------------------
def _make_temporal_forward(module: nn.Module, original: Callable):
    @functools.wraps(original)
    def forward(hidden_states, encoder_hidden_states=None, timestep=None, class_labels=None, num_frames: int = 1,
                cross_attention_kwargs=None, return_dict: bool = True):
        simple = (_fast_ok(hidden_states, module) and _layout_ok(hidden_states, module.norm)
                  and hidden_states.shape[0] % max(num_frames, 1) == 0 and isinstance(module.proj_in, nn.Linear))
        nhwc = (_fast_ok(hidden_states, module) and _nhwc_ok(hidden_states, module.norm)
                and hidden_states.shape[0] % max(num_frames, 1) == 0 and isinstance(module.proj_in, nn.Linear))
        if not simple and not nhwc:
            return _fallback("temporal_model", original, hidden_states, encoder_hidden_states=encoder_hidden_states,
                             timestep=timestep, class_labels=class_labels, num_frames=num_frames,
                             cross_attention_kwargs=cross_attention_kwargs, return_dict=return_dict)
        norm = module.norm
        # GroupNorm over (C/G, F, h, w) per video, then (BF, C, h, w) -> (B*S, F, C) in the same pass
        if nhwc:
            x = ops.group_norm_nhwc(hidden_states, norm.weight, norm.bias, norm.num_groups, norm.eps, num_frames,
                                    to_positions=True)
        else:
            x = ops.group_norm_tokens(hidden_states, norm.weight, norm.bias, norm.num_groups, norm.eps, num_frames)
        x = module.proj_in(x)
        for block in module.transformer_blocks:
            block.__dict__[OWNED_INPUT_FLAG] = True
            x = block(x, encoder_hidden_states=encoder_hidden_states, timestep=timestep,
                      cross_attention_kwargs=cross_attention_kwargs, class_labels=class_labels)
        x = module.proj_out(x)
        if nhwc:
            output = ops.positions_to_nhwc_residual(x.contiguous(), hidden_states, num_frames)
        else:
            output = ops.tokens_to_nchw_residual(x.contiguous(), hidden_states, num_frames)
        if not return_dict:
            return (output,)
        return _Sample(output)

    return forward


# ----------------------------------------------------------------------------------------------------------------
# ResnetBlock2D in channels-last: only its two GroupNorm + SiLU prologues (and the time-embedding add) move into the
# library; the convolutions stay PyTorch / cuDNN calls (north_star), now fed the layout cuDNN computes in.
# ----------------------------------------------------------------------------------------------------------------
def _is_silu(fn) -> bool:
    return isinstance(fn, nn.SiLU) or fn is F.silu


def _resnet_simple(module: nn.Module, x: torch.Tensor, temb) -> bool:
    """The ResnetBlock2D configurations the channels-last forward below covers (everything SD1.5 uses)."""
    return (_fast_ok(x, module) and temb is not None and _nhwc_ok(x, module.norm1)
            and isinstance(module.norm2, nn.GroupNorm) and module.norm2.affine and _norm_dtype_ok(module.norm2, x)
            and _is_silu(getattr(module, "nonlinearity", None))
            and getattr(module, "time_emb_proj", None) is not None
            and getattr(module, "time_embedding_norm", "default") in (None, "default")
            and not getattr(module, "up", False) and not getattr(module, "down", False)
            and getattr(module, "upsample", None) is None and getattr(module, "downsample", None) is None
            and getattr(module.dropout, "p", 0.0) == 0.0 and module.conv1.out_channels % 8 == 0
            and isinstance(module.conv1, nn.Conv2d) and isinstance(module.conv2, nn.Conv2d)
            and module.conv1.padding_mode == "zeros" and module.conv2.padding_mode == "zeros"
            and module.conv1.out_channels <= 4096)


def _pair_ok(module: nn.Module, x: torch.Tensor, x2: torch.Tensor, temb) -> bool:
    """(hidden, skip) pair an up block would concatenate: the resnet can read the two tensors in place when its
    shortcut is a 1x1 convolution (it always is there: the concatenation changes the channel count)."""
    cs = getattr(module, "conv_shortcut", None)
    return (_resnet_simple(module, x, temb) and x2.is_cuda and x2.dtype == x.dtype and ops.is_channels_last(x2)
            and x2.shape[0] == x.shape[0] and x2.shape[2:] == x.shape[2:] and x2.shape[1] % 8 == 0
            and x.shape[1] + x2.shape[1] <= 4096 and (x.shape[1] + x2.shape[1]) % module.norm1.num_groups == 0
            and isinstance(cs, nn.Conv2d) and cs.kernel_size == (1, 1) and cs.stride == (1, 1) and cs.padding == (0, 0)
            and cs.groups == 1 and cs.in_channels == x.shape[1] + x2.shape[1])


def _summed_bias(module: nn.Module, slot: str, a: Optional[torch.Tensor], b: torch.Tensor) -> torch.Tensor:
    """``a + b`` of two per-channel bias vectors, cached on the module (keyed on both parameters' storage and version)."""
    if a is None:
        return b
    cache = module.__dict__.get(slot)
    if cache is None:
        from .processors import _PackedWeights

        cache = module.__dict__[slot] = _PackedWeights()
    return cache.get([a, b], lambda: a + b)


def _resnet_forward_nhwc(module: nn.Module, x: torch.Tensor, temb: torch.Tensor,
                         x2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ResnetBlock2D on channels-last activations; with ``x2`` the input is ``torch.cat([x, x2], 1)`` read in place."""
    n1, n2 = module.norm1, module.norm2
    c1, c2 = module.conv1, module.conv2
    h = ops.group_norm_nhwc(x, n1.weight, n1.bias, n1.num_groups, n1.eps, 1, silu=True, x2=x2)
    # cuDNN applies a convolution bias as a separate broadcast-add pass; both biases of the block are per-channel
    # constants that the next bandwidth kernel can carry instead: conv1's joins the time-embedding term of norm2,
    # conv2's the residual add
    h = F.conv2d(h, c1.weight, None, c1.stride, c1.padding, c1.dilation, c1.groups)
    # SiLU(temb) is the same tensor for every ResnetBlock2D of a forward: computed once and kept on `temb` itself; conv1's
    # bias joins the projection's own bias (cached sum) instead of a separate broadcast add
    cached = getattr(temb, "_b200_silu", None)
    if cached is not None and cached[0] == temb._version:
        st = cached[1]
    else:
        st = F.silu(temb)
        try:
            temb._b200_silu = (temb._version, st)   # (an in-place write to temb bumps its version: recomputed then)
        except Exception:  # noqa: BLE001  (a tensor subclass that refuses attributes: just recompute next time)
            pass
    tp = module.time_emb_proj
    tb = tp.bias
    if c1.bias is not None:
        tb = _summed_bias(module, "_b200_temb_bias", tp.bias, c1.bias)
    t = F.linear(st, tp.weight, tb)
    if not ops.is_channels_last(h):
        h = h.contiguous(memory_format=torch.channels_last)
    h = ops.group_norm_nhwc(h, n2.weight, n2.bias, n2.num_groups, n2.eps, 1, silu=True, add=t)
    h = F.conv2d(h, c2.weight, None, c2.stride, c2.padding, c2.dilation, c2.groups)
    cs = getattr(module, "conv_shortcut", None)
    bias = c2.bias
    if x2 is not None:
        # 1x1 shortcut convolution of the concatenation = two token GEMMs on the two sources (channels-last rows)
        N, Ca, hh, ww = x.shape
        Cb = x2.shape[1]
        wmat = cs.weight.reshape(cs.out_channels, Ca + Cb)
        xa = x.permute(0, 2, 3, 1).reshape(N * hh * ww, Ca)
        xb = x2.permute(0, 2, 3, 1).reshape(N * hh * ww, Cb)
        sc = torch.mm(xa, wmat[:, :Ca].t())
        sc.addmm_(xb, wmat[:, Ca:].t())
        x = sc.view(N, hh, ww, cs.out_channels).permute(0, 3, 1, 2)
        if cs.bias is not None:
            bias = cs.bias if bias is None else _summed_bias(module, "_b200_out_bias", bias, cs.bias)
    elif cs is not None:
        if isinstance(cs, nn.Conv2d) and cs.padding_mode == "zeros" and ops.is_channels_last(h):
            # the shortcut convolution's bias rides in the same residual pass as conv2's
            x = F.conv2d(x, cs.weight, None, cs.stride, cs.padding, cs.dilation, cs.groups)
            if cs.bias is not None:
                bias = cs.bias if bias is None else _summed_bias(module, "_b200_out_bias", bias, cs.bias)
        else:
            x = cs(x)
    if ops.is_channels_last(x) and ops.is_channels_last(h):
        out = ops.nhwc_add(x, h, bias)             # x + h + bias[c] in one pass
    else:
        out = x + (h if bias is None else h + bias[None, :, None, None])
    osf = getattr(module, "output_scale_factor", 1.0)
    return out if osf == 1.0 else out / osf


def _make_resnet_forward(module: nn.Module, original: Callable):
    @functools.wraps(original)
    def forward(x, temb=None, *args, **kwargs):
        if not _resnet_simple(module, x, temb):
            return _fallback("resnet", original, x, temb, *args, **kwargs)
        return _resnet_forward_nhwc(module, x, temb)

    return forward


def _make_upblock_forward(module: nn.Module, original: Callable, cross: bool):
    """CrossFrameAttnUpBlockMotion (reference unet_motion_cross_frame_attn.py:439-529) / UpBlockMotion: the same
    sequencing, but ``torch.cat([hidden_states, res_hidden_states], dim=1)`` (:457) is not materialised -- the resnet's
    GroupNorm and shortcut convolution read the two tensors in place."""
    @functools.wraps(original)
    def forward(hidden_states, res_hidden_states_tuple, temb=None, *args, **kwargs):
        names = (("enable_cross_frame_attn", "encoder_hidden_states", "cross_attention_kwargs", "upsample_size",
                  "attention_mask", "encoder_attention_mask", "num_frames") if cross
                 else ("upsample_size", "scale", "num_frames"))
        # FreeU (enable_freeu sets s1/s2/b1/b2 on the up blocks) rescales hidden / skip before the concatenation and
        # training / gradient checkpointing need the stock loop: those calls keep the original forward
        freeu = all(getattr(module, a, None) for a in ("s1", "s2", "b1", "b2"))
        if (len(args) > len(names) or not set(kwargs) <= set(names) or not _fast_ok(hidden_states, module) or freeu
                or getattr(module, "gradient_checkpointing", False) and module.training):
            return _fallback("up_block", original, hidden_states, res_hidden_states_tuple, temb, *args, **kwargs)
        kw = dict(zip(names, args))
        kw.update(kwargs)
        num_frames = kw.get("num_frames", 1)
        layers = (zip(module.resnets, module.attentions, module.motion_modules) if cross
                  else zip(module.resnets, [None] * len(module.resnets), module.motion_modules))
        for resnet, attn, motion_module in layers:
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            if "forward" in resnet.__dict__ and _pair_ok(resnet, hidden_states, res, temb):
                hidden_states = _resnet_forward_nhwc(resnet, hidden_states, temb, res)
            else:
                FALLBACKS["up_block_concat"] = FALLBACKS.get("up_block_concat", 0) + 1
                hidden_states = resnet(torch.cat([hidden_states, res], dim=1), temb)
            if attn is not None:
                hidden_states = attn(hidden_states, enable_cross_frame_attn=kw.get("enable_cross_frame_attn", False),
                                     num_frames=num_frames, encoder_hidden_states=kw.get("encoder_hidden_states"),
                                     cross_attention_kwargs=kw.get("cross_attention_kwargs"),
                                     attention_mask=kw.get("attention_mask"),
                                     encoder_attention_mask=kw.get("encoder_attention_mask"),
                                     return_dict=False)[0]                                       # reference :499-508
            hidden_states = motion_module(hidden_states, num_frames=num_frames)[0]
        if module.upsamplers is not None:
            for u in module.upsamplers:
                hidden_states = u(hidden_states, kw.get("upsample_size"))
        return hidden_states

    return forward


def _conv_bias_nhwc(conv: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """``conv(x)`` with the bias applied by one full-bandwidth in-place pass (channels-last bf16)."""
    y = F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
    if conv.bias is None:
        return y
    if ops.is_channels_last(y) and y.shape[1] % 8 == 0:
        return ops.nhwc_bias_add_(y, conv.bias)
    return y + conv.bias[None, :, None, None]


def _sampler_ok(module: nn.Module, x: torch.Tensor) -> bool:
    conv = getattr(module, "conv", None)
    # Downsample2D(padding=0) pads (0, 1, 0, 1) before its convolution (diffusers): not the plain conv call below
    return (_fast_ok(x, module) and isinstance(conv, nn.Conv2d) and conv.padding_mode == "zeros" and x.dim() == 4
            and ops.is_channels_last(x) and x.shape[1] % 8 == 0
            and (type(module).__name__ != "Downsample2D" or conv.padding not in ((0, 0), 0)))


def _make_upsample_forward(module: nn.Module, original: Callable):
    """Upsample2D: nearest 2x interpolation as one read / four writes per vector, then the convolution."""
    @functools.wraps(original)
    def forward(x, output_size=None, *args, **kwargs):
        if output_size is not None or not _sampler_ok(module, x):
            return _fallback("upsample", original, x, output_size, *args, **kwargs)
        return _conv_bias_nhwc(module.conv, ops.upsample2x_nhwc(x))

    return forward


def _make_downsample_forward(module: nn.Module, original: Callable):
    @functools.wraps(original)
    def forward(x, *args, **kwargs):
        if not _sampler_ok(module, x):
            return _fallback("downsample", original, x, *args, **kwargs)
        return _conv_bias_nhwc(module.conv, x)

    return forward


def _make_out_norm_forwards(norm: nn.GroupNorm, act: nn.Module, norm_forward: Callable, act_forward: Callable):
    """The UNet's tail ``conv_act(conv_norm_out(sample))`` (reference unet_motion_cross_frame_attn.py:1437-1439) as one
    GroupNorm + SiLU pass of the channels-last kernels; the activation module becomes the identity for that call."""
    state = {"fused": False}

    @functools.wraps(norm_forward)
    def norm_fwd(x):
        if _fast_ok(x, norm) and norm.affine and x.dim() == 4 and _nhwc_ok(x, norm):
            state["fused"] = True
            return ops.group_norm_nhwc(x, norm.weight, norm.bias, norm.num_groups, norm.eps, 1, silu=True)
        state["fused"] = False
        return _fallback("conv_norm_out", norm_forward, x)

    @functools.wraps(act_forward)
    def act_fwd(x):
        if state["fused"]:
            state["fused"] = False
            return x
        return act_forward(x)

    return norm_fwd, act_fwd


class _Sample:
    """``.sample`` attribute and tuple-style ``[0]`` (what the call sites use, e.g. unet_motion_cross_frame_attn.py:326)."""

    def __init__(self, sample):
        self.sample = sample

    def __getitem__(self, idx):
        return (self.sample,)[idx]


# ----------------------------------------------------------------------------------------------------------------
# installation
# ----------------------------------------------------------------------------------------------------------------
def install_fast_forwards(root: nn.Module) -> List[Callable[[], None]]:
    """Swap the forwards of every supported module under ``root``; returns the undo callbacks."""
    undo: List[Callable[[], None]] = []

    def patch(module: nn.Module, new_forward: Callable):
        had_own = "forward" in module.__dict__
        previous = module.__dict__.get("forward")
        module.forward = new_forward

        def restore(m=module, had=had_own, prev=previous):
            if had:
                m.forward = prev
            else:
                m.__dict__.pop("forward", None)

        undo.append(restore)

    out_norm, out_act = getattr(root, "conv_norm_out", None), getattr(root, "conv_act", None)
    if isinstance(out_norm, nn.GroupNorm) and _is_silu(out_act) and isinstance(out_act, nn.Module):
        norm_fwd, act_fwd = _make_out_norm_forwards(out_norm, out_act, out_norm.forward, out_act.forward)
        patch(out_norm, norm_fwd)
        patch(out_act, act_fwd)

    for module in list(root.modules()):
        name = type(module).__name__
        if name == "I2VAdapterTransformerBlock":
            patch(module, _make_block_forward(module, module.forward, True))
        elif name == "BasicTransformerBlock":
            patch(module, _make_block_forward(module, module.forward, False))
        elif name == "I2VAdapterTransformer2DModel":
            patch(module, _make_transformer2d_forward(module, module.forward))
        elif name == "TransformerTemporalModel":
            patch(module, _make_temporal_forward(module, module.forward))
        elif name == "ResnetBlock2D":
            patch(module, _make_resnet_forward(module, module.forward))
        elif name == "CrossFrameAttnUpBlockMotion" and hasattr(module, "motion_modules"):
            patch(module, _make_upblock_forward(module, module.forward, True))
        elif name == "UpBlockMotion" and hasattr(module, "motion_modules"):
            patch(module, _make_upblock_forward(module, module.forward, False))
        elif name == "Upsample2D":
            patch(module, _make_upsample_forward(module, module.forward))
        elif name == "Downsample2D":
            patch(module, _make_downsample_forward(module, module.forward))
    return undo
