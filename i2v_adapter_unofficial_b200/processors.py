"""B200 attention processors — drop-in replacements for the diffusers processors the reference installs
(``AttnProcessor2_0`` / ``IPAdapterAttnProcessor2_0``, ``/root/reference/src/models/unet_motion_cross_frame_attn.py:
1259-1272``), with the same calling convention::

    processor(attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0, **kw)

and the same return value (attention output *after* ``attn.to_out``, before the block's residual add).  Weights stay
owned by the ``Attention`` modules and are read on every call, so ``load_i2v_adapter`` / ``load_motion_modules`` /
``.to()`` keep working; packed weight copies are a cache keyed on the parameters' storage and version.

Projections (``to_q/k/v/out``) stay on cuBLAS through ``F.linear`` but are packed so each block issues
one input GEMM and one output GEMM; the attention arithmetic runs in ``libi2v_attn_b200.so``.

``install(unet)`` walks ``unet.attn_processors`` and swaps every processor:

=====================================  =======================================================================
attention                              processor
=====================================  =======================================================================
``...transformer_blocks.N.attn1``      ``B200SpatialAttnProcessor``   (self-attention; fused with the block's
                                       ``i2v_adapter`` when cross-frame attention is enabled for the call)
``...transformer_blocks.N.i2v_adapter`` ``B200CrossFrameAttnProcessor`` (frame-0 K/V projected once per video)
``...transformer_blocks.N.attn2``      ``B200IPAdapterAttnProcessor`` (adopts ``to_k_ip/to_v_ip``) or
                                       ``B200AttnProcessor`` when no IP-Adapter is loaded
``...motion_modules...attn1/attn2``    ``B200TemporalAttnProcessor``
=====================================  =======================================================================
"""
from __future__ import annotations

import inspect
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from ._lib import MODE_AUTO, MODE_GENERIC


class RuntimeContext:
    """Facts the AttnProcessor boundary does not pass: set by forward hooks, read by processors."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.num_frames: Optional[int] = None
        self.ctx_replicated: bool = False  # encoder_hidden_states rows are repeat_interleave'd per frame (UNet :1355)


class BlockState:
    """Per I2VAdapterTransformerBlock: what this call of the block was asked to do."""

    def __init__(self):
        self.enable_cross_frame = False
        self.num_frames: Optional[int] = None
        self.cross_done = False  # attn1's processor already produced self + cross-frame output
        # frame-sharded runs (partition.FramePartitioner): frame 0 lives on another rank; the projected frame-0 K/V
        # come from its owner through one broadcast instead of from the local rows b*F
        self.first_frame_source = None


_CACHE_EPOCH = [0]   # bumped by invalidate_caches(): part of every packed-weight cache key (here and in fastpath.py)


def invalidate_caches() -> None:
    """Drop every packed / augmented weight copy held by the processors and the module-level fast path.

    The caches key on ``(data_ptr, _version)`` of their source parameters, which catches ``load_state_dict``,
    ``.to()``, optimizer steps and ordinary in-place ops -- but **not** writes through ``.data`` (``p.data.zero_()``,
    ``p.data.copy_()`` as EMA ``copy_to`` does; the reference's own ``from_transformer2d_model`` zero-initialises
    ``to_out`` that way, src/modules/i2v_adapter.py:142-143): those leave ``_version`` untouched.  Call this (or
    ``Installation.invalidate_caches()``) after such an update.  ``install`` also registers a ``load_state_dict``
    post-hook that calls it."""
    _CACHE_EPOCH[0] += 1


class _PackedWeights:
    """Cache of concatenated weight matrices, invalidated when any source parameter changes (see
    ``invalidate_caches`` for the one kind of change it cannot see)."""

    def __init__(self):
        self._key = None
        self._val = None

    @staticmethod
    def _sig(t: Optional[torch.Tensor]):
        return None if t is None else (t.data_ptr(), t._version, t.dtype, t.device, tuple(t.shape))

    def get(self, sources: List[Optional[torch.Tensor]], build):
        key = (_CACHE_EPOCH[0],) + tuple(self._sig(s) for s in sources)
        if key != self._key:
            with torch.no_grad():   # the packed copy is a constant of the forward-only kernels, never a graph node
                self._val = build()
            self._key = key
        return self._val


def _cat_bias(biases: List[Optional[torch.Tensor]], widths: List[int], like: torch.Tensor) -> Optional[torch.Tensor]:
    if all(b is None for b in biases):
        return None
    return torch.cat([b if b is not None else like.new_zeros(w) for b, w in zip(biases, widths)])


_LOG2E = 1.4426950408889634


def _augmented_projection(mods, kinds, heads: int, d: int, scale: float):
    """Weights of a packed projection that emits the augmented operand layout of ``ops.fused_self_xframe_aug``:
    per part and head ``AUG_DPAD`` output rows -- the ``d`` rows of the module's weight (query parts pre-multiplied by
    ``scale * log2(e)`` in fp32, then rounded once), zero rows for the padding -- and a bias that is 1 at column ``d``
    of every key / value head (the ones column) on top of the module's own bias.  ``kinds``: 'q' or 'kv' per part."""
    dp = ops.AUG_DPAD
    ws, bs = [], []
    for m, kind in zip(mods, kinds):
        w = m.weight.float().view(heads, d, -1)
        b = m.bias.float().view(heads, d) if m.bias is not None else w.new_zeros(heads, d)
        if kind == "q":
            w = w * (scale * _LOG2E)
            b = b * (scale * _LOG2E)
        wp = w.new_zeros(heads, dp, w.shape[-1])
        wp[:, :d] = w
        bp = w.new_zeros(heads, dp)
        bp[:, :d] = b
        if kind == "kv":
            bp[:, d] = 1.0
        ws.append(wp.view(heads * dp, -1))
        bs.append(bp.view(heads * dp))
    dt = mods[0].weight.dtype
    return torch.cat(ws).to(dt).contiguous(), torch.cat(bs).to(dt).contiguous()


def _check_supported(attn, attention_mask, what: str, hidden_states: Optional[torch.Tensor] = None) -> None:
    if torch.is_grad_enabled() and ((hidden_states is not None and hidden_states.requires_grad)
                                    or any(p.requires_grad for p in attn.parameters())):
        # the kernels are forward-only (no autograd.Function): with grad mode on, to_q / to_out of the trainable
        # I2V-Adapter (src/train_i2v_adapter.py) would silently receive no gradient
        raise RuntimeError(
            f"{what}: the B200 attention kernels are inference-only. Call the model under torch.no_grad() / "
            f"torch.inference_mode(), or uninstall() the B200 processors before training.")
    if attention_mask is not None:
        raise NotImplementedError(
            f"{what}: attention_mask is not supported by the B200 kernels (the reference UNet always passes None, "
            f"src/pipelines/pipeline_i2v_adapter.py:676-683)")
    for name in ("group_norm", "spatial_norm", "norm_cross"):
        if getattr(attn, name, None):
            raise NotImplementedError(f"{what}: Attention.{name} is not used on the SD1.5 path and is not supported")


RESIDUAL_KW = "_b200_residual"   # kwarg the module-level fast path uses to hand a processor the block's residual stream
RESIDUAL_INPLACE_KW = "_b200_residual_inplace"   # ... and to say that the processor may accumulate into it in place


def _out_proj(proc, attn, o: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], kwargs) -> torch.Tensor:
    """``to_out`` of an Attention (Linear + Dropout(0)) on o [B, S, K].

    When the caller passes the residual stream (``fastpath._block_forward_fast``), the residual add rides in the GEMM
    as its beta = 1 term (``torch.addmm``: one cuBLAS launch, no separate add pass) and the bias is *not* applied:
    it is left in ``proc.deferred_bias`` for the caller, which folds the deferred biases of a block into the next
    LayerNorm read and into the feed-forward's last GEMM.  ``proc.fused_residual`` tells the
    caller which of the two results it got."""
    res = kwargs.get(RESIDUAL_KW)
    drop = attn.to_out[1] if len(attn.to_out) > 1 else None
    if (res is not None and o.dim() == 3 and res.shape[:-1] == o.shape[:-1] and res.shape[-1] == weight.shape[0]
            and res.dtype == o.dtype and not getattr(attn, "residual_connection", False)
            and getattr(attn, "rescale_output_factor", 1.0) == 1.0 and getattr(drop, "p", 0.0) == 0.0):
        if kwargs.get(RESIDUAL_INPLACE_KW, False) and res.is_contiguous():
            # the caller owns `res` (an intermediate of its own): accumulate in place, cuBLAS reads and writes C once
            # (out of place, ATen first copies the residual into the output: a full extra read + write pass)
            out = res.view(-1, res.shape[-1]).addmm_(o.reshape(-1, o.shape[-1]), weight.t()).view(res.shape)
        else:
            out = torch.addmm(res.reshape(-1, res.shape[-1]), o.reshape(-1, o.shape[-1]), weight.t()).view(res.shape)
        proc.deferred_bias = bias
        proc.fused_residual = True
        return out
    proc.fused_residual = False
    proc.deferred_bias = None
    out = F.linear(o, weight, bias)
    return out if drop is None else drop(out)


def _finish(attn, o: torch.Tensor, residual: torch.Tensor) -> torch.Tensor:
    if getattr(attn, "residual_connection", False):
        o = o + residual
    r = getattr(attn, "rescale_output_factor", 1.0)
    return o if r == 1.0 else o / r


def _to_3d(hidden_states: torch.Tensor):
    if hidden_states.dim() == 4:
        b, c, h, w = hidden_states.shape
        return hidden_states.view(b, c, h * w).transpose(1, 2), (b, c, h, w)
    return hidden_states, None


def _from_3d(o: torch.Tensor, shape4):
    if shape4 is None:
        return o
    b, c, h, w = shape4
    return o.transpose(-1, -2).reshape(b, c, h, w)


class B200AttnProcessor:
    """Replacement for ``AttnProcessor2_0``: self- or cross-attention of any ``Attention`` module.

    ``kv_replicated`` (set by ``install`` for spatial attn2 without IP-Adapter): the context rows of the frames of
    a video are identical when the UNet-level hook says so, so K/V are projected once per video."""

    fused_residual = False   # set per call by _out_proj: the result already contains the caller's residual
    deferred_bias = None     # ... and this output-projection bias has not been applied

    def __init__(self, mode: int = MODE_AUTO, context: Optional[RuntimeContext] = None,
                 state: Optional[BlockState] = None):
        self.mode = mode
        self.context = context
        self.state = state
        self._w_qkv = _PackedWeights()
        self._w_kv = _PackedWeights()

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 scale: float = 1.0, **kwargs):
        _check_supported(attn, attention_mask, type(self).__name__, hidden_states)
        residual = hidden_states
        x, shape4 = _to_3d(hidden_states)
        B, S, _ = x.shape
        H = attn.heads
        inner = attn.to_q.out_features
        d = inner // H
        if encoder_hidden_states is None:
            w = self._w_qkv.get([attn.to_q.weight, attn.to_k.weight, attn.to_v.weight,
                                 attn.to_q.bias, attn.to_k.bias, attn.to_v.bias],
                                lambda: (torch.cat([attn.to_q.weight, attn.to_k.weight, attn.to_v.weight]),
                                         _cat_bias([attn.to_q.bias, attn.to_k.bias, attn.to_v.bias], [inner] * 3,
                                                   attn.to_q.weight)))
            qkv = F.linear(x, w[0], w[1]).view(B, S, 3, H, d)
            o = ops.sdpa(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], 1, attn.scale, self.mode)
        else:
            ctx = encoder_hidden_states
            group = 1
            if (self.context is not None and self.context.ctx_replicated and self.context.num_frames
                    and B % self.context.num_frames == 0 and ctx.shape[0] == B):
                group = self.context.num_frames
                ctx = ctx[0::group]
            w = self._w_kv.get([attn.to_k.weight, attn.to_v.weight, attn.to_k.bias, attn.to_v.bias],
                               lambda: (torch.cat([attn.to_k.weight, attn.to_v.weight]),
                                        _cat_bias([attn.to_k.bias, attn.to_v.bias], [inner] * 2, attn.to_k.weight)))
            q = attn.to_q(x).view(B, S, H, d)
            kv = F.linear(ctx, w[0], w[1]).view(ctx.shape[0], ctx.shape[1], 2, H, d)
            o = ops.sdpa(q, kv[:, :, 0], kv[:, :, 1], group, attn.scale, self.mode)
        if shape4 is None:
            return _finish(attn, _out_proj(self, attn, o.view(B, S, inner), attn.to_out[0].weight, attn.to_out[0].bias,
                                           kwargs), residual)
        self.fused_residual = False
        o = attn.to_out[0](o.view(B, S, inner))
        o = attn.to_out[1](o)
        return _finish(attn, _from_3d(o, shape4), residual)


class B200SpatialAttnProcessor(B200AttnProcessor):
    """attn1 of an ``I2VAdapterTransformerBlock`` (reference src/modules/i2v_adapter.py:468-473).

    When the block was called with ``enable_cross_frame_attn=True`` (known through the block's forward hook) the
    processor also computes the block's I2V-Adapter cross-frame attention (:484-492) in the same launch:
    one packed input GEMM ``[Wq_self; Wk_self; Wv_self; Wq_x]`` over the normalised hidden states, one small GEMM
    ``[Wk_x; Wv_x]`` over frame 0 of each video, one fused attention kernel, one stacked output GEMM
    ``[Wo_self | Wo_x]`` with the summed biases.  It returns ``self + cross`` and flags the block state so that the
    sibling ``B200CrossFrameAttnProcessor`` contributes zero to the block's ``attn_output + cross`` (:494)."""

    def __init__(self, sibling=None, state: Optional[BlockState] = None, mode: int = MODE_AUTO, fuse: bool = True):
        super().__init__(mode=mode, state=state)
        self.sibling = sibling  # the block's i2v_adapter Attention module
        self.fuse = fuse
        self._w_in = _PackedWeights()
        self._w_x = _PackedWeights()
        self._w_out = _PackedWeights()
        self._w_in_aug = _PackedWeights()
        self._w_x_aug = _PackedWeights()
        self.augmented = True  # use the augmented operand layout where the kernel has it (d = 40, bf16)

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 scale: float = 1.0, **kwargs):
        st = self.state
        fused = (self.fuse and self.sibling is not None and st is not None and st.enable_cross_frame
                 and st.num_frames and encoder_hidden_states is None and hidden_states.dim() == 3
                 and hidden_states.shape[0] % st.num_frames == 0
                 and not getattr(attn, "residual_connection", False)
                 and getattr(attn, "rescale_output_factor", 1.0) == 1.0)
        if not fused:
            return super().__call__(attn, hidden_states, encoder_hidden_states, attention_mask, temb, scale, **kwargs)
        _check_supported(attn, attention_mask, type(self).__name__, hidden_states)
        xa = self.sibling
        _check_supported(xa, None, type(self).__name__)
        x = hidden_states
        BF, S, C = x.shape
        Fr = st.num_frames
        H = attn.heads
        inner = attn.to_q.out_features
        d = inner // H
        srcs = [attn.to_q, attn.to_k, attn.to_v, xa.to_q]
        w_in = self._w_in.get([m.weight for m in srcs] + [m.bias for m in srcs],
                              lambda: (torch.cat([m.weight for m in srcs]),
                                       _cat_bias([m.bias for m in srcs], [inner] * 4, attn.to_q.weight)))
        w_x = self._w_x.get([xa.to_k.weight, xa.to_v.weight, xa.to_k.bias, xa.to_v.bias],
                            lambda: (torch.cat([xa.to_k.weight, xa.to_v.weight]),
                                     _cat_bias([xa.to_k.bias, xa.to_v.bias], [inner] * 2, xa.to_k.weight)))
        so, xo = attn.to_out[0], xa.to_out[0]
        w_out = self._w_out.get([so.weight, xo.weight, so.bias, xo.bias],
                                lambda: (torch.cat([so.weight, xo.weight], dim=1),
                                         None if so.bias is None and xo.bias is None else
                                         (so.bias if so.bias is not None else 0) +
                                         (xo.bias if xo.bias is not None else 0)))
        first = x[0::Fr]  # frame 0 of every video: rows b*F (reference :484), no F-times repeat (:485)
        if self.augmented and d == ops.AUG_D and x.dtype == torch.bfloat16 and self.mode != MODE_GENERIC:
            # level 0 (d = 40): the projections emit the augmented layout (scale folded into the query weights, ones
            # column in K and V, head dim padded to 48) that lets the kernel skip the scale FMA and the row-sum adds
            dp = ops.AUG_DPAD
            wa = self._w_in_aug.get([m.weight for m in srcs] + [m.bias for m in srcs],
                                    lambda: _augmented_projection(srcs, ["q", "kv", "kv", "q"], H, d, attn.scale))
            wxa = self._w_x_aug.get([xa.to_k.weight, xa.to_v.weight, xa.to_k.bias, xa.to_v.bias],
                                    lambda: _augmented_projection([xa.to_k, xa.to_v], ["kv", "kv"], H, d, attn.scale))
            y = F.linear(x, wa[0], wa[1]).view(BF, S, 4, H, dp)
            if st.first_frame_source is not None:  # frame shard: the owner of global frame 0 projects, one broadcast
                kvx = st.first_frame_source.broadcast_from_first_frame_owner(
                    lambda: F.linear(first, wxa[0], wxa[1]), (BF // Fr, S, 2 * H * dp), x.dtype, x.device)
            else:
                kvx = F.linear(first, wxa[0], wxa[1])
            kvx = kvx.view(BF // Fr, S, 2, H, dp)
            o = ops.fused_self_xframe_aug(y[:, :, 0], y[:, :, 1], y[:, :, 2], y[:, :, 3], kvx[:, :, 0], kvx[:, :, 1],
                                          Fr, d)
            out = _out_proj(self, attn, o.view(BF, S, 2 * inner), w_out[0], w_out[1], kwargs)
            st.cross_done = True
            return out
        y = F.linear(x, w_in[0], w_in[1]).view(BF, S, 4, H, d)
        if st.first_frame_source is not None:  # frame shard: K/V of global frame 0 are broadcast by their owner
            kvx = st.first_frame_source.broadcast_from_first_frame_owner(
                lambda: F.linear(first, w_x[0], w_x[1]), (BF // Fr, S, 2 * inner), x.dtype, x.device
            ).view(BF // Fr, S, 2, H, d)
        else:
            kvx = F.linear(first, w_x[0], w_x[1]).view(BF // Fr, S, 2, H, d)
        o = ops.fused_self_xframe(y[:, :, 0], y[:, :, 1], y[:, :, 2], y[:, :, 3], kvx[:, :, 0], kvx[:, :, 1], Fr,
                                  attn.scale, self.mode)
        out = _out_proj(self, attn, o.view(BF, S, 2 * inner), w_out[0], w_out[1], kwargs)
        st.cross_done = True
        return out


class B200CrossFrameAttnProcessor:
    """i2v_adapter of an ``I2VAdapterTransformerBlock`` (reference src/modules/i2v_adapter.py:487-492).

    The block hands over ``encoder_hidden_states`` = frame 0 of each video repeated ``num_frames`` times (:485).
    With the block state available only every ``num_frames``-th row is projected and the kernel indexes it with
    ``kv_group = num_frames``.  If attn1's processor already produced the fused result this returns a zero-stride
    zeros view (the block adds it to ``attn_output``)."""

    def __init__(self, state: Optional[BlockState] = None, mode: int = MODE_AUTO):
        self.state = state
        self.mode = mode
        self._w_kv = _PackedWeights()

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 scale: float = 1.0, **kwargs):
        st = self.state
        if st is not None and st.cross_done:
            st.cross_done = False
            return hidden_states.new_zeros((1,) * hidden_states.dim()).expand_as(hidden_states)
        _check_supported(attn, attention_mask, type(self).__name__, hidden_states)
        x = hidden_states
        B, S, _ = x.shape
        H = attn.heads
        inner = attn.to_q.out_features
        d = inner // H
        ctx = x if encoder_hidden_states is None else encoder_hidden_states
        group = 1
        if (st is not None and st.enable_cross_frame and st.num_frames and B % st.num_frames == 0
                and ctx.shape[0] == B and encoder_hidden_states is not None):
            group = st.num_frames
            ctx = ctx[0::group]
        w = self._w_kv.get([attn.to_k.weight, attn.to_v.weight, attn.to_k.bias, attn.to_v.bias],
                           lambda: (torch.cat([attn.to_k.weight, attn.to_v.weight]),
                                    _cat_bias([attn.to_k.bias, attn.to_v.bias], [inner] * 2, attn.to_k.weight)))
        q = attn.to_q(x).view(B, S, H, d)
        if (st is not None and st.first_frame_source is not None and st.enable_cross_frame
                and encoder_hidden_states is not None and ctx.shape[0] * group == B):
            # frame shard: the rows of `ctx` are this rank's *local* first frames; global frame 0 lives on its owner.
            # (Also with one frame per rank, where group == 1 and the local rows would otherwise pass for frame 0.)
            kv = st.first_frame_source.broadcast_from_first_frame_owner(
                lambda: F.linear(ctx, w[0], w[1]), (ctx.shape[0], ctx.shape[1], 2 * inner), x.dtype, x.device)
        else:
            kv = F.linear(ctx, w[0], w[1])
        kv = kv.view(ctx.shape[0], ctx.shape[1], 2, H, d)
        o = ops.sdpa(q, kv[:, :, 0], kv[:, :, 1], group, attn.scale, self.mode)
        o = attn.to_out[0](o.view(B, S, inner))
        o = attn.to_out[1](o)
        return _finish(attn, o, hidden_states)


class B200IPAdapterAttnProcessor(nn.Module):
    """Replacement for ``IPAdapterAttnProcessor2_0`` (text + image-prompt decoupled cross-attention).

    Owns ``to_k_ip`` / ``to_v_ip`` like the original (same parameter names, so the IP-Adapter checkpoint keys
    ``<id>.to_k_ip.weight`` load unchanged); ``from_processor`` adopts the parameters of an installed original."""

    def __init__(self, hidden_size: int, cross_attention_dim: Optional[int] = None, num_tokens: int = 4,
                 scale: float = 1.0, mode: int = MODE_AUTO, context: Optional[RuntimeContext] = None):
        super().__init__()
        self.hidden_size = hidden_size
        self.cross_attention_dim = cross_attention_dim
        self.num_tokens = num_tokens
        self.scale = scale
        self.mode = mode
        self.context = context
        self.to_k_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)
        self.to_v_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)
        self._w_txt = _PackedWeights()
        self._w_ip = _PackedWeights()

    @classmethod
    def from_processor(cls, proc, mode: int = MODE_AUTO, context: Optional[RuntimeContext] = None):
        new = cls(proc.hidden_size, proc.cross_attention_dim, proc.num_tokens, proc.scale, mode, context)
        new.to_k_ip, new.to_v_ip = proc.to_k_ip, proc.to_v_ip  # share the parameters, no copy
        return new

    def forward(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                scale: float = 1.0, **kwargs):
        _check_supported(attn, attention_mask, type(self).__name__, hidden_states)
        residual = hidden_states
        x, shape4 = _to_3d(hidden_states)
        B, S, _ = x.shape
        H = attn.heads
        inner = attn.to_q.out_features
        d = inner // H
        ctx = x if encoder_hidden_states is None else encoder_hidden_states
        group = 1
        c = self.context
        if c is not None and c.ctx_replicated and c.num_frames and B % c.num_frames == 0 and ctx.shape[0] == B:
            group = c.num_frames
            ctx = ctx[0::group]
        end = ctx.shape[1] - self.num_tokens
        w_txt = self._w_txt.get([attn.to_k.weight, attn.to_v.weight, attn.to_k.bias, attn.to_v.bias],
                                lambda: (torch.cat([attn.to_k.weight, attn.to_v.weight]),
                                         _cat_bias([attn.to_k.bias, attn.to_v.bias], [inner] * 2, attn.to_k.weight)))
        w_ip = self._w_ip.get([self.to_k_ip.weight, self.to_v_ip.weight],
                              lambda: torch.cat([self.to_k_ip.weight, self.to_v_ip.weight]))
        q = attn.to_q(x).view(B, S, H, d)
        Bk, T = ctx.shape[0], ctx.shape[1]
        kv = torch.empty((Bk, T, 2, H, d), dtype=x.dtype, device=x.device)
        kv2 = kv.view(Bk, T, 2 * inner)
        kv2[:, :end] = F.linear(ctx[:, :end], w_txt[0], w_txt[1])
        kv2[:, end:] = F.linear(ctx[:, end:], w_ip)
        o = ops.ip_xattn(q, kv[:, :, 0], kv[:, :, 1], end, self.scale, group, attn.scale, self.mode)
        if shape4 is None:
            return _finish(attn, _out_proj(self, attn, o.view(B, S, inner), attn.to_out[0].weight, attn.to_out[0].bias,
                                           kwargs), residual)
        self.fused_residual = False
        o = attn.to_out[0](o.view(B, S, inner))
        o = attn.to_out[1](o)
        return _finish(attn, _from_3d(o, shape4), residual)


class B200TemporalAttnProcessor:
    """Motion-module temporal self-attention (both attn1 and attn2 of the temporal BasicTransformerBlock are
    self-attentions: ``double_self_attention=True``).  hidden_states is [B*S, F, C]."""

    fused_residual = False   # set per call by _out_proj: the result already contains the caller's residual
    deferred_bias = None     # ... and this output-projection bias has not been applied

    def __init__(self, mode: int = MODE_AUTO):
        self.mode = mode
        self._w_qkv = _PackedWeights()
        self._generic = B200AttnProcessor(mode=mode)

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 scale: float = 1.0, **kwargs):
        if encoder_hidden_states is not None or hidden_states.dim() != 3:
            out = self._generic(attn, hidden_states, encoder_hidden_states, attention_mask, temb, scale, **kwargs)
            self.fused_residual, self.deferred_bias = self._generic.fused_residual, self._generic.deferred_bias
            return out
        _check_supported(attn, attention_mask, type(self).__name__, hidden_states)
        x = hidden_states
        N, Fr, _ = x.shape
        H = attn.heads
        inner = attn.to_q.out_features
        d = inner // H
        w = self._w_qkv.get([attn.to_q.weight, attn.to_k.weight, attn.to_v.weight,
                             attn.to_q.bias, attn.to_k.bias, attn.to_v.bias],
                            lambda: (torch.cat([attn.to_q.weight, attn.to_k.weight, attn.to_v.weight]),
                                     _cat_bias([attn.to_q.bias, attn.to_k.bias, attn.to_v.bias], [inner] * 3,
                                               attn.to_q.weight)))
        qkv = F.linear(x, w[0], w[1]).view(N, Fr, 3, H, d)
        o = ops.temporal_attn(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], attn.scale, self.mode)
        return _finish(attn, _out_proj(self, attn, o.view(N, Fr, inner), attn.to_out[0].weight, attn.to_out[0].bias,
                                       kwargs), hidden_states)


# ----------------------------------------------------------------------------------------------------------
# installation
# ----------------------------------------------------------------------------------------------------------
class Installation:
    """Handle returned by ``install``: keeps the previous processors and hook handles for ``uninstall``."""

    def __init__(self, unet, previous: Dict[str, Any], hooks: List[Any], context: RuntimeContext,
                 processors: Dict[str, Any], undo_forwards: Optional[List[Any]] = None):
        self.unet = unet
        self.previous = previous
        self.hooks = hooks
        self.context = context
        self.processors = processors
        self.undo_forwards = undo_forwards or []

    @staticmethod
    def invalidate_caches() -> None:
        """See ``processors.invalidate_caches``: needed after parameter writes through ``.data`` only."""
        invalidate_caches()

    def uninstall(self) -> None:
        for fn in reversed(self.undo_forwards):
            fn()
        self.undo_forwards = []
        for h in self.hooks:
            h.remove()
        self.hooks = []
        apply_attn_processors(self.unet, self.previous)


def collect_attn_processors(module: nn.Module) -> Dict[str, Any]:
    """``attn_processors`` for any module tree (same traversal and key format as the UNet property, reference
    :1118-1136); lets ``install`` work on a single block as well as on the whole UNet."""
    if hasattr(module, "attn_processors"):
        return dict(module.attn_processors)
    found: Dict[str, Any] = {}

    def walk(name: str, m: nn.Module):
        if hasattr(m, "get_processor"):
            found[f"{name}.processor"] = m.get_processor(return_deprecated_lora=True)
        for sub, child in m.named_children():
            if sub != "processor":
                walk(f"{name}.{sub}", child)

    for name, child in module.named_children():
        walk(name, child)
    return found


def apply_attn_processors(module: nn.Module, processors: Dict[str, Any]) -> None:
    if hasattr(module, "set_attn_processor"):
        module.set_attn_processor(dict(processors))
        return
    for name, proc in processors.items():
        _module_by_path(module, name[: -len(".processor")]).set_processor(proc)


def _module_by_path(root: nn.Module, path: str) -> nn.Module:
    m = root
    for part in (path.split(".") if path else []):
        m = getattr(m, part) if not part.isdigit() else m[int(part)]
    return m


def _bind_forward_args(module: nn.Module, args: Tuple, kwargs: Dict[str, Any]) -> Dict[str, Any]:
    try:
        return inspect.signature(module.forward).bind_partial(*args, **kwargs).arguments
    except TypeError:
        return dict(kwargs)


def install(unet: nn.Module, mode: int = MODE_AUTO, fuse_cross_frame: bool = True,
            fast_path: bool = True, channels_last: bool = True) -> Installation:
    """Swap every attention processor of ``unet`` (a ``UNetMotionCrossFrameAttnModel`` — the reference's or
    ``hostmodel``'s — or any sub-module exposing ``attn_processors`` / ``set_attn_processor``) for the B200 ones.

    ``fast_path=True`` additionally swaps the ``forward`` of the transformer wrappers / blocks around the attention
    calls (``fastpath.py``: GroupNorm + layout change, LayerNorm (+ positional embedding), GEGLU, layout change +
    residual in the library's bandwidth kernels); ``fast_path=False`` changes nothing but the processors.

    Must run *after* ``load_ip_adapter`` / ``_load_ip_adapter_weights`` because that call replaces all processors
    (reference :1246-1281); the IP-Adapter processors found are adopted, parameters shared."""
    ops.register_torch_ops()
    previous = collect_attn_processors(unet)
    context = RuntimeContext()
    hooks: List[Any] = []
    states: Dict[str, BlockState] = {}
    new: Dict[str, Any] = {}

    for name, old in previous.items():
        path = name[: -len(".processor")]
        parent_path, _, leaf = path.rpartition(".")  # parent_path == "" when `unet` is itself the transformer block
        if ".motion_modules." in f".{path}." or path.startswith("motion_modules."):
            new[name] = B200TemporalAttnProcessor(mode)
            continue
        block = _module_by_path(unet, parent_path)
        is_i2v_block = hasattr(block, "i2v_adapter") and hasattr(block, "attn1")
        st = states.setdefault(parent_path, BlockState()) if is_i2v_block else None
        if leaf == "attn1" and is_i2v_block:
            new[name] = B200SpatialAttnProcessor(block.i2v_adapter, st, mode, fuse_cross_frame)
        elif leaf == "i2v_adapter":
            new[name] = B200CrossFrameAttnProcessor(st, mode)
        elif hasattr(old, "to_k_ip") and hasattr(old, "to_v_ip"):
            new[name] = B200IPAdapterAttnProcessor.from_processor(old, mode, context)
        else:
            new[name] = B200AttnProcessor(mode, context, st)

    # per-block hooks: enable_cross_frame_attn / num_frames are forward arguments of the block, not of the processor
    for parent_path, st in states.items():
        block = _module_by_path(unet, parent_path)

        def pre(module, args, kwargs, _st=st):
            bound = _bind_forward_args(module, args, kwargs)
            _st.enable_cross_frame = bool(bound.get("enable_cross_frame_attn", False))
            _st.num_frames = bound.get("num_frames", None)
            _st.cross_done = False
            if context.num_frames is None and _st.num_frames:
                context.num_frames = _st.num_frames

        def post(module, args, output, _st=st):
            _st.enable_cross_frame = False
            _st.cross_done = False

        hooks.append(block.register_forward_pre_hook(pre, with_kwargs=True))
        hooks.append(block.register_forward_hook(post))

    # UNet-level hook: the UNet forward replicates the text / image tokens per frame itself (:1355)
    params = inspect.signature(unet.forward).parameters
    if "sample" in params and "enable_cross_frame_attn" in params:
        def unet_pre(module, args, kwargs):
            bound = _bind_forward_args(module, args, kwargs)
            sample = bound.get("sample")
            context.reset()
            if sample is not None and sample.dim() == 5:
                context.num_frames = int(sample.shape[1])
                context.ctx_replicated = True

        def unet_post(module, args, output):
            context.reset()

        hooks.append(unet.register_forward_pre_hook(unet_pre, with_kwargs=True))
        hooks.append(unet.register_forward_hook(unet_post))

    if hasattr(unet, "register_load_state_dict_post_hook"):
        # load_state_dict copies through no-grad copy_ (bumps _version) on most paths, but assign=True and custom
        # loaders swap .data: drop the packed copies after any load so stale weights can never be used
        hooks.append(unet.register_load_state_dict_post_hook(lambda module, incompatible: invalidate_caches()))

    apply_attn_processors(unet, new)
    undo_forwards = []
    if fast_path and channels_last and any(p.is_cuda and p.dtype == torch.bfloat16 for p in unet.parameters()):
        # Convolution weights to channels-last: the activations then flow through the UNet as [N, h, w, C], which is
        # both what cuDNN computes in (its NCHW <-> NHWC conversion kernels disappear) and the token-major layout of
        # the transformer blocks (the spatial wrapper's layout changes disappear, the motion module's become a row
        # permutation).  Values, state-dict keys and shapes are untouched; only strides of 4-D parameters change.
        converted = [p for p in unet.parameters()
                     if p.dim() == 4 and not p.is_contiguous(memory_format=torch.channels_last)]
        unet.to(memory_format=torch.channels_last)

        def restore_format(params=converted):   # uninstall gives the stock path back bit for bit
            for p in params:
                p.data = p.data.contiguous()

        undo_forwards.append(restore_format)
    if fast_path:
        from .fastpath import install_fast_forwards

        undo_forwards = undo_forwards + install_fast_forwards(unet)
    return Installation(unet, previous, hooks, context, new, undo_forwards)
