"""Torch-facing wrappers of the C ABI (``include/i2v_attn_b200.h``).

PyTorch is used for device memory and streams only: each function validates its tensors, allocates the output,
passes raw device pointers + element strides to ``libi2v_attn_b200.so`` and returns.  The work is enqueued on
``torch.cuda.current_stream()``; nothing synchronises.  CPU tensors are rejected — there is no fallback.

The same functions are registered as ``torch.ops.i2v_b200.*`` (CUDA dispatch key only).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import I2VTensor, MODE_AUTO, MODE_FAST, MODE_GENERIC  # noqa: F401  (re-exported)

_DTYPES = {torch.bfloat16: _lib.I2V_BF16, torch.float32: _lib.I2V_F32}


def _dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"i2v_b200 kernels take bfloat16 or float32 tensors, got {t.dtype}") from None


def _require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = tensors[0].device
    for t in tensors:
        if t.device.type != "cuda":
            raise RuntimeError(
                "i2v_b200 kernels run on CUDA (sm_100a) tensors only; got a tensor on "
                f"'{t.device}'. There is no CPU fallback.")
        if t.device != dev:
            raise RuntimeError("all tensors must live on the same CUDA device")
    return dev


def _desc(t: torch.Tensor) -> I2VTensor:
    """[b, s, h, d] view -> i2v_tensor."""
    if t.dim() != 4:
        raise ValueError(f"expected a 4-D [batch, seq, heads, d] view, got shape {tuple(t.shape)}")
    if t.shape[-1] > 1 and t.stride(-1) != 1:
        raise ValueError("the head dimension must be contiguous")
    return I2VTensor(t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))


def _stream(dev: torch.device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _on_device:
    """Make ``dev`` current for the duration of the call (cheap when it already is)."""

    def __init__(self, dev: torch.device):
        self.dev = dev
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if self.dev.index is not None and self.dev.index != cur:
            self.prev = cur
            torch.cuda.set_device(self.dev)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


def sdpa(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, kv_group: int = 1, scale: Optional[float] = None,
         mode: int = MODE_AUTO, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [B, Sq, H, d]; k, v [B / kv_group, Skv, H, d] (strided views allowed) -> o [B, Sq, H, d] contiguous."""
    dev = _require_cuda(q, k, v)
    B, Sq, H, d = q.shape
    Skv = k.shape[1]
    if k.shape != v.shape or k.shape[0] * kv_group != B or k.shape[2] != H or k.shape[3] != d:
        raise ValueError(f"shape mismatch: q {tuple(q.shape)} k {tuple(k.shape)} v {tuple(v.shape)} kv_group {kv_group}")
    if k.dtype != q.dtype or v.dtype != q.dtype:
        raise TypeError("q, k, v must share a dtype")
    o = torch.empty((B, Sq, H, d), dtype=q.dtype, device=dev) if out is None else out
    scale = float(d) ** -0.5 if scale is None else float(scale)
    lib = _lib.load()
    with _on_device(dev):
        _lib.check(lib.i2v_sdpa_fwd(_desc(q), _desc(k), _desc(v), _desc(o), B, H, Sq, Skv, d, kv_group, scale,
                                    _dtype_code(q), mode, _stream(dev)))
    return o


def fused_self_xframe(q_self, k_self, v_self, q_x, k_x, v_x, num_frames: int, scale: Optional[float] = None,
                      mode: int = MODE_AUTO, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Spatial self-attention + I2V-Adapter cross-frame attention in one launch.

    q_self, k_self, v_self, q_x: [B*F, S, H, d]; k_x, v_x: [B, S, H, d] (frame 0 of each video).
    Returns o [B*F, S, 2, H, d]: o[:, :, 0] = self-attention, o[:, :, 1] = cross-frame attention, i.e. a
    [B*F, S, 2*H*d] matrix ready for a single stacked output projection."""
    dev = _require_cuda(q_self, k_self, v_self, q_x, k_x, v_x)
    BF, S, H, d = q_self.shape
    if BF % num_frames:
        raise ValueError(f"Batch size {BF} must be divisible by the number of frames {num_frames}.")
    o = torch.empty((BF, S, 2, H, d), dtype=q_self.dtype, device=dev) if out is None else out
    scale = float(d) ** -0.5 if scale is None else float(scale)
    lib = _lib.load()
    with _on_device(dev):
        _lib.check(lib.i2v_fused_self_xframe_fwd(
            _desc(q_self), _desc(k_self), _desc(v_self), _desc(o[:, :, 0]), _desc(q_x), _desc(k_x), _desc(v_x),
            _desc(o[:, :, 1]), BF, H, S, d, num_frames, scale, _dtype_code(q_self), mode, _stream(dev)))
    return o


AUG_D, AUG_DPAD = 40, 48  # the augmented layout exists for SD1.5 level 0 (d = 40 stored padded to 48)


def fused_self_xframe_aug(q_self, k_self, v_self, q_x, k_x, v_x, num_frames: int, d: int = AUG_D,
                          out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`fused_self_xframe` on the augmented operand layout (include/i2v_attn_b200.h, i2v_fused_self_xframe_aug_fwd).

    All six operands are [.., S, H, 48] views: q[..., :d] pre-multiplied by scale*log2(e) with q[..., d:] = 0,
    k[..., d] = v[..., d] = 1, k/v[..., d+1:] = 0 (`augment_qkv` builds them from plain tensors; the fused block's
    packed projection produces them directly).  Returns o [B*F, S, 2, H, d]."""
    dev = _require_cuda(q_self, k_self, v_self, q_x, k_x, v_x)
    BF, S, H, dp = q_self.shape
    if BF % num_frames:
        raise ValueError(f"Batch size {BF} must be divisible by the number of frames {num_frames}.")
    o = torch.empty((BF, S, 2, H, d), dtype=q_self.dtype, device=dev) if out is None else out
    lib = _lib.load()
    with _on_device(dev):
        _lib.check(lib.i2v_fused_self_xframe_aug_fwd(
            _desc(q_self), _desc(k_self), _desc(v_self), _desc(o[:, :, 0]), _desc(q_x), _desc(k_x), _desc(v_x),
            _desc(o[:, :, 1]), BF, H, S, d, dp, num_frames, _dtype_code(q_self), _stream(dev)))
    return o


def augment_qkv(q, k, v, scale: Optional[float] = None):
    """Plain [.., S, H, 40] q, k, v -> the augmented [.., S, H, 48] operands (test / reference helper: the product
    path gets this layout from the projection GEMM at no extra pass)."""
    d = q.shape[-1]
    scale = float(d) ** -0.5 if scale is None else float(scale)
    pad = AUG_DPAD - d

    def _pad(x, one):
        z = torch.zeros(*x.shape[:-1], pad, dtype=x.dtype, device=x.device)
        if one:
            z[..., 0] = 1
        return torch.cat([x, z], dim=-1)

    qa = _pad((q.float() * (scale * 1.4426950408889634)).to(q.dtype), False)
    return qa, _pad(k, True), _pad(v, True)


def ip_xattn(q, k, v, n_txt: int, ip_scale: float, kv_group: int = 1, scale: Optional[float] = None,
             mode: int = MODE_AUTO, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """IP-Adapter decoupled cross-attention.  q [B, Sq, H, d]; k, v [B / kv_group, n_txt + n_ip, H, d] hold the text
    tokens followed by the image-prompt tokens."""
    dev = _require_cuda(q, k, v)
    B, Sq, H, d = q.shape
    n_ip = k.shape[1] - n_txt
    if n_ip <= 0 or n_txt <= 0:
        raise ValueError(f"need text and image tokens, got n_txt={n_txt}, total={k.shape[1]}")
    o = torch.empty((B, Sq, H, d), dtype=q.dtype, device=dev) if out is None else out
    scale = float(d) ** -0.5 if scale is None else float(scale)
    lib = _lib.load()
    with _on_device(dev):
        _lib.check(lib.i2v_ip_xattn_fwd(
            _desc(q), _desc(k[:, :n_txt]), _desc(v[:, :n_txt]), _desc(k[:, n_txt:]), _desc(v[:, n_txt:]), _desc(o),
            B, H, Sq, n_txt, n_ip, d, kv_group, scale, float(ip_scale), _dtype_code(q), mode, _stream(dev)))
    return o


def temporal_attn(q, k, v, scale: Optional[float] = None, mode: int = MODE_AUTO,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q, k, v [n_pos, F, H, d] (strided views allowed, head stride == d) -> o [n_pos, F, H, d] contiguous."""
    dev = _require_cuda(q, k, v)
    N, Fr, H, d = q.shape
    if k.shape != q.shape or v.shape != q.shape:
        raise ValueError("temporal attention is self-attention: q, k, v must have the same shape")
    o = torch.empty((N, Fr, H, d), dtype=q.dtype, device=dev) if out is None else out
    scale = float(d) ** -0.5 if scale is None else float(scale)
    lib = _lib.load()
    with _on_device(dev):
        _lib.check(lib.i2v_temporal_attn_fwd(_desc(q), _desc(k), _desc(v), _desc(o), N, H, Fr, d, scale,
                                             _dtype_code(q), mode, _stream(dev)))
    return o


def reshard_pack(x: torch.Tensor, world: int, inverse: bool = False, out: Optional[torch.Tensor] = None):
    """x [V, f, S, C] -> chunks [G, V, f, S/G, C] (inverse: chunks -> x)."""
    dev = _require_cuda(x)
    if not x.is_contiguous():
        raise ValueError("reshard_pack needs a contiguous tensor")
    if not inverse:
        V, f, S, C = x.shape
        o = torch.empty((world, V, f, S // world, C), dtype=x.dtype, device=dev) if out is None else out
    else:
        G, V, f, Sl, C = x.shape
        S = Sl * G
        o = torch.empty((V, f, S, C), dtype=x.dtype, device=dev) if out is None else out
    lib = _lib.load()
    with _on_device(dev):
        _lib.check(lib.i2v_reshard_pack(x.data_ptr(), o.data_ptr(), V, f, S, C, world, x.element_size(),
                                        int(inverse), _stream(dev)))
    return o


def reshard_unpack(x: torch.Tensor, world: int, inverse: bool = False, out: Optional[torch.Tensor] = None):
    """recv [G, V, f, Sl, C] -> y [V, G*f, Sl, C] (inverse: y -> recv)."""
    dev = _require_cuda(x)
    if not x.is_contiguous():
        raise ValueError("reshard_unpack needs a contiguous tensor")
    if not inverse:
        G, V, f, Sl, C = x.shape
        o = torch.empty((V, G * f, Sl, C), dtype=x.dtype, device=dev) if out is None else out
    else:
        V, Fall, Sl, C = x.shape
        f = Fall // world
        o = torch.empty((world, V, f, Sl, C), dtype=x.dtype, device=dev) if out is None else out
    lib = _lib.load()
    with _on_device(dev):
        _lib.check(lib.i2v_reshard_unpack(x.data_ptr(), o.data_ptr(), V, f, Sl, C, world, x.element_size(),
                                          int(inverse), _stream(dev)))
    return o


# ----------------------------------------------------------------------------------------------------------
# normalisation prologues, layout changes, residual epilogue (bf16, contiguous tensors)
# ----------------------------------------------------------------------------------------------------------
def _require_bf16_contig(*tensors: torch.Tensor) -> torch.device:
    dev = _require_cuda(*tensors)
    for t in tensors:
        if t.dtype != torch.bfloat16:
            raise TypeError(f"this kernel takes bfloat16 tensors, got {t.dtype}")
        if not t.is_contiguous():
            raise ValueError("this kernel takes contiguous tensors")
    return dev


def layernorm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5,
              pe: Optional[torch.Tensor] = None, pre: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LayerNorm over the last axis; ``pe`` [F, C] is added per row with ``row % F`` (rows of a [N, F, C] tensor);
    ``pre`` [C] is added to x before the statistics (a deferred per-channel bias of the producer)."""
    dev = _require_bf16_contig(*[t for t in (x, weight, bias, pe, pre) if t is not None])
    C = x.shape[-1]
    rows = x.numel() // C
    if pe is not None and (pe.dim() != 2 or pe.shape[1] != C or x.dim() < 2 or x.shape[-2] != pe.shape[0]):
        raise ValueError(f"pe {tuple(pe.shape)} does not match x {tuple(x.shape)}: need pe [x.shape[-2], C]")
    if pre is not None and tuple(pre.shape) != (C,):
        raise ValueError(f"pre {tuple(pre.shape)} must be [{C}]")
    y = torch.empty_like(x)
    lib = _lib.load()
    with _on_device(dev):
        _lib.check(lib.i2v_layernorm_pre_fwd(x.data_ptr(), None if pre is None else pre.data_ptr(), weight.data_ptr(),
                                             bias.data_ptr(), None if pe is None else pe.data_ptr(), y.data_ptr(), rows,
                                             C, 0 if pe is None else pe.shape[0], float(eps), _stream(dev)))
    return y


def geglu(x: torch.Tensor, ones_column: bool = False) -> torch.Tensor:
    """x [..., 2*D] -> x[..., :D] * gelu(x[..., D:]).  With ``ones_column`` the result is [..., D + 8] whose extra
    columns are (1, 0, ..., 0): the following GEMM then carries its bias as weight column D."""
    dev = _require_bf16_contig(x)
    D = x.shape[-1] // 2
    rows = x.numel() // (2 * D)
    ld = D + 8 if ones_column else D
    y = torch.empty(x.shape[:-1] + (ld,), dtype=x.dtype, device=dev)
    lib = _lib.load()
    with _on_device(dev):
        _lib.check(lib.i2v_geglu_ld_fwd(x.data_ptr(), y.data_ptr(), rows, D, ld, _stream(dev)))
    return y


def ff_geglu(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
             ones_column: bool = False) -> torch.Tensor:
    """GEGLU feed-forward input projection fused with the activation: ``h, g = F.linear(x, weight, bias).chunk(2, -1);
    h * gelu(g)`` in one tcgen05 GEMM (include/i2v_attn_b200.h, i2v_ff_geglu_fwd).  x [..., K], weight [2N, K],
    bias [2N] -> [..., N] (or [..., N + 8] with the ones column of `geglu`)."""
    dev = _require_bf16_contig(x)
    K = x.shape[-1]
    N = weight.shape[0] // 2
    if weight.shape != (2 * N, K) or not weight.is_contiguous() or weight.dtype != torch.bfloat16:
        raise ValueError(f"ff_geglu: weight must be a contiguous bf16 [2N, K] matrix, got {tuple(weight.shape)} {weight.dtype}")
    if bias is not None and (bias.shape != (2 * N,) or bias.dtype != torch.bfloat16 or not bias.is_contiguous()):
        raise ValueError("ff_geglu: bias must be a contiguous bf16 [2N] vector")
    rows = x.numel() // K
    ld = N + 8 if ones_column else N
    y = torch.empty(x.shape[:-1] + (ld,), dtype=x.dtype, device=dev)
    lib = _lib.load()
    with _on_device(dev):
        _lib.check(lib.i2v_ff_geglu_fwd(x.data_ptr(), weight.data_ptr(), None if bias is None else bias.data_ptr(),
                                        y.data_ptr(), rows, K, N, ld, _stream(dev)))
    return y


def linear_supported(x: torch.Tensor, weight: torch.Tensor) -> bool:
    """Shapes `linear` takes: CUDA bf16, rows of x contiguous, K and N multiples of 8, 16-byte aligned bases."""
    return (x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and weight.dim() == 2
            and weight.is_contiguous() and x.shape[-1] == weight.shape[1] and x.shape[-1] % 8 == 0
            and weight.shape[0] % 8 == 0 and x.stride(-1) == 1 and x.data_ptr() % 16 == 0
            and weight.data_ptr() % 16 == 0)


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
           residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``F.linear(x, weight, bias) [+ residual]`` on the library's tcgen05 token GEMM (i2v_linear_fwd).

    x [..., K] (the leading axes must flatten to rows with one pitch), weight [N, K] contiguous, bias [N];
    ``residual`` [..., N] is added in the epilogue (after the product is rounded to bf16, like the reference's separate
    add); ``out`` may be the residual itself (in-place accumulation into the residual stream)."""
    dev = _require_cuda(x, weight)
    K, N = x.shape[-1], weight.shape[0]
    if not linear_supported(x, weight):
        raise ValueError(f"linear: unsupported operands x {tuple(x.shape)} {x.dtype} weight {tuple(weight.shape)}")
    x2 = x.reshape(-1, K)
    if x2.stride(1) != 1 or (x2.shape[0] > 1 and x2.stride(0) % 8):
        x2 = x2.contiguous()
    rows = x2.shape[0]
    ld_x = x2.stride(0) if rows > 1 else K
    if bias is not None and (bias.shape != (N,) or bias.dtype != torch.bfloat16 or not bias.is_contiguous()):
        raise ValueError("linear: bias must be a contiguous bf16 [N] vector")
    res2 = None
    if residual is not None:
        if residual.dtype != torch.bfloat16 or residual.shape[-1] != N or residual.numel() != rows * N:
            raise ValueError(f"linear: residual {tuple(residual.shape)} does not match [{rows}, {N}]")
        res2 = residual.reshape(rows, N)
        if not res2.is_contiguous():
            res2 = res2.contiguous()
    if out is None:
        o = torch.empty(x.shape[:-1] + (N,), dtype=x.dtype, device=dev)
    else:
        o = out
        if o.dtype != torch.bfloat16 or not o.is_contiguous() or o.numel() != rows * N:
            raise ValueError("linear: out must be a contiguous bf16 tensor of rows * N elements")
    lib = _lib.load()
    with _on_device(dev):
        _lib.check(lib.i2v_linear_fwd(x2.data_ptr(), weight.data_ptr(), None if bias is None else bias.data_ptr(),
                                      None if res2 is None else res2.data_ptr(), o.data_ptr(), rows, K, N, ld_x, N, N,
                                      _stream(dev)))
    return o


def ff_geglu_supported(x: torch.Tensor, weight: torch.Tensor) -> bool:
    """Shapes the fused kernel takes (K % 64 == 0, N % 128 == 0, bf16, contiguous)."""
    return (x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and weight.is_contiguous()
            and x.shape[-1] % 64 == 0 and (weight.shape[0] // 2) % 128 == 0 and weight.shape[0] % 2 == 0)


def group_norm_tokens(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, groups: int, eps: float,
                      frames_per_stat: int = 1) -> torch.Tensor:
    """GroupNorm of x [N, C, h, w] with statistics shared by ``frames_per_stat`` consecutive batch entries, written
    token-major: [N, h*w, C] (frames_per_stat = 1) or [N/F * h*w, F, C] (frames_per_stat = F)."""
    dev = _require_bf16_contig(x, weight, bias)
    N, C, h, w = x.shape
    S, fg = h * w, frames_per_stat
    partial = torch.empty((N, groups, 2), dtype=torch.float32, device=dev)
    out = (torch.empty((N, S, C), dtype=x.dtype, device=dev) if fg == 1
           else torch.empty((N // fg * S, fg, C), dtype=x.dtype, device=dev))
    lib = _lib.load()
    with _on_device(dev):
        st = _stream(dev)
        _lib.check(lib.i2v_gn_stats(x.data_ptr(), partial.data_ptr(), N, C, S, groups, st))
        _lib.check(lib.i2v_gn_apply_transpose(x.data_ptr(), partial.data_ptr(), weight.data_ptr(), bias.data_ptr(),
                                              out.data_ptr(), N, C, S, groups, fg, float(eps), st))
    return out


def tokens_to_nchw_residual(y: torch.Tensor, residual: torch.Tensor, frames_per_stat: int = 1) -> torch.Tensor:
    """Inverse layout of ``group_norm_tokens`` plus the residual: y token-major, residual [N, C, h, w] -> [N, C, h, w]."""
    dev = _require_bf16_contig(y, residual)
    N, C, h, w = residual.shape
    if y.numel() != residual.numel():
        raise ValueError(f"y {tuple(y.shape)} and residual {tuple(residual.shape)} differ in size")
    out = torch.empty_like(residual)
    lib = _lib.load()
    with _on_device(dev):
        _lib.check(lib.i2v_untranspose_residual(y.data_ptr(), residual.data_ptr(), out.data_ptr(), N, C, h * w,
                                                frames_per_stat, _stream(dev)))
    return out


def is_channels_last(x: torch.Tensor) -> bool:
    """(N, C, h, w) stored as [N, h, w, C] (torch.channels_last), dense."""
    return x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)


def group_norm_nhwc(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, groups: int, eps: float,
                    frames_per_stat: int = 1, silu: bool = False, add: Optional[torch.Tensor] = None,
                    to_positions: bool = False, x2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """GroupNorm of a channels-last activation x (N, C, h, w) [+ per-(n, c) ``add`` before, + SiLU after].

    Returns a channels-last (N, C, h, w) tensor, or with ``to_positions`` the motion module's [N/F * h*w, F, C]
    token layout (statistics shared by the F = ``frames_per_stat`` frames of a video either way).  With ``x2`` the
    input is the channel concatenation ``torch.cat([x, x2], 1)``, read from the two tensors in place."""
    if not (x.is_cuda and x.dtype == torch.bfloat16 and is_channels_last(x)):
        raise RuntimeError("group_norm_nhwc: needs a CUDA bf16 channels-last tensor (no CPU fallback)")
    dev = x.device
    N, C, h, w = x.shape
    C1 = C
    if x2 is not None:
        if not (x2.is_cuda and x2.dtype == x.dtype and is_channels_last(x2) and x2.shape[0] == N
                and x2.shape[2:] == x.shape[2:] and C1 % 8 == 0 and x2.shape[1] % 8 == 0):
            raise RuntimeError("group_norm_nhwc: the second source must be a matching channels-last tensor, C % 8 == 0")
        C = C1 + x2.shape[1]
    S, fg = h * w, frames_per_stat
    lib = _lib.load()
    scratch = torch.empty((int(lib.i2v_gn_nhwc_scratch_floats(N, groups)),), dtype=torch.float32, device=dev)
    if to_positions:
        out = torch.empty((N // fg * S, fg, C), dtype=x.dtype, device=dev)
    else:
        out = torch.empty((N, C, h, w), dtype=x.dtype, device=dev, memory_format=torch.channels_last)
    if add is not None:
        add = add.to(dtype=x.dtype).reshape(N, C).contiguous()
    with _on_device(dev):
        _lib.check(lib.i2v_gn_nhwc_cat(x.data_ptr(), None if x2 is None else x2.data_ptr(), C1,
                                       add.data_ptr() if add is not None else None, weight.data_ptr(), bias.data_ptr(),
                                       out.data_ptr(), scratch.data_ptr(), N, S, C, groups, fg, float(eps), int(silu),
                                       int(to_positions), _stream(dev)))
    return out


def group_norm_nhwc_sums(x: torch.Tensor, groups: int, frames_per_stat: int) -> torch.Tensor:
    """Raw per-(video, group) (sum, sum of squares) over the ``frames_per_stat`` local frames of a channels-last
    activation x (N, C, h, w): fp32 [N / F, groups, 2].  First half of the frame-sharded GroupNorm."""
    if not (x.is_cuda and x.dtype == torch.bfloat16 and is_channels_last(x)):
        raise RuntimeError("group_norm_nhwc_sums: needs a CUDA bf16 channels-last tensor (no CPU fallback)")
    N, C, h, w = x.shape
    lib = _lib.load()
    scratch = torch.empty((int(lib.i2v_gn_nhwc_scratch_floats(N, groups)),), dtype=torch.float32, device=x.device)
    sums = torch.empty((N // frames_per_stat, groups, 2), dtype=torch.float32, device=x.device)
    with _on_device(x.device):
        _lib.check(lib.i2v_gn_nhwc_sums(x.data_ptr(), sums.data_ptr(), scratch.data_ptr(), N, h * w, C, groups,
                                        frames_per_stat, _stream(x.device)))
    return sums


def group_norm_nhwc_apply(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, mean_rstd: torch.Tensor,
                          groups: int, frames_per_stat: int, world: int = 1, silu: bool = False) -> torch.Tensor:
    """Second half: y = (x - mean) * rstd * w + b with caller-provided fp32 statistics [N / F, groups, 2].
    ``world == 1``: the motion module's [N/F * S, F, C] token layout; ``world > 1``: the frame partitioner's
    all-to-all send buffer [W, V, S / W, F_local, C]."""
    if not (x.is_cuda and x.dtype == torch.bfloat16 and is_channels_last(x)):
        raise RuntimeError("group_norm_nhwc_apply: needs a CUDA bf16 channels-last tensor (no CPU fallback)")
    N, C, h, w = x.shape
    S, fg = h * w, frames_per_stat
    V = N // fg
    if mean_rstd.shape != (V, groups, 2) or mean_rstd.dtype != torch.float32 or not mean_rstd.is_contiguous():
        raise ValueError(f"mean_rstd must be a contiguous fp32 [{V}, {groups}, 2] tensor")
    if world > 1:
        out = torch.empty((world, V, S // world, fg, C), dtype=x.dtype, device=x.device)
    else:
        out = torch.empty((V * S, fg, C), dtype=x.dtype, device=x.device)
    lib = _lib.load()
    with _on_device(x.device):
        _lib.check(lib.i2v_gn_nhwc_apply(x.data_ptr(), None, weight.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                         mean_rstd.data_ptr(), N, S, C, groups, fg, int(silu), 2 if world > 1 else 1,
                                         world, _stream(x.device)))
    return out


def sharded_positions_to_nhwc_residual(y: torch.Tensor, residual: torch.Tensor, frames_per_stat: int,
                                       world: int) -> torch.Tensor:
    """Frame partitioner, way back: y [W, V, S / W, F_local, C] (all-to-all receive buffer) + channels-last residual
    (N, C, h, w) -> channels-last (N, C, h, w)."""
    if not (y.is_cuda and y.dtype == torch.bfloat16 and y.is_contiguous() and is_channels_last(residual)):
        raise RuntimeError("sharded_positions_to_nhwc_residual: needs CUDA bf16 tensors, residual channels-last")
    N, C, h, w = residual.shape
    if y.numel() != residual.numel():
        raise ValueError(f"y {tuple(y.shape)} and residual {tuple(residual.shape)} differ in size")
    out = torch.empty_like(residual)
    lib = _lib.load()
    with _on_device(y.device):
        _lib.check(lib.i2v_rows_residual_sharded(y.data_ptr(), residual.data_ptr(), out.data_ptr(), N, h * w, C,
                                                 frames_per_stat, world, _stream(y.device)))
    return out


def positions_to_nhwc_residual(y: torch.Tensor, residual: torch.Tensor, frames_per_stat: int) -> torch.Tensor:
    """Motion module, way back: y [N/F * h*w, F, C] + channels-last residual (N, C, h, w) -> channels-last (N, C, h, w)."""
    if not (y.is_cuda and y.dtype == torch.bfloat16 and y.is_contiguous() and is_channels_last(residual)):
        raise RuntimeError("positions_to_nhwc_residual: needs CUDA bf16 tensors, residual channels-last (no CPU fallback)")
    N, C, h, w = residual.shape
    if y.numel() != residual.numel():
        raise ValueError(f"y {tuple(y.shape)} and residual {tuple(residual.shape)} differ in size")
    out = torch.empty_like(residual)
    lib = _lib.load()
    with _on_device(y.device):
        _lib.check(lib.i2v_rows_residual(y.data_ptr(), residual.data_ptr(), out.data_ptr(), N, h * w, C, frames_per_stat,
                                         _stream(y.device)))
    return out


def nhwc_add(x: torch.Tensor, y: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x + (y + bias[c]) for two channels-last (N, C, h, w) tensors, one pass (ResnetBlock2D's residual add with the
    second convolution's bias folded in)."""
    if not (x.is_cuda and x.dtype == torch.bfloat16 and is_channels_last(x) and is_channels_last(y)
            and x.shape == y.shape and y.dtype == x.dtype):
        raise RuntimeError("nhwc_add: needs two CUDA bf16 channels-last tensors of one shape (no CPU fallback)")
    N, C, h, w = x.shape
    out = torch.empty_like(x)
    if bias is not None:
        bias = bias.to(dtype=x.dtype).contiguous()
    lib = _lib.load()
    with _on_device(x.device):
        _lib.check(lib.i2v_rows_residual_bias(y.data_ptr(), x.data_ptr(), None if bias is None else bias.data_ptr(),
                                              out.data_ptr(), N, h * w, C, 1, _stream(x.device)))
    return out


def nhwc_bias_add_(x: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """In place x += bias[c] on a channels-last (N, C, h, w) tensor: a convolution bias as one full-bandwidth pass
    (cuDNN's bf16 channels-last convolutions apply it with a separate strided elementwise kernel)."""
    if not (x.is_cuda and x.dtype == torch.bfloat16 and is_channels_last(x) and x.shape[1] % 8 == 0):
        raise RuntimeError("nhwc_bias_add_: needs a CUDA bf16 channels-last tensor with C % 8 == 0 (no CPU fallback)")
    N, C, h, w = x.shape
    bias = bias.to(dtype=x.dtype).contiguous()
    lib = _lib.load()
    with _on_device(x.device):
        _lib.check(lib.i2v_rows_residual_bias(x.data_ptr(), None, bias.data_ptr(), x.data_ptr(), N, h * w, C, 1,
                                              _stream(x.device)))
    return x


def upsample2x_nhwc(x: torch.Tensor) -> torch.Tensor:
    """``F.interpolate(x, scale_factor=2.0, mode="nearest")`` for a channels-last bf16 (N, C, h, w) tensor."""
    if not (x.is_cuda and x.dtype == torch.bfloat16 and is_channels_last(x) and x.shape[1] % 8 == 0):
        raise RuntimeError("upsample2x_nhwc: needs a CUDA bf16 channels-last tensor with C % 8 == 0 (no CPU fallback)")
    N, C, h, w = x.shape
    out = torch.empty((N, C, 2 * h, 2 * w), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    lib = _lib.load()
    with _on_device(x.device):
        _lib.check(lib.i2v_upsample2x_nhwc(x.data_ptr(), out.data_ptr(), N, h, w, C, _stream(x.device)))
    return out


# ----------------------------------------------------------------------------------------------------------
# torch.library registration: torch.ops.i2v_b200.*  (CUDA key only -> CPU tensors raise NotImplementedError)
# ----------------------------------------------------------------------------------------------------------
_torch_lib = None


def register_torch_ops() -> None:
    global _torch_lib
    if _torch_lib is not None:
        return
    lib = torch.library.Library("i2v_b200", "DEF")
    lib.define("sdpa(Tensor q, Tensor k, Tensor v, int kv_group, float scale, int mode) -> Tensor")
    lib.define("fused_self_xframe(Tensor q_self, Tensor k_self, Tensor v_self, Tensor q_x, Tensor k_x, Tensor v_x, "
               "int num_frames, float scale, int mode) -> Tensor")
    lib.define("ip_xattn(Tensor q, Tensor k, Tensor v, int n_txt, float ip_scale, int kv_group, float scale, "
               "int mode) -> Tensor")
    lib.define("temporal_attn(Tensor q, Tensor k, Tensor v, float scale, int mode) -> Tensor")
    lib.impl("sdpa", lambda q, k, v, g, s, m: sdpa(q, k, v, g, s, m), "CUDA")
    lib.impl("fused_self_xframe",
             lambda qs, ks, vs, qx, kx, vx, f, s, m: fused_self_xframe(qs, ks, vs, qx, kx, vx, f, s, m), "CUDA")
    lib.impl("ip_xattn", lambda q, k, v, n, ips, g, s, m: ip_xattn(q, k, v, n, ips, g, s, m), "CUDA")
    lib.impl("temporal_attn", lambda q, k, v, s, m: temporal_attn(q, k, v, s, m), "CUDA")
    _torch_lib = lib
