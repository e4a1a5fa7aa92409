"""Frame / batch partitioner: how the attention hot path is split over the GPUs of one 8xB200 box.

The reference samples on a single GPU (``/root/reference/src/pipelines/pipeline_i2v_adapter.py:748, 781-785``); the
partitions below are new functionality that follows from the data layout the reference's UNet fixes
(``/root/reference/src/models/unet_motion_cross_frame_attn.py:1358``: batch row = ``video * num_frames + frame``).

``BatchPartition`` — videos and CFG halves are independent through the whole UNet (guidance combines the halves
    only after the forward, pipeline ``:686-688``): a contiguous split of the video axis, **no collective**.

``FramePartitioner`` — long clips: every rank holds ``f = F / G`` consecutive frames of every video.  Convolutions,
    spatial self-attention, IP-Adapter cross-attention and their norms are per frame and need nothing.  Three
    couplings each get one exchange step:

    * I2V-Adapter cross-frame attention (``src/modules/i2v_adapter.py:484-492``): the rank owning frame 0 projects
      its K/V once and shares them with ``ncclBroadcast`` (``torch.distributed.broadcast``), every rank then runs
      the fused kernel with ``kv_group = f``;
    * motion-module temporal attention (``TransformerTemporalModel``, constructed at
      ``unet_motion_cross_frame_attn.py:232-244``): an all-to-all re-shards ``[V, f, S, C]`` (my frames, all
      positions) to ``[V, F, S/G, C]`` (all frames, my positions) before the module's transformer and back after it;
      the two local copies around the collective are the C-ABI kernels ``i2v_reshard_pack / i2v_reshard_unpack``;
    * the motion module's GroupNorm statistics span all frames of a video: per-rank (mean, M2) are all-gathered
      (``2 * V * groups`` floats) and merged with the parallel-variance formula.

Everything here is host-side orchestration on ``torch.distributed`` (NCCL on GPUs; gloo in the CPU tests).  The layout
copies run in ``libi2v_attn_b200.so`` for CUDA tensors.  CPU tensors are rejected unless the caller opts into the
plain-PyTorch index permutation with ``allow_torch_layout=True`` — that switch exists for the gloo tests of the
index math and is never set by the library itself.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import nn

from . import ops


# ----------------------------------------------------------------------------------------------------------------
# data-parallel split over videos / CFG halves
# ----------------------------------------------------------------------------------------------------------------
class BatchPartition:
    """Contiguous split of ``n_items`` (videos, or CFG halves x videos) over ``world`` ranks; no communication in the
    step.  Ranks may hold different counts when ``n_items % world != 0`` (the first ``n_items % world`` get one
    more)."""

    def __init__(self, n_items: int, world: int, rank: int):
        if n_items <= 0 or world <= 0 or not (0 <= rank < world):
            raise ValueError(f"bad partition: n_items={n_items} world={world} rank={rank}")
        self.n_items, self.world, self.rank = n_items, world, rank
        base, extra = divmod(n_items, world)
        self.counts = [base + (1 if r < extra else 0) for r in range(world)]
        self.offsets = [sum(self.counts[:r]) for r in range(world)]

    @property
    def local_slice(self) -> slice:
        return slice(self.offsets[self.rank], self.offsets[self.rank] + self.counts[self.rank])

    def split(self, tensor: torch.Tensor, dim: int = 0) -> torch.Tensor:
        if tensor.shape[dim] != self.n_items:
            raise ValueError(f"axis {dim} has {tensor.shape[dim]} entries, partition was built for {self.n_items}")
        return tensor.narrow(dim, self.offsets[self.rank], self.counts[self.rank])

    def split_cfg(self, tensor: torch.Tensor) -> torch.Tensor:
        """For ``[negative | positive]`` stacked embeddings (2 * n_items rows, pipeline ``:613-614``): this rank's rows
        of both halves, still stacked ``[negative | positive]``."""
        if tensor.shape[0] != 2 * self.n_items:
            raise ValueError(f"expected {2 * self.n_items} rows, got {tensor.shape[0]}")
        neg, pos = tensor[: self.n_items], tensor[self.n_items:]
        return torch.cat([neg[self.local_slice], pos[self.local_slice]])

    def gather(self, local: torch.Tensor, group=None) -> torch.Tensor:
        """All-gather the per-rank results along dim 0 (final latents; not part of the timed step)."""
        if self.world == 1:
            return local
        if len(set(self.counts)) == 1:
            out = local.new_empty((self.n_items,) + tuple(local.shape[1:]))
            dist.all_gather_into_tensor(out, local.contiguous(), group=group)
            return out
        # uneven counts: pad every contribution to the largest one (collectives need equal sizes), trim after
        cmax = max(self.counts)
        padded = local.new_zeros((cmax,) + tuple(local.shape[1:]))
        padded[: local.shape[0]] = local
        out = local.new_empty((self.world * cmax,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, padded, group=group)
        out = out.view((self.world, cmax) + tuple(local.shape[1:]))
        return torch.cat([out[r, :c] for r, c in enumerate(self.counts)])


# ----------------------------------------------------------------------------------------------------------------
# layout copies around the all-to-all
# ----------------------------------------------------------------------------------------------------------------
def _pack(x: torch.Tensor, world: int, inverse: bool, allow_torch_layout: bool) -> torch.Tensor:
    if x.is_cuda:
        return ops.reshard_pack(x, world, inverse)
    if not allow_torch_layout:
        raise RuntimeError("re-shard layout copies run in libi2v_attn_b200.so on CUDA tensors; CPU tensors need "
                           "allow_torch_layout=True (test-only index permutation)")
    if not inverse:   # [V, f, S, C] -> [G, V, f, S/G, C]
        V, f, S, C = x.shape
        return x.view(V, f, world, S // world, C).permute(2, 0, 1, 3, 4).contiguous()
    G, V, f, Sl, C = x.shape  # [G, V, f, S/G, C] -> [V, f, S, C]
    return x.permute(1, 2, 0, 3, 4).reshape(V, f, G * Sl, C)


def _unpack(x: torch.Tensor, world: int, inverse: bool, allow_torch_layout: bool) -> torch.Tensor:
    if x.is_cuda:
        return ops.reshard_unpack(x, world, inverse)
    if not allow_torch_layout:
        raise RuntimeError("re-shard layout copies run in libi2v_attn_b200.so on CUDA tensors; CPU tensors need "
                           "allow_torch_layout=True (test-only index permutation)")
    if not inverse:   # [G, V, f, Sl, C] -> [V, G*f, Sl, C]
        G, V, f, Sl, C = x.shape
        return x.permute(1, 0, 2, 3, 4).reshape(V, G * f, Sl, C)
    V, Fall, Sl, C = x.shape  # [V, G*f, Sl, C] -> [G, V, f, Sl, C]
    return x.view(V, world, Fall // world, Sl, C).permute(1, 0, 2, 3, 4).contiguous()


def frames_to_positions(x: torch.Tensor, group=None, allow_torch_layout: bool = False) -> torch.Tensor:
    """``[V, f, S, C]`` (this rank's frames, every position) -> ``[V, G*f, S/G, C]`` (every frame, this rank's
    positions): pack, one all-to-all, unpack."""
    world = dist.get_world_size(group)
    if x.shape[2] % world:
        raise ValueError(f"{x.shape[2]} spatial positions are not divisible by the {world} ranks of the group")
    send = _pack(x.contiguous(), world, False, allow_torch_layout)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return _unpack(recv, world, False, allow_torch_layout)


def positions_to_frames(y: torch.Tensor, group=None, allow_torch_layout: bool = False) -> torch.Tensor:
    """Inverse of ``frames_to_positions``: ``[V, G*f, S/G, C]`` -> ``[V, f, S, C]``."""
    world = dist.get_world_size(group)
    if y.shape[1] % world:
        raise ValueError(f"{y.shape[1]} frames are not divisible by the {world} ranks of the group")
    send = _unpack(y.contiguous(), world, True, allow_torch_layout)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return _pack(recv, world, True, allow_torch_layout)


def sharded_group_norm(x: torch.Tensor, norm: nn.GroupNorm, group=None) -> torch.Tensor:
    """GroupNorm of the motion module over ``(C/groups, F, h, w)`` per video when the F axis is sharded.
    x: ``[V, f, C, h, w]`` (local frames).  Per-rank mean / M2 are all-gathered and merged exactly."""
    V, f, C, h, w = x.shape
    g = norm.num_groups
    xf = x.float().view(V, f, g, C // g, h * w)
    n_local = f * (C // g) * h * w
    mean_l = xf.mean(dim=(1, 3, 4))                                       # [V, g]
    m2_l = (xf - mean_l.view(V, 1, g, 1, 1)).square().sum(dim=(1, 3, 4))  # [V, g]
    world = dist.get_world_size(group)
    stats = torch.stack([mean_l, m2_l])                                    # [2, V, g]
    allstats = stats.new_empty((world * 2, V, g))                          # ranks concatenated along dim 0
    dist.all_gather_into_tensor(allstats, stats.contiguous(), group=group)
    allstats = allstats.view(world, 2, V, g)
    means, m2s = allstats[:, 0], allstats[:, 1]                            # [G, V, g]
    mean = means.mean(dim=0)
    m2 = m2s.sum(dim=0) + n_local * (means - mean).square().sum(dim=0)
    rstd = torch.rsqrt(m2 / (n_local * world) + norm.eps)
    y = (xf - mean.view(V, 1, g, 1, 1)) * rstd.view(V, 1, g, 1, 1)
    y = y.view(V, f, C, h, w)
    if norm.affine:
        y = y * norm.weight.float().view(1, 1, C, 1, 1) + norm.bias.float().view(1, 1, C, 1, 1)
    return y.to(x.dtype)


# ----------------------------------------------------------------------------------------------------------------
# frame partitioner
# ----------------------------------------------------------------------------------------------------------------
class FramePartitioner:
    """Installs the frame-sharded execution of the three cross-frame couplings into a UNet (the reference's
    ``UNetMotionCrossFrameAttnModel`` or the host mirror) whose ranks each hold ``F / G`` frames of every video.

        part = FramePartitioner(unet, group=None)     # after install(unet) if the B200 processors are used
        part.install()
        local = part.shard_frames(latents)            # [B, F, 4, h, w] -> [B, F/G, 4, h, w]
        noise = unet(local, t, enable_cross_frame_attn=True, ...)
        full = part.gather_frames(noise.sample)

    The rank owning global frame 0 is group rank 0 (frames are dealt out in order)."""

    def __init__(self, unet: nn.Module, group=None, allow_torch_layout: bool = False):
        self.unet = unet
        self.group = group
        self.allow_torch_layout = allow_torch_layout
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.root = dist.get_global_rank(group, 0) if group is not None else 0
        self._undo: List[Callable[[], None]] = []
        self.stats = {"broadcasts": 0, "all_to_alls": 0, "stat_gathers": 0}
        self._timing: Optional[Dict[str, list]] = None   # bench.py: CUDA-event pairs around every collective

    # -- collective timing (bench.py's frame-sharded leg) -------------------------------------------------------
    def start_timing(self) -> None:
        self._timing = {"broadcast": [], "all_to_all": [], "stat_gather": []}

    def _collective(self, kind: str, fn: Callable[[], Any], nbytes: float, buffer_bytes: int = 0):
        """Run one collective; when timing is on, bracket it with events on the current stream (the NCCL work is
        stream-ordered with it) and remember the bytes this rank moves."""
        if self._timing is None or not torch.cuda.is_available():
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        self._timing[kind].append((e0, e1, float(nbytes), int(buffer_bytes)))
        return out

    def _isolated_ms(self, kind: str, buffer_bytes: int, device, reps: int = 5) -> float:
        """One collective of this kind and size on its own, ranks synchronised, back to back: the NCCL / NVLink time
        without the skew between ranks that the in-step brackets include."""
        n = max(buffer_bytes // 2, self.world)
        n -= n % self.world
        a = torch.zeros(n, dtype=torch.bfloat16, device=device)
        b = torch.empty_like(a)
        if kind == "all_to_all":
            fn = lambda: dist.all_to_all_single(b, a, group=self.group)  # noqa: E731
        elif kind == "broadcast":
            fn = lambda: dist.broadcast(a, src=self.root, group=self.group)  # noqa: E731
        else:
            small = torch.zeros(max(buffer_bytes // 4, 1), dtype=torch.float32, device=device)
            fn = lambda: dist.all_reduce(small, group=self.group)  # noqa: E731
        for _ in range(2):
            fn()
        dist.barrier(group=self.group)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def timing_summary(self, steps: int) -> Dict[str, Any]:
        """Per step and per kind: calls, device ms inside the collectives (includes waiting for the slowest rank),
        bytes this rank sends, and the resulting GB/s."""
        torch.cuda.synchronize()
        out: Dict[str, Any] = {}
        timing, self._timing = self._timing or {}, None
        device = torch.device("cuda", torch.cuda.current_device())
        for kind, pairs in timing.items():
            ms = sum(a.elapsed_time(b) for a, b, _, _ in pairs)
            nbytes = sum(n for _, _, n, _ in pairs)
            # the same collectives on their own (per distinct size, ranks synchronised): what NVLink / NCCL take
            iso_cache: Dict[int, float] = {}
            iso = 0.0
            for _, _, _, buf in pairs:
                if buf not in iso_cache:
                    iso_cache[buf] = self._isolated_ms(kind, buf, device)
                iso += iso_cache[buf]
            out[kind] = dict(calls_per_step=len(pairs) / max(steps, 1), in_step_ms_per_step=ms / max(steps, 1),
                             bytes_sent_per_step_per_rank=nbytes / max(steps, 1),
                             isolated_ms_per_step=iso / max(steps, 1),
                             isolated_gb_per_s_per_rank=(nbytes / 1e9) / (iso / 1e3) if iso > 0 else None)
        out["note"] = ("in_step: event brackets around each collective inside eager steps (include waiting for the slowest "
                       "rank); isolated: the same calls replayed alone with synchronised ranks")
        return out

    # -- frame axis helpers ------------------------------------------------------------------------------------
    @property
    def owns_first_frame(self) -> bool:
        return self.rank == 0

    def shard_frames(self, x: torch.Tensor, dim: int = 1) -> torch.Tensor:
        F = x.shape[dim]
        if F % self.world:
            raise ValueError(f"{F} frames are not divisible by {self.world} ranks")
        f = F // self.world
        return x.narrow(dim, self.rank * f, f)

    def gather_frames(self, x: torch.Tensor, dim: int = 1) -> torch.Tensor:
        parts = [torch.empty_like(x) for _ in range(self.world)]
        dist.all_gather(parts, x.contiguous(), group=self.group)
        return torch.cat(parts, dim=dim)

    # -- cross-frame attention: frame-0 tensors come from the owner ---------------------------------------------
    def broadcast_from_first_frame_owner(self, compute: Callable[[], torch.Tensor], shape, dtype, device):
        """Owner: ``t = compute()``; others: an empty tensor of ``shape``.  One broadcast, returns ``t`` everywhere.
        Used for the projected frame-0 K/V (B200 processors) or frame 0's normalised hidden states (stock ones)."""
        t = compute().contiguous() if self.owns_first_frame else torch.empty(shape, dtype=dtype, device=device)
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"first-frame tensor has shape {tuple(t.shape)}, expected {tuple(shape)}")
        self._collective("broadcast", lambda: dist.broadcast(t, src=self.root, group=self.group),
                         t.numel() * t.element_size() if self.owns_first_frame else 0, t.numel() * t.element_size())
        self.stats["broadcasts"] += 1
        return t

    # -- motion module ------------------------------------------------------------------------------------------
    def temporal_forward(self, module: nn.Module, hidden_states: torch.Tensor, num_frames: int = 1, **kw):
        """Frame-sharded ``TransformerTemporalModel.forward``: same arithmetic, the frame
        axis is completed by an all-to-all so each rank runs the temporal transformer on ``S / G`` positions."""
        bf, C, h, w = hidden_states.shape
        f = num_frames
        if bf % f:
            raise ValueError(f"Batch size {bf} must be divisible by the number of frames {f}.")
        V, S, G = bf // f, h * w, self.world
        if self._library_path_ok(module, hidden_states, S):
            return self._temporal_forward_library(module, hidden_states, f, **kw)
        residual = hidden_states
        x = sharded_group_norm(hidden_states.view(V, f, C, h, w), module.norm, self.group)
        self.stats["stat_gathers"] += 1
        tokens = x.permute(0, 1, 3, 4, 2).reshape(V, f, S, C)
        full = frames_to_positions(tokens, self.group, self.allow_torch_layout)          # [V, F, S/G, C]
        self.stats["all_to_alls"] += 1
        t = full.permute(0, 2, 1, 3).reshape(V * (S // G), f * G, C)
        t = module.proj_in(t)
        for block in module.transformer_blocks:
            t = block(t, encoder_hidden_states=kw.get("encoder_hidden_states"), timestep=kw.get("timestep"),
                      cross_attention_kwargs=kw.get("cross_attention_kwargs"), class_labels=kw.get("class_labels"))
        t = module.proj_out(t)
        back = t.view(V, S // G, f * G, C).permute(0, 2, 1, 3).contiguous()               # [V, F, S/G, C]
        local = positions_to_frames(back, self.group, self.allow_torch_layout)            # [V, f, S, C]
        self.stats["all_to_alls"] += 1
        out = local.view(V, f, h, w, C).permute(0, 1, 4, 2, 3).reshape(bf, C, h, w) + residual
        if kw.get("return_dict", True):
            from .hostmodel.i2v_adapter import _Sample
            return _Sample(out)
        return (out,)

    def _library_path_ok(self, module: nn.Module, x: torch.Tensor, S: int) -> bool:
        norm = module.norm
        return (x.is_cuda and x.dtype == torch.bfloat16 and ops.is_channels_last(x) and not x.is_contiguous()
                and isinstance(norm, nn.GroupNorm) and norm.affine and norm.weight.dtype == x.dtype
                and x.shape[1] % 8 == 0 and x.shape[1] <= 4096 and S % self.world == 0
                and isinstance(module.proj_in, nn.Linear) and not torch.is_grad_enabled())

    def _temporal_forward_library(self, module: nn.Module, hidden_states: torch.Tensor, f: int, **kw):
        """The same exchange on the library's channels-last kernels: raw GroupNorm sums -> all-reduce of
        ``2 * V * groups`` floats -> the apply pass writes the all-to-all send buffer ``[G, V, S/G, f, C]`` directly
        (position-major rows, no pack pass) -> one row permutation into ``[V * S/G, F, C]`` -> temporal transformer
        -> inverse permutation -> all-to-all -> the residual pass reads the receive buffer in place."""
        from .fastpath import OWNED_INPUT_FLAG

        bf, C, h, w = hidden_states.shape
        V, S, G = bf // f, h * w, self.world
        Sl = S // G
        norm = module.norm
        groups = norm.num_groups
        sums = ops.group_norm_nhwc_sums(hidden_states, groups, f)                         # [V, groups, 2] fp32
        self._collective("stat_gather", lambda: dist.all_reduce(sums, group=self.group), sums.numel() * 4, sums.numel() * 4)
        self.stats["stat_gathers"] += 1
        cnt = float(f * G) * float(C // groups) * float(S)
        mean = sums[..., 0] / cnt
        rstd = torch.rsqrt((sums[..., 1] / cnt - mean * mean).clamp_min_(0.0) + norm.eps)
        send = ops.group_norm_nhwc_apply(hidden_states, norm.weight, norm.bias,
                                         torch.stack([mean, rstd], dim=-1).contiguous(), groups, f, world=G)
        recv = torch.empty_like(send)                                                     # [G, V, Sl, f, C]
        nbytes = send.numel() * send.element_size() * (G - 1) / G
        self._collective("all_to_all", lambda: dist.all_to_all_single(recv, send, group=self.group), nbytes,
                         send.numel() * send.element_size())
        self.stats["all_to_alls"] += 1
        t = ops.reshard_unpack(recv.view(G, V * Sl, f, 1, C), G).view(V * Sl, G * f, C)   # every frame, my positions
        t = module.proj_in(t)
        for block in module.transformer_blocks:
            block.__dict__[OWNED_INPUT_FLAG] = True
            t = block(t, encoder_hidden_states=kw.get("encoder_hidden_states"), timestep=kw.get("timestep"),
                      cross_attention_kwargs=kw.get("cross_attention_kwargs"), class_labels=kw.get("class_labels"))
            block.__dict__.pop(OWNED_INPUT_FLAG, None)
        t = module.proj_out(t)
        send2 = ops.reshard_unpack(t.view(V * Sl, G * f, 1, C).contiguous(), G, inverse=True)   # [G, V*Sl, f, 1, C]
        recv2 = torch.empty_like(send2)
        self._collective("all_to_all", lambda: dist.all_to_all_single(recv2, send2, group=self.group), nbytes,
                         send2.numel() * send2.element_size())
        self.stats["all_to_alls"] += 1
        out = ops.sharded_positions_to_nhwc_residual(recv2, hidden_states, f, G)
        if kw.get("return_dict", True):
            from .hostmodel.i2v_adapter import _Sample
            return _Sample(out)
        return (out,)

    # -- installation -------------------------------------------------------------------------------------------
    def install(self) -> "FramePartitioner":
        from .processors import B200CrossFrameAttnProcessor, B200SpatialAttnProcessor

        for name, module in self.unet.named_modules():
            if type(module).__name__ == "TransformerTemporalModel":
                self._patch_temporal(module)
            elif hasattr(module, "i2v_adapter") and hasattr(module, "attn1"):
                p1 = module.attn1.get_processor() if hasattr(module.attn1, "get_processor") else None
                px = module.i2v_adapter.get_processor() if hasattr(module.i2v_adapter, "get_processor") else None
                if isinstance(p1, B200SpatialAttnProcessor) or isinstance(px, B200CrossFrameAttnProcessor):
                    for p in (p1, px):
                        if getattr(p, "state", None) is not None:
                            p.state.first_frame_source = self
                            self._undo.append(lambda st=p.state: setattr(st, "first_frame_source", None))
                else:
                    self._hook_stock_cross_frame(module.i2v_adapter)
        return self

    def uninstall(self) -> None:
        for fn in reversed(self._undo):
            fn()
        self._undo = []

    def _patch_temporal(self, module: nn.Module) -> None:
        original = module.forward

        def sharded(hidden_states, encoder_hidden_states=None, timestep=None, class_labels=None, num_frames=1,
                    cross_attention_kwargs=None, return_dict=True):
            return self.temporal_forward(module, hidden_states, num_frames, encoder_hidden_states=encoder_hidden_states,
                                         timestep=timestep, class_labels=class_labels,
                                         cross_attention_kwargs=cross_attention_kwargs, return_dict=return_dict)

        module.forward = sharded
        self._undo.append(lambda: setattr(module, "forward", original))

    def _hook_stock_cross_frame(self, attn: nn.Module) -> None:
        """Stock processors: the block passes ``encoder_hidden_states`` = its *local* first frame repeated f times
        (``src/modules/i2v_adapter.py:484-485``); replace it by the owner's frame 0."""

        def pre(module, args, kwargs):
            ctx = kwargs.get("encoder_hidden_states")
            if ctx is None:
                return None
            hidden = args[0] if args else kwargs["hidden_states"]
            bf = hidden.shape[0]
            videos = self._videos_hint(bf)
            f = bf // videos
            first = self.broadcast_from_first_frame_owner(lambda: ctx[0::f], (videos,) + tuple(ctx.shape[1:]), ctx.dtype,
                                                          ctx.device)
            kwargs = dict(kwargs)
            kwargs["encoder_hidden_states"] = first.repeat_interleave(f, dim=0)
            return args, kwargs

        h = attn.register_forward_pre_hook(pre, with_kwargs=True)
        self._undo.append(h.remove)

    # the stock-processor hook sees only [B*f, S, C]; the UNet-level hook below records f for it
    _frames_local: Optional[int] = None

    def _videos_hint(self, bf: int) -> int:
        f = self._frames_local
        if not f or bf % f:
            raise ValueError("FramePartitioner: call unet(sample[B, f, 4, h, w]) (5-D sample) so the local frame count "
                             "is known to the cross-frame hook")
        return bf // f

    def watch_unet_forward(self) -> "FramePartitioner":
        """Record the local frame count from the UNet's 5-D ``sample`` (needed by the stock-processor hook only)."""
        import inspect

        def pre(module, args, kwargs):
            try:
                bound = inspect.signature(module.forward).bind_partial(*args, **kwargs).arguments
            except TypeError:
                bound = kwargs
            sample = bound.get("sample")
            if sample is not None and sample.dim() == 5:
                self._frames_local = int(sample.shape[1])

        h = self.unet.register_forward_pre_hook(pre, with_kwargs=True)
        self._undo.append(h.remove)
        return self


def sharded_denoise_step(part: FramePartitioner, unet, scheduler, latents_local, t, prompt_embeds,
                         guidance_scale: float = 7.5, condition_image_latents=None, image_embeds=None):
    """``hostmodel.denoise_step`` on a frame shard: identical except that the first-frame re-imposition
    (pipeline ``:668-669``) happens only on the rank that owns frame 0."""
    from .hostmodel.pipeline import denoise_step

    return denoise_step(unet, scheduler, latents_local, t, prompt_embeds, guidance_scale, condition_image_latents,
                        image_embeds, impose_first_frame=part.owns_first_frame)
