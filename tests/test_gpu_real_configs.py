"""Parity at BASELINE.json's real configurations (-m gpu): configs[0] exactly, the 25-step DDIM gate at the SD1.5
widths (the kernels that carry the benchmark: d = 40 pipelined / augmented, d = 80 / 160, IP-Adapter streaming, temporal,
fused feed-forward GEMM), the augmented-layout entry at the full level-0 sizes of configs[1] and configs[3], the
frame-sharded GroupNorm kernels, and the library's launch timing.

Tolerances are BASELINE.json's: max-abs <= 2e-2 on bf16 attention outputs, cosine >= 0.999 on the final latent."""
import pytest
import torch
import torch.nn.functional as F

from helpers import fake_ip_adapter_state_dict, randomize_zero_init
from i2v_adapter_unofficial_b200 import _lib, fastpath, install, ops
from i2v_adapter_unofficial_b200.hostmodel import (
    CrossFrameAttnDownBlockMotion,
    DDIMScheduler,
    IPAdapterAttnProcessor2_0,
    UNetMotionCrossFrameAttnModel,
    denoise,
)
from oracle.attention_oracle import attention_oracle, ip_adapter_attention_oracle
from oracle.unet_oracle import down_block_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF16_TOL = 2e-2
SD15 = dict(block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, cross_attention_dim=768,
            num_attention_heads=8, motion_num_attention_heads=8, motion_max_seq_length=32, norm_num_groups=32)


def _bf16(x):
    return x.to(DEV, torch.bfloat16)


def _nonzero_adapter_out(module, seed=9, std=0.02):
    """The I2V-Adapter's output projection is zero-initialised (reference src/modules/i2v_adapter.py:142-143): give it
    values so the cross-frame branch contributes to what is compared."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if ".i2v_adapter.to_out.0." in name:
                p.copy_(torch.randn(p.shape, generator=g) * std)
    return module


# ---------------------------------------------------------------------------------------------------------------
# configs[0]: one SD1.5 level-0 down block, 1 x 16 frames of 32 x 32 latents (test/test_unet_motion_cross_frame_attn.py)
# ---------------------------------------------------------------------------------------------------------------
def _c1_block():
    torch.manual_seed(0)
    block = CrossFrameAttnDownBlockMotion(in_channels=320, out_channels=320, temb_channels=1280, num_layers=2,
                                          resnet_eps=1e-5, resnet_groups=32, num_attention_heads=8,
                                          cross_attention_dim=768, add_downsample=True,
                                          temporal_num_attention_heads=8, temporal_max_seq_length=32).eval()
    for tr in block.attentions:
        for tb in tr.transformer_blocks:
            tb.attn2.set_processor(IPAdapterAttnProcessor2_0(320, 768, num_tokens=4, scale=1.0))
    return _nonzero_adapter_out(block)


def test_config1_down_block_matches_cpu_oracle():
    """BASELINE.json configs[0]: 320 channels, 8 heads (d = 40), 2 layers, 1 video x 16 frames, 32 x 32 latent, time
    embedding 1280, 77 text (+ 4 image) tokens of width 768.

    (1) processors only: the output of every attention call (attn1 + i2v_adapter, attn2, temporal attn1 / attn2) is
        compared with the oracle's attention on the *same* inputs (captured from the run): max-abs <= 2e-2.
    (2) whole block with the module-level fast path vs the fp32 CPU oracle: cosine >= 0.999 and max-abs within
        3e-2 of the output range (bf16 storage of 12 residual sub-layers)."""
    block = _c1_block()
    V, Fr, hw = 1, 16, 32
    g = torch.Generator().manual_seed(1)
    x = torch.randn(V * Fr, 320, hw, hw, generator=g)
    temb = torch.randn(V * Fr, 1280, generator=g)
    ctx = torch.randn(V, 77 + 4, 768, generator=g).repeat_interleave(Fr, dim=0)
    sd = {f"blk.{k}": v.float() for k, v in block.state_dict().items()}
    cfg = dict(norm_num_groups=32, norm_eps=1e-5, motion_num_attention_heads=8)
    with torch.no_grad():
        ref, _ = down_block_oracle(sd, "blk", x, temb, ctx, cfg, 8, True, Fr, 4, 1.0)

    block = block.to(DEV, torch.bfloat16)
    # (1) per-attention outputs, processors only
    handle = install(block, fast_path=False)
    captured = []

    def hook(name):
        def fn(module, args, kwargs, output):
            captured.append((name, args[0].detach().float().cpu(),
                             None if kwargs.get("encoder_hidden_states") is None
                             else kwargs["encoder_hidden_states"].detach().float().cpu(),
                             output.detach().float().cpu()))
        return fn

    hooks = []
    for name, m in block.named_modules():
        if name.endswith(("attn1", "attn2", "i2v_adapter")):
            hooks.append(m.register_forward_hook(hook(name), with_kwargs=True))
    n0 = _lib.launch_count()
    with torch.no_grad():
        block(_bf16(x), _bf16(temb), enable_cross_frame_attn=True, encoder_hidden_states=_bf16(ctx), num_frames=Fr)
    torch.cuda.synchronize()
    assert _lib.launch_count() - n0 >= 2 * (2 + 2)   # per layer: fused spatial, IP, two temporal
    for h in hooks:
        h.remove()
    handle.uninstall()
    by_name = {}
    for name, xin, enc, out in captured:
        by_name.setdefault(name, []).append((xin, enc, out))
    checked = 0
    with torch.no_grad():
        for name, calls in by_name.items():
            for xin, enc, out in calls:
                pre = f"blk.{name}"
                if name.endswith("i2v_adapter"):
                    continue   # contributes zero: attn1's processor returned self + cross-frame in one launch
                if ".motion_modules." in f".{name}":
                    want = attention_oracle(sd, pre, xin, None, 8)
                elif name.endswith("attn1"):
                    first = xin[0::Fr].repeat_interleave(Fr, dim=0)
                    want = attention_oracle(sd, pre, xin, None, 8) + \
                        attention_oracle(sd, pre[:-len("attn1")] + "i2v_adapter", xin, first, 8)
                else:
                    want = ip_adapter_attention_oracle(sd, pre, xin, enc, 8, 4, 1.0)
                err = (out - want).abs().max().item()
                assert err <= BF16_TOL, (name, err)
                checked += 1
    assert checked == 2 * 4   # two layers x (attn1 + attn2 + temporal attn1 + temporal attn2)

    # (2) the whole block on the fast path
    handle = install(block)
    fastpath.reset_fallback_counts()
    with torch.no_grad():   # channels-last input, as the UNet's conv_in hands it to its first block
        out, _ = block(_bf16(x).contiguous(memory_format=torch.channels_last), _bf16(temb),
                       enable_cross_frame_attn=True, encoder_hidden_states=_bf16(ctx), num_frames=Fr)
    out = out.float().cpu()
    assert fastpath.fallback_counts() == {}
    handle.uninstall()
    cos = F.cosine_similarity(out.flatten(), ref.flatten(), dim=0).item()
    err, scale = (out - ref).abs().max().item(), ref.abs().max().item()
    assert cos >= 0.999, cos
    assert err <= 3e-2 * scale, (err, scale)


# ---------------------------------------------------------------------------------------------------------------
# 25-step DDIM at the SD1.5 widths
# ---------------------------------------------------------------------------------------------------------------
def test_25_step_ddim_at_sd15_widths_cosine():
    """BASELINE.json's end-to-end gate on the target architecture: full SD1.5 UNetMotion (320/640/1280/1280, two layers
    per block, 8 heads -> d = 40/80/160) + motion modules + I2V-Adapter + IP-Adapter, 16 frames, 32 x 32 latent, CFG
    7.5, 25 DDIM steps with first-frame re-imposition (pipeline :666-697).  B200 processors + fast path in bf16
    against the stock SDPA processors in fp32 on the same GPU and weights: cosine >= 0.999 on the final latent."""
    torch.manual_seed(0)
    unet = UNetMotionCrossFrameAttnModel(**SD15).eval()
    unet._load_ip_adapter_weights(fake_ip_adapter_state_dict(unet, 3, image_embed_dim=64))
    unet = _nonzero_adapter_out(randomize_zero_init(unet))
    g = torch.Generator().manual_seed(2)
    frames, size = 16, 32
    sample = torch.randn(1, frames, 4, size, size, generator=g)
    ctx = torch.randn(1, 77, 768, generator=g)
    img = torch.randn(1, 64, generator=g)
    ctx2 = torch.cat([torch.randn(1, 77, 768, generator=g), ctx])
    img2 = torch.cat([torch.zeros_like(img), img])
    cond = torch.randn(1, 4, size, size, generator=g)

    def run(dtype, b200):
        model = unet.to(DEV, dtype)
        handle = install(model) if b200 else None
        if b200:
            fastpath.reset_fallback_counts()
        n0 = _lib.launch_count()
        out = denoise(model, DDIMScheduler(), sample.clone().to(DEV, dtype), ctx2.to(DEV, dtype), 25, 7.5,
                      cond.to(DEV, dtype), img2.to(DEV, dtype)).float().cpu()
        if b200:
            assert _lib.launch_count() - n0 > 25 * 100 and fastpath.fallback_counts() == {}
            handle.uninstall()
        return out

    ref32 = run(torch.float32, False)     # the reference's processors, fp32 (the pipeline's default dtype)
    stock16 = run(torch.bfloat16, False)  # the same processors in bf16: what bf16 storage alone costs
    ours16 = run(torch.bfloat16, True)
    cos = lambda a, b: F.cosine_similarity(a.flatten(), b.flatten(), dim=0).item()  # noqa: E731
    c_ours, c_stock = cos(ours16, ref32), cos(stock16, ref32)
    print(f"25-step DDIM at SD1.5 widths: cosine ours-bf16 vs fp32 {c_ours:.6f}, stock-bf16 vs fp32 {c_stock:.6f}, "
          f"ours vs stock bf16 {cos(ours16, stock16):.6f}")
    assert torch.isfinite(ours16).all()
    assert c_ours >= 0.999, (c_ours, c_stock)


# ---------------------------------------------------------------------------------------------------------------
# the augmented-layout entry (the kernel of bench.py's roofline line) at the full level-0 sizes
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [(2, 16, 4096), (1, 2, 9216)], ids=["c2_BF32_S4096", "c4_B2_S9216"])
def test_fused_augmented_layout_full_sizes_vs_sdpa(case):
    """`i2v_fused_self_xframe_aug_fwd` at configs[1] level 0 (32 frames x 8 heads x 4096 tokens) and at the configs[3]
    sequence length (9216 tokens) against torch's own SDPA in fp32 on the same device, every frame and head."""
    V, Fr, S = case
    H, d = 8, 40
    BF = V * Fr
    g = torch.Generator(device=DEV).manual_seed(21)
    mk = lambda b: torch.randn(b, S, H, d, device=DEV, generator=g).to(torch.bfloat16)  # noqa: E731
    q, k, v, qx, kx, vx = mk(BF), mk(BF), mk(BF), mk(BF), mk(V), mk(V)
    qa, ka, va = ops.augment_qkv(q, k, v)
    qxa, kxa, vxa = ops.augment_qkv(qx, kx, vx)
    o = ops.fused_self_xframe_aug(qa, ka, va, qxa, kxa, vxa, Fr)
    t = lambda x: x.permute(0, 2, 1, 3).float()  # noqa: E731
    worst_s = worst_x = 0.0
    for b in range(BF):
        ref_s = F.scaled_dot_product_attention(t(q[b:b + 1]), t(k[b:b + 1]), t(v[b:b + 1])).permute(0, 2, 1, 3)
        ref_x = F.scaled_dot_product_attention(t(qx[b:b + 1]), t(kx[b // Fr:b // Fr + 1]),
                                               t(vx[b // Fr:b // Fr + 1])).permute(0, 2, 1, 3)
        worst_s = max(worst_s, (o[b:b + 1, :, 0].float() - ref_s).abs().max().item())
        worst_x = max(worst_x, (o[b:b + 1, :, 1].float() - ref_x).abs().max().item())
    assert worst_s <= BF16_TOL and worst_x <= BF16_TOL, (worst_s, worst_x)


# ---------------------------------------------------------------------------------------------------------------
# frame-sharded GroupNorm kernels (one GPU plays the ranks in turn)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [(2, 4, 2, 320, 16, 16), (1, 8, 4, 640, 8, 8), (2, 2, 2, 1280, 4, 8)],
                         ids=lambda c: "V{}F{}W{}C{}h{}w{}".format(*c))
def test_sharded_group_norm_kernels_match_unsharded(case):
    """i2v_gn_nhwc_sums / i2v_gn_nhwc_apply (perm = 2) / i2v_rows_residual_sharded: the ranks' raw sums added up give
    the unsharded statistics, the send buffers concatenated give the unsharded token layout, and the way back
    reproduces `positions_to_nhwc_residual`."""
    V, Fall, W, C, h, w = case
    f, S, G = Fall // W, h * w, 32
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(V, Fall, C, h, w, generator=g) * 2 + 0.5).to(DEV, torch.bfloat16)
    wgt = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV, torch.bfloat16)
    bias = (0.1 * torch.randn(C, generator=g)).to(DEV, torch.bfloat16)
    full = x.view(V * Fall, C, h, w).contiguous(memory_format=torch.channels_last)
    want = ops.group_norm_nhwc(full, wgt, bias, G, 1e-6, Fall, to_positions=True)        # [V*S, Fall, C]
    shards = [x[:, r * f:(r + 1) * f].reshape(V * f, C, h, w).contiguous(memory_format=torch.channels_last)
              for r in range(W)]
    sums = sum(ops.group_norm_nhwc_sums(s, G, f) for s in shards)                        # the all-reduce
    cnt = float(Fall) * (C // G) * S
    mean = sums[..., 0] / cnt
    rstd = torch.rsqrt((sums[..., 1] / cnt - mean * mean).clamp_min(0) + 1e-6)
    stats = torch.stack([mean, rstd], -1).contiguous()
    sends = [ops.group_norm_nhwc_apply(s, wgt, bias, stats, G, f, world=W) for s in shards]   # [W, V, S/W, f, C] each
    Sl = S // W
    for dst in range(W):   # what rank `dst` holds after the all-to-all and the row permutation
        recv = torch.stack([sends[src][dst] for src in range(W)])                           # [W_src, V, Sl, f, C]
        t = ops.reshard_unpack(recv.view(W, V * Sl, f, 1, C), W).view(V, Sl, Fall, C)
        ref = want.view(V, S, Fall, C)[:, dst * Sl:(dst + 1) * Sl]
        assert (t.float() - ref.float()).abs().max().item() <= 2 ** -6 * max(1.0, ref.float().abs().max().item())
    # way back: y [V*S, Fall, C] -> per rank the receive buffer [W, V, Sl, f, C] -> + residual
    y = torch.randn(V * S, Fall, C, generator=g).to(DEV, torch.bfloat16)
    want_back = ops.positions_to_nhwc_residual(y, full, Fall)                               # (V*Fall, C, h, w)
    for r in range(W):
        yr = y.view(V, S, Fall, C)[:, :, r * f:(r + 1) * f]                                 # frames of rank r
        recv2 = torch.stack([yr[:, gs * Sl:(gs + 1) * Sl] for gs in range(W)]).contiguous()  # [W, V, Sl, f, C]
        got = ops.sharded_positions_to_nhwc_residual(recv2, shards[r], f, W)
        ref = want_back.view(V, Fall, C, h, w)[:, r * f:(r + 1) * f].reshape(V * f, C, h, w)
        assert torch.equal(got, ref)


def test_library_launch_timing_brackets_the_kernel():
    """i2v_prof_arm / i2v_prof_read: eager launches and launches replayed from a CUDA graph."""
    H, d, S, B = 8, 40, 1024, 4
    q = torch.randn(B, S, H, d, device=DEV).to(torch.bfloat16)
    ops.sdpa(q, q, q, 1, None, ops.MODE_FAST)
    torch.cuda.synchronize()
    _lib.prof_arm(_lib.PROF_DENSE, S, B, 8)
    for _ in range(3):
        ops.sdpa(q, q, q, 1, None, ops.MODE_FAST)
    ops.sdpa(q[:2], q[:2], q[:2], 1, None, ops.MODE_FAST)     # other batch: not matched
    torch.cuda.synchronize()
    ms = _lib.prof_read(_lib.PROF_DENSE)
    assert len(ms) == 3 and all(0.0 < m < 50.0 for m in ms)
    # captured: the pairs become external event nodes, re-recorded by each replay
    _lib.prof_arm(_lib.PROF_DENSE, S, B, 8)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        o = ops.sdpa(q, q, q, 1, None, ops.MODE_FAST)
        o2 = ops.sdpa(q, q, o, 1, None, ops.MODE_FAST)
    _lib.prof_arm(_lib.PROF_DENSE, 0, 0, 0)
    ops.sdpa(q, q, q, 1, None, ops.MODE_FAST)                 # disarmed: takes no pair
    for _ in range(2):
        graph.replay()
        torch.cuda.synchronize()
        ms = _lib.prof_read(_lib.PROF_DENSE)
        assert len(ms) == 2 and all(0.0 < m < 50.0 for m in ms), ms
    assert torch.isfinite(o2.float()).all()


# ---------------------------------------------------------------------------------------------------------------
# IP-Adapter attention on the tcgen05 kernel (d = 40, 77 + 4 tokens)
# ---------------------------------------------------------------------------------------------------------------
def _ip_ref(q, k, v, nt, ip_scale, g):
    t = lambda x: x.permute(0, 2, 1, 3).float()  # noqa: E731
    rep = lambda x: t(x).repeat_interleave(g, 0)  # noqa: E731
    a = F.scaled_dot_product_attention(t(q), rep(k[:, :nt]), rep(v[:, :nt]))
    b = F.scaled_dot_product_attention(t(q), rep(k[:, nt:]), rep(v[:, nt:]))
    return (a + ip_scale * b).permute(0, 2, 1, 3)


@pytest.mark.parametrize("case", [(32, 4096, 8, 16, 1.0), (2, 77, 8, 1, 0.5), (6, 129, 2, 3, 2.0), (160, 128, 8, 1, 1.0)],
                         ids=["c2_level0", "one_ragged_tile", "two_heads_three_frames", "more_groups_than_sms"])
def test_ip_adapter_tcgen05_kernel(case):
    """`i2v_ip_xattn_fwd` at d = 40 with the pipeline's 77 + 4 tokens: the tcgen05 kernel with K / V resident per (video,
    head) at the full configs[1] level-0 size, on ragged / tiny query tiles, and the dispatch back to the streaming kernel
    when there are more (video, head) groups than SMs -- against torch SDPA in fp32 and against the streaming kernel."""
    B, S, H, g, ip_scale = case
    d, nt, ni = 40, 77, 4
    gen = torch.Generator(device=DEV).manual_seed(B + S)
    q = (torch.randn(B, S, H, d, device=DEV, generator=gen) * 1.5).to(torch.bfloat16)
    k = (torch.randn(B // g, nt + ni, H, d, device=DEV, generator=gen) * 1.5).to(torch.bfloat16)
    v = (torch.randn(B // g, nt + ni, H, d, device=DEV, generator=gen) * 3.0).to(torch.bfloat16)   # ones columns vs large V
    o = ops.ip_xattn(q, k, v, nt, ip_scale, g, None, ops.MODE_FAST)
    ref = _ip_ref(q, k, v, nt, ip_scale, g)
    scale = max(1.0, ref.abs().max().item())
    assert (o.float() - ref).abs().max().item() <= BF16_TOL * scale
    lib = _lib.load()
    lib.i2v_set_tuning(5, 4)      # the streaming kernel on the same operands
    try:
        o2 = ops.ip_xattn(q, k, v, nt, ip_scale, g, None, ops.MODE_FAST)
    finally:
        lib.i2v_set_tuning(5, 0)
    assert (o.float() - o2.float()).abs().max().item() <= BF16_TOL * scale


# ---------------------------------------------------------------------------------------------------------------
# level 1 (d = 80) on the pipelined kernel
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [(32, 1024, 16), (8, 2304, 4), (4, 1000, 2), (2, 130, 1)],
                         ids=["c2_level1", "c4_level1", "ragged", "tiny"])
def test_fused_self_xframe_d80_pipelined_kernel(case):
    """`i2v_fused_self_xframe_fwd` at d = 80: the pipelined kernel with two swizzle sub-tiles per row (the default since
    round 2) at the configs[1] / configs[3] level-1 sizes and on ragged query / key tiles, against torch SDPA in fp32 and
    against the first tcgen05 kernel (tuning key 3 = 1) on the same operands."""
    BF, S, Fr = case
    H, d = 8, 80
    gen = torch.Generator(device=DEV).manual_seed(BF * S)
    mk = lambda b: torch.randn(b, S, H, d, device=DEV, generator=gen).to(torch.bfloat16)  # noqa: E731
    q, k, v, qx, kx, vx = mk(BF), mk(BF), mk(BF), mk(BF), mk(BF // Fr), mk(BF // Fr)
    o = ops.fused_self_xframe(q, k, v, qx, kx, vx, Fr)
    t = lambda x: x.transpose(1, 2).float()  # noqa: E731
    n = min(BF, 4)   # the fp32 reference of a few frames is enough (and all of them for the small cases)
    ref_s = F.scaled_dot_product_attention(t(q[:n]), t(k[:n]), t(v[:n])).transpose(1, 2)
    ref_x = F.scaled_dot_product_attention(t(qx[-n:]), t(kx[(BF - n) // Fr:].repeat_interleave(Fr, 0)[-n:]),
                                           t(vx[(BF - n) // Fr:].repeat_interleave(Fr, 0)[-n:])).transpose(1, 2)
    assert (o[:n, :, 0].float() - ref_s).abs().max().item() <= BF16_TOL
    assert (o[-n:, :, 1].float() - ref_x).abs().max().item() <= BF16_TOL
    lib = _lib.load()
    lib.i2v_set_tuning(3, 1)
    try:
        o_old = ops.fused_self_xframe(q, k, v, qx, kx, vx, Fr)
    finally:
        lib.i2v_set_tuning(3, 0)
    assert (o.float() - o_old.float()).abs().max().item() <= BF16_TOL


@pytest.mark.parametrize("case", [(32, 320, 64, 64, 1, True), (32, 320, 32, 32, 16, False), (6, 1280, 8, 8, 1, True)],
                         ids=["resnet_level0", "motion_module_frames", "mid_block"])
def test_group_norm_fused_finalize_matches_the_separate_launch(case):
    """The one-call channels-last GroupNorm lets every apply CTA reduce the partial sums itself (one launch less); the
    reduction uses the lane assignment and shuffle tree of `gn_finalize_kernel`, so the result must equal the
    three-launch form (tuning key 9 = 1) bit for bit -- the shared atomics of the statistics pass make two runs differ
    in the last bits of the partials, so both forms are compared against fp32 GroupNorm instead when they do."""
    N, C, h, w, fg, silu = case
    gen = torch.Generator(device=DEV).manual_seed(N * C)
    x = (torch.randn(N, C, h, w, device=DEV, generator=gen) * 2 + 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wgt = torch.randn(C, device=DEV, generator=gen).to(torch.bfloat16)
    b = torch.randn(C, device=DEV, generator=gen).to(torch.bfloat16)
    lib = _lib.load()
    y_fused = ops.group_norm_nhwc(x, wgt, b, 32, 1e-5, fg, silu=silu)
    lib.i2v_set_tuning(9, 1)
    try:
        y_sep = ops.group_norm_nhwc(x, wgt, b, 32, 1e-5, fg, silu=silu)
    finally:
        lib.i2v_set_tuning(9, 0)
    V = N // fg
    ref = F.group_norm(x.float().view(V, fg, C, h, w).transpose(1, 2).reshape(V, C, fg * h, w), 32, wgt.float(), b.float(), 1e-5)
    ref = ref.view(V, C, fg, h, w).transpose(1, 2).reshape(N, C, h, w)
    if silu:
        ref = F.silu(ref.to(torch.bfloat16).float())
    scale = max(1.0, ref.abs().max().item())
    assert (y_fused.float() - ref).abs().max().item() <= BF16_TOL * scale
    assert (y_sep.float() - ref).abs().max().item() <= BF16_TOL * scale
    # same partials -> same statistics -> same output, up to the run-to-run jitter of the statistics pass
    assert (y_fused.float() - y_sep.float()).abs().max().item() <= 2 * 2.0 ** -8 * scale


# ---------------------------------------------------------------------------------------------------------------
# one UNet forward at the benchmark shapes themselves (every kernel at the size bench.py times it at, in the system)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [(16, 64), (16, 96)], ids=["configs1_64x64", "configs3_96x96"])
def test_unet_forward_at_benchmark_shapes_matches_fp32_stock_processors(case):
    """BASELINE.json configs[1] / configs[3] exactly as bench.py runs them (CFG batch of one video, 16 frames, 64 x 64 /
    96 x 96 latent, 77 text + 4 image tokens): one UNet forward with the B200 processors + fast path in bf16 against the
    stock SDPA processors in fp32 on the same GPU and weights.  Level 0 then runs at S = 4096 / 9216 (pipelined d = 40
    kernel, IP-Adapter tcgen05 kernel with 32 / 72 query tiles per frame), level 1 at S = 1024 / 2304 (pipelined d = 80
    kernel), the temporal kernel at 8192 / 18432 positions."""
    frames, size = case
    torch.manual_seed(0)
    unet = UNetMotionCrossFrameAttnModel(**SD15).eval()
    unet._load_ip_adapter_weights(fake_ip_adapter_state_dict(unet, 3, image_embed_dim=64))
    unet = _nonzero_adapter_out(randomize_zero_init(unet))
    g = torch.Generator().manual_seed(5)
    sample = torch.randn(2, frames, 4, size, size, generator=g)
    ctx = torch.randn(2, 77, 768, generator=g)
    img = torch.randn(2, 64, generator=g)

    def run(dtype, b200):
        model = unet.to(DEV, dtype)
        handle = install(model) if b200 else None
        if b200:
            fastpath.reset_fallback_counts()
        with torch.no_grad():
            out = model(sample.to(DEV, dtype), 481, enable_cross_frame_attn=True, encoder_hidden_states=ctx.to(DEV, dtype),
                        added_cond_kwargs={"image_embeds": img.to(DEV, dtype)}).sample.float().cpu()
        if b200:
            assert fastpath.fallback_counts() == {}
            handle.uninstall()
        torch.cuda.empty_cache()
        return out

    ref32 = run(torch.float32, False)
    stock16 = run(torch.bfloat16, False)
    ours16 = run(torch.bfloat16, True)
    cos = lambda a, b: F.cosine_similarity(a.flatten(), b.flatten(), dim=0).item()  # noqa: E731
    e_ours = (ours16 - ref32).abs().max().item() / ref32.abs().max().item()
    e_stock = (stock16 - ref32).abs().max().item() / ref32.abs().max().item()
    print(f"UNet forward {frames} f {size} x {size}: cosine ours-bf16 vs fp32 {cos(ours16, ref32):.6f} (stock bf16 {cos(stock16, ref32):.6f}), "
          f"max-abs / range ours {e_ours:.4f} stock {e_stock:.4f}")
    assert torch.isfinite(ours16).all()
    assert cos(ours16, ref32) >= 0.999
    assert e_ours <= max(2e-2, 2.0 * e_stock)   # no worse than twice what bf16 storage alone costs the stock path
