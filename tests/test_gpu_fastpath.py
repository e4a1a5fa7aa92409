"""Normalisation / layout / epilogue kernels and the module-level fast path (-m gpu): each kernel through the C ABI
against the CPU oracle arithmetic (fp32 PyTorch on the same bf16 inputs), then whole modules with
``install(..., fast_path=True)`` against ``fast_path=False`` and the fp32 oracle."""
import pytest
import torch
import torch.nn.functional as F

from helpers import SD15_HEADDIM_CFG, make_unet, randomize_zero_init, unet_inputs
from i2v_adapter_unofficial_b200 import _lib, install, ops
from i2v_adapter_unofficial_b200.hostmodel import (
    I2VAdapterTransformer2DModel,
    IPAdapterAttnProcessor2_0,
    TransformerTemporalModel,
)
from oracle.attention_oracle import temporal_model_oracle, transformer2d_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand(shape, seed, scale=1.0, shift=0.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale + shift).to(torch.bfloat16)


@pytest.mark.parametrize("C", [320, 640, 1280, 64, 2048])
@pytest.mark.parametrize("with_pe", [False, True])
def test_layernorm_kernel(C, with_pe):
    N, Fr = 37, 16
    x = _rand((N, Fr, C), 1, 2.0, 0.5)
    w, b = _rand((C,), 2, 0.2, 1.0), _rand((C,), 3, 0.2)
    pe = _rand((Fr, C), 4) if with_pe else None
    ref = F.layer_norm(x.float(), (C,), w.float(), b.float(), 1e-5)
    if with_pe:
        ref = ref.to(torch.bfloat16).float() + pe.float()   # the reference adds pos_embed to the bf16 LayerNorm output
    got = ops.layernorm(x.to(DEV), w.to(DEV), b.to(DEV), 1e-5, None if pe is None else pe.to(DEV)).float().cpu()
    assert got.shape == x.shape
    assert (got - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item() / 4)


@pytest.mark.parametrize("D", [1280, 2560, 5120, 8])
def test_geglu_kernel(D):
    x = _rand((3, 50, 2 * D), 5, 1.5)
    h, g = x.float().chunk(2, dim=-1)
    ref = h * F.gelu(g)
    got = ops.geglu(x.to(DEV)).float().cpu()
    assert (got - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item() / 4)


@pytest.mark.parametrize("case", [(6, 320, 16, 16, 1), (6, 320, 16, 16, 3), (4, 640, 8, 8, 2), (2, 1280, 12, 12, 2),
                                  (32, 320, 64, 64, 16), (3, 64, 4, 2, 1)], ids=str)
def test_group_norm_tokens_and_back(case):
    N, C, h, w, fg = case
    x = _rand((N, C, h, w), 6, 1.5, 0.3)
    wt, bs = _rand((C,), 7, 0.2, 1.0), _rand((C,), 8, 0.2)
    V, S = N // fg, h * w
    xf = x.float().view(V, fg, C, h, w).permute(0, 2, 1, 3, 4)          # statistics over (C/G, fg, h, w)
    ref = F.group_norm(xf, 32, wt.float(), bs.float(), 1e-6)             # [V, C, fg, h, w]
    ref_tok = ref.permute(0, 3, 4, 2, 1).reshape(V * S, fg, C)          # (B*S, F, C) as TransformerTemporalModel does
    tok = ops.group_norm_tokens(x.to(DEV), wt.to(DEV), bs.to(DEV), 32, 1e-6, fg)
    exp_shape = (N, S, C) if fg == 1 else (V * S, fg, C)
    assert tuple(tok.shape) == exp_shape
    assert (tok.float().cpu().view(V * S, fg, C) - ref_tok).abs().max().item() <= 3e-2
    # way back + residual: out[n, c, s] = y[token layout] + res
    y = _rand(tuple(tok.shape), 9)
    res = _rand((N, C, h, w), 10)
    back = ops.tokens_to_nchw_residual(y.to(DEV), res.to(DEV), fg).float().cpu()
    ref_back = (y.float().view(V, h, w, fg, C).permute(0, 3, 4, 1, 2).reshape(N, C, h, w) + res.float())
    assert (back - ref_back).abs().max().item() <= 2e-2


def test_layout_kernels_reject_unsupported_shapes():
    x = torch.zeros(2, 100, 4, 4, device=DEV, dtype=torch.bfloat16)     # C % 64 != 0
    w = torch.ones(100, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(_lib.I2VLibraryError) as e:
        ops.group_norm_tokens(x, w, w, 4, 1e-6, 1)
    assert e.value.code == -2
    with pytest.raises(TypeError):
        ops.layernorm(torch.zeros(4, 64, device=DEV), torch.ones(64, device=DEV), torch.zeros(64, device=DEV))


def _sd(module, prefix):
    return {f"{prefix}.{k}": v.float().cpu() for k, v in module.state_dict().items()}


def test_spatial_transformer_fast_path_matches_oracle_and_slow_path():
    torch.manual_seed(0)
    C, H, Fr, V, hw = 320, 8, 4, 2, 16
    m = I2VAdapterTransformer2DModel(num_attention_heads=H, attention_head_dim=C // H, in_channels=C,
                                     norm_num_groups=32, cross_attention_dim=768).eval()
    for tb in m.transformer_blocks:
        tb.attn2.set_processor(IPAdapterAttnProcessor2_0(C, 768, num_tokens=4, scale=1.0))
    randomize_zero_init(m)
    x = torch.randn(V * Fr, C, hw, hw)
    ctx = torch.randn(V, 81, 768).repeat_interleave(Fr, dim=0)
    with torch.no_grad():
        ref = transformer2d_oracle(_sd(m, "t"), "t", x, ctx, H, 32, True, Fr, 4, 1.0)
    m = m.to(DEV, torch.bfloat16)
    outs = {}
    for fast in (False, True):
        handle = install(m, fast_path=fast)
        n0 = _lib.launch_count()
        with torch.no_grad():
            outs[fast] = m(x.to(DEV, torch.bfloat16), enable_cross_frame_attn=True,
                           encoder_hidden_states=ctx.to(DEV, torch.bfloat16), num_frames=Fr).sample.float().cpu()
        launches = _lib.launch_count() - n0
        handle.uninstall()
        # fast path: + gn_stats, gn_apply_transpose, 3 layernorm, geglu, untranspose_residual
        assert launches == (2 + 7 if fast else 2)
    for fast in (False, True):
        cos = F.cosine_similarity(outs[fast].flatten(), ref.flatten(), dim=0).item()
        assert cos >= 0.999, (fast, cos)
    scale = ref.abs().max().item()
    assert (outs[True] - ref).abs().max().item() <= 1.5 * max((outs[False] - ref).abs().max().item(), 0.01 * scale)


def test_temporal_module_fast_path_matches_oracle_and_slow_path():
    torch.manual_seed(1)
    m = TransformerTemporalModel(num_attention_heads=8, attention_head_dim=40, in_channels=320, norm_num_groups=32,
                                 positional_embeddings="sinusoidal", num_positional_embeddings=32).eval()
    randomize_zero_init(m)
    x = torch.randn(2 * 16, 320, 8, 8)
    with torch.no_grad():
        ref = temporal_model_oracle(_sd(m, "m"), "m", x, 16, 8, 32)
    m = m.to(DEV, torch.bfloat16)
    outs = {}
    for fast in (False, True):
        handle = install(m, fast_path=fast)
        n0 = _lib.launch_count()
        with torch.no_grad():
            outs[fast] = m(x.to(DEV, torch.bfloat16), num_frames=16)[0].float().cpu()
        launches = _lib.launch_count() - n0
        handle.uninstall()
        assert launches == (2 + 7 if fast else 2)
        cos = F.cosine_similarity(outs[fast].flatten(), ref.flatten(), dim=0).item()
        assert cos >= 0.999 and (outs[fast] - ref).abs().max().item() <= 5e-2, (fast, cos)


def test_unet_fast_path_equals_processor_only_path():
    """Whole UNet (SD1.5 head dims), bf16: fast path vs processors only vs the fp32 oracle."""
    from oracle.unet_oracle import unet_oracle

    unet = randomize_zero_init(make_unet(SD15_HEADDIM_CFG, ip_adapter=True))
    sample, ctx, img = unet_inputs(unet, videos=1, frames=4, size=32, tokens=77, image_embed_dim=64)
    with torch.no_grad():
        ref = unet_oracle(dict(unet.state_dict()), dict(unet.config), sample, 37, True, ctx, img)
    unet = unet.to(DEV, torch.bfloat16)
    outs = {}
    for fast in (False, True):
        handle = install(unet, fast_path=fast)
        with torch.no_grad():
            outs[fast] = unet(sample.to(DEV, torch.bfloat16), 37, True, ctx.to(DEV, torch.bfloat16),
                              added_cond_kwargs={"image_embeds": img.to(DEV, torch.bfloat16)}).sample.float().cpu()
        handle.uninstall()
    cos = {k: F.cosine_similarity(v.flatten(), ref.flatten(), dim=0).item() for k, v in outs.items()}
    assert cos[True] >= 0.999 and cos[False] >= 0.999, cos
    assert cos[True] >= cos[False] - 2e-4


# ---- channels-last (NHWC) GroupNorm family ----------------------------------------------------------------------
def _gn_ref(x, w, b, G, eps, fg, add=None, silu=False):
    """fp32 restatement on the bf16 inputs with the reference's rounding points (each PyTorch op rounds to bf16)."""
    xf = x.float()
    if add is not None:
        xf = (xf + add.float()[:, :, None, None]).to(torch.bfloat16).float()
    N, C, h, w_ = xf.shape
    V = N // fg
    y = xf.view(V, fg, C, h, w_).permute(0, 2, 1, 3, 4)                       # (V, C, F, h, w): statistics per video
    y = F.group_norm(y, G, w.float(), b.float(), eps).permute(0, 2, 1, 3, 4).reshape(N, C, h, w_)
    if silu:
        y = F.silu(y.to(torch.bfloat16).float())
    return y


@pytest.mark.parametrize("case", [(4, 320, 16, 16, 32, 1), (6, 640, 8, 8, 32, 3), (2, 2560, 8, 8, 32, 1),
                                  (2, 1920, 4, 4, 32, 2), (3, 64, 5, 7, 8, 1), (32, 320, 32, 32, 32, 16)],
                         ids=lambda c: "N{}C{}h{}w{}G{}fg{}".format(*c))
@pytest.mark.parametrize("variant", ["plain", "silu", "add_silu"])
def test_group_norm_nhwc(case, variant):
    N, C, h, w_, G, fg = case
    x = _rand((N, C, h, w_), 3, 1.5, 0.3)
    wt, bs = _rand((C,), 4, 0.2, 1.0), _rand((C,), 5, 0.3)
    add = _rand((N, C), 6, 0.7) if variant == "add_silu" else None
    silu = variant != "plain"
    ref = _gn_ref(x, wt, bs, G, 1e-6, fg, add, silu)
    xc = x.to(DEV).contiguous(memory_format=torch.channels_last)
    y = ops.group_norm_nhwc(xc, wt.to(DEV), bs.to(DEV), G, 1e-6, fg, silu=silu,
                            add=None if add is None else add.to(DEV))
    assert ops.is_channels_last(y) or h * w_ == 1
    assert (y.float().cpu() - ref).abs().max().item() <= 4e-2
    if variant == "plain":
        # motion-module layout: [V*S, F, C] with row (v*S + s)*F + f
        yp = ops.group_norm_nhwc(xc, wt.to(DEV), bs.to(DEV), G, 1e-6, fg, to_positions=True)
        V, S = N // fg, h * w_
        want = ref.view(V, fg, C, S).permute(0, 3, 1, 2).reshape(V * S, fg, C)
        assert (yp.float().cpu() - want).abs().max().item() <= 4e-2
        # and back, with the residual
        res = _rand((N, C, h, w_), 7)
        back = ops.positions_to_nhwc_residual(yp, res.to(DEV).contiguous(memory_format=torch.channels_last), fg)
        assert ops.is_channels_last(back) or h * w_ == 1
        want_back = yp.float().cpu().view(V, S, fg, C).permute(0, 2, 3, 1).reshape(N, C, h, w_) + res.float()
        assert (back.float().cpu() - want_back).abs().max().item() <= 2e-2


def test_nhwc_kernels_reject_what_they_do_not_support():
    x = torch.zeros(2, 64, 4, 4, device=DEV, dtype=torch.bfloat16)
    w = torch.ones(64, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="channels-last"):
        ops.group_norm_nhwc(x, w, w, 8, 1e-5)                       # NCHW-contiguous input
    xc = x.contiguous(memory_format=torch.channels_last)
    with pytest.raises(RuntimeError, match="bad shape"):
        ops.group_norm_nhwc(xc, w, w, 7, 1e-5)                      # C % G != 0


def test_unet_channels_last_equals_nchw_fast_path():
    """Whole UNet, bf16: the channels-last data flow (default) against the NCHW fast path and the fp32 oracle."""
    from oracle.unet_oracle import unet_oracle

    unet = randomize_zero_init(make_unet(SD15_HEADDIM_CFG, ip_adapter=True))
    sample, ctx, img = unet_inputs(unet, videos=1, frames=4, size=32, tokens=77, image_embed_dim=64)
    with torch.no_grad():
        ref = unet_oracle(dict(unet.state_dict()), dict(unet.config), sample, 37, True, ctx, img)
    outs = {}
    for cl in (False, True):
        u = randomize_zero_init(make_unet(SD15_HEADDIM_CFG, ip_adapter=True))
        u.load_state_dict(unet.state_dict())
        u = u.to(DEV, torch.bfloat16)
        handle = install(u, channels_last=cl)
        n0 = _lib.launch_count()
        with torch.no_grad():
            outs[cl] = u(sample.to(DEV, torch.bfloat16), 37, True, ctx.to(DEV, torch.bfloat16),
                         added_cond_kwargs={"image_embeds": img.to(DEV, torch.bfloat16)}).sample.float().cpu()
        launches = _lib.launch_count() - n0
        handle.uninstall()
        if cl:
            assert launches > 200, launches     # the resnet GroupNorms run in the library too
    cos = {k: F.cosine_similarity(v.flatten(), ref.flatten(), dim=0).item() for k, v in outs.items()}
    assert cos[True] >= 0.999 and cos[False] >= 0.999, cos
    assert cos[True] >= cos[False] - 3e-4, cos


# ---- deferred-bias / fused-residual building blocks ------------------------------------------------------------
@pytest.mark.parametrize("C", [320, 1280])
def test_layernorm_pre_bias(C):
    N, Fr = 19, 16
    x, pre = _rand((N, Fr, C), 11, 2.0, 0.5), _rand((C,), 12, 0.5)
    w, b = _rand((C,), 13, 0.3, 1.0), _rand((C,), 14, 0.3)
    xs = (x.float() + pre.float()).to(torch.bfloat16).float()           # the reference's separate add rounds to bf16
    ref = F.layer_norm(xs, (C,), w.float(), b.float(), 1e-5)
    y = ops.layernorm(x.to(DEV), w.to(DEV), b.to(DEV), 1e-5, None, pre.to(DEV)).float().cpu()
    assert (y - ref).abs().max().item() <= 3e-2


def test_geglu_ones_column():
    rows, D = 77, 1280
    x = _rand((rows, 2 * D), 15, 1.5)
    y = ops.geglu(x.to(DEV), ones_column=True).cpu()
    assert y.shape == (rows, D + 8)
    assert torch.equal(y[:, :D], ops.geglu(x.to(DEV)).cpu())
    pad = torch.zeros(rows, 8, dtype=torch.bfloat16)
    pad[:, 0] = 1
    assert torch.equal(y[:, D:], pad)


def test_nhwc_add_with_bias():
    N, C, h, w_ = 3, 640, 8, 8
    x, y, b = _rand((N, C, h, w_), 16), _rand((N, C, h, w_), 17), _rand((C,), 18)
    cl = lambda t: t.to(DEV).contiguous(memory_format=torch.channels_last)  # noqa: E731
    out = ops.nhwc_add(cl(x), cl(y), b.to(DEV))
    ref = (y.float() + b.float()[None, :, None, None]).to(torch.bfloat16).float() + x.float()
    assert ops.is_channels_last(out)
    assert (out.float().cpu() - ref).abs().max().item() <= 4e-2     # one bf16 rounding of a sum of magnitude <= 8
    out0 = ops.nhwc_add(cl(x), cl(y))
    assert (out0.float().cpu() - (x.float() + y.float())).abs().max().item() <= 4e-2


def test_fused_residual_block_equals_separate_adds():
    """Residual adds riding in the output GEMMs (deferred biases) against the same block with processors only."""
    torch.manual_seed(3)
    m = TransformerTemporalModel(num_attention_heads=8, attention_head_dim=40, in_channels=320, norm_num_groups=32,
                                 positional_embeddings="sinusoidal", num_positional_embeddings=32).eval()
    randomize_zero_init(m)
    for p in m.parameters():      # non-zero biases everywhere, so a dropped or doubled bias cannot hide
        if p.dim() == 1:
            torch.nn.init.normal_(p, 0.0, 0.5)
    x = torch.randn(2 * 16, 320, 8, 8)
    m = m.to(DEV, torch.bfloat16)
    outs = {}
    for fast in (False, True):
        handle = install(m, fast_path=fast)
        with torch.no_grad():
            outs[fast] = m(x.to(DEV, torch.bfloat16), num_frames=16)[0].float().cpu()
        handle.uninstall()
    cos = F.cosine_similarity(outs[True].flatten(), outs[False].flatten(), dim=0).item()
    assert cos >= 0.9995, cos
    assert (outs[True] - outs[False]).abs().max().item() <= 0.15 * outs[False].abs().max().item()


def test_adapter_checkpoints_loaded_after_install_take_effect(tmp_path):
    """Weights read from disk (I2V-Adapter / motion-adapter directories, reference load_i2v_adapter :1038-1041 and
    load_motion_modules :1028-1036) into a UNet whose B200 processors and fast path are already installed: the packed
    projection weights are caches keyed on the parameters' version and must follow."""
    from i2v_adapter_unofficial_b200.hostmodel import I2VAdapterModule, MotionAdapter

    trained = randomize_zero_init(make_unet(SD15_HEADDIM_CFG, seed=11))
    trained.save_i2v_adapter_modules(str(tmp_path / "i2v"))
    trained.save_motion_modules(str(tmp_path / "motion"))
    sample, ctx, _ = unet_inputs(trained, videos=1, frames=4, size=16, tokens=77)

    def run(u):
        with torch.no_grad():
            return u(sample.to(DEV, torch.bfloat16), 37, True, ctx.to(DEV, torch.bfloat16)).sample.float().cpu()

    trained = trained.to(DEV, torch.bfloat16)
    install(trained)
    want = run(trained)

    other = randomize_zero_init(make_unet(SD15_HEADDIM_CFG, seed=11))
    with torch.no_grad():
        for name, p in other.named_parameters():
            if "i2v_adapter" in name or "motion_modules" in name:
                p.mul_(0.25)
    other = other.to(DEV, torch.bfloat16)
    install(other)
    before = run(other)                       # builds every packed-weight cache from the perturbed weights
    assert F.cosine_similarity(before.flatten(), want.flatten(), dim=0).item() < 0.999
    other.load_i2v_adapter(I2VAdapterModule.from_pretrained(str(tmp_path / "i2v"), torch_dtype=torch.bfloat16))
    other.load_motion_modules(MotionAdapter.from_pretrained(str(tmp_path / "motion"), torch_dtype=torch.bfloat16))
    assert torch.equal(run(other), want)


# ---- feed-forward input projection fused with GEGLU (tcgen05 GEMM) ------------------------------------------------
@pytest.mark.parametrize("case", [(256, 320, 1280, True), (1000, 64, 128, False), (384, 640, 2560, True),
                                  (130, 1280, 5120, True), (5000, 320, 1280, False),
                                  (100, 320, 1280, True),       # one row block: the single-CTA kernel
                                  (2048 + 77, 256, 384, True)],  # odd number of row blocks, three n-tiles
                         ids=lambda c: "rows{}K{}N{}bias{}".format(*c))
@pytest.mark.parametrize("ones", [False, True])
def test_ff_geglu_gemm_matches_fp32_reference(case, ones):
    rows, K, N, with_bias = case
    x = _rand((rows, K), 41, 1.0)
    w = _rand((2 * N, K), 42, K ** -0.5)
    b = _rand((2 * N,), 43, 0.5) if with_bias else None
    proj = F.linear(x.float(), w.float(), None if b is None else b.float())
    ref = proj[:, :N] * F.gelu(proj[:, N:])
    y = ops.ff_geglu(x.to(DEV), w.to(DEV), None if b is None else b.to(DEV), ones_column=ones).float().cpu()
    assert y.shape == (rows, N + 8 if ones else N)
    err = (y[:, :N] - ref).abs()
    assert err.max().item() <= 2e-2 * max(1.0, ref.abs().max().item()), err.max().item()
    assert (err / (ref.abs() + 0.1)).mean().item() <= 4e-3
    if ones:
        pad = torch.zeros(rows, 8)
        pad[:, 0] = 1
        assert torch.equal(y[:, N:], pad)


def test_ff_geglu_rejects_unsupported_shapes():
    x, w = _rand((64, 72), 44).to(DEV), _rand((256, 72), 45).to(DEV)
    with pytest.raises(RuntimeError, match="K % 64"):
        ops.ff_geglu(x, w)
    x, w = _rand((64, 64), 46).to(DEV), _rand((192, 64), 47).to(DEV)
    with pytest.raises(RuntimeError, match="N % 128"):
        ops.ff_geglu(x, w)


@pytest.mark.parametrize("shape", [(3, 64, 5, 7), (2, 1280, 8, 8), (32, 640, 16, 16)])
def test_upsample2x_nhwc_is_exact(shape):
    x = _rand(shape, 51).to(DEV).contiguous(memory_format=torch.channels_last)
    y = ops.upsample2x_nhwc(x)
    ref = F.interpolate(x, scale_factor=2.0, mode="nearest")
    assert ops.is_channels_last(y) and y.shape == ref.shape
    assert torch.equal(y, ref)


def test_nhwc_bias_add_in_place():
    x = _rand((3, 320, 6, 5), 52, 2.0).to(DEV).contiguous(memory_format=torch.channels_last)
    b = _rand((320,), 53).to(DEV)
    ref = (x.float() + b.float()[None, :, None, None]).to(torch.bfloat16)
    out = ops.nhwc_bias_add_(x, b)
    assert out.data_ptr() == x.data_ptr() and torch.equal(out, ref)


def test_sampler_fast_paths_match_the_modules():
    from i2v_adapter_unofficial_b200.hostmodel.layers import Downsample2D, Upsample2D

    torch.manual_seed(5)
    for cls, shape in ((Upsample2D, (4, 64, 8, 8)), (Downsample2D, (4, 64, 16, 16))):
        m = cls(64).eval().to(DEV, torch.bfloat16).to(memory_format=torch.channels_last)
        x = _rand(shape, 54).to(DEV).contiguous(memory_format=torch.channels_last)
        with torch.no_grad():
            ref = m(x).float()
            handle = install(m)
            out = m(x).float()
            handle.uninstall()
        assert out.shape == ref.shape
        assert (out - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("case", [(4, 320, 320, 16, 16, 32), (2, 1280, 640, 8, 8, 32), (3, 640, 320, 12, 12, 32)])
@pytest.mark.parametrize("silu", [False, True])
def test_group_norm_nhwc_two_sources_equals_concatenation(case, silu):
    N, C1, C2, h, w_, G = case
    cl = lambda t: t.to(DEV).contiguous(memory_format=torch.channels_last)  # noqa: E731
    a, b = cl(_rand((N, C1, h, w_), 61, 1.5, 0.3)), cl(_rand((N, C2, h, w_), 62, 0.7, -0.2))
    w, bias = _rand((C1 + C2,), 63, 0.3, 1.0).to(DEV), _rand((C1 + C2,), 64, 0.3).to(DEV)
    cat = cl(torch.cat([a, b], 1))
    ref = ops.group_norm_nhwc(cat, w, bias, G, 1e-5, 1, silu=silu)
    out = ops.group_norm_nhwc(a, w, bias, G, 1e-5, 1, silu=silu, x2=b)
    assert ops.is_channels_last(out) and out.shape == ref.shape
    # channel sums accumulate in a different order: at most one bf16 ulp apart (0.031 for |y| in [4, 8))
    err = (out.float() - ref.float()).abs()
    assert (err <= 2e-2 * torch.clamp(ref.float().abs() / 2, min=1.0)).all(), err.max().item()


def test_up_block_without_concatenation_matches_module():
    """CrossFrameAttnUpBlockMotion fast forward (GroupNorm and shortcut convolution read hidden and skip in place)
    against the same block with processors only."""
    from i2v_adapter_unofficial_b200.hostmodel import CrossFrameAttnUpBlockMotion

    torch.manual_seed(9)
    blk = CrossFrameAttnUpBlockMotion(in_channels=320, out_channels=320, prev_output_channel=640, temb_channels=1280,
                                      num_layers=3, resnet_eps=1e-5, resnet_groups=32, num_attention_heads=8,
                                      cross_attention_dim=768, add_upsample=False, temporal_num_attention_heads=8,
                                      temporal_max_seq_length=32).eval()
    randomize_zero_init(blk)
    Fr, hw = 4, 16
    x = torch.randn(Fr, 640, hw, hw)
    skips = tuple(torch.randn(Fr, c, hw, hw) for c in (320, 320, 320))
    temb, ctx = torch.randn(Fr, 1280), torch.randn(Fr, 77, 768)
    blk = blk.to(DEV, torch.bfloat16)
    args = [t.to(DEV, torch.bfloat16) for t in (x, temb, ctx)]
    sk = tuple(t.to(DEV, torch.bfloat16) for t in skips)
    outs = {}
    for fast in (False, True):
        handle = install(blk, fast_path=fast)
        with torch.no_grad():
            outs[fast] = blk(args[0], sk, args[1], enable_cross_frame_attn=True, encoder_hidden_states=args[2],
                             num_frames=Fr).float().cpu()
        handle.uninstall()
    cos = F.cosine_similarity(outs[True].flatten(), outs[False].flatten(), dim=0).item()
    assert cos >= 0.9995, cos
