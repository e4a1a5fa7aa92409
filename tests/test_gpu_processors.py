"""Drop-in tests (-m gpu): the B200 processors installed into the host mirror of the reference's UNet classes,
compared with the CPU oracle (fp32) and with the stock SDPA processors on the same device."""
import pytest
import torch
import torch.nn.functional as F

from helpers import TINY_CFG, make_unet, randomize_zero_init, unet_inputs
from i2v_adapter_unofficial_b200 import _lib, install, ops
from i2v_adapter_unofficial_b200.hostmodel import (
    DDIMScheduler,
    I2VAdapterTransformerBlock,
    IPAdapterAttnProcessor2_0,
    TransformerTemporalModel,
    denoise,
)
from oracle.attention_oracle import (
    attention_oracle,
    i2v_block_oracle,
    ip_adapter_attention_oracle,
    temporal_model_oracle,
)
from oracle.unet_oracle import unet_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _sd(module, prefix):
    return {f"{prefix}.{k}": v.float().cpu() for k, v in module.state_dict().items()}


def _bf16(x):
    return x.to(DEV, torch.bfloat16)


@pytest.mark.parametrize("dims", [(320, 8, 40), (640, 8, 80), (1280, 8, 160)], ids=lambda d: f"C{d[0]}d{d[2]}")
@pytest.mark.parametrize("fuse", [True, False])
def test_block_attention_outputs_bf16(dims, fuse):
    """attn1 + i2v_adapter and attn2 (IP-Adapter) outputs of one I2VAdapterTransformerBlock: max-abs <= 2e-2."""
    C, H, d = dims
    V, Fr, S, T = 2, 3, 200, 77
    torch.manual_seed(0)
    blk = I2VAdapterTransformerBlock(C, H, d, cross_attention_dim=768).eval()
    blk.attn2.set_processor(IPAdapterAttnProcessor2_0(C, 768, num_tokens=4, scale=0.8))
    x = torch.randn(V * Fr, S, C)
    ctx = torch.randn(V, T + 4, 768).repeat_interleave(Fr, dim=0)
    sd = _sd(blk, "b")
    with torch.no_grad():
        nh = F.layer_norm(x, (C,), sd["b.norm1.weight"], sd["b.norm1.bias"], 1e-5)
        first = nh[0::Fr].repeat_interleave(Fr, dim=0)
        ref_self = attention_oracle(sd, "b.attn1", nh, None, H)
        ref_cross = attention_oracle(sd, "b.i2v_adapter", nh, first, H)
        ref_ip = ip_adapter_attention_oracle(sd, "b.attn2", nh, ctx, H, 4, 0.8)
        ref_block = i2v_block_oracle(sd, "b", x, ctx, H, True, Fr, 4, 0.8)

    blk = blk.to(DEV, torch.bfloat16)
    handle = install(blk, fuse_cross_frame=fuse)
    n0 = _lib.launch_count()
    with torch.no_grad():
        # what the block does at src/modules/i2v_adapter.py:468-494, processor by processor
        st = blk.attn1.processor.state
        st.enable_cross_frame, st.num_frames = True, Fr
        a1 = blk.attn1(_bf16(nh))
        cx = blk.i2v_adapter(_bf16(nh), encoder_hidden_states=_bf16(first))
        got = (a1 + cx).float().cpu()
        st.enable_cross_frame = False
        ip = blk.attn2(_bf16(nh), encoder_hidden_states=_bf16(ctx)).float().cpu()
        out = blk(_bf16(x), enable_cross_frame_attn=True, num_frames=Fr, encoder_hidden_states=_bf16(ctx)).float().cpu()
    assert _lib.launch_count() > n0
    assert (got - (ref_self + ref_cross)).abs().max().item() <= 2e-2
    assert (ip - ref_ip).abs().max().item() <= 2e-2
    cos = F.cosine_similarity(out.flatten(), ref_block.flatten(), dim=0).item()
    assert cos >= 0.999
    handle.uninstall()


def test_temporal_module_bf16():
    torch.manual_seed(1)
    m = TransformerTemporalModel(num_attention_heads=8, attention_head_dim=40, in_channels=320, norm_num_groups=32,
                                 positional_embeddings="sinusoidal", num_positional_embeddings=32).eval()
    x = torch.randn(2 * 16, 320, 12, 12)
    with torch.no_grad():
        ref = temporal_model_oracle(_sd(m, "m"), "m", x, 16, 8, 32)
    m = m.to(DEV, torch.bfloat16)
    install(m, fast_path=False)
    n0 = _lib.launch_count()
    with torch.no_grad():
        out = m(_bf16(x), num_frames=16)[0].float().cpu()
    assert _lib.launch_count() == n0 + 2  # attn1 and attn2 of the temporal block
    assert (out - ref).abs().max().item() <= 5e-2 and F.cosine_similarity(out.flatten(), ref.flatten(), dim=0) >= 0.999


@pytest.mark.parametrize("ip", [False, True])
def test_unet_fp32_check_mode_matches_oracle(ip):
    """Whole (toy-width) UNet in fp32 through the library's fp32-math kernels: <= 1e-3 relative."""
    unet = randomize_zero_init(make_unet(ip_adapter=ip))
    sample, ctx, img = unet_inputs(unet, videos=2, frames=3, size=16, tokens=6, image_embed_dim=64 if ip else None)
    with torch.no_grad():
        ref = unet_oracle(dict(unet.state_dict()), dict(unet.config), sample, 37, True, ctx, img)
    unet = unet.to(DEV)
    install(unet)
    n0 = _lib.launch_count()
    # the out-of-scope PyTorch layers (cuDNN convolutions, cuBLAS linears) must not drop to TF32 in this check
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            out = unet(sample.to(DEV), 37, True, ctx.to(DEV),
                       added_cond_kwargs={"image_embeds": img.to(DEV)} if ip else None).sample.cpu()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    assert _lib.launch_count() > n0
    err, scale = (out - ref).abs().max().item(), ref.abs().max().item()
    assert err <= 1e-3 * scale, f"max-abs {err:.3e} vs 1e-3 * {scale:.3e}"


def test_unet_bf16_b200_vs_stock_processors_and_oracle():
    """SD1.5 head dims (40 / 80 / 160) at two levels: B200 processors vs the stock SDPA processors (same device,
    same bf16 weights) and vs the fp32 CPU oracle."""
    cfg = dict(block_out_channels=(320, 640), down_block_types=("CrossFrameAttnDownBlockMotion", "DownBlockMotion"),
               up_block_types=("UpBlockMotion", "CrossFrameAttnUpBlockMotion"), cross_attention_dim=768,
               num_attention_heads=8, motion_num_attention_heads=8, norm_num_groups=32, layers_per_block=1)
    torch.manual_seed(0)
    from i2v_adapter_unofficial_b200.hostmodel import UNetMotionCrossFrameAttnModel
    from helpers import fake_ip_adapter_state_dict

    unet = UNetMotionCrossFrameAttnModel(**cfg).eval()
    unet._load_ip_adapter_weights(fake_ip_adapter_state_dict(unet, 3))
    sample, ctx, img = unet_inputs(unet, videos=2, frames=4, size=16, tokens=77, image_embed_dim=64)
    with torch.no_grad():
        ref = unet_oracle(dict(unet.state_dict()), dict(unet.config), sample, 500, True, ctx, img)
    unet = unet.to(DEV, torch.bfloat16)
    args = (_bf16(sample), 500, True, _bf16(ctx))
    kw = dict(added_cond_kwargs={"image_embeds": _bf16(img)})
    with torch.no_grad():
        stock = unet(*args, **kw).sample.float().cpu()
        handle = install(unet)
        n0 = _lib.launch_count()
        ours = unet(*args, **kw).sample.float().cpu()
        launches = _lib.launch_count() - n0
        handle.uninstall()
        again = unet(*args, **kw).sample.float().cpu()
    assert launches >= 3 * 2 + 4 * 2  # spatial (fused + ip) x 3 blocks, temporal x 2 per motion module
    assert torch.equal(again, stock)  # uninstall restores the stock path exactly
    cos = lambda a, b: F.cosine_similarity(a.flatten(), b.flatten(), dim=0).item()  # noqa: E731
    assert cos(ours, stock) >= 0.9995
    assert cos(ours, ref) >= 0.999
    assert (ours - ref).abs().max().item() <= max(3 * (stock - ref).abs().max().item(), 2e-2)


def test_25_step_ddim_final_latent_cosine():
    """BASELINE.json: cosine >= 0.999 on the final latent after 25 DDIM steps (CFG 7.5, first-frame re-imposition)."""
    unet = randomize_zero_init(make_unet(ip_adapter=True))
    sample, ctx, img = unet_inputs(unet, videos=1, frames=4, size=16, tokens=10, image_embed_dim=64)
    ctx2 = torch.cat([torch.randn_like(ctx), ctx])
    img2 = torch.cat([torch.zeros_like(img), img])
    cond = torch.randn(1, 4, 16, 16)

    def run(model, dtype, b200):
        model = model.to(DEV, dtype)
        handle = install(model) if b200 else None
        out = denoise(model, DDIMScheduler(), sample.clone().to(DEV, dtype), ctx2.to(DEV, dtype), 25, 7.5,
                      cond.to(DEV, dtype), img2.to(DEV, dtype)).float().cpu()
        if handle:
            handle.uninstall()
        return out

    ref32 = run(unet, torch.float32, False)          # stock SDPA processors, fp32
    ours32 = run(unet, torch.float32, True)          # library fp32 check mode
    stock16 = run(unet, torch.bfloat16, False)
    ours16 = run(unet, torch.bfloat16, True)
    cos = lambda a, b: F.cosine_similarity(a.flatten(), b.flatten(), dim=0).item()  # noqa: E731
    assert cos(ours32, ref32) >= 0.9999
    assert cos(ours16, stock16) >= 0.999
    assert cos(ours16, ref32) >= 0.999
