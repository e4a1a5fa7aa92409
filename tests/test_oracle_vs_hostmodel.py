"""Host model (stock SDPA processors, the path the reference runs) against the functional CPU oracle, fp32.
Also the reference's own shape tests (test/test_i2v_adapter.py, test/test_unet_motion_cross_frame_attn.py) replayed
against the host mirror, and the reference's error behaviour."""
import pytest
import torch

from helpers import TINY_CFG, make_unet, randomize_zero_init, unet_inputs
from i2v_adapter_unofficial_b200.hostmodel import (
    CrossFrameAttnDownBlockMotion,
    DDIMScheduler,
    I2VAdapterModule,
    I2VAdapterTransformer2DModel,
    I2VAdapterTransformerBlock,
    TransformerTemporalModel,
    UNetMotionCrossFrameAttnModel,
    denoise_step,
)
from oracle.attention_oracle import i2v_block_oracle, temporal_model_oracle, transformer2d_oracle
from oracle.unet_oracle import (
    add_noise_oracle,
    ddim_alphas_cumprod_oracle,
    ddim_timesteps_oracle,
    denoise_step_oracle,
    unet_oracle,
)

TOL = 2e-5


def _sd(module, prefix):
    return {f"{prefix}.{k}": v for k, v in module.state_dict().items()}


@pytest.mark.parametrize("enable", [True, False])
def test_i2v_block_matches_oracle(enable):
    torch.manual_seed(0)
    blk = I2VAdapterTransformerBlock(80, 2, 40, cross_attention_dim=48).eval()
    x, ctx = torch.randn(6, 20, 80), torch.randn(6, 7, 48)
    with torch.no_grad():
        y = blk(x, enable_cross_frame_attn=enable, num_frames=3, encoder_hidden_states=ctx)
        ref = i2v_block_oracle(_sd(blk, "b"), "b", x, ctx, 2, enable, 3)
    assert (y - ref).abs().max().item() <= TOL


def test_cross_frame_branch_changes_output_and_zero_out_proj_is_noop():
    # from_transformer2d_model zero-initialises i2v_adapter.to_out (reference src/modules/i2v_adapter.py:181-182):
    # then the cross-frame branch contributes exactly 0.
    torch.manual_seed(1)
    blk = I2VAdapterTransformerBlock(64, 4, 16, cross_attention_dim=32).eval()
    x, ctx = torch.randn(4, 9, 64), torch.randn(4, 5, 32)
    with torch.no_grad():
        on = blk(x, enable_cross_frame_attn=True, num_frames=2, encoder_hidden_states=ctx)
        off = blk(x, enable_cross_frame_attn=False, encoder_hidden_states=ctx)
        assert (on - off).abs().max().item() > 1e-3
        blk.i2v_adapter.to_out[0].weight.zero_()
        blk.i2v_adapter.to_out[0].bias.zero_()
        on0 = blk(x, enable_cross_frame_attn=True, num_frames=2, encoder_hidden_states=ctx)
    assert torch.equal(on0, off)


def test_block_error_behaviour_matches_reference():
    blk = I2VAdapterTransformerBlock(32, 4, 8, cross_attention_dim=16).eval()
    x, ctx = torch.randn(5, 4, 32), torch.randn(5, 3, 16)
    with pytest.raises(ValueError, match="`num_frames` must be provided"):
        blk(x, enable_cross_frame_attn=True, encoder_hidden_states=ctx)
    with pytest.raises(ValueError, match="must be divisible by the number of frames"):
        blk(x, enable_cross_frame_attn=True, num_frames=2, encoder_hidden_states=ctx)


def test_transformer2d_matches_oracle_and_from_transformer2d_init():
    torch.manual_seed(2)
    m = I2VAdapterTransformer2DModel(4, 16, in_channels=64, cross_attention_dim=32, norm_num_groups=8).eval()
    x, ctx = torch.randn(4, 64, 6, 5), torch.randn(4, 5, 32)
    with torch.no_grad():
        y = m(x, enable_cross_frame_attn=True, num_frames=2, encoder_hidden_states=ctx, return_dict=False)[0]
        ref = transformer2d_oracle(_sd(m, "t"), "t", x, ctx, 4, 8, True, 2)
    assert (y - ref).abs().max().item() <= TOL
    other = I2VAdapterTransformer2DModel(4, 16, in_channels=64, cross_attention_dim=32, norm_num_groups=8)
    other.from_transformer2d_model(m)
    tb = other.transformer_blocks[0]
    assert torch.equal(tb.i2v_adapter.to_q.weight, m.transformer_blocks[0].attn1.to_q.weight)
    assert tb.i2v_adapter.to_out[0].weight.abs().sum().item() == 0.0


def test_temporal_model_matches_oracle():
    torch.manual_seed(3)
    m = TransformerTemporalModel(num_attention_heads=4, attention_head_dim=16, in_channels=64, norm_num_groups=8,
                                 positional_embeddings="sinusoidal", num_positional_embeddings=32).eval()
    x = torch.randn(2 * 5, 64, 4, 3)
    with torch.no_grad():
        y = m(x, num_frames=5)[0]
        ref = temporal_model_oracle(_sd(m, "m"), "m", x, 5, 4, 8)
    assert (y - ref).abs().max().item() <= TOL


@pytest.mark.parametrize("ip", [False, True])
def test_unet_matches_oracle(ip):
    unet = randomize_zero_init(make_unet(ip_adapter=ip))
    sample, ctx, img = unet_inputs(unet, videos=2, frames=3, size=16, tokens=6, image_embed_dim=64 if ip else None)
    with torch.no_grad():
        y = unet(sample, 37, True, ctx, added_cond_kwargs={"image_embeds": img} if ip else None).sample
        ref = unet_oracle(dict(unet.state_dict()), dict(unet.config), sample, 37, True, ctx, img)
    assert y.shape == sample.shape
    assert (y - ref).abs().max().item() <= 1e-4


def test_unet_has_90_processors_in_reference_order():
    unet = UNetMotionCrossFrameAttnModel(**{**TINY_CFG, "layers_per_block": 2})
    keys = list(unet.attn_processors.keys())
    assert len(keys) == 16 * 3 + 21 * 2 == 90
    # down -> up -> mid, attn1, attn2, i2v_adapter inside a block
    assert keys[0] == "down_blocks.0.attentions.0.transformer_blocks.0.attn1.processor"
    assert keys[1].endswith("attn2.processor") and keys[2].endswith("i2v_adapter.processor")
    first_up = next(i for i, k in enumerate(keys) if k.startswith("up_blocks"))
    first_mid = next(i for i, k in enumerate(keys) if k.startswith("mid_block"))
    assert first_up < first_mid
    with pytest.raises(ValueError, match="does not match the number of attention layers"):
        unet.set_attn_processor({keys[0]: unet.attn_processors[keys[0]]})


def test_reference_shape_tests_replayed():
    # test/test_i2v_adapter.py:11-71
    m = I2VAdapterTransformer2DModel(8, 64, in_channels=512, out_channels=512, num_layers=1,
                                     cross_attention_dim=1024, norm_num_groups=32).eval()
    with torch.no_grad():
        out = m(torch.randn(8, 512, 8, 8), enable_cross_frame_attn=True, num_frames=4,
                encoder_hidden_states=torch.randn(8, 77, 1024), return_dict=False)[0]
    assert out.shape == (8, 512, 8, 8)
    # test/test_unet_motion_cross_frame_attn.py:18-92
    blk = CrossFrameAttnDownBlockMotion(in_channels=64, out_channels=128, temb_channels=512, cross_attention_dim=768,
                                        num_attention_heads=8, num_layers=2).eval()
    with torch.no_grad():
        h, states = blk(torch.randn(2 * 8, 64, 16, 16), torch.randn(16, 512), enable_cross_frame_attn=True,
                        encoder_hidden_states=torch.randn(16, 42, 768), num_frames=8)
    assert h.shape == (16, 128, 8, 8) and len(states) == 3


def test_adapter_module_roundtrip():
    unet = make_unet()
    adapter = unet.obtain_i2v_adapter_modules()
    assert isinstance(adapter, I2VAdapterModule)
    other = make_unet(seed=5)
    other.load_i2v_adapter(adapter)
    k = "down_blocks.0.attentions.0.transformer_blocks.0.i2v_adapter.to_q.weight"
    assert torch.equal(other.state_dict()[k], unet.state_dict()[k])
    unet.freeze_unet_params()
    trainable = [n for n, p in unet.named_parameters() if p.requires_grad]
    assert trainable and all(".i2v_adapter.to_q." in n or ".i2v_adapter.to_out." in n for n in trainable)


def test_ddim_and_first_frame_noise_identity():
    s = DDIMScheduler()
    s.set_timesteps(25)
    assert torch.equal(s.timesteps, ddim_timesteps_oracle(25))
    assert torch.allclose(s.alphas_cumprod.double(), ddim_alphas_cumprod_oracle(), atol=1e-6)
    # reference test/test_first_frame_pertubation.py:39 — zero noise on frame 0 => sqrt(alpha_cumprod) * x0
    x0 = torch.randn(2, 4, 4, 8, 8)
    noise = torch.randn_like(x0)
    noise[:, 0] = 0
    t = torch.tensor([500, 500])
    noisy = s.add_noise(x0, noise, t)
    assert torch.allclose(noisy[:, 0], s.alphas_cumprod[500] ** 0.5 * x0[:, 0], atol=1e-6)
    assert torch.allclose(noisy, add_noise_oracle(x0, noise, 500).float(), atol=1e-5)


def test_denoise_step_matches_oracle():
    unet = randomize_zero_init(make_unet(ip_adapter=True))
    sample, ctx, img = unet_inputs(unet, videos=1, frames=3, size=16, tokens=6, image_embed_dim=64)
    ctx2 = torch.cat([torch.randn_like(ctx), ctx])
    img2 = torch.cat([torch.zeros_like(img), img])
    cond = torch.randn(1, 4, 16, 16)
    s = DDIMScheduler()
    s.set_timesteps(25)
    t = int(s.timesteps[3])
    out = denoise_step(unet, s, sample.clone(), t, ctx2, 7.5, cond, img2)
    ref = denoise_step_oracle(dict(unet.state_dict()), dict(unet.config), sample, t, ctx2, 25, 7.5, cond, img2)
    assert (out - ref).abs().max().item() <= 5e-4
