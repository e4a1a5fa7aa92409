"""install(): processor mapping, IP-Adapter weight adoption, hooks, uninstall — host logic, CPU only."""
import pytest
import torch

from helpers import TINY_CFG, make_unet
from i2v_adapter_unofficial_b200 import (
    B200AttnProcessor,
    B200CrossFrameAttnProcessor,
    B200IPAdapterAttnProcessor,
    B200SpatialAttnProcessor,
    B200TemporalAttnProcessor,
    install,
)
from i2v_adapter_unofficial_b200.hostmodel import AttnProcessor2_0, IPAdapterAttnProcessor2_0
from i2v_adapter_unofficial_b200.processors import _PackedWeights


@pytest.mark.parametrize("ip", [False, True])
def test_install_maps_every_processor(ip):
    unet = make_unet({**TINY_CFG, "layers_per_block": 2}, ip_adapter=ip)
    old = dict(unet.attn_processors)
    handle = install(unet)
    new = unet.attn_processors
    assert list(new.keys()) == list(old.keys()) and len(new) == 90
    counts = {}
    for name, proc in new.items():
        counts[type(proc).__name__] = counts.get(type(proc).__name__, 0) + 1
        if "motion_modules" in name:
            assert isinstance(proc, B200TemporalAttnProcessor)
        elif name.endswith("attn1.processor"):
            assert isinstance(proc, B200SpatialAttnProcessor) and proc.sibling is not None
        elif name.endswith("i2v_adapter.processor"):
            assert isinstance(proc, B200CrossFrameAttnProcessor)
        elif ip:
            assert isinstance(proc, B200IPAdapterAttnProcessor)
            assert proc.to_k_ip.weight is old[name].to_k_ip.weight  # parameters shared, not copied
        else:
            assert type(proc) is B200AttnProcessor
    assert counts["B200TemporalAttnProcessor"] == 42 and counts["B200SpatialAttnProcessor"] == 16
    handle.uninstall()
    restored = unet.attn_processors
    for name in old:
        assert restored[name] is old[name]
    assert isinstance(restored["down_blocks.0.motion_modules.0.transformer_blocks.0.attn1.processor"], AttnProcessor2_0)


def test_installed_processors_fail_loudly_on_cpu():
    unet = make_unet()
    install(unet)
    x = torch.randn(1, 2, 4, 16, 16)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        unet(x, 10, True, torch.randn(1, 5, TINY_CFG["cross_attention_dim"]))


def test_installed_processors_refuse_grad_mode():
    """The kernels are forward-only: with grad mode on and trainable parameters (the reference trains to_q / to_out of
    the adapter, src/train_i2v_adapter.py) the processors must raise instead of silently dropping the gradient."""
    unet = make_unet()
    install(unet)
    x = torch.randn(1, 2, 4, 16, 16)
    with pytest.raises(RuntimeError, match="inference-only"):
        unet(x, 10, True, torch.randn(1, 5, TINY_CFG["cross_attention_dim"]))


def test_block_hooks_carry_num_frames_and_enable_flag():
    unet = make_unet()
    handle = install(unet)
    seen = {}
    blk = unet.down_blocks[0].attentions[0].transformer_blocks[0]
    proc = blk.attn1.processor

    def spy(attn, hidden_states, **kw):
        seen["enable"], seen["frames"] = proc.state.enable_cross_frame, proc.state.num_frames
        raise KeyboardInterrupt  # stop before any kernel is needed

    blk.attn1.processor = spy
    with pytest.raises(KeyboardInterrupt):
        unet(torch.randn(1, 3, 4, 16, 16), 10, True, torch.randn(1, 5, TINY_CFG["cross_attention_dim"]))
    assert seen == {"enable": True, "frames": 3}
    assert handle.context.num_frames == 3 and handle.context.ctx_replicated


def test_packed_weight_cache_invalidates_on_update():
    w = torch.nn.Parameter(torch.randn(4, 4))
    cache = _PackedWeights()
    calls = []
    build = lambda: calls.append(1) or w.detach().clone()  # noqa: E731
    a = cache.get([w], build)
    assert cache.get([w], build) is a and len(calls) == 1
    with torch.no_grad():
        w.add_(1.0)  # in-place update bumps the version counter (load_state_dict does the same)
    b = cache.get([w], build)
    assert len(calls) == 2 and not torch.equal(a, b)
    w.data = torch.randn(4, 4)  # storage replaced (.to(), new checkpoint)
    cache.get([w], build)
    assert len(calls) == 3


def test_invalidate_caches_catches_writes_through_data():
    """``p.data.zero_()`` (the reference's own zero-init, src/modules/i2v_adapter.py:142-143) does not bump
    ``_version``: the explicit epoch is the documented way to drop the packed copies."""
    from i2v_adapter_unofficial_b200.processors import Installation, invalidate_caches

    w = torch.nn.Parameter(torch.randn(4, 4))
    cache = _PackedWeights()
    calls = []
    build = lambda: calls.append(1) or w.detach().clone()  # noqa: E731
    a = cache.get([w], build)
    assert not a.requires_grad
    w.data.zero_()
    assert cache.get([w], build) is a          # invisible to the (data_ptr, _version) key ...
    invalidate_caches()
    b = cache.get([w], build)                  # ... until the epoch moves
    assert len(calls) == 2 and torch.count_nonzero(b) == 0
    Installation.invalidate_caches()
    cache.get([w], build)
    assert len(calls) == 3


def test_packed_weights_are_built_without_autograd_graph():
    w = torch.nn.Parameter(torch.randn(4, 4))
    cache = _PackedWeights()
    with torch.enable_grad():
        v = cache.get([w], lambda: torch.cat([w, w]))
    assert v.grad_fn is None and not v.requires_grad


def test_load_state_dict_after_install_invalidates_caches():
    from i2v_adapter_unofficial_b200 import processors

    unet = make_unet()
    install(unet)
    before = processors._CACHE_EPOCH[0]
    unet.load_state_dict(unet.state_dict())
    assert processors._CACHE_EPOCH[0] > before


def test_fast_path_fallbacks_are_counted():
    """A swapped forward that cannot take its fast path (here: CPU tensors) hands the call to the stock forward and
    says so in `fastpath.fallback_counts()`; bench.py asserts the dictionary stays empty on the measured path."""
    from i2v_adapter_unofficial_b200 import fastpath
    from i2v_adapter_unofficial_b200.hostmodel.layers import Downsample2D, ResnetBlock2D

    fastpath.reset_fallback_counts()
    res = ResnetBlock2D(in_channels=32, out_channels=32, temb_channels=64, groups=8).eval()
    down = Downsample2D(32, padding=0).eval()
    undo = fastpath.install_fast_forwards(torch.nn.ModuleList([res, down]))
    x, temb = torch.randn(2, 32, 8, 8), torch.randn(2, 64)
    with torch.no_grad():
        y = res(x, temb)
        z = down(x)
    assert y.shape == x.shape and z.shape == (2, 32, 4, 4)      # padding = 0 pads right / bottom first (diffusers)
    counts = fastpath.fallback_counts()
    assert counts.get("resnet") == 1 and counts.get("downsample") == 1
    for fn in undo:
        fn()
    fastpath.reset_fallback_counts()
    assert fastpath.fallback_counts() == {}


def test_host_attention_takes_a_per_batch_mask():
    """diffusers' prepare_attention_mask: a (B, 1, L) additive mask is repeated per head."""
    from i2v_adapter_unofficial_b200.hostmodel.attention import Attention

    torch.manual_seed(0)
    attn = Attention(64, heads=4, dim_head=16).eval()
    x = torch.randn(2, 5, 64)
    mask = torch.zeros(2, 1, 5)
    mask[:, :, 3:] = -1e4
    with torch.no_grad():
        masked = attn(x, attention_mask=mask)
        ref = attn(x[:, :3], encoder_hidden_states=None)      # the same as dropping the masked keys ... for the kept rows
        full = attn(x, encoder_hidden_states=x[:, :3])
    assert masked.shape == (2, 5, 64)
    assert torch.allclose(masked, full, atol=1e-5)
    assert torch.allclose(masked[:, :3], ref, atol=1e-5)
