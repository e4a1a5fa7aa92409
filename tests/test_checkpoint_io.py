"""Adapter checkpoint I/O: the I2V-Adapter and motion-adapter directories and the IP-Adapter
file in the layouts the reference reads (``I2VAdapterModule.from_pretrained``, ``MotionAdapter.from_pretrained``,
``_load_ip_adapter_weights``).  Host logic, CPU only."""
import json
import os

import pytest
import torch

from helpers import TINY_CFG, fake_ip_adapter_state_dict, make_unet, randomize_zero_init, unet_inputs
from i2v_adapter_unofficial_b200.hostmodel import (
    I2VAdapterModule,
    IPAdapterAttnProcessor2_0,
    MotionAdapter,
    load_ip_adapter_file,
    save_ip_adapter_file,
)


def _forward(unet, seed=1, image_embed_dim=None):
    sample, ctx, img = unet_inputs(unet, videos=1, frames=2, size=16, seed=seed, image_embed_dim=image_embed_dim)
    added = None if img is None else {"image_embeds": img}
    with torch.no_grad():
        return unet(sample, 10, True, ctx, added_cond_kwargs=added)[0]


@pytest.mark.parametrize("safe", [True, False], ids=["safetensors", "bin"])
def test_i2v_adapter_directory_round_trip(tmp_path, safe):
    src = randomize_zero_init(make_unet(seed=3))
    src.save_i2v_adapter_modules(str(tmp_path), safe_serialization=safe)
    files = sorted(os.listdir(tmp_path))
    assert files == ["config.json", "diffusion_pytorch_model.safetensors" if safe else "diffusion_pytorch_model.bin"]
    cfg = json.load(open(tmp_path / "config.json"))
    assert cfg["_class_name"] == "I2VAdapterModule" and cfg["block_depth"] == TINY_CFG["layers_per_block"]
    assert cfg["block_out_channels"] == list(TINY_CFG["block_out_channels"])

    adapter = I2VAdapterModule.from_pretrained(str(tmp_path))
    keys = list(adapter.state_dict())
    # the reference's key layout: only the blocks with attention carry an adapter, the first up block is a dummy
    assert "down_blocks.0.attentions.0.transformer_blocks.0.i2v_adapter.to_q.weight" in keys
    assert "up_blocks.1.attentions.1.transformer_blocks.0.i2v_adapter.to_out.0.bias" in keys
    assert "mid_block.attentions.0.transformer_blocks.0.i2v_adapter.to_v.weight" in keys
    assert not any(k.startswith("up_blocks.0.") for k in keys)
    assert all("i2v_adapter" in k for k in keys)

    dst = randomize_zero_init(make_unet(seed=3))
    with torch.no_grad():   # perturb the adapter weights of the destination, then restore them from disk
        for name, p in dst.named_parameters():
            if "i2v_adapter" in name:
                p.add_(1.0)
    assert not torch.allclose(_forward(dst), _forward(src))
    dst.load_i2v_adapter(adapter)
    assert torch.equal(_forward(dst), _forward(src))


def test_motion_adapter_directory_round_trip(tmp_path):
    src = randomize_zero_init(make_unet(seed=4))
    src.save_motion_modules(str(tmp_path))
    cfg = json.load(open(tmp_path / "config.json"))
    assert cfg["_class_name"] == "MotionAdapter" and cfg["motion_max_seq_length"] == 32
    adapter = MotionAdapter.from_pretrained(str(tmp_path))
    keys = list(adapter.state_dict())
    assert "down_blocks.3.motion_modules.0.transformer_blocks.0.attn1.to_q.weight" in keys
    assert "up_blocks.0.motion_modules.1.proj_in.weight" in keys
    assert "mid_block.motion_modules.0.norm.weight" in keys
    want = {k: v for k, v in src.state_dict().items() if "motion_modules" in k}
    assert set(keys) == set(want)

    dst = randomize_zero_init(make_unet(seed=4))
    with torch.no_grad():
        for name, p in dst.named_parameters():
            if "motion_modules" in name:
                p.mul_(0.5)
    assert not torch.allclose(_forward(dst), _forward(src))
    dst.load_motion_modules(adapter)
    assert torch.equal(_forward(dst), _forward(src))


def test_from_pretrained_errors(tmp_path):
    with pytest.raises(EnvironmentError, match="not a directory"):
        I2VAdapterModule.from_pretrained(str(tmp_path / "missing"))
    (tmp_path / "config.json").write_text(json.dumps({"_class_name": "I2VAdapterModule", "block_depth": 1,
                                                      "block_out_channels": [32, 64, 64, 64],
                                                      "num_attention_heads": 4}))
    with pytest.raises(EnvironmentError, match="no file named"):
        I2VAdapterModule.from_pretrained(str(tmp_path))
    # a weights file of another architecture is rejected (strict load, as ModelMixin)
    make_unet({**TINY_CFG, "block_out_channels": (32, 32, 64, 64)}).save_i2v_adapter_modules(str(tmp_path / "other"))
    os.replace(tmp_path / "other" / "diffusion_pytorch_model.safetensors", tmp_path / "diffusion_pytorch_model.safetensors")
    with pytest.raises(RuntimeError, match="size mismatch"):
        I2VAdapterModule.from_pretrained(str(tmp_path))


@pytest.mark.parametrize("ext", ["bin", "safetensors"])
def test_ip_adapter_file_round_trip(tmp_path, ext):
    unet = make_unet(seed=5)
    state = fake_ip_adapter_state_dict(unet, seed=9)
    path = str(tmp_path / f"ip-adapter_sd15.{ext}")
    save_ip_adapter_file(state, path)
    loaded = load_ip_adapter_file(path)
    assert set(loaded) == {"image_proj", "ip_adapter"}
    assert set(loaded["ip_adapter"]) == set(state["ip_adapter"]) and "1.to_k_ip.weight" in loaded["ip_adapter"]
    for group in state:
        for k, v in state[group].items():
            assert torch.equal(loaded[group][k], v)

    ref = make_unet(seed=5)
    ref._load_ip_adapter_weights(state)
    unet.load_ip_adapter(path)
    procs = unet.attn_processors
    assert isinstance(procs["down_blocks.0.attentions.0.transformer_blocks.0.attn2.processor"], IPAdapterAttnProcessor2_0)
    assert unet.config.encoder_hid_dim_type == "ip_image_proj"
    assert torch.equal(_forward(unet, image_embed_dim=64), _forward(ref, image_embed_dim=64))
