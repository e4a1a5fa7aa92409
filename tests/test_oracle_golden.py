"""The oracle's attention leaf against golden vectors produced by the REFERENCE's own importable implementation
(``/root/reference/src/modules/attention.py`` BasicAttention / BasicTransformerBlock; generator:
``oracle/make_golden.py``).  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.attention_oracle import _layer_norm, attention_oracle, sdpa_oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _load(path):
    z = np.load(path)
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")}
    return z, sd


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "basic_attention_*.npz"))))
def test_attention_leaf_matches_reference_basic_attention(path):
    z, sd = _load(path)
    sd = {f"attn.{k}": v for k, v in sd.items()}
    x = torch.from_numpy(z["x"])
    ctx = torch.from_numpy(z["ctx"]) if "ctx" in z.files else None
    y = attention_oracle(sd, "attn", x, ctx, int(z["heads"]))
    ref = torch.from_numpy(z["y"])
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() <= 2e-6, os.path.basename(path)


def test_golden_fixture_set_is_complete():
    names = {os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "*.npz"))}
    assert {"basic_attention_self_d40.npz", "basic_attention_cross_text77.npz",
            "basic_attention_temporal_f16_d80.npz", "basic_attention_self_d160.npz",
            "basic_block_self_cross.npz"} <= names


def test_two_attention_block_matches_reference_basic_transformer_block():
    z, sd = _load(os.path.join(GOLDEN, "basic_block_self_cross.npz"))
    x, ctx, heads = torch.from_numpy(z["x"]), torch.from_numpy(z["ctx"]), int(z["heads"])
    # reference src/modules/attention.py:74-77: x = attn1(norm1(x)) + x ; x = attn2(norm2(x), ctx) + x
    h = attention_oracle(sd, "attn1", _layer_norm(x, sd, "norm1"), None, heads) + x
    h = attention_oracle(sd, "attn2", _layer_norm(h, sd, "norm2"), ctx, heads) + h
    assert (h - torch.from_numpy(z["y"])).abs().max().item() <= 5e-6


@pytest.mark.parametrize("shape", [(2, 3, 17, 9, 40), (1, 8, 16, 16, 80), (2, 2, 5, 81, 160)])
def test_sdpa_oracle_equals_torch_sdpa(shape):
    b, h, sq, skv, d = shape
    g = torch.Generator().manual_seed(sum(shape))
    q, k, v = (torch.randn(b, h, s, d, generator=g) for s in (sq, skv, skv))
    assert (sdpa_oracle(q, k, v) - F.scaled_dot_product_attention(q, k, v)).abs().max().item() <= 2e-6


def test_sdpa_oracle_fp64_agrees_with_fp32():
    g = torch.Generator().manual_seed(3)
    q, k, v = (torch.randn(1, 2, 33, 40, generator=g) for _ in range(3))
    o64 = sdpa_oracle(q.double(), k.double(), v.double())
    assert (sdpa_oracle(q, k, v).double() - o64).abs().max().item() <= 1e-6
