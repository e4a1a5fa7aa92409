"""Generate tests/golden/*.npz from the REFERENCE ITSELF (run in the build container, where /root/reference exists).

The only part of the reference's attention path that imports without ``diffusers`` is the in-tree stack
``src/modules/attention.py``: ``BasicAttention`` (:26-62) computes exactly the diffusers ``Attention`` +
``AttnProcessor2_0`` arithmetic (bias-free q/k/v Linear, F.scaled_dot_product_attention with the default scale,
Linear+bias), and ``BasicTransformerBlock`` (:64-77) chains a self- and a cross-attention with LayerNorm + residual.
This script instantiates them with seeded default init, runs seeded inputs in fp32 on CPU and stores weights, inputs
and outputs; ``tests/test_oracle_golden.py`` replays them through ``oracle/attention_oracle.py``.

    python oracle/make_golden.py            # writes tests/golden/basic_attention_*.npz, basic_block_*.npz

/root/reference is read-only and does not exist on the GPU box: nothing at test time imports it.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main() -> None:
    sys.path.insert(0, REF)
    from src.modules.attention import BasicAttention, BasicTransformerBlock  # noqa: E402  (the reference's code)

    os.makedirs(OUT, exist_ok=True)
    # (query_dim, context_dim, head_dim, heads, batch, seq_q, seq_kv).  Head dims are the SD1.5 ones (40 / 80 / 160);
    # two heads keep the stored weight matrices small.  Geometries: spatial self-attention, text cross-attention
    # (77 tokens), temporal attention (16 frames as the sequence).
    cases = {
        "self_d40": (80, None, 40, 2, 2, 64, None),
        "cross_text77": (80, 96, 40, 2, 2, 48, 77),
        "temporal_f16_d80": (160, None, 80, 2, 6, 16, None),
        "self_d160": (320, None, 160, 2, 1, 24, None),
    }
    for i, (name, (qd, cd, hd, nh, b, sq, skv)) in enumerate(cases.items()):
        torch.manual_seed(100 + i)
        m = BasicAttention(qd, context_dim=cd, head_dim=hd, num_heads=nh).eval()
        x = torch.randn(b, sq, qd)
        ctx = None if cd is None else torch.randn(b, skv, cd)
        with torch.no_grad():
            y = m(x, ctx)
        arrs = {f"w.{k}": v.numpy() for k, v in m.state_dict().items()}
        arrs.update(x=x.numpy(), y=y.numpy(), heads=np.int64(nh))
        if ctx is not None:
            arrs["ctx"] = ctx.numpy()
        np.savez_compressed(os.path.join(OUT, f"basic_attention_{name}.npz"), **arrs)
        print(name, tuple(y.shape), float(y.abs().mean()))

    torch.manual_seed(200)
    blk = BasicTransformerBlock(128, context_dim=96, head_dim=32, num_heads=4).eval()
    x = torch.randn(3, 40, 128)
    ctx = torch.randn(3, 11, 96)
    with torch.no_grad():
        y = blk(x, ctx)
    arrs = {f"w.{k}": v.numpy() for k, v in blk.state_dict().items()}
    arrs.update(x=x.numpy(), ctx=ctx.numpy(), y=y.numpy(), heads=np.int64(4))
    np.savez_compressed(os.path.join(OUT, "basic_block_self_cross.npz"), **arrs)
    print("block", tuple(y.shape), float(y.abs().mean()))


if __name__ == "__main__":
    main()
