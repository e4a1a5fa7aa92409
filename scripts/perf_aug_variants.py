"""Level-0 fused self + cross-frame launch on the augmented layout: time and check the variants behind tuning key 7
(0 = three query tiles, P in TMEM; 1 / 2 = four query tiles, P through shared memory, 2 / 4 MMA warps)."""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from i2v_adapter_unofficial_b200 import _lib, ops  # noqa: E402

keys = [int(a) for a in sys.argv[1:]] or [0, 1, 2]
torch.manual_seed(3)
Bv, Fr, H, S, d = 2, 16, 8, 4096, 40
mk = lambda b: torch.randn(b, S, H, d, device="cuda", dtype=torch.bfloat16)  # noqa: E731
q, k, v, qx, kx, vx = mk(Bv * Fr), mk(Bv * Fr), mk(Bv * Fr), mk(Bv * Fr), mk(Bv), mk(Bv)
qa, ka, va = ops.augment_qkv(q, k, v)
qxa, kxa, vxa = ops.augment_qkv(qx, kx, vx)
lib = _lib.load()
flops = 2 * 4.0 * Bv * Fr * H * S * S * d


def sdpa(q_, k_, v_):
    t = lambda x: x.transpose(1, 2).float()  # noqa: E731
    return torch.nn.functional.scaled_dot_product_attention(t(q_), t(k_), t(v_)).transpose(1, 2)


ref_s = sdpa(q[:2], k[:2], v[:2])
ref_x = sdpa(qx[15:17], kx[[0, 1]], vx[[0, 1]])
for key in keys:
    lib.i2v_set_tuning(7, key)
    for _ in range(3):
        o = ops.fused_self_xframe_aug(qa, ka, va, qxa, kxa, vxa, Fr)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n = 10
    ev[0].record()
    for _ in range(n):
        o = ops.fused_self_xframe_aug(qa, ka, va, qxa, kxa, vxa, Fr)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / n
    o_s, o_x = o[:, :, 0], o[:, :, 1]
    e_s = (o_s[:2, ..., :d].float() - ref_s).abs().max().item()
    e_x = (o_x[15:17, ..., :d].float() - ref_x).abs().max().item()
    print(f"key7={key}: {ms:.3f} ms  {flops / ms / 1e9:.0f} TFLOP/s  max|err| self {e_s:.4f} xframe {e_x:.4f}", flush=True)
lib.i2v_set_tuning(7, 0)
