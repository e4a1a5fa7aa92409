"""compute-sanitizer target: one small launch of every kernel family through the C ABI (results checked loosely so a
corrupted output also fails).  Run as `compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_small.py`."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.getcwd())
from i2v_adapter_unofficial_b200 import _lib, ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)  # noqa: E731


def sdpa(q, k, v, g=1):
    t = lambda x: x.transpose(1, 2).float()  # noqa: E731
    return F.scaled_dot_product_attention(t(q), t(k).repeat_interleave(g, 0), t(v).repeat_interleave(g, 0)).transpose(1, 2)


def check(name, got, want, tol=3e-2):
    err = (got.float() - want.float()).abs().max().item()
    print(f"{name:28s} max|err| {err:.4f}", flush=True)
    assert err <= tol, name


# pipelined dense kernel, augmented layout (d = 40): two videos x two frames, ragged S
V, Fr, H, S, d = 2, 2, 8, 200, 40
q, k, v, qx, kx, vx = bf(V * Fr, S, H, d), bf(V * Fr, S, H, d), bf(V * Fr, S, H, d), bf(V * Fr, S, H, d), bf(V, S, H, d), bf(V, S, H, d)
o = ops.fused_self_xframe_aug(*ops.augment_qkv(q, k, v), *ops.augment_qkv(qx, kx, vx), Fr)
check("dense pipe (aug) self", o[:, :, 0], sdpa(q, k, v))
check("dense pipe (aug) xframe", o[:, :, 1], sdpa(qx, kx, vx, Fr))
# first tcgen05 kernel: d = 80 and d = 160
for dd in (80, 160):
    q, k, v = bf(2, 300, 8, dd), bf(2, 300, 8, dd), bf(2, 300, 8, dd)
    check(f"dense d={dd}", ops.sdpa(q, k, v, 1, None, ops.MODE_FAST), sdpa(q, k, v))
# temporal
for dd in (40, 160):
    q, k, v = bf(96, 16, 8, dd), bf(96, 16, 8, dd), bf(96, 16, 8, dd)
    check(f"temporal d={dd}", ops.temporal_attn(q, k, v, None, ops.MODE_FAST), sdpa(q, k, v))
# IP-Adapter streaming kernel
q, kv = bf(4, 256, 8, 40), bf(2, 81, 2, 8, 40)
want = sdpa(q, kv[:, :77, 0], kv[:, :77, 1], 2) + 0.5 * sdpa(q, kv[:, 77:, 0], kv[:, 77:, 1], 2)
check("ip tcgen05", ops.ip_xattn(q, kv[:, :, 0], kv[:, :, 1], 77, 0.5, 2, None, ops.MODE_FAST), want)
q = bf(4, 200, 8, 40)   # ragged last query tile: clamped cp.async source rows, guarded stores
want = sdpa(q, kv[:, :77, 0], kv[:, :77, 1], 2) + 0.5 * sdpa(q, kv[:, 77:, 0], kv[:, 77:, 1], 2)
check("ip tcgen05 ragged", ops.ip_xattn(q, kv[:, :, 0], kv[:, :, 1], 77, 0.5, 2, None, ops.MODE_FAST), want)
_lib.load().i2v_set_tuning(5, 4)
check("ip stream", ops.ip_xattn(q, kv[:, :, 0], kv[:, :, 1], 77, 0.5, 2, None, ops.MODE_FAST), want)
_lib.load().i2v_set_tuning(5, 0)
# generic fp32-math kernel
q, k, v = bf(2, 70, 2, 24), bf(2, 70, 2, 24), bf(2, 70, 2, 24)
check("generic", ops.sdpa(q, k, v, 1, None, ops.MODE_GENERIC), sdpa(q, k, v))
# feed-forward GEMM + GEGLU, token GEMM
x, w, b = bf(300, 320), (torch.randn(2560, 320, device=dev) * 0.05).to(torch.bfloat16), bf(2560) * 0.1
h, g = F.linear(x.float(), w.float(), b.float()).chunk(2, -1)
check("ff geglu", ops.ff_geglu(x, w, b), h * F.gelu(g), 6e-2)
w2, r = (torch.randn(320, 640, device=dev) * 0.04).to(torch.bfloat16), bf(300, 320)
check("token gemm", ops.linear(bf(300, 640) * 0 + x.repeat(1, 2), w2, None, r), F.linear(x.repeat(1, 2).float(), w2.float()) + r.float(), 6e-2)
# norm / layout kernels
x = bf(64, 320)
check("layernorm", ops.layernorm(x, torch.ones(320, device=dev).bfloat16(), torch.zeros(320, device=dev).bfloat16(), 1e-5),
      F.layer_norm(x.float(), (320,)))
xi = bf(4, 64, 8, 8).contiguous(memory_format=torch.channels_last)
wg, bg = torch.ones(64, device=dev).bfloat16(), torch.zeros(64, device=dev).bfloat16()
check("group norm nhwc", ops.group_norm_nhwc(xi, wg, bg, 8, 1e-5, 1), F.group_norm(xi.float(), 8, None, None, 1e-5))
check("group norm nhwc (frames)", ops.group_norm_nhwc(xi, wg, bg, 8, 1e-5, 2, to_positions=True).view(2, 64, 2, 64),
      F.group_norm(xi.float().view(2, 2, 64, 8, 8).transpose(1, 2), 8, None, None, 1e-5).permute(0, 3, 4, 2, 1).reshape(2, 64, 2, 64))
torch.cuda.synchronize()
print("sanitize_small: all kernels ran", flush=True)
