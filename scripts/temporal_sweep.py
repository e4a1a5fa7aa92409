import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/scripts")
from i2v_adapter_unofficial_b200 import ops, _lib
from gpu_selftest import _time
lib = _lib.load()
for (N, Fr, H, d) in ((8192, 16, 8, 40), (2048, 16, 8, 80), (512, 16, 8, 160), (65536, 16, 8, 40)):
    qkv = torch.randn(N, Fr, 3, H, d, device="cuda", dtype=torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    by = 4.0 * N * Fr * H * d * 2
    for stages, per_sm, hg in ((0, 0, 0), (2, 3, 0), (2, 2, 0), (4, 1, 0), (2, 3, 4), (2, 2, 4), (3, 2, 4), (2, 3, 2), (2, 2, 2), (3, 2, 2), (2, 4, 2)):
        if (hg == 4 and d not in (80, 64)) or (hg == 2 and d not in (160, 128)):
            continue
        lib.i2v_set_tuning(0, stages); lib.i2v_set_tuning(1, per_sm); lib.i2v_set_tuning(4, hg)
        try:
            ms = _time(lambda: ops.temporal_attn(q, k, v, None, ops.MODE_FAST), iters=20)
            print(f"N{N} d{d} stages{stages} cta/sm{per_sm} hg{hg}: {ms*1e3:.1f} us = {by/ms/1e6:.0f} GB/s", flush=True)
        except Exception as e:
            print(f"N{N} d{d} stages{stages} cta/sm{per_sm}: {str(e)[:80]}")
