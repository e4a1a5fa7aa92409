"""Developer tool: IP-Adapter cross-attention (K3) at the C2 level shapes: the default dispatch (d = 40: the tcgen05
kernel with resident K / V, ip_xattn_tc_sm100.cuh; else the streaming kernel), the streaming kernel (tuning key 5 = 4), the
streaming kernel with run-time token counts (5 = 2) and the dense kernel's single-tile mode (5 = 1); achieved GB/s on the
algorithmic bytes (read Q, write O once)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2v_adapter_unofficial_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
torch.manual_seed(0)
for (B, S, d) in [(32, 4096, 40), (32, 1024, 80), (32, 256, 160), (32, 64, 160)]:
    H, g, nt, ni = 8, 16, 77, 4
    q = torch.randn(B, S, H, d, device="cuda", dtype=torch.bfloat16)
    kv = torch.randn(B // g, nt + ni, 2, H, d, device="cuda", dtype=torch.bfloat16)
    k, v = kv[:, :, 0], kv[:, :, 1]
    nbytes = 2 * q.numel() * 2
    res = {}
    for name, key, cfg in (("default", 0, 0), ("stream", 4, 0), ("stream-rt", 2, 0), ("tcgen05", 1, 0)):
        lib.i2v_set_tuning(5, key)
        lib.i2v_set_tuning(6, cfg)
        for _ in range(3):
            o = ops.ip_xattn(q, k, v, nt, 1.0, g, None, ops.MODE_FAST)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            o = ops.ip_xattn(q, k, v, nt, 1.0, g, None, ops.MODE_FAST)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res[name] = (ms, o.float())
        print(f"[perf-ip] S{S} d{d} {name:8s}: {ms * 1e3:7.1f} us = {nbytes / ms / 1e6:7.0f} GB/s", flush=True)
    print(f"[perf-ip] S{S} d{d} max |default - stream| = {(res['default'][1] - res['stream'][1]).abs().max().item():.3e}", flush=True)
lib.i2v_set_tuning(5, 0)
lib.i2v_set_tuning(6, 0)
