"""Token GEMM against cuBLAS (F.linear / addmm_) at the projection shapes of one denoise step (C2: BF = 32)."""
import os
import statistics
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.getcwd())
from i2v_adapter_unofficial_b200 import _lib, ops  # noqa: E402

torch.manual_seed(0)
SHAPES = [  # (rows, K, N, bias, residual, what)
    (131072, 320, 1536, True, False, "L0 packed qkv+qx (aug)"), (131072, 640, 320, False, True, "L0 stacked out + res"),
    (131072, 320, 320, False, False, "L0 attn2 to_q"), (131072, 320, 320, False, True, "L0 attn2 out + res"),
    (131072, 1288, 320, False, True, "L0 ff.net.2 + res"), (131072, 320, 960, True, False, "L0 temporal qkv"),
    (8192, 320, 768, True, False, "L0 frame-0 kv"),
    (32768, 640, 2560, True, False, "L1 packed qkv+qx"), (32768, 1280, 640, False, True, "L1 stacked out + res"),
    (32768, 2568, 640, False, True, "L1 ff.net.2 + res"), (32768, 640, 1920, True, False, "L1 temporal qkv"),
    (8192, 1280, 5120, True, False, "L2 packed qkv+qx"), (8192, 2560, 1280, False, True, "L2 stacked out + res"),
    (8192, 5128, 1280, False, True, "L2 ff.net.2 + res"), (8192, 1280, 3840, True, False, "L2 temporal qkv"),
]


def timeit(fn, n=20):
    """GPU time per call: the launches are queued behind a spin kernel, so host-side launch cost (tensor-map encoding
    through ctypes for the own kernel, cuBLAS heuristics for F.linear) is not in the bracket -- as in a replayed graph."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(4_000_000)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1) / n)
    return statistics.median(best)


tot_a = tot_b = 0.0
for rows, K, N, wb, wr, what in SHAPES:
    x = torch.randn(rows, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    b = torch.randn(N, device="cuda").to(torch.bfloat16) if wb else None
    r = torch.randn(rows, N, device="cuda").to(torch.bfloat16) if wr else None
    if wr:
        cub = lambda: r.addmm_(x, w.t())  # noqa: E731
        own = lambda: ops.linear(x, w, b, r, out=r)  # noqa: E731
    else:
        cub = lambda: F.linear(x, w, b)  # noqa: E731
        own = lambda: ops.linear(x, w, b)  # noqa: E731
    ta, tb = timeit(cub), timeit(own)
    tot_a += ta; tot_b += tb
    fl = 2.0 * rows * K * N
    print(f"{what:26s} [{rows:6d} x {K:4d}] x [{N:4d}]: cuBLAS {ta * 1e3:7.1f} us ({fl / ta / 1e9:6.0f} TF)   own {tb * 1e3:7.1f} us "
          f"({fl / tb / 1e9:6.0f} TF)   {ta / tb:5.2f}x", flush=True)
print(f"sum: cuBLAS {tot_a * 1e3:.0f} us, own {tot_b * 1e3:.0f} us")
