"""Developer tool: GEGLU feed-forward input projection at the SD1.5 level shapes -- cuBLAS GEMM + geglu kernel against the
fused tcgen05 GEMM (i2v_ff_geglu_fwd)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2v_adapter_unofficial_b200 import ops  # noqa: E402

torch.manual_seed(0)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for rows, C in [(131072, 320), (32768, 640), (8192, 1280), (2048, 1280)]:
    x = torch.randn(rows, C, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(8 * C, C, device="cuda") * C ** -0.5).to(torch.bfloat16)
    b = (torch.randn(8 * C, device="cuda") * 0.1).to(torch.bfloat16)
    t_gemm = timeit(lambda: F.linear(x, w, b))
    t_old = timeit(lambda: ops.geglu(F.linear(x, w, b), ones_column=True))
    t_new = timeit(lambda: ops.ff_geglu(x, w, b, ones_column=True))
    a, c = ops.geglu(F.linear(x, w, b), ones_column=True), ops.ff_geglu(x, w, b, ones_column=True)
    flops = 2.0 * rows * 8 * C * C
    print(f"[perf-ff] rows {rows} C {C}: cuBLAS {t_gemm * 1e3:7.1f} us  cuBLAS+geglu {t_old * 1e3:7.1f} us  fused "
          f"{t_new * 1e3:7.1f} us = {flops / t_new / 1e9:6.0f} TFLOP/s  max|diff| {(a.float() - c.float()).abs().max().item():.3e}",
          flush=True)
