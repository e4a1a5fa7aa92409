"""Developer tool: GEGLU kernel at the C2 level sizes, GB/s on read [rows, 2D] + write [rows, D]."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2v_adapter_unofficial_b200 import ops
for rows, D in [(131072, 1280), (32768, 2560), (8192, 5120)]:
    x = torch.randn(rows, 2 * D, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        y = ops.geglu(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        y = ops.geglu(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    h, g = x[:, :D].float(), x[:, D:].float()
    ref = h * torch.nn.functional.gelu(g).to(torch.bfloat16).float()
    err = (y.float() - ref).abs().max().item()
    print(f"[perf-geglu] rows {rows} D {D}: {ms*1e3:.1f} us = {(3*rows*D*2)/ms/1e6:.0f} GB/s  max err vs torch {err:.3e}")
