"""Developer tool: kernel durations (torch profiler) of the LayerNorm / NHWC GroupNorm / residual kernels at C2 sizes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2v_adapter_unofficial_b200 import ops
dev = "cuda"
bf = torch.bfloat16
def run():
    for (rows, C) in [(131072, 320), (32768, 640), (8192, 1280)]:
        x = torch.randn(rows, C, device=dev, dtype=bf); w = torch.ones(C, device=dev, dtype=bf); b = torch.zeros(C, device=dev, dtype=bf)
        for _ in range(5):
            ops.layernorm(x, w, b, 1e-5)
    for (N, C, h) in [(32, 320, 64), (32, 640, 32), (32, 1280, 16)]:
        x = torch.randn(N, C, h, h, device=dev, dtype=bf).contiguous(memory_format=torch.channels_last)
        w = torch.ones(C, device=dev, dtype=bf); b = torch.zeros(C, device=dev, dtype=bf)
        t = torch.randn(N, C, device=dev, dtype=bf)
        for _ in range(5):
            y = ops.group_norm_nhwc(x, w, b, 32, 1e-5, 1, silu=True, add=t)
            yp = ops.group_norm_nhwc(x, w, b, 32, 1e-6, 16, to_positions=True)
            ops.positions_to_nhwc_residual(yp, x, 16)
            ops.nhwc_add(x, y, b)
run(); torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    run(); torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if "i2v::" in e.key:
        print(f"{e.key[:70]:70s} n={e.count:3d} avg {e.device_time_total / e.count:7.1f} us")
