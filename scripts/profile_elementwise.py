"""Developer tool: which ATen elementwise ops (add / copy / cat) are still in the benchmarked step, by input shapes
and Python call site."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from i2v_adapter_unofficial_b200 import install  # noqa: E402
from i2v_adapter_unofficial_b200.hostmodel import DDIMScheduler, denoise_step  # noqa: E402

dev = torch.device("cuda:0")
unet = bench.build_unet(dev, torch.bfloat16)
install(unet)
sched = DDIMScheduler()
sched.set_timesteps(25)
ts = [int(t) for t in sched.timesteps]
d_in = bench.make_inputs(1, bench.FRAMES, bench.LATENT, 1, torch.bfloat16, device=dev)
lat = d_in["latents"].clone()
for i in range(2):
    lat = denoise_step(unet, sched, lat, ts[i], d_in["prompt"], 7.5, d_in["cond"], d_in["image"])
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU],
                            record_shapes=True, with_stack=True) as prof:
    lat = denoise_step(unet, sched, lat, ts[3], d_in["prompt"], 7.5, d_in["cond"], d_in["image"])
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=6):
    if e.key in ("aten::add", "aten::add_", "aten::copy_", "aten::cat", "aten::contiguous", "aten::clone", "aten::mul",
                 "aten::upsample_nearest2d", "aten::_to_copy"):
        stack = [s for s in e.stack if "i2v_adapter_unofficial_b200" in s or "bench.py" in s][:2]
        rows.append((e.device_time_total, e.count, e.key, str(e.input_shapes)[:90], " <- ".join(s.split("/")[-1][:60] for s in stack)))
for t, n, k, shp, st in sorted(rows, reverse=True)[:40]:
    print(f"{t / 1e3:7.3f} ms n={n:3d} {k:22s} {shp:90s} {st}")
