"""Developer tool: torch-profiler kernel table of the benchmarked denoise step (where does the step time go)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from i2v_adapter_unofficial_b200 import install  # noqa: E402
from i2v_adapter_unofficial_b200.hostmodel import DDIMScheduler, denoise_step  # noqa: E402


def main():
    stock = "--stock" in sys.argv
    dev = torch.device("cuda:0")
    unet = bench.build_unet(dev, torch.bfloat16)
    if os.environ.get("CHANNELS_LAST") == "1":
        unet = unet.to(memory_format=torch.channels_last)
    if not stock:
        install(unet)
    sched = DDIMScheduler()
    sched.set_timesteps(25)
    ts = [int(t) for t in sched.timesteps]
    d_in = bench.make_inputs(1, 16, 64, 1, torch.bfloat16, device=dev)
    lat = d_in["latents"].clone()
    for i in range(3):
        lat = denoise_step(unet, sched, lat, ts[i], d_in["prompt"], 7.5, d_in["cond"], d_in["image"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3):
        lat = denoise_step(unet, sched, lat, ts[3 + i], d_in["prompt"], 7.5, d_in["cond"], d_in["image"])
    e1.record()
    torch.cuda.synchronize()
    print(f"{'stock SDPA' if stock else 'B200'} processors: {e0.elapsed_time(e1) / 3:.2f} ms/step")
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA,
                                            torch.profiler.ProfilerActivity.CPU]) as prof:
        for i in range(2):
            lat = denoise_step(unet, sched, lat, ts[6 + i], d_in["prompt"], 7.5, d_in["cond"], d_in["image"])
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))


if __name__ == "__main__":
    main()
