"""Developer tool: how much of the graph-replayed denoise step is kernel time and how much is gaps between kernels.
torch.profiler (CUPTI) records the kernels of three graph replays; the step time comes from CUDA events without the profiler."""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from i2v_adapter_unofficial_b200 import install  # noqa: E402
from i2v_adapter_unofficial_b200.hostmodel import DDIMScheduler  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    unet = bench.build_unet(dev, torch.bfloat16)
    install(unet)
    sched = DDIMScheduler()
    sched.set_timesteps(25)
    with torch.no_grad():
        r = bench.StepRunner(unet, sched, dev, 0, 1, 1, 16, 64, graph=True, prof=False)
        ms = r.timed(10, 3) / 10
        print(f"graph replay: {ms:.2f} ms per step")
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            for i in range(3):
                r.step(20 + i)
            torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    n = len(evs) // 3
    step = evs[n:2 * n]   # the middle replay
    busy = sum(e.time_range.end - e.time_range.start for e in step)
    span = step[-1].time_range.end - step[0].time_range.start
    gaps = [step[i + 1].time_range.start - step[i].time_range.end for i in range(len(step) - 1)]
    print(f"middle replay under the profiler: {len(step)} kernels, span {span / 1e3:.2f} ms, kernel time {busy / 1e3:.2f} ms, "
          f"gaps {sum(g for g in gaps if g > 0) / 1e3:.2f} ms (median gap {sorted(gaps)[len(gaps) // 2]:.2f} us)")
    by = collections.defaultdict(lambda: [0, 0.0])
    for e in step:
        k = e.name[:90]
        by[k][0] += 1
        by[k][1] += e.time_range.end - e.time_range.start
    for k, (c, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:28]:
        print(f"{t / 1e3:8.3f} ms {c:5d}  {k}")
    # gaps by the kernel that follows
    after = collections.defaultdict(lambda: [0, 0.0])
    for i, g in enumerate(gaps):
        k = step[i + 1].name[:70]
        after[k][0] += 1
        after[k][1] += max(g, 0)
    print("--- gap time by the kernel that follows the gap")
    for k, (c, t) in sorted(after.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"{t / 1e3:8.3f} ms {c:5d}  (mean {t / c:.2f} us)  {k}")


if __name__ == "__main__":
    main()
