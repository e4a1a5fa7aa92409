"""Developer tool: which Python lines still launch PyTorch elementwise adds inside the benchmarked denoise step."""
import collections
import os
import sys
import traceback

import torch
from torch.utils._python_dispatch import TorchDispatchMode

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from i2v_adapter_unofficial_b200 import install  # noqa: E402
from i2v_adapter_unofficial_b200.hostmodel import DDIMScheduler, denoise_step  # noqa: E402

WATCH = ("aten.add", "aten.add_", "aten.mul", "aten.cat", "aten.copy_", "aten.clone", "aten._to_copy", "aten.silu", "aten.sub", "aten.div")


class Spy(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.hits = collections.Counter()

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = str(func).rsplit(".", 1)[0]
        if name in WATCH:
            numel = max([a.numel() for a in args if isinstance(a, torch.Tensor)] + [0])
            if numel >= 1 << 10:
                st = [f for f in traceback.extract_stack() if "i2v_adapter_unofficial_b200" in f.filename or "bench.py" in f.filename]
                where = " <- ".join(f"{os.path.basename(f.filename)}:{f.lineno}" for f in st[-3:][::-1])
                self.hits[(name, where, numel)] += 1
        return func(*args, **(kwargs or {}))


def main():
    dev = torch.device("cuda:0")
    unet = bench.build_unet(dev, torch.bfloat16)
    install(unet)
    sched = DDIMScheduler()
    sched.set_timesteps(25)
    ts = [int(t) for t in sched.timesteps]
    d_in = bench.make_inputs(1, 16, 64, 1, torch.bfloat16, device=dev)
    lat = d_in["latents"].clone()
    with torch.no_grad():
        lat = denoise_step(unet, sched, lat, ts[0], d_in["prompt"], 7.5, d_in["cond"], d_in["image"])
        spy = Spy()
        with spy:
            lat = denoise_step(unet, sched, lat, ts[1], d_in["prompt"], 7.5, d_in["cond"], d_in["image"])
    for (name, where, numel), n in sorted(spy.hits.items(), key=lambda kv: -kv[1] * kv[0][2]):
        print(f"{n:4d} x {name:14s} numel {numel:10d}  {where}")


if __name__ == "__main__":
    main()
