"""Developer tool: SM clock and board power while one kernel variant runs back to back for ~1 s (is the dense kernel
power-limited?).  python scripts/power_probe.py [key=value ...]"""
import os
import statistics
import sys
import threading
import time

import torch

sys.path.insert(0, os.getcwd())
from i2v_adapter_unofficial_b200 import _lib, ops  # noqa: E402
import pynvml  # noqa: E402

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
lib = _lib.load()
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    lib.i2v_set_tuning(int(k), int(v))
torch.manual_seed(3)
Bv, Fr, H, S, d = 2, 16, 8, 4096, 40
mk = lambda b: torch.randn(b, S, H, d, device="cuda", dtype=torch.bfloat16)  # noqa: E731
qa, ka, va = ops.augment_qkv(mk(Bv * Fr), mk(Bv * Fr), mk(Bv * Fr))
qxa, kxa, vxa = ops.augment_qkv(mk(Bv * Fr), mk(Bv), mk(Bv))
fn = lambda: ops.fused_self_xframe_aug(qa, ka, va, qxa, kxa, vxa, Fr)  # noqa: E731
for _ in range(3):
    fn()
torch.cuda.synchronize()
samples, stop = [], threading.Event()


def poll():
    while not stop.is_set():
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.005)


print("idle: clock", pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), "MHz, power",
      pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, "W, limit", pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0, "W, max clock",
      pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
th = threading.Thread(target=poll, daemon=True)
th.start()
n = 600
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    fn()
e1.record()
torch.cuda.synchronize()
t1 = time.perf_counter()
stop.set()
th.join()
ms = e0.elapsed_time(e1) / n
win = [s for s in samples if t0 + 0.3 * (t1 - t0) <= s[0] <= t1]
first = [s for s in samples if s[0] <= t0 + 0.1 * (t1 - t0)]
print(f"{' '.join(sys.argv[1:]) or 'default'}: {ms:.3f} ms per launch over {n} launches; late window: SM clock median "
      f"{statistics.median(s[1] for s in win)} MHz, power median {statistics.median(s[2] for s in win):.0f} W (max "
      f"{max(s[2] for s in win):.0f}); first 10 %: clock {statistics.median(s[1] for s in first)} MHz, power "
      f"{statistics.median(s[2] for s in first):.0f} W")
