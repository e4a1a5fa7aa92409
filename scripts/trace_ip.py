"""Developer tool: in-kernel timeline of one CTA of the IP-Adapter tcgen05 kernel (needs a -DI2V_TRACE build:
I2V_ATTN_LIB=build/libtrace.so).  Mean duration of every phase per query tile for the softmax warpgroups (warp 0 of each),
the MMA warps and the TMA warp."""
import ctypes
import os
import statistics
import sys

import torch

sys.path.insert(0, os.getcwd())
from i2v_adapter_unofficial_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    lib.i2v_set_tuning(int(k), int(v))
B, S, d, H, g, nt, ni = 32, 4096, 40, 8, 16, 77, 4
q = torch.randn(B, S, H, d, device="cuda", dtype=torch.bfloat16)
kvt = torch.randn(B // g, nt + ni, 2, H, d, device="cuda", dtype=torch.bfloat16)
fn = lambda: ops.ip_xattn(q, kvt[:, :, 0], kvt[:, :, 1], nt, 1.0, g, None, ops.MODE_FAST)  # noqa: E731
for _ in range(3):
    fn()
buf = torch.zeros(16 * 1024, dtype=torch.int64, device="cuda")
lib.i2v_debug_set_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.i2v_debug_set_trace(ctypes.c_void_p(buf.data_ptr()), int(os.environ.get("TRACE_CTA", "70")))
fn()
torch.cuda.synchronize()
lib.i2v_debug_set_trace(None, 0)
b = buf.cpu().view(16, 1024)
ev = {s: [(int(x) >> 48, int(x) & 0xffffffffffff) for x in b[s] if int(x) != 0] for s in range(16)}
NAMES = {0x1: "wait S", 0x2: "S ready", 0x3: "P computed", 0x4: "P stored", 0x5: "O ready", 0x6: "O in regs, S freed", 0x7: "stores issued",
         0x10: "tile start", 0x11: "Q landed", 0x12: "slot free", 0x13: "QK issued", 0x14: "P full", 0x15: "PV issued",
         0x20: "wait Q slot", 0x21: "Q slot free"}


def phases(events, first_tag):
    steps, cur = [], []
    for tag, clk in events:
        if tag == first_tag and cur:
            steps.append(cur)
            cur = []
        cur.append((tag, clk))
    steps = [s for s in steps[3:-2] if len(s) == len(steps[3])]
    out = []
    for i in range(len(steps[0])):
        durs = []
        for k, s in enumerate(steps):
            if i + 1 < len(s):
                durs.append(s[i + 1][1] - s[i][1])
            elif k + 1 < len(steps):
                durs.append(steps[k + 1][0][1] - s[i][1])
        out.append((NAMES.get(steps[0][i][0], hex(steps[0][i][0])), statistics.mean(durs), min(durs), max(durs)))
    total = statistics.mean([steps[k + 1][0][1] - steps[k][0][1] for k in range(len(steps) - 1)])
    return out, total, len(steps)


for slot, first, what in [(0, 0x1, "softmax warpgroup 0"), (1, 0x1, "softmax warpgroup 1"), (2, 0x1, "softmax warpgroup 2"),
                          (8, 0x10, "MMA warp 0"), (9, 0x10, "MMA warp 1"), (10, 0x10, "MMA warp 2"), (12, 0x20, "TMA warp")]:
    if len(ev[slot]) < 20:
        continue
    out, total, n = phases(ev[slot], first)
    print(f"--- {what}: {total:7.0f} clk per tile ({n} tiles)")
    for name, mean, lo, hi in out:
        print(f"      after '{name:20s}' {mean:7.0f} clk (min {lo:6d} max {hi:6d})")
t0 = min(c for s in ev.values() for _, c in s)
t1 = max(c for s in ev.values() for _, c in s)
print(f"traced CTA: {t1 - t0} clk from first to last event")
