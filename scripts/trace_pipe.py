"""Developer tool: in-kernel timeline of one CTA of the pipelined dense kernel on the augmented layout (needs a
-DI2V_TRACE build: I2V_ATTN_LIB=build/libtrace.so).  Prints the mean duration of every phase of a KV step for the traced
softmax warps (warp 0 and warp 3 of each tile) and MMA warps, and the hand-off latencies between them."""
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
torch.manual_seed(3)
Bv, Fr, H, S, d = 2, 16, 8, 4096, 40
mk = lambda b: torch.randn(b, S, H, d, device="cuda", dtype=torch.bfloat16)  # noqa: E731
qa, ka, va = ops.augment_qkv(mk(Bv * Fr), mk(Bv * Fr), mk(Bv * Fr))
qxa, kxa, vxa = ops.augment_qkv(mk(Bv * Fr), mk(Bv), mk(Bv))
fn = lambda: ops.fused_self_xframe_aug(qa, ka, va, qxa, kxa, vxa, Fr)  # noqa: E731
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
SM = {0x1: "wait S", 0x2: "S ready", 0x3: "S in regs", 0x4: "freed S, masks", 0x5: "max + exp done", 0x6: "PV(j-1) done",
      0x7: "P stored"}
MM = {0x10: "qk: start", 0x11: "K landed", 0x12: "S free", 0x13: "QK issued", 0x14: "V landed", 0x15: "P full"}


def phases(events, first_tag, names):
    """split into steps at first_tag; mean duration from each event to the next one, steady-state steps only"""
    steps, cur = [], []
    for tag, clk in events:
        if tag == first_tag and cur:
            steps.append(cur)
            cur = []
        cur.append((tag, clk))
    steps = [s for s in steps[8:56] if len(s) == len(steps[8])]
    out = []
    for i in range(len(steps[0])):
        tag = steps[0][i][0]
        nxt = [(s[i + 1][1] if i + 1 < len(s) else None) for s in steps]
        durs = []
        for k, s in enumerate(steps):
            if nxt[k] is not None:
                durs.append(nxt[k] - s[i][1])
            elif k + 1 < len(steps):
                durs.append(steps[k + 1][0][1] - s[i][1])
        out.append((names.get(tag, hex(tag)), statistics.mean(durs), min(durs), max(durs)))
    total = statistics.mean([steps[k + 1][0][1] - steps[k][0][1] for k in range(len(steps) - 1)])
    return out, total, steps


for slot in (0, 1, 2, 4, 5, 6):
    if len(ev[slot]) < 100:
        continue
    out, total, _ = phases(ev[slot], 0x1, SM)
    print(f"--- softmax warp {0 if slot < 4 else 3} of tile {slot % 4}: {total:7.0f} clk per KV step")
    for name, mean, lo, hi in out:
        print(f"      after '{name:16s}' {mean:7.0f} clk (min {lo:6d} max {hi:6d})")
for slot in (8, 9, 10):
    if len(ev[slot]) < 100:
        continue
    out, total, _ = phases(ev[slot], 0x10, MM)
    print(f"--- MMA warp of tile {slot - 8}: {total:7.0f} clk per KV step")
    for name, mean, lo, hi in out:
        print(f"      after '{name:16s}' {mean:7.0f} clk (min {lo:6d} max {hi:6d})")
# hand-off latencies for tile 0: softmax "S in regs" (s_free arrive follows at once) -> MMA "S free" -> MMA "QK issued" -> softmax "S ready"
s0 = [(t, c) for t, c in ev[0]]
s3 = [(t, c) for t, c in ev[4]]
m0 = [(t, c) for t, c in ev[8]]
inregs0 = [c for t, c in s0 if t == 0x3]
inregs3 = [c for t, c in s3 if t == 0x3]
sfree = [c for t, c in m0 if t == 0x12]
issued = [c for t, c in m0 if t == 0x13]
ready = [c for t, c in s0 if t == 0x2]
pst0 = [c for t, c in s0 if t == 0x7]
pst3 = [c for t, c in s3 if t == 0x7]
pfull = [c for t, c in m0 if t == 0x15]
n = min(len(inregs0), len(inregs3), len(sfree) - 1, len(issued) - 1, len(ready) - 1, len(pst0), len(pst3), len(pfull)) - 2
win = range(10, min(n, 56))
# QK(j+1) is the (j+1)-th qk_step: index j+1 in sfree / issued; S(j+1) ready is index j+1 in ready
print("--- tile 0 hand-offs (mean clk over steady-state steps)")
print(f"      warp 0 vs warp 3 'S in regs' skew        {statistics.mean([abs(inregs0[j] - inregs3[j]) for j in win]):7.0f}")
print(f"      last 'S in regs' -> MMA sees 'S free'    {statistics.mean([sfree[j + 1] - max(inregs0[j], inregs3[j]) for j in win]):7.0f}")
print(f"      MMA 'QK issued' -> softmax 'S ready'     {statistics.mean([ready[j + 1] - issued[j + 1] for j in win]):7.0f}   (negative: S was ready before the warp asked)")
print(f"      warp 0 vs warp 3 'P stored' skew         {statistics.mean([abs(pst0[j] - pst3[j]) for j in win]):7.0f}")
print(f"      last 'P stored' -> MMA sees 'P full'     {statistics.mean([pfull[j] - max(pst0[j], pst3[j]) for j in win]):7.0f}")
