"""Developer tool: timeline of one CTA of the dense kernel (needs a -DI2V_TRACE build, I2V_ATTN_LIB=...)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from i2v_adapter_unofficial_b200 import _lib, ops  # noqa: E402

variant = int(os.environ.get("I2V_VARIANTS", "2"))
emu = int(os.environ.get("I2V_EMUS", "1"))
lib = _lib.load()
lib.i2v_set_tuning(3, variant + 1)
lib.i2v_set_tuning(2, emu)
Bv, Fr, H, S, d = 2, 16, 8, 4096, 40
torch.manual_seed(0)
y = torch.randn(Bv * Fr, S, 4, H, d, device="cuda", dtype=torch.bfloat16)
kvx = torch.randn(Bv, S, 2, H, d, device="cuda", dtype=torch.bfloat16)
fn = lambda: ops.fused_self_xframe(y[:, :, 0], y[:, :, 1], y[:, :, 2], y[:, :, 3], kvx[:, :, 0], kvx[:, :, 1], Fr,
                                   None, ops.MODE_FAST)
for _ in range(2):
    fn()
buf = torch.zeros(16 * 1024, dtype=torch.int64, device="cuda")
lib.i2v_debug_set_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.i2v_debug_set_trace(ctypes.c_void_p(buf.data_ptr()), 2000)
fn()
torch.cuda.synchronize()
lib.i2v_debug_set_trace(None, 0)
b = buf.cpu().view(16, 1024)
names = {0x1: "S.wait", 0x2: "S.got", 0x3: "S.ld", 0x4: "S.max", 0x5: "S.exp", 0x6: "S.arr", 0x10: "M.wait", 0x11: "M.got",
         0x12: "M.iss"}
t0 = None
for slot in range(16):
    ev = [(int(x) >> 48, int(x) & 0xffffffffffff) for x in b[slot] if int(x) != 0]
    if not ev:
        continue
    if t0 is None:
        t0 = min(e[1] for e in ev)
    print(f"--- slot {slot} ({'softmax wg%d' % slot if slot < 8 else 'mma warp %d' % (slot - 8)}), {len(ev)} events")
    # steady-state window: events 200..290
    prev = None
    for tag, clk in ev[120:120 + int(os.environ.get("NEV", "48"))]:
        nm = names.get(tag >> 4, hex(tag))
        print(f"   {nm:7s} t{tag & 0xf}  @{clk - t0:8d}  +{0 if prev is None else clk - prev:5d}")
        prev = clk
