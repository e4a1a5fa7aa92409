"""Level-1 fused self + cross-frame launch (d = 80, S = 1024, 32 frames x 8 heads): time (median of 20, library-side event
pairs) and check the kernel variants selected by `key=value` tuning arguments (default: pipelined kernel, two query tiles x 64 keys; 3=6: three x 48 keys; 3=1: the first tcgen05 kernel):

    python scripts/perf_dense80.py 3=1 3=0 3=6 "3=0,2=3" "3=0,2=5" "3=0,2=6"
"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.getcwd())
from i2v_adapter_unofficial_b200 import _lib, ops  # noqa: E402

combos = sys.argv[1:] or ["3=0"]
torch.manual_seed(3)
Bv, Fr, H, d = 2, 16, 8, 80
S = int(os.environ.get("SWEEP_S", "1024"))
mk = lambda b: torch.randn(b, S, H, d, device="cuda", dtype=torch.bfloat16)  # noqa: E731
q, k, v, qx, kx, vx = mk(Bv * Fr), mk(Bv * Fr), mk(Bv * Fr), mk(Bv * Fr), mk(Bv), mk(Bv)
lib = _lib.load()
flops = 2 * 4.0 * Bv * Fr * H * S * S * d


def sdpa(q_, k_, v_):
    t = lambda x: x.transpose(1, 2).float()  # noqa: E731
    return torch.nn.functional.scaled_dot_product_attention(t(q_), t(k_), t(v_)).transpose(1, 2)


ref_s = sdpa(q, k, v)
ref_x = sdpa(qx, kx.repeat_interleave(Fr, 0), vx.repeat_interleave(Fr, 0))
for combo in combos:
    pairs = [tuple(int(x) for x in kv.split("=")) for kv in combo.split(",")]
    for key, val in pairs:
        lib.i2v_set_tuning(key, val)
    try:
        for _ in range(3):
            o = ops.fused_self_xframe(q, k, v, qx, kx, vx, Fr)
        torch.cuda.synchronize()
        _lib.prof_arm(_lib.PROF_DENSE, S, Bv * Fr, 20)
        for _ in range(20):
            o = ops.fused_self_xframe(q, k, v, qx, kx, vx, Fr)
        torch.cuda.synchronize()
        ms = _lib.prof_read(_lib.PROF_DENSE)
        _lib.prof_arm(_lib.PROF_DENSE, 0, 0, 0)
        e_s = (o[:, :, 0].float() - ref_s).abs().max().item()
        e_x = (o[:, :, 1].float() - ref_x).abs().max().item()
        med = statistics.median(ms)
        print(f"{combo:>16}: median {med * 1e3:.1f} us (min {min(ms) * 1e3:.1f} max {max(ms) * 1e3:.1f})  {flops / med / 1e9:.0f} TFLOP/s  "
              f"max|err| self {e_s:.4f} xframe {e_x:.4f}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"{combo:>16}: FAILED {type(e).__name__}: {e}", flush=True)
    for key, _ in pairs:
        lib.i2v_set_tuning(key, 0)
