"""Level-0 fused self + cross-frame launch on the augmented layout: time (median of 20, library-side event pairs) and
check the kernel variants selected by the tuning keys given as `key=value` arguments, one combination per line:

    python scripts/perf_dense_sweep.py 2=0 2=1 2=3 2=4 2=5 2=6 "8=1" "8=2,2=5"
"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.getcwd())
from i2v_adapter_unofficial_b200 import _lib, ops  # noqa: E402

combos = sys.argv[1:] or ["2=0"]
torch.manual_seed(3)
Bv, Fr, H, S, d = 2, 16, 8, int(os.environ.get("SWEEP_S", "4096")), 40
mk = lambda b: torch.randn(b, S, H, d, device="cuda", dtype=torch.bfloat16)  # noqa: E731
q, k, v, qx, kx, vx = mk(Bv * Fr), mk(Bv * Fr), mk(Bv * Fr), mk(Bv * Fr), mk(Bv), mk(Bv)
qa, ka, va = ops.augment_qkv(q, k, v)
qxa, kxa, vxa = ops.augment_qkv(qx, kx, vx)
lib = _lib.load()
flops = 2 * 4.0 * Bv * Fr * H * S * S * d


def sdpa(q_, k_, v_):
    t = lambda x: x.transpose(1, 2).float()  # noqa: E731
    return torch.nn.functional.scaled_dot_product_attention(t(q_), t(k_), t(v_)).transpose(1, 2)


ref_s = sdpa(q[:2], k[:2], v[:2])
ref_x = sdpa(qx[15:17], kx[[0, 1]], vx[[0, 1]])
for combo in combos:
    pairs = [tuple(int(x) for x in kv.split("=")) for kv in combo.split(",")]
    for key, val in pairs:
        lib.i2v_set_tuning(key, val)
    try:
        for _ in range(3):
            o = ops.fused_self_xframe_aug(qa, ka, va, qxa, kxa, vxa, Fr)
        torch.cuda.synchronize()
        _lib.prof_arm(_lib.PROF_DENSE, S, Bv * Fr, 20)
        for _ in range(20):
            o = ops.fused_self_xframe_aug(qa, ka, va, qxa, kxa, vxa, Fr)
        torch.cuda.synchronize()
        ms = _lib.prof_read(_lib.PROF_DENSE)
        _lib.prof_arm(_lib.PROF_DENSE, 0, 0, 0)
        o_s, o_x = o[:, :, 0], o[:, :, 1]
        e_s = (o_s[:2, ..., :d].float() - ref_s).abs().max().item()
        e_x = (o_x[15:17, ..., :d].float() - ref_x).abs().max().item()
        med = statistics.median(ms)
        print(f"{combo:>16}: median {med:.3f} ms (min {min(ms):.3f} max {max(ms):.3f})  {flops / med / 1e9:.0f} TFLOP/s  "
              f"max|err| self {e_s:.4f} xframe {e_x:.4f}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"{combo:>16}: FAILED {type(e).__name__}: {e}", flush=True)
    for key, _ in pairs:
        lib.i2v_set_tuning(key, 0)
