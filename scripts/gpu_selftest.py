"""Developer GPU self-test: every C-ABI kernel family against a torch fp32 reference on the same device.

Each case group runs in its own subprocess under a timeout so that a device-side trap (the kernels turn protocol
deadlocks into traps) or a hang in one group cannot take the others down.  Usage (on a B200 box):

    python scripts/gpu_selftest.py [group ...]        # groups: generic temporal dense fused ip perf
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ref_sdpa(q, k, v, kv_group=1, scale=None):
    import torch
    import torch.nn.functional as F

    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    if kv_group > 1:
        kf = kf.repeat_interleave(kv_group, dim=0)
        vf = vf.repeat_interleave(kv_group, dim=0)
    o = F.scaled_dot_product_attention(qf, kf, vf, scale=scale)
    return o.permute(0, 2, 1, 3)


def report(name, out, ref, tol):
    import torch

    diff = (out.float() - ref.float()).abs()
    err = diff.max().item()
    rel = err / max(ref.float().abs().max().item(), 1e-9)
    bad = not (err <= tol) or torch.isnan(out.float()).any().item()
    print(f"[{'FAIL' if bad else ' ok '}] {name}: max_abs={err:.3e} rel_to_max={rel:.3e} tol={tol:g}", flush=True)
    if bad:
        # error structure helps to tell a descriptor bug from a softmax bug
        d = diff
        while d.dim() > 2:
            d = d.amax(dim=0)
        print("   per-row-block max err (first dims reduced):", [f"{x:.2e}" for x in d.amax(dim=1)[:16].tolist()])
        print("   per-col max err:", [f"{x:.2e}" for x in d.amax(dim=0)[:48].tolist()])
        print("   nan count:", torch.isnan(out.float()).sum().item(), "out absmax:", out.float().abs().max().item())
    return not bad


def group_generic():
    import torch
    from i2v_adapter_unofficial_b200 import ops

    torch.manual_seed(0)
    ok = True
    dev = "cuda"
    for dt, tol in ((torch.float32, 2e-5), (torch.bfloat16, 2e-2)):
        for (B, H, Sq, Skv, d, g) in ((2, 8, 200, 77, 40, 1), (4, 4, 65, 130, 160, 2), (3, 2, 33, 31, 16, 1)):
            q = torch.randn(B, Sq, H, d, device=dev, dtype=dt)
            k = torch.randn(B // g, Skv, H, d, device=dev, dtype=dt)
            v = torch.randn(B // g, Skv, H, d, device=dev, dtype=dt)
            o = ops.sdpa(q, k, v, g, None, ops.MODE_GENERIC)
            ok &= report(f"generic sdpa {dt} B{B} H{H} Sq{Sq} Skv{Skv} d{d} g{g}", o, ref_sdpa(q, k, v, g), tol)
        # ip
        B, H, Sq, d, g = 4, 8, 100, 40, 2
        q = torch.randn(B, Sq, H, d, device=dev, dtype=dt)
        k = torch.randn(B // g, 81, H, d, device=dev, dtype=dt)
        v = torch.randn(B // g, 81, H, d, device=dev, dtype=dt)
        o = ops.ip_xattn(q, k, v, 77, 0.7, g, None, ops.MODE_GENERIC)
        ref = ref_sdpa(q, k[:, :77], v[:, :77], g) + 0.7 * ref_sdpa(q, k[:, 77:], v[:, 77:], g)
        ok &= report(f"generic ip {dt}", o, ref, tol)
    return ok


def group_temporal():
    import torch
    from i2v_adapter_unofficial_b200 import ops

    torch.manual_seed(1)
    ok = True
    for (N, Fr, H, d) in ((512, 16, 8, 40), (300, 16, 8, 80), (200, 16, 8, 160), (256, 8, 8, 40), (128, 32, 8, 40),
                          (64, 24, 8, 80), (77, 16, 8, 64), (50, 32, 8, 160), (8192, 16, 8, 40)):
        qkv = torch.randn(N, Fr, 3, H, d, device="cuda", dtype=torch.bfloat16)
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        o = ops.temporal_attn(q, k, v, None, ops.MODE_FAST)
        torch.cuda.synchronize()
        ok &= report(f"temporal N{N} F{Fr} H{H} d{d}", o, ref_sdpa(q, k, v), 2e-2)
    return ok


def group_dense(cases=None):
    import torch
    from i2v_adapter_unofficial_b200 import ops

    torch.manual_seed(2)
    ok = True
    cases = cases or ((1, 1, 128, 128, 64, 1), (1, 1, 256, 256, 64, 1), (2, 2, 512, 384, 64, 1),
                      (1, 2, 256, 256, 40, 1), (2, 8, 1024, 1024, 40, 1), (2, 8, 200, 77, 40, 1),
                      (2, 8, 1024, 1024, 80, 1), (2, 8, 256, 256, 160, 1), (4, 8, 64, 64, 160, 1),
                      (4, 8, 576, 576, 32, 2), (2, 4, 300, 300, 16, 1), (2, 8, 4096, 4096, 40, 1))
    for (B, H, Sq, Skv, d, g) in cases:
        q = torch.randn(B, Sq, H, d, device="cuda", dtype=torch.bfloat16)
        k = torch.randn(B // g, Skv, H, d, device="cuda", dtype=torch.bfloat16)
        v = torch.randn(B // g, Skv, H, d, device="cuda", dtype=torch.bfloat16)
        o = ops.sdpa(q, k, v, g, None, ops.MODE_FAST)
        torch.cuda.synchronize()
        ok &= report(f"dense sdpa B{B} H{H} Sq{Sq} Skv{Skv} d{d} g{g}", o, ref_sdpa(q, k, v, g), 2e-2)
    return ok


def group_dense64():
    return group_dense(((1, 1, 128, 128, 64, 1), (1, 1, 256, 256, 64, 1), (2, 2, 512, 384, 64, 1)))


def group_fused():
    import torch
    from i2v_adapter_unofficial_b200 import ops

    torch.manual_seed(3)
    ok = True
    for (Bv, Fr, H, S, d) in ((2, 4, 8, 256, 40), (1, 16, 8, 1024, 40), (2, 3, 8, 320, 80), (2, 2, 8, 64, 160)):
        BF = Bv * Fr
        y = torch.randn(BF, S, 4, H, d, device="cuda", dtype=torch.bfloat16)
        kvx = torch.randn(Bv, S, 2, H, d, device="cuda", dtype=torch.bfloat16)
        o = ops.fused_self_xframe(y[:, :, 0], y[:, :, 1], y[:, :, 2], y[:, :, 3], kvx[:, :, 0], kvx[:, :, 1], Fr,
                                  None, ops.MODE_FAST)
        torch.cuda.synchronize()
        ok &= report(f"fused self  Bv{Bv} F{Fr} S{S} d{d}", o[:, :, 0], ref_sdpa(y[:, :, 0], y[:, :, 1], y[:, :, 2]),
                     2e-2)
        ok &= report(f"fused xframe Bv{Bv} F{Fr} S{S} d{d}", o[:, :, 1],
                     ref_sdpa(y[:, :, 3], kvx[:, :, 0], kvx[:, :, 1], Fr), 2e-2)
    return ok


def group_ip():
    import torch
    from i2v_adapter_unofficial_b200 import ops

    torch.manual_seed(4)
    ok = True
    for (B, H, Sq, d, g, nt, ni) in ((4, 8, 256, 40, 2, 77, 4), (2, 8, 1024, 80, 1, 77, 4), (2, 8, 64, 160, 2, 50, 14),
                                     (4, 8, 256, 160, 2, 77, 4), (2, 8, 200, 160, 1, 100, 28),
                                     (32, 8, 4096, 40, 16, 77, 4)):
        q = torch.randn(B, Sq, H, d, device="cuda", dtype=torch.bfloat16)
        k = torch.randn(B // g, nt + ni, H, d, device="cuda", dtype=torch.bfloat16)
        v = torch.randn(B // g, nt + ni, H, d, device="cuda", dtype=torch.bfloat16)
        o = ops.ip_xattn(q, k, v, nt, 0.6, g, None, ops.MODE_FAST)
        torch.cuda.synchronize()
        ref = ref_sdpa(q, k[:, :nt], v[:, :nt], g) + 0.6 * ref_sdpa(q, k[:, nt:], v[:, nt:], g)
        ok &= report(f"ip B{B} Sq{Sq} d{d} g{g} n{nt}+{ni}", o, ref, 2e-2)
    return ok


def _time(fn, iters=10, warm=3):
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def group_perf():
    import torch
    import torch.nn.functional as F
    from i2v_adapter_unofficial_b200 import ops

    torch.manual_seed(5)
    # dense, C2 level-0 shape: BF=32, H=8, S=4096, d=40
    for (B, H, S, d) in ((32, 8, 4096, 40), (32, 8, 1024, 80), (32, 8, 256, 160)):
        q = torch.randn(B, S, H, d, device="cuda", dtype=torch.bfloat16)
        k = torch.randn(B, S, H, d, device="cuda", dtype=torch.bfloat16)
        v = torch.randn(B, S, H, d, device="cuda", dtype=torch.bfloat16)
        ms = _time(lambda: ops.sdpa(q, k, v, 1, None, ops.MODE_FAST))
        fl = 4.0 * B * H * S * S * d
        qt, kt, vt = (t.permute(0, 2, 1, 3) for t in (q, k, v))
        ms_t = _time(lambda: F.scaled_dot_product_attention(qt, kt, vt))
        print(f"[perf] dense S{S} d{d}: {ms:.3f} ms = {fl / ms / 1e9:.1f} TFLOP/s | torch SDPA {ms_t:.3f} ms = "
              f"{fl / ms_t / 1e9:.1f} TFLOP/s", flush=True)
    # temporal, C2 shapes
    for (N, Fr, H, d) in ((8192, 16, 8, 40), (2048, 16, 8, 80), (512, 16, 8, 160), (8192, 32, 8, 40)):
        qkv = torch.randn(N, Fr, 3, H, d, device="cuda", dtype=torch.bfloat16)
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        by = 4.0 * N * Fr * H * d * 2
        for stages in (2, 3, 4):
            for per_sm in (1, 2):
                from i2v_adapter_unofficial_b200 import _lib
                _lib.load().i2v_set_tuning(0, stages)
                _lib.load().i2v_set_tuning(1, per_sm)
                try:
                    ms = _time(lambda: ops.temporal_attn(q, k, v, None, ops.MODE_FAST), iters=20)
                    print(f"[perf] temporal N{N} F{Fr} d{d} stages{stages} cta/sm{per_sm}: {ms * 1e3:.1f} us = "
                          f"{by / ms / 1e6:.0f} GB/s", flush=True)
                except Exception as e:  # noqa: BLE001
                    print(f"[perf] temporal N{N} F{Fr} d{d} stages{stages} cta/sm{per_sm}: {e}", flush=True)
        qt, kt, vt = (t.permute(0, 2, 1, 3) for t in (q, k, v))
        ms_t = _time(lambda: F.scaled_dot_product_attention(qt, kt, vt), iters=20)
        print(f"[perf] temporal torch SDPA N{N} F{Fr} d{d}: {ms_t * 1e3:.1f} us = {by / ms_t / 1e6:.0f} GB/s (alg. bytes)",
              flush=True)
    return True


def group_perfdense():
    """A few launches of the level-0 fused self + cross-frame kernel (C2 shape) for ncu captures."""
    import torch
    from i2v_adapter_unofficial_b200 import ops

    torch.manual_seed(3)
    Bv, Fr, H, S, d = 2, 16, 8, 4096, 40
    y = torch.randn(Bv * Fr, S, 4, H, d, device="cuda", dtype=torch.bfloat16)
    kvx = torch.randn(Bv, S, 2, H, d, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.fused_self_xframe(y[:, :, 0], y[:, :, 1], y[:, :, 2], y[:, :, 3], kvx[:, :, 0], kvx[:, :, 1], Fr,
                                       None, ops.MODE_FAST)
    ms = _time(fn, iters=3, warm=2)
    fl = 2 * 4.0 * Bv * Fr * H * S * S * d
    print(f"[perf] fused self+xframe L0 (C2): {ms:.3f} ms = {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
    return True


def group_tunedense():
    """Variant sweep of the level-0 fused kernel: tuning key 3 (tile variant) x key 2 (exp2 split)."""
    import torch
    from i2v_adapter_unofficial_b200 import _lib, ops

    torch.manual_seed(3)
    Bv, Fr, H, S, d = 2, 16, 8, 4096, 40
    y = torch.randn(Bv * Fr, S, 4, H, d, device="cuda", dtype=torch.bfloat16)
    kvx = torch.randn(Bv, S, 2, H, d, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.fused_self_xframe(y[:, :, 0], y[:, :, 1], y[:, :, 2], y[:, :, 3], kvx[:, :, 0], kvx[:, :, 1], Fr,
                                       None, ops.MODE_FAST)
    fl = 2 * 4.0 * Bv * Fr * H * S * S * d
    ref_self = ref_sdpa(y[:2, :, 0], y[:2, :, 1], y[:2, :, 2])
    lib = _lib.load()
    variants = [int(a) for a in os.environ.get("I2V_VARIANTS", "0,1").split(",")]
    for variant in variants:
        for emu in [int(a) for a in os.environ.get("I2V_EMUS", "0,1,3,4,5").split(",")]:
            lib.i2v_set_tuning(3, variant + 1)
            lib.i2v_set_tuning(2, emu)
            ms = _time(fn, iters=5, warm=2)
            o = fn()
            err = (o[:2, :, 0].float() - ref_self).abs().max().item()
            print(f"[tune] variant {variant} emu-key {emu}: {ms:.3f} ms = {fl / ms / 1e9:.1f} TFLOP/s  max_abs_err {err:.2e}",
                  flush=True)
    lib.i2v_set_tuning(3, 0)
    lib.i2v_set_tuning(2, 0)
    return True


def group_tuneaug():
    """Augmented-layout fused kernel (C2 level-0 shape): variant x exp2-split sweep, error against torch SDPA."""
    import torch
    from i2v_adapter_unofficial_b200 import _lib, ops

    torch.manual_seed(3)
    Bv, Fr, H, S, d = 2, 16, 8, int(os.environ.get("I2V_S", "4096")), 40
    sc = float(os.environ.get("I2V_QSCALE", "1.0"))
    q = torch.randn(Bv * Fr, S, H, d, device="cuda", dtype=torch.bfloat16) * sc
    k = torch.randn(Bv * Fr, S, H, d, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(Bv * Fr, S, H, d, device="cuda", dtype=torch.bfloat16)
    qx = torch.randn(Bv * Fr, S, H, d, device="cuda", dtype=torch.bfloat16) * sc
    kx = torch.randn(Bv, S, H, d, device="cuda", dtype=torch.bfloat16)
    vx = torch.randn(Bv, S, H, d, device="cuda", dtype=torch.bfloat16)
    qa, ka, va = ops.augment_qkv(q, k, v)
    qxa, kxa, vxa = ops.augment_qkv(qx, kx, vx)
    fn = lambda: ops.fused_self_xframe_aug(qa, ka, va, qxa, kxa, vxa, Fr)
    fl = 2 * 4.0 * Bv * Fr * H * S * S * d
    ref_self = ref_sdpa(q[:2], k[:2], v[:2])
    ref_x = ref_sdpa(qx[Fr:Fr + 2], kx[1:2], vx[1:2], kv_group=2)
    lib = _lib.load()
    ok = True
    for variant in [int(a) for a in os.environ.get("I2V_VARIANTS", "4").split(",")]:
        for emu in [int(a) for a in os.environ.get("I2V_EMUS", "1,3,4,5").split(",")]:
            lib.i2v_set_tuning(3, variant + 1)
            lib.i2v_set_tuning(2, emu)
            ms = _time(fn, iters=5, warm=2)
            o = fn()
            e1 = (o[:2, :, 0].float() - ref_self).abs().max().item()
            e2 = (o[Fr:Fr + 2, :, 1].float() - ref_x).abs().max().item()
            ok = ok and e1 < 2e-2 and e2 < 2e-2
            print(f"[tuneaug] variant {variant} emu-key {emu}: {ms:.3f} ms = {fl / ms / 1e9:.1f} TFLOP/s  max_abs_err self {e1:.2e} xframe {e2:.2e}",
                  flush=True)
    lib.i2v_set_tuning(3, 0)
    lib.i2v_set_tuning(2, 0)
    return ok


def group_perftemporal():
    import torch
    from i2v_adapter_unofficial_b200 import ops

    torch.manual_seed(5)
    N, Fr, H, d = 8192, 16, 8, 40
    qkv = torch.randn(N, Fr, 3, H, d, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.temporal_attn(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], None, ops.MODE_FAST)
    ms = _time(fn, iters=3, warm=2)
    print(f"[perf] temporal L0 (C2): {ms * 1e3:.1f} us = {4.0 * N * Fr * H * d * 2 / ms / 1e6:.0f} GB/s", flush=True)
    return True


GROUPS = {"tunedense": group_tunedense, "tuneaug": group_tuneaug, "perfdense": group_perfdense, "perftemporal": group_perftemporal, "generic": group_generic, "temporal": group_temporal, "dense64": group_dense64, "dense": group_dense,
          "fused": group_fused, "ip": group_ip, "perf": group_perf}


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--run":
        ok = GROUPS[sys.argv[2]]()
        sys.exit(0 if ok else 1)
    groups = sys.argv[1:] or ["generic", "temporal", "dense64", "dense", "fused", "ip"]
    summary = {}
    for g in groups:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", g], timeout=240)
            summary[g] = "ok" if r.returncode == 0 else f"FAILED (exit {r.returncode})"
        except subprocess.TimeoutExpired:
            summary[g] = "TIMEOUT"
        print(f"== group {g}: {summary[g]} ({time.time() - t0:.1f}s)", flush=True)
    print("SUMMARY", summary)
    sys.exit(0 if all(v == "ok" for v in summary.values()) else 1)


if __name__ == "__main__":
    main()
