"""ctypes binding of ``libi2v_attn_b200.so`` (C ABI declared in ``include/i2v_attn_b200.h``).

This module is deliberately thin: it loads the in-tree shared library, declares the argument types of every exported
symbol and turns negative status codes into Python exceptions.  There is no fallback of any kind: if the library is
missing, cannot be loaded, or the device is not sm_100, the error propagates to the caller.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libi2v_attn_b200.so"
LIB_PATH = os.environ.get("I2V_ATTN_LIB", os.path.join(_HERE, LIB_NAME))  # env override: developer A/B builds

I2V_BF16, I2V_F32 = 0, 1
MODE_AUTO, MODE_FAST, MODE_GENERIC = 0, 1, 2

#: every symbol include/i2v_attn_b200.h declares (checked by tests/test_cabi_symbols.py)
EXPORTED_SYMBOLS = (
    "i2v_version",
    "i2v_last_error",
    "i2v_device_supported",
    "i2v_launch_count",
    "i2v_sdpa_fwd",
    "i2v_fused_self_xframe_fwd",
    "i2v_fused_self_xframe_aug_fwd",
    "i2v_ip_xattn_fwd",
    "i2v_temporal_attn_fwd",
    "i2v_reshard_pack",
    "i2v_reshard_unpack",
    "i2v_set_tuning",
    "i2v_prof_arm",
    "i2v_prof_read",
    "i2v_layernorm_fwd",
    "i2v_layernorm_pre_fwd",
    "i2v_geglu_ld_fwd",
    "i2v_ff_geglu_fwd",
    "i2v_linear_fwd",
    "i2v_upsample2x_nhwc",
    "i2v_geglu_fwd",
    "i2v_gn_stats",
    "i2v_gn_apply_transpose",
    "i2v_untranspose_residual",
    "i2v_gn_nhwc_scratch_floats",
    "i2v_gn_nhwc",
    "i2v_gn_nhwc_cat",
    "i2v_rows_residual",
    "i2v_rows_residual_bias",
    "i2v_gn_nhwc_sums",
    "i2v_gn_nhwc_apply",
    "i2v_rows_residual_sharded",
)


class I2VTensor(ctypes.Structure):
    """Mirror of ``i2v_tensor``: device pointer + element strides (batch, seq, head); d is contiguous."""

    _fields_ = [
        ("data", ctypes.c_void_p),
        ("stride_b", ctypes.c_int64),
        ("stride_s", ctypes.c_int64),
        ("stride_h", ctypes.c_int64),
    ]


class I2VLibraryError(RuntimeError):
    """A C-ABI call returned a negative status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libi2v_attn_b200 error {code}: {message}")
        self.code = code
        self.message = message


_lib: Optional[ctypes.CDLL] = None


def _declare(lib: ctypes.CDLL) -> None:
    T = ctypes.POINTER(I2VTensor)
    i, f, p = ctypes.c_int, ctypes.c_float, ctypes.c_void_p
    lib.i2v_version.restype = i
    lib.i2v_version.argtypes = []
    lib.i2v_last_error.restype = ctypes.c_char_p
    lib.i2v_last_error.argtypes = []
    lib.i2v_device_supported.restype = i
    lib.i2v_device_supported.argtypes = []
    lib.i2v_launch_count.restype = ctypes.c_int64
    lib.i2v_launch_count.argtypes = []
    lib.i2v_sdpa_fwd.restype = i
    lib.i2v_sdpa_fwd.argtypes = [T, T, T, T, i, i, i, i, i, i, f, i, i, p]
    lib.i2v_fused_self_xframe_fwd.restype = i
    lib.i2v_fused_self_xframe_fwd.argtypes = [T, T, T, T, T, T, T, T, i, i, i, i, i, f, i, i, p]
    lib.i2v_fused_self_xframe_aug_fwd.restype = i
    lib.i2v_fused_self_xframe_aug_fwd.argtypes = [T, T, T, T, T, T, T, T, i, i, i, i, i, i, i, p]
    lib.i2v_ip_xattn_fwd.restype = i
    lib.i2v_ip_xattn_fwd.argtypes = [T, T, T, T, T, T, i, i, i, i, i, i, i, f, f, i, i, p]
    lib.i2v_temporal_attn_fwd.restype = i
    lib.i2v_temporal_attn_fwd.argtypes = [T, T, T, T, i, i, i, i, f, i, i, p]
    lib.i2v_reshard_pack.restype = i
    lib.i2v_reshard_pack.argtypes = [p, p, i, i, i, i, i, i, i, p]
    lib.i2v_reshard_unpack.restype = i
    lib.i2v_reshard_unpack.argtypes = [p, p, i, i, i, i, i, i, i, p]
    lib.i2v_set_tuning.restype = i
    lib.i2v_set_tuning.argtypes = [i, i]
    ll = ctypes.c_longlong
    lib.i2v_prof_arm.restype = i
    lib.i2v_prof_arm.argtypes = [i, ll, ll, i]
    lib.i2v_prof_read.restype = i
    lib.i2v_prof_read.argtypes = [i, ctypes.POINTER(ctypes.c_float), i]
    lib.i2v_layernorm_fwd.restype = i
    lib.i2v_layernorm_fwd.argtypes = [p, p, p, p, p, ll, i, i, f, p]
    lib.i2v_geglu_fwd.restype = i
    lib.i2v_geglu_fwd.argtypes = [p, p, ll, i, p]
    lib.i2v_layernorm_pre_fwd.restype = i
    lib.i2v_layernorm_pre_fwd.argtypes = [p, p, p, p, p, p, ll, i, i, f, p]
    lib.i2v_geglu_ld_fwd.restype = i
    lib.i2v_geglu_ld_fwd.argtypes = [p, p, ll, i, i, p]
    lib.i2v_upsample2x_nhwc.restype = i
    lib.i2v_upsample2x_nhwc.argtypes = [p, p, i, i, i, i, p]
    lib.i2v_ff_geglu_fwd.restype = i
    lib.i2v_ff_geglu_fwd.argtypes = [p, p, p, p, ll, i, i, i, p]
    lib.i2v_linear_fwd.restype = i
    lib.i2v_linear_fwd.argtypes = [p, p, p, p, p, ll, i, i, i, i, i, p]
    lib.i2v_gn_stats.restype = i
    lib.i2v_gn_stats.argtypes = [p, p, i, i, i, i, p]
    lib.i2v_gn_apply_transpose.restype = i
    lib.i2v_gn_apply_transpose.argtypes = [p, p, p, p, p, i, i, i, i, i, f, p]
    lib.i2v_untranspose_residual.restype = i
    lib.i2v_untranspose_residual.argtypes = [p, p, p, i, i, i, i, p]
    lib.i2v_gn_nhwc_scratch_floats.restype = ctypes.c_longlong
    lib.i2v_gn_nhwc_scratch_floats.argtypes = [i, i]
    lib.i2v_gn_nhwc.restype = i
    lib.i2v_gn_nhwc.argtypes = [p, p, p, p, p, p, i, i, i, i, i, f, i, i, p]
    lib.i2v_gn_nhwc_cat.restype = i
    lib.i2v_gn_nhwc_cat.argtypes = [p, p, i, p, p, p, p, p, i, i, i, i, i, f, i, i, p]
    lib.i2v_rows_residual.restype = i
    lib.i2v_rows_residual.argtypes = [p, p, p, i, i, i, i, p]
    lib.i2v_gn_nhwc_sums.restype = i
    lib.i2v_gn_nhwc_sums.argtypes = [p, p, p, i, i, i, i, i, p]
    lib.i2v_gn_nhwc_apply.restype = i
    lib.i2v_gn_nhwc_apply.argtypes = [p, p, p, p, p, p, i, i, i, i, i, i, i, i, p]
    lib.i2v_rows_residual_sharded.restype = i
    lib.i2v_rows_residual_sharded.argtypes = [p, p, p, i, i, i, i, i, p]
    lib.i2v_rows_residual_bias.restype = i
    lib.i2v_rows_residual_bias.argtypes = [p, p, p, p, i, i, i, i, p]


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises ``FileNotFoundError`` / ``OSError`` when it is absent or broken."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU or PyTorch fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        _declare(lib)
        _lib = lib
    return _lib


def check(code: int) -> None:
    if code != 0:
        msg = load().i2v_last_error()
        raise I2VLibraryError(code, msg.decode("utf-8", "replace") if msg else "")


def launch_count() -> int:
    return int(load().i2v_launch_count())


PROF_DENSE, PROF_TEMPORAL, PROF_IP, PROF_GEMM = 1, 2, 3, 4


def prof_arm(kind: int, match_a: int = 0, match_b: int = 0, max_pairs: int = 64) -> None:
    """Bracket the next ``max_pairs`` matching launches of a kernel class with CUDA events (``i2v_prof_arm``)."""
    check(load().i2v_prof_arm(kind, match_a, match_b, max_pairs))


def prof_read(kind: int, capacity: int = 256):
    """Durations (ms) of the recorded pairs; synchronise first."""
    buf = (ctypes.c_float * capacity)()
    n = load().i2v_prof_read(kind, buf, capacity)
    if n < 0:
        check(n)
    return [float(buf[i]) for i in range(n)]
