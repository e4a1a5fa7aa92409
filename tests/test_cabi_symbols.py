"""The C-ABI library loads without a GPU, exports every symbol include/i2v_attn_b200.h declares, and fails loudly
(no CPU fallback) when there is no sm_100 device or when handed CPU tensors."""
import ctypes
import os
import re

import pytest
import torch

from i2v_adapter_unofficial_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "i2v_attn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(i2v_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS), (declared, _lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.i2v_version() >= 100


def test_library_is_native_sm100_code():
    data = open(_lib.LIB_PATH, "rb").read()
    assert b"sm_100a" in data or b"sm_100" in data


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_no_device_is_a_loud_error_not_a_fallback():
    lib = _lib.load()
    assert lib.i2v_device_supported() == 0
    assert lib.i2v_last_error()
    t = _lib.I2VTensor(ctypes.c_void_p(256), 64, 8, 8)
    rc = lib.i2v_sdpa_fwd(t, t, t, t, 1, 1, 8, 8, 8, 1, 1.0, _lib.I2V_BF16, _lib.MODE_AUTO, None)
    assert rc == -6  # I2V_ERR_NO_DEVICE
    with pytest.raises(_lib.I2VLibraryError):
        _lib.check(rc)


def test_argument_validation_happens_before_any_device_work():
    lib = _lib.load()
    t = _lib.I2VTensor(ctypes.c_void_p(256), 64, 8, 8)
    assert lib.i2v_sdpa_fwd(t, t, t, t, 0, 1, 8, 8, 8, 1, 1.0, 0, 0, None) == -1      # bad shape
    assert lib.i2v_sdpa_fwd(t, t, t, t, 1, 1, 8, 8, 8, 1, 1.0, 7, 0, None) == -5      # bad dtype
    assert lib.i2v_sdpa_fwd(t, t, t, t, 3, 1, 8, 8, 8, 2, 1.0, 0, 0, None) == -1      # batch % kv_group
    rc = lib.i2v_fused_self_xframe_fwd(t, t, t, t, t, t, t, t, 5, 1, 8, 8, 2, 1.0, 0, 0, None)
    assert rc == -1 and b"must be divisible by the number of frames" in lib.i2v_last_error()
    assert lib.i2v_reshard_pack(None, None, 1, 1, 7, 8, 2, 2, 0, None) == -1          # seq % world


def test_ops_reject_cpu_tensors():
    q = torch.randn(1, 8, 2, 16, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.sdpa(q, q, q)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.temporal_attn(q, q, q)
    ops.register_torch_ops()
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.i2v_b200.sdpa(q, q, q, 1, 1.0, 0)
