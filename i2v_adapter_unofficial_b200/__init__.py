"""i2v_adapter_unofficial_b200 — B200 (sm_100a) implementation of the denoising-UNet attention hot path of
xUhEngwAng/I2V-Adapter-Unofficial behind the reference's diffusers ``AttnProcessor`` boundary.

    from i2v_adapter_unofficial_b200 import install
    handle = install(unet)          # after load_ip_adapter(...); unet = UNetMotionCrossFrameAttnModel
    ...                             # pipeline / unet forward unchanged
    handle.uninstall()

Layout: ``csrc/`` CUDA kernels + C ABI (``include/i2v_attn_b200.h``), ``_lib`` ctypes binding, ``ops`` tensor-level
wrappers (also ``torch.ops.i2v_b200.*``), ``processors`` the drop-in processors, ``partition`` the frame / batch
partitioner, ``hostmodel`` a minimal diffusers-0.25-compatible mirror of the reference's UNet classes (diffusers is
not installable in this environment).
"""
from . import _lib, ops, partition  # noqa: F401
from .processors import (  # noqa: F401
    B200AttnProcessor,
    B200CrossFrameAttnProcessor,
    B200IPAdapterAttnProcessor,
    B200SpatialAttnProcessor,
    B200TemporalAttnProcessor,
    install,
)

__all__ = [
    "install",
    "ops",
    "partition",
    "B200AttnProcessor",
    "B200SpatialAttnProcessor",
    "B200CrossFrameAttnProcessor",
    "B200IPAdapterAttnProcessor",
    "B200TemporalAttnProcessor",
]
