"""Minimal diffusers-0.25-compatible mirror of the module classes the reference's UNet is assembled from
(``diffusers`` cannot be installed in this environment).  Host / plumbing code on stock PyTorch."""
from .attention import Attention, AttnProcessor2_0, IPAdapterAttnProcessor2_0  # noqa: F401
from .ddim import DDIMScheduler  # noqa: F401
from .i2v_adapter import I2VAdapterModule, I2VAdapterTransformer2DModel, I2VAdapterTransformerBlock  # noqa: F401
from .layers import BasicTransformerBlock, ImageProjection  # noqa: F401
from .pipeline import denoise, denoise_step  # noqa: F401
from .checkpoint import load_ip_adapter_file, save_ip_adapter_file  # noqa: F401
from .temporal import DownBlockMotion, MotionAdapter, MotionModules, TransformerTemporalModel, UpBlockMotion  # noqa: F401
from .unet import (  # noqa: F401
    CrossFrameAttnDownBlockMotion,
    CrossFrameAttnUpBlockMotion,
    UNetMidBlockCrossFrameAttnMotion,
    UNetMotionCrossFrameAttnModel,
)
