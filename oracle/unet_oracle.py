"""CPU oracle for the callers of the attention hot path — TEST INFRASTRUCTURE (see attention_oracle.py header).

Functional restatement of the UNet wiring around the attention operators: the cross-frame blocks, the UNet forward
(the frame / batch layout owner) and the pipeline's denoise step.  Parity at this level is UNPINNED by the reference
(it ships no numeric tests and its ``diffusers`` dependency is not installable here); each function follows the
reference lines cited in its docstring and SURVEY.md Appendix A.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

from .attention_oracle import StateDict, _linear, temporal_model_oracle, transformer2d_oracle


def resnet_oracle(sd: StateDict, prefix: str, x, temb, groups: int, eps: float):
    """diffusers ``ResnetBlock2D``: GN-SiLU-conv3x3, + time projection, GN-SiLU-conv3x3,
    + (1x1-conv) shortcut."""
    h = F.group_norm(x, groups, sd[f"{prefix}.norm1.weight"], sd[f"{prefix}.norm1.bias"], eps)
    h = F.conv2d(F.silu(h), sd[f"{prefix}.conv1.weight"], sd[f"{prefix}.conv1.bias"], padding=1)
    h = h + _linear(F.silu(temb), sd, f"{prefix}.time_emb_proj")[:, :, None, None]
    h = F.group_norm(h, groups, sd[f"{prefix}.norm2.weight"], sd[f"{prefix}.norm2.bias"], eps)
    h = F.conv2d(F.silu(h), sd[f"{prefix}.conv2.weight"], sd[f"{prefix}.conv2.bias"], padding=1)
    if f"{prefix}.conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[f"{prefix}.conv_shortcut.weight"], sd[f"{prefix}.conv_shortcut.bias"])
    return x + h


def _count(sd: StateDict, prefix: str, leaf: str) -> int:
    n = 0
    while f"{prefix}.{n}.{leaf}" in sd:
        n += 1
    return n


def down_block_oracle(sd: StateDict, prefix: str, x, temb, ctx, cfg, heads: int, enable: bool, num_frames: int,
                      ip_tokens: int, ip_scale: float):
    """``CrossFrameAttnDownBlockMotion.forward`` (src/models/unet_motion_cross_frame_attn.py:265-340) or, when the
    block has no ``attentions``, diffusers ``DownBlockMotion.forward``:
    per layer resnet -> [I2V spatial transformer] -> motion module; then the stride-2 conv."""
    outs = ()
    has_attn = f"{prefix}.attentions.0.norm.weight" in sd
    for i in range(_count(sd, f"{prefix}.resnets", "norm1.weight")):
        x = resnet_oracle(sd, f"{prefix}.resnets.{i}", x, temb, cfg["norm_num_groups"], cfg["norm_eps"])
        if has_attn:
            x = transformer2d_oracle(sd, f"{prefix}.attentions.{i}", x, ctx, heads, cfg["norm_num_groups"], enable,
                                     num_frames, ip_tokens, ip_scale)
        x = temporal_model_oracle(sd, f"{prefix}.motion_modules.{i}", x, num_frames,
                                  cfg["motion_num_attention_heads"], cfg["norm_num_groups"])
        outs += (x,)
    if f"{prefix}.downsamplers.0.conv.weight" in sd:
        x = F.conv2d(x, sd[f"{prefix}.downsamplers.0.conv.weight"], sd[f"{prefix}.downsamplers.0.conv.bias"], stride=2,
                     padding=1)
        outs += (x,)
    return x, outs


def up_block_oracle(sd: StateDict, prefix: str, x, res: Sequence[torch.Tensor], temb, ctx, cfg, heads: int,
                    enable: bool, num_frames: int, ip_tokens: int, ip_scale: float):
    """``CrossFrameAttnUpBlockMotion.forward`` (:439-529) / diffusers ``UpBlockMotion.forward``: pop skip, concat,
    resnet -> [I2V spatial transformer] -> motion module; nearest-2x upsample + conv."""
    res = tuple(res)
    has_attn = f"{prefix}.attentions.0.norm.weight" in sd
    for i in range(_count(sd, f"{prefix}.resnets", "norm1.weight")):
        x = torch.cat([x, res[-1]], dim=1)
        res = res[:-1]
        x = resnet_oracle(sd, f"{prefix}.resnets.{i}", x, temb, cfg["norm_num_groups"], cfg["norm_eps"])
        if has_attn:
            x = transformer2d_oracle(sd, f"{prefix}.attentions.{i}", x, ctx, heads, cfg["norm_num_groups"], enable,
                                     num_frames, ip_tokens, ip_scale)
        x = temporal_model_oracle(sd, f"{prefix}.motion_modules.{i}", x, num_frames,
                                  cfg["motion_num_attention_heads"], cfg["norm_num_groups"])
    if f"{prefix}.upsamplers.0.conv.weight" in sd:
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        x = F.conv2d(x, sd[f"{prefix}.upsamplers.0.conv.weight"], sd[f"{prefix}.upsamplers.0.conv.bias"], padding=1)
    return x


def mid_block_oracle(sd: StateDict, prefix: str, x, temb, ctx, cfg, heads: int, enable: bool, num_frames: int,
                     ip_tokens: int, ip_scale: float):
    """``UNetMidBlockCrossFrameAttnMotion.forward`` (:627-694): resnet, then (transformer, motion module, resnet)."""
    x = resnet_oracle(sd, f"{prefix}.resnets.0", x, temb, cfg["norm_num_groups"], cfg["norm_eps"])
    for i in range(_count(sd, f"{prefix}.attentions", "norm.weight")):
        x = transformer2d_oracle(sd, f"{prefix}.attentions.{i}", x, ctx, heads, cfg["norm_num_groups"], enable,
                                 num_frames, ip_tokens, ip_scale)
        x = temporal_model_oracle(sd, f"{prefix}.motion_modules.{i}", x, num_frames,
                                  cfg["motion_num_attention_heads"], cfg["norm_num_groups"])
        x = resnet_oracle(sd, f"{prefix}.resnets.{i + 1}", x, temb, cfg["norm_num_groups"], cfg["norm_eps"])
    return x


def timestep_embedding_oracle(timesteps: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers ``Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0)``."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    e = timesteps[:, None].float() * freqs[None]
    return torch.cat([torch.cos(e), torch.sin(e)], dim=-1)


def unet_oracle(sd: StateDict, cfg: Dict, sample, timestep, enable_cross_frame_attn: bool, encoder_hidden_states,
                image_embeds: Optional[torch.Tensor] = None, ip_scale: float = 1.0):
    """``UNetMotionCrossFrameAttnModel.forward`` (src/models/unet_motion_cross_frame_attn.py:1289-1451).
    sample (B,F,4,h,w); :1333 num_frames; :1336-1344 time embedding repeated per frame; :1346-1353 IP image tokens
    appended to the text tokens; :1355 context repeated per frame; :1358 frames folded into the batch
    (row = video*F + frame); :1363-1436 down / mid / up; :1439-1446 output head and un-fold."""
    b, num_frames = sample.shape[0], sample.shape[1]
    ch = cfg["block_out_channels"]
    heads = cfg["num_attention_heads"]
    t = torch.as_tensor(timestep).reshape(-1).expand(b)
    emb = timestep_embedding_oracle(t, ch[0]).to(sample.dtype)
    emb = _linear(F.silu(_linear(emb, sd, "time_embedding.linear_1")), sd, "time_embedding.linear_2")
    emb = emb.repeat_interleave(num_frames, dim=0)

    ip_tokens = 0
    ctx = encoder_hidden_states
    if image_embeds is not None:
        tok = _linear(image_embeds, sd, "encoder_hid_proj.image_embeds").reshape(b, 4, -1)
        tok = F.layer_norm(tok, (tok.shape[-1],), sd["encoder_hid_proj.norm.weight"], sd["encoder_hid_proj.norm.bias"])
        ctx = torch.cat([ctx, tok.to(ctx.dtype)], dim=1)
        ip_tokens = 4
    ctx = ctx.repeat_interleave(num_frames, dim=0)

    x = sample.reshape((b * num_frames, -1) + tuple(sample.shape[3:]))
    x = F.conv2d(x, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    skips = (x,)
    for i in range(len(ch)):
        x, outs = down_block_oracle(sd, f"down_blocks.{i}", x, emb, ctx, cfg, heads, enable_cross_frame_attn,
                                    num_frames, ip_tokens, ip_scale)
        skips += outs
    x = mid_block_oracle(sd, "mid_block", x, emb, ctx, cfg, heads, enable_cross_frame_attn, num_frames, ip_tokens,
                         ip_scale)
    for i in range(len(ch)):
        n = _count(sd, f"up_blocks.{i}.resnets", "norm1.weight")
        res, skips = skips[-n:], skips[:-n]
        x = up_block_oracle(sd, f"up_blocks.{i}", x, res, emb, ctx, cfg, heads, enable_cross_frame_attn, num_frames,
                            ip_tokens, ip_scale)
    x = F.group_norm(x, cfg["norm_num_groups"], sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], cfg["norm_eps"])
    x = F.conv2d(F.silu(x), sd["conv_out.weight"], sd["conv_out.bias"], padding=1)
    return x.reshape((b, num_frames) + tuple(x.shape[1:]))


# --------------------------------------------------------------------------------------------------------------
# scheduler + denoise step
# --------------------------------------------------------------------------------------------------------------
def ddim_alphas_cumprod_oracle(num_train: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012):
    """SD1.5 ``scaled_linear`` betas."""
    betas = torch.linspace(beta_start**0.5, beta_end**0.5, num_train, dtype=torch.float64) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def ddim_timesteps_oracle(num_inference_steps: int, num_train: int = 1000):
    """``timestep_spacing="linspace"``: round(linspace(0, T-1, N))[::-1] (pipeline :755-757)."""
    return torch.linspace(0, num_train - 1, num_inference_steps, dtype=torch.float64).round().flip(0).long()


def ddim_step_oracle(noise_pred, t: int, sample, num_inference_steps: int, alphas_cumprod=None):
    """DDIM update with eta = 0, epsilon prediction, clip_sample=False, set_alpha_to_one=False."""
    ac = ddim_alphas_cumprod_oracle() if alphas_cumprod is None else alphas_cumprod
    prev_t = int(t) - 1000 // num_inference_steps
    a_t = float(ac[int(t)])
    a_prev = float(ac[prev_t]) if prev_t >= 0 else float(ac[0])
    x0 = (sample - (1 - a_t) ** 0.5 * noise_pred) / a_t**0.5
    return a_prev**0.5 * x0 + (1 - a_prev) ** 0.5 * noise_pred


def add_noise_oracle(x0, noise, t: int, alphas_cumprod=None):
    """``scheduler.add_noise``; with zero noise on frame 0 it reduces to sqrt(alpha_cumprod_t) * x0, the identity the
    reference asserts in test/test_first_frame_pertubation.py:39."""
    ac = ddim_alphas_cumprod_oracle() if alphas_cumprod is None else alphas_cumprod
    a = float(ac[int(t)])
    return a**0.5 * x0 + (1 - a) ** 0.5 * noise


def denoise_step_oracle(sd: StateDict, cfg: Dict, latents, t: int, prompt_embeds, num_inference_steps: int,
                        guidance_scale: float = 7.5, condition_image_latents=None, image_embeds=None):
    """One iteration of the pipeline loop (src/pipelines/pipeline_i2v_adapter.py:666-691): :668-669 re-impose the
    condition latent on frame 0; :672 CFG duplication; :676-683 UNet; :686-688 guidance; :691 scheduler step."""
    latents = latents.clone()
    if condition_image_latents is not None:
        latents[:, 0] = condition_image_latents
    do_cfg = guidance_scale > 1.0
    x = torch.cat([latents] * 2) if do_cfg else latents
    eps = unet_oracle(sd, cfg, x, t, condition_image_latents is not None, prompt_embeds, image_embeds)
    if do_cfg:
        un, tx = eps.chunk(2)
        eps = un + guidance_scale * (tx - un)
    return ddim_step_oracle(eps, t, latents, num_inference_steps)
