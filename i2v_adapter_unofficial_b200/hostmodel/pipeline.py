"""Latent-space denoising loop of ``I2VAdapterPipeline.__call__``
(``/root/reference/src/pipelines/pipeline_i2v_adapter.py:666-700``) without the VAE / CLIP stages, which are out of
the hot path: the caller passes ``prompt_embeds``, ``image_embeds`` and ``condition_image_latents`` directly.

``denoise_step`` is one "UNet denoise step" of BASELINE.json's metric: first-frame re-imposition, CFG duplication,
the UNet forward, guidance and the scheduler update.
"""
from __future__ import annotations

from typing import Optional

import torch


@torch.no_grad()
def denoise_step(unet, scheduler, latents, t, prompt_embeds, guidance_scale: float = 7.5,
                 condition_image_latents: Optional[torch.Tensor] = None, image_embeds: Optional[torch.Tensor] = None,
                 cross_attention_kwargs=None, impose_first_frame: bool = True):
    """latents (B, F, 4, h, w); prompt_embeds (2B, 77, D) = [negative, positive] when guidance_scale > 1.
    ``impose_first_frame=False`` is for frame shards that do not hold frame 0 (``partition.sharded_denoise_step``)."""
    has_condition = condition_image_latents is not None
    do_cfg = guidance_scale > 1.0
    if has_condition and impose_first_frame:
        latents[:, 0] = condition_image_latents
    latent_model_input = torch.cat([latents] * 2) if do_cfg else latents
    latent_model_input = scheduler.scale_model_input(latent_model_input, t)
    added = {"image_embeds": image_embeds} if image_embeds is not None else None
    noise_pred = unet(latent_model_input, t, enable_cross_frame_attn=has_condition,
                      encoder_hidden_states=prompt_embeds, cross_attention_kwargs=cross_attention_kwargs,
                      added_cond_kwargs=added).sample
    if do_cfg:
        uncond, text = noise_pred.chunk(2)
        noise_pred = uncond + guidance_scale * (text - uncond)
    return scheduler.step(noise_pred, t, latents, eta=0.0).prev_sample


@torch.no_grad()
def denoise(unet, scheduler, latents, prompt_embeds, num_inference_steps: int = 25, guidance_scale: float = 7.5,
            condition_image_latents: Optional[torch.Tensor] = None, image_embeds: Optional[torch.Tensor] = None,
            frame_similarity_sample_ratio: float = 1.0):
    scheduler.set_timesteps(num_inference_steps, device=latents.device)
    init = min(int(num_inference_steps * frame_similarity_sample_ratio), num_inference_steps)
    timesteps = scheduler.timesteps[max(num_inference_steps - init, 0):]
    for t in timesteps:
        latents = denoise_step(unet, scheduler, latents, t, prompt_embeds, guidance_scale,
                               condition_image_latents, image_embeds)
    if condition_image_latents is not None:
        latents[:, 0] = condition_image_latents
    return latents
