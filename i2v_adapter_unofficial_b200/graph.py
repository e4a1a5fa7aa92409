"""CUDA-graph capture of one pipeline denoise iteration.

One iteration of the reference's sampling loop (``/root/reference/src/pipelines/pipeline_i2v_adapter.py:666-691``:
first-frame re-imposition, CFG duplication, UNet forward, guidance, DDIM update) is ~1700 kernel launches whose CPU
issue time (~60 ms) is close to their GPU time; replaying a captured graph removes the host from the loop.  Everything
that varies between iterations lives in device buffers that are refreshed before each replay:

* the latents (updated in place by the captured DDIM step),
* the timestep (a 0-dim int64 tensor handed to the UNet) and the four DDIM coefficients
  ``sqrt(a_t), sqrt(1 - a_t), sqrt(a_prev), sqrt(1 - a_prev)`` (a 4-element fp32 tensor), copied from a pinned host
  table indexed by the iteration number.

The captured arithmetic is that of ``hostmodel.pipeline.denoise_step`` with the scheduler scalars read from that
tensor instead of Python floats; the DDIM update itself is evaluated in fp32 and rounded once (the eager step rounds
every intermediate to the latents' dtype), so the two agree to one bf16 ulp per step, not bit for bit.  Prompt / image embeddings and the condition latents are static buffers that
``load_inputs`` overwrites (host or device sources).
"""
from __future__ import annotations

from typing import Optional

import torch


class GraphedDenoiser:
    def __init__(self, unet, scheduler, latents: torch.Tensor, prompt_embeds: torch.Tensor,
                 guidance_scale: float = 7.5, condition_image_latents: Optional[torch.Tensor] = None,
                 image_embeds: Optional[torch.Tensor] = None, warmup: int = 2, before_capture=None,
                 impose_first_frame: bool = True):
        if not latents.is_cuda:
            raise RuntimeError("GraphedDenoiser captures a CUDA graph: the buffers must live on a CUDA device")
        if scheduler.num_inference_steps is None:
            raise ValueError("call scheduler.set_timesteps(...) first")
        self.unet, self.scheduler, self.guidance_scale = unet, scheduler, float(guidance_scale)
        # frame shards (partition.FramePartitioner): only the rank that owns global frame 0 re-imposes the condition
        # latents; the collectives of the sharded forward (NCCL broadcast / all-to-all / all-reduce) are captured
        # into the graph like any other kernel, every rank capturing the same sequence
        self.impose_first_frame = bool(impose_first_frame)
        dev = latents.device
        self.latents = latents.clone()
        self.prompt = prompt_embeds.clone()
        self.cond = None if condition_image_latents is None else condition_image_latents.clone()
        self.image = None if image_embeds is None else image_embeds.clone()
        # per-iteration scalars: pinned host table -> device
        ts = [int(t) for t in scheduler.timesteps]
        ac = scheduler.alphas_cumprod.double()
        rows = []
        for t in ts:
            prev_t = t - scheduler.num_train_timesteps // scheduler.num_inference_steps
            a_t = float(ac[t])
            a_prev = float(ac[prev_t]) if prev_t >= 0 else float(scheduler.final_alpha_cumprod)
            rows.append([a_t ** 0.5, (1 - a_t) ** 0.5, a_prev ** 0.5, (1 - a_prev) ** 0.5])
        self._coef_host = torch.tensor(rows, dtype=torch.float32).pin_memory()
        self._t_host = torch.tensor(ts, dtype=torch.int64).pin_memory()
        self.coef = torch.zeros(4, dtype=torch.float32, device=dev)
        self.t = torch.zeros((), dtype=torch.int64, device=dev)
        self.num_iterations = len(ts)
        self.launches_per_step = 0
        self._set_iteration(0)
        # warm-up on a side stream (lazy initialisation, packed-weight caches, cuBLAS workspaces), then capture
        keep = self.latents.clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.latents.copy_(keep)
        from . import _lib

        if before_capture is not None:   # e.g. arm the library's launch timing so the event pairs land in the graph
            before_capture()
        n0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self._body()
        self.launches_per_step = _lib.launch_count() - n0   # library kernels recorded into the graph
        self.latents.copy_(keep)

    def _set_iteration(self, i: int) -> None:
        i %= self.num_iterations
        self.coef.copy_(self._coef_host[i], non_blocking=True)
        self.t.copy_(self._t_host[i], non_blocking=True)

    def _body(self) -> None:
        lat = self.latents
        has_condition = self.cond is not None
        do_cfg = self.guidance_scale > 1.0
        if has_condition and self.impose_first_frame:
            lat[:, 0] = self.cond                                                            # pipeline :668-669
        inp = torch.cat([lat] * 2) if do_cfg else lat                                        # :672
        added = {"image_embeds": self.image} if self.image is not None else None
        noise = self.unet(inp, self.t, enable_cross_frame_attn=has_condition, encoder_hidden_states=self.prompt,
                          added_cond_kwargs=added).sample                                    # :676-683
        if do_cfg:
            uncond, text = noise.chunk(2)
            noise = uncond + self.guidance_scale * (text - uncond)                           # :686-688
        # DDIM, eta = 0 (:691).  The coefficients stay fp32 and so does the update: rounded to bf16 first,
        # sqrt(alpha) = 0.99912 would become exactly 1.0 and bias every one of the 25 steps.  One rounding, on the copy.
        c = self.coef
        noise32 = noise.float()
        x0 = (lat.float() - c[1] * noise32) / c[0]
        lat.copy_(c[2] * x0 + c[3] * noise32)

    def load_inputs(self, latents=None, prompt_embeds=None, condition_image_latents=None, image_embeds=None) -> None:
        """Overwrite the static buffers (sources may be pinned host tensors: these are the step's H2D copies)."""
        for dst, src in ((self.latents, latents), (self.prompt, prompt_embeds), (self.cond, condition_image_latents),
                         (self.image, image_embeds)):
            if src is not None:
                if dst is None:
                    raise ValueError("this input was not part of the captured step")
                dst.copy_(src, non_blocking=True)

    def step(self, iteration: int) -> torch.Tensor:
        """Run iteration ``iteration`` of the schedule on the current latents; returns the (static) latents buffer."""
        self._set_iteration(iteration)
        self.graph.replay()
        return self.latents
