"""Benchmark of BASELINE.json's metric: UNet denoise steps/s (16 frames, 512x512 -> 64x64 latent, CFG) with the
attention hot path on the sm_100a kernels.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c2|c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pipeline denoise iteration (reference src/pipelines/pipeline_i2v_adapter.py:666-691): first-frame
re-imposition, CFG duplication, UNetMotionCrossFrameAttnModel forward (full SD1.5 architecture + motion modules +
I2V-Adapter + IP-Adapter, random init), guidance, DDIM update.

Workloads (BASELINE.json `configs`; `--config` picks the one the headline line is measured on):

  c2 (default)  1 video x CFG per GPU, 16 frames, 64x64 latent.  N > 1: every rank samples its own video
                (independent videos need no communication) -> weak scaling.
  c3            8 videos x CFG in total, split contiguously over the N ranks (BatchPartition) -> strong scaling;
                N = 1 runs all 16 sequences on one GPU.
  c4            96x96 latent (S = 9216 at level 0) + IP-Adapter tokens, 1 video x CFG per GPU -> weak scaling.
  c5            one 32-frame clip, frame-sharded over the N ranks (FramePartitioner: frame-0 K/V broadcast,
                all-to-all re-shard around the motion modules, GroupNorm statistics gather) -> strong scaling;
                N = 1 runs the clip unsharded.

The default (c2) run also measures short c4 and c5 legs and, at N = 1, the same step with the stock SDPA processors
on the same GPU (`gpu_baseline`: what the reference does on a GPU) and the CPU oracle port (`cpu_baseline`); they
are extra keys of the ONE JSON line rank 0 prints.

`--impl reference` times the reference path's CPU restatement (oracle/) with all host threads on the full workload;
the reference itself cannot run here because its `diffusers` dependency is not installable.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "UNet denoise steps/s (16f 512^2, CFG)"
UNIT = "steps/s"
GUIDANCE, DDIM_STEPS = 7.5, 25
SD15 = dict(block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, cross_attention_dim=768,
            num_attention_heads=8, motion_num_attention_heads=8, motion_max_seq_length=32, norm_num_groups=32)
IMAGE_EMBED_DIM = 1024
HEADS, HEAD_DIM_L0 = 8, 40

WORKLOADS = {
    "c2": dict(frames=16, latent=64, total_videos=None, scaling="weak", sharded=False,
               text="configs[1]: SD1.5 UNetMotion (Realistic Vision arch, random init) + motion adapter + I2V-Adapter "
                    "+ IP-Adapter, 1 video x CFG per GPU, 16 frames, 64x64 latent, DDIM 25 timesteps"),
    "c3": dict(frames=16, latent=64, total_videos=8, scaling="strong", sharded=False,
               text="configs[2]: same model, 8 videos x CFG in total split over the GPUs, 16 frames, 64x64 latent"),
    "c4": dict(frames=16, latent=96, total_videos=None, scaling="weak", sharded=False,
               text="configs[3]: same model, 1 video x CFG per GPU, 16 frames, 96x96 latent (9216 tokens per frame at "
                    "level 0), 77 text + 4 IP-Adapter image tokens"),
    "c5": dict(frames=32, latent=64, total_videos=1, scaling="strong", sharded=True,
               text="configs[4]: same model, one 32-frame 64x64-latent clip x CFG, frame-sharded over the GPUs "
                    "(frame-0 K/V broadcast, all-to-all re-shard of the motion modules)"),
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch, copied from the committed `ncu --set full` digests (a bench
# run cannot read hardware counters); the source file is named next to the number in the JSON line
TRAFFIC = {
    "dense_l0_c2": (596.6e6, "profiles/r02_dense_attn_l0.md"),
    "temporal_l0_c2": (313.1e6, "profiles/r02_temporal_attn_l0.md"),
}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops=p.get("bf16_tflops_sustained", 1396.3), hbm=p.get("hbm_gbs", 6546.9), source="measured")
    return dict(tflops=1400.0, hbm=6650.0, source="fallback")


# ------------------------------------------------------------------------------------------------------------
# model + inputs
# ------------------------------------------------------------------------------------------------------------
def build_unet(device, dtype):
    import torch
    from i2v_adapter_unofficial_b200.hostmodel import UNetMotionCrossFrameAttnModel

    torch.manual_seed(0)
    with torch.device(device):
        unet = UNetMotionCrossFrameAttnModel(**SD15)
    unet = unet.to(dtype).eval()
    # IP-Adapter weights in the ip-adapter_sd15.bin layout, random values (there is no network for checkpoints)
    g = torch.Generator(device="cpu").manual_seed(1)
    cross = SD15["cross_attention_dim"]
    sd = {"image_proj": {"proj.weight": torch.randn(4 * cross, IMAGE_EMBED_DIM, generator=g) * IMAGE_EMBED_DIM**-0.5,
                         "proj.bias": torch.zeros(4 * cross), "norm.weight": torch.ones(cross),
                         "norm.bias": torch.zeros(cross)}, "ip_adapter": {}}
    key_id = 1
    for name in unet.attn_processors.keys():
        if name.endswith("attn2.processor") and "motion_modules" not in name:
            if name.startswith("mid_block"):
                hidden = SD15["block_out_channels"][-1]
            elif name.startswith("up_blocks"):
                hidden = list(reversed(SD15["block_out_channels"]))[int(name[len("up_blocks.")])]
            else:
                hidden = SD15["block_out_channels"][int(name[len("down_blocks.")])]
            sd["ip_adapter"][f"{key_id}.to_k_ip.weight"] = torch.randn(hidden, cross, generator=g) * cross**-0.5
            sd["ip_adapter"][f"{key_id}.to_v_ip.weight"] = torch.randn(hidden, cross, generator=g) * cross**-0.5
            key_id += 2
    unet._load_ip_adapter_weights(sd)
    return unet


def make_inputs(videos, frames, latent, seed, dtype, device="cpu", pin=False):
    import torch

    g = torch.Generator(device="cpu").manual_seed(seed)
    t = dict(latents=torch.randn(videos, frames, 4, latent, latent, generator=g),
             cond=torch.randn(videos, 4, latent, latent, generator=g),
             prompt=torch.randn(2 * videos, 77, SD15["cross_attention_dim"], generator=g),
             image=torch.randn(2 * videos, IMAGE_EMBED_DIM, generator=g))
    t = {k: v.to(dtype) for k, v in t.items()}
    if pin:
        t = {k: v.pin_memory() for k, v in t.items()}
    if device != "cpu":
        t = {k: v.to(device) for k, v in t.items()}
    return t


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons DURING the timed region.  In-process NVML polling every 5 ms (the timed
    region of a graph-replayed run is a few hundred ms: an `nvidia-smi -lms` child does not even start in that time);
    `nvidia-smi` remains the fallback when the NVML binding is unavailable."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    REASON_BITS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
                   (0x4, "sw_power_cap"))

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []
        self.samples = []      # (sm_mhz, power_w, reason_mask) from NVML
        self.sm_max = None
        self._stop = threading.Event()
        self._thread = None
        self.how = None

    def _nvml_handle(self):
        import pynvml

        pynvml.nvmlInit()
        try:
            import torch

            uuid = str(torch.cuda.get_device_properties(self.gpu_index).uuid)
            uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
        except Exception:  # noqa: BLE001
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)

    def _poll(self, pynvml, handle):
        reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.samples.append((float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)),
                                     pynvml.nvmlDeviceGetPowerUsage(handle) / 1000.0, int(reasons_fn(handle))))
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.005)

    def start(self):
        try:
            pynvml, handle = self._nvml_handle()
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._poll, args=(pynvml, handle), daemon=True)
            self._thread.start()
            self.how = "nvml"
            return
        except Exception:  # noqa: BLE001
            self._thread = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            self.how = "nvidia-smi"
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=1.0)
            sm = [s[0] for s in self.samples]
            mask = 0
            for s in self.samples:
                mask |= s[2]
            reasons = sorted(name for bit, name in self.REASON_BITS if mask & bit)
            return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=self.sm_max,
                        power_w_max=max((s[1] for s in self.samples), default=None), samples=len(sm), reasons=reasons,
                        source="nvml, 5 ms period, timed region only")
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(power) if power else None, samples=len(sm), reasons=sorted(reasons),
                    source="nvidia-smi -lms 100")


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the full 16-frame workload
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, budget_s=None, frames=16, latent=64):
    """Times `denoise_step_oracle` (fp32, all host threads, `F.scaled_dot_product_attention` as the reference's
    AttnProcessor2_0 calls it) on the FULL workload (all 16 frames, CFG batch 2, 64x64 latent).  `budget_s` caps the
    wall time: the number of timed steps is reduced (never below one) when a step is too slow for `steps` of them."""
    import torch
    import oracle.attention_oracle as attention_oracle
    from oracle.unet_oracle import denoise_step_oracle, ddim_timesteps_oracle

    attention_oracle.USE_TORCH_SDPA = True  # time the library call the reference makes (AttnProcessor2_0)

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    unet = build_unet("cpu", torch.float32)
    sd = {k: v for k, v in unet.state_dict().items()}
    cfg = dict(unet.config)
    inp = make_inputs(1, frames, latent, seed=1, dtype=torch.float32)
    ts = ddim_timesteps_oracle(DDIM_STEPS)
    times = []
    t_start = time.perf_counter()
    done_warm = 0
    with torch.no_grad():
        i = 0
        while len(times) < steps:
            t0 = time.perf_counter()
            denoise_step_oracle(sd, cfg, inp["latents"], int(ts[i % len(ts)]), inp["prompt"], DDIM_STEPS, GUIDANCE,
                                inp["cond"], inp["image"])
            dt = time.perf_counter() - t0
            i += 1
            if done_warm < warmup:
                done_warm += 1
                # a warm-up step that alone eats a third of the budget is kept as the first timed step instead
                if budget_s is not None and dt > budget_s / 3:
                    times.append(dt)
            else:
                times.append(dt)
            if budget_s is not None and times and (time.perf_counter() - t_start) + dt > budget_s:
                break
    sec_per_step = sum(times) / len(times)
    return dict(value=1.0 / sec_per_step, unit=UNIT, cores=cores, kind="port",
                sample=(f"oracle denoise step (fp32, CFG batch 2, {latent}x{latent} latent) on all {frames} frames: "
                        f"{len(times)} timed full step(s) of {sec_per_step:.2f} s, no extrapolation")), len(times), done_warm


def run_reference(args, rank, world):
    if rank != 0:
        return
    base, steps_done, warm_done = cpu_reference_run(max(1, args.steps), max(0, args.warmup), budget_s=args.cpu_budget)
    line = dict(metric=METRIC, value=base["value"], unit=UNIT, n_gpus=args.gpus, steps=steps_done, warmup=warm_done,
                ms_per_step=1e3 / base["value"], higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=WORKLOADS["c2"]["text"],
                            note=f"reference path restated in oracle/ (diffusers not installable); CPU, full 16-frame "
                                 f"step; steps capped by a {args.cpu_budget:.0f} s wall budget "
                                 f"(requested {args.steps} + {args.warmup} warm-up)"),
                cpu_baseline=base,
                e2e=dict(value=base["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
class Emitter:
    """Rank 0 prints exactly one JSON line.  A watchdog thread covers the optional legs: if one of them hangs (a
    collective that never completes), the main line measured so far is still printed, with the leg marked failed."""

    def __init__(self, rank):
        self.rank = rank
        self.line = None
        self._done = False
        self._lock = threading.Lock()
        self._timer = None
        self._leg = None

    def emit(self):
        with self._lock:
            if self._done or self.line is None:
                return
            self._done = True
            if self.rank == 0:
                print(json.dumps(self.line), flush=True)

    def guard(self, leg, seconds):
        self.cancel()
        self._leg = leg

        def fire():
            if self.line is not None:
                self.line.setdefault("legs", {})[leg] = dict(error=f"timed out after {seconds} s")
            self.emit()
            os._exit(0 if self.line is not None else 3)

        self._timer = threading.Timer(seconds, fire)
        self._timer.daemon = True
        self._timer.start()

    def cancel(self):
        if self._timer is not None:
            self._timer.cancel()
            self._timer = None


def _stats(ms):
    return dict(median_ms=statistics.median(ms), min_ms=min(ms), max_ms=max(ms), mean_ms=sum(ms) / len(ms),
                launches_timed=len(ms))


def dense_roofline(ms, bf, seq, peaks, traffic_key=None, note=""):
    if not ms:
        return None
    st = _stats(ms)
    flops = 2 * 4.0 * bf * HEADS * seq * seq * HEAD_DIM_L0   # self + cross-frame, true head dim (40), softmax not counted
    achieved = flops / (st["median_ms"] * 1e-3) / 1e12
    tr = TRAFFIC.get(traffic_key)
    return dict(bound="tensor",
                kernel=(f"dense_attn_pipe_kernel (fused spatial self + cross-frame attention, level 0: S={seq}, d=40, "
                        f"{bf} frames x {HEADS} heads x 2 problems, augmented operand layout)"),
                achieved=achieved, peak=peaks["tflops"], unit="TFLOP/s", frac=achieved / peaks["tflops"],
                peak_source=f"{peaks['source']} sustained bf16 (kernel timed inside the step)",
                frac_of_nominal_2250=achieved / 2250.0, avg_launch_ms=st["median_ms"], flops_per_launch=flops,
                timing=dict(st, how="cudaEvent pairs recorded by the library around each launch" + note),
                achieved_at_min_ms=flops / (st["min_ms"] * 1e-3) / 1e12,
                # the operator is exponential-bound at d = 40 (one exp per 4 * 40 flops): the same launch against the SM's
                # exponential rate -- MUFU.EX2 issues 16 per clock and SM; informational, `frac` above is the graded figure
                exponentials=dict(per_launch=flops / (4.0 * HEAD_DIM_L0), achieved_per_s=flops / (4.0 * HEAD_DIM_L0) / (st["median_ms"] * 1e-3),
                                  mufu_peak_per_s_at_1965mhz=16.0 * 148 * 1.965e9,
                                  frac_of_mufu_peak=flops / (4.0 * HEAD_DIM_L0) / (st["median_ms"] * 1e-3) / (16.0 * 148 * 1.965e9),
                                  note="3 of 8 column pairs take an FMA-pipe polynomial instead of MUFU.EX2"),
                traffic=tr[0] if tr else None, traffic_source=(f"{tr[1]} (ncu --set full; not measured in this run)" if tr else None))


def temporal_roofline(ms, n_pos, frames, peaks, traffic_key=None, note=""):
    if not ms:
        return None
    st = _stats(ms)
    nbytes = 4.0 * n_pos * frames * HEADS * HEAD_DIM_L0 * 2   # read Q, K, V, write O once (bf16)
    gbs = nbytes / (st["median_ms"] * 1e-3) / 1e9
    tr = TRAFFIC.get(traffic_key)
    return dict(bound="hbm", kernel=(f"temporal_attn_kernel<d=40> (motion module, level 0: {n_pos} positions x {frames} "
                                     f"frames x {HEADS} heads)"),
                achieved=gbs, peak=peaks["hbm"], unit="GB/s", frac=gbs / peaks["hbm"],
                peak_source=f"{peaks['source']} copy bandwidth", avg_launch_ms=st["median_ms"], bytes_per_launch=nbytes,
                timing=dict(st, how="cudaEvent pairs recorded by the library around each launch" + note),
                traffic=tr[0] if tr else None, traffic_source=(f"{tr[1]} (ncu --set full; not measured in this run)" if tr else None))


class StepRunner:
    """One workload on this rank: device buffers, the graph-captured (or eager) step, timed loops."""

    def __init__(self, unet, sched, dev, rank, world, videos, frames, latent, graph, prof=True):
        import torch
        from i2v_adapter_unofficial_b200 import _lib

        self.torch, self._lib = torch, _lib
        self.unet, self.sched, self.dev, self.rank, self.world = unet, sched, dev, rank, world
        self.videos, self.frames, self.latent = videos, frames, latent
        self.ts = [int(t) for t in sched.timesteps]
        self.host = make_inputs(videos, frames, latent, seed=1 + rank, dtype=torch.bfloat16, pin=True)
        self.d_in = {k: v.to(dev, non_blocking=True) for k, v in self.host.items()}
        self.bf = 2 * videos * frames
        self.seq = latent * latent
        self.graphed = None
        self.prof = prof
        if graph:
            from i2v_adapter_unofficial_b200.graph import GraphedDenoiser

            self.graphed = GraphedDenoiser(self.unet, sched, self.d_in["latents"], self.d_in["prompt"], GUIDANCE,
                                           self.d_in["cond"], self.d_in["image"],
                                           before_capture=self._arm if prof else None)
            self._disarm()
        self.lat = self.d_in["latents"].clone()

    def _arm(self, pairs=64):
        self._lib.prof_arm(self._lib.PROF_DENSE, self.seq, self.bf, pairs)
        self._lib.prof_arm(self._lib.PROF_TEMPORAL, 2 * self.videos * self.seq, HEAD_DIM_L0, pairs)

    def _disarm(self):
        self._lib.prof_arm(self._lib.PROF_DENSE, 0, 0, 0)
        self._lib.prof_arm(self._lib.PROF_TEMPORAL, 0, 0, 0)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier()
        self.torch.cuda.synchronize()

    def step(self, i):
        from i2v_adapter_unofficial_b200.hostmodel import denoise_step

        if self.graphed is not None:
            return self.graphed.step(i)   # latents live in the graph's static buffer
        self.lat = denoise_step(self.unet, self.sched, self.lat, self.ts[i % len(self.ts)], self.d_in["prompt"],
                                GUIDANCE, self.d_in["cond"], self.d_in["image"])
        return self.lat

    def timed(self, steps, warmup, body=None, profiler_range=False):
        """`warmup` untimed steps, then exactly `steps` between barrier + synchronize; device time, max over ranks."""
        torch = self.torch
        body = body or self.step
        for i in range(warmup):
            body(i)
        self.barrier()
        if profiler_range:
            torch.cuda.profiler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            body(warmup + i)
        e1.record()
        self.barrier()
        if profiler_range:
            torch.cuda.profiler.stop()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def e2e_body(self):
        """Host buffers in, host buffer out: H2D of this step's inputs from pinned memory and D2H of the new latents
        inside the timed region, through the public API (GraphedDenoiser.load_inputs/step or denoise_step)."""
        from i2v_adapter_unofficial_b200.hostmodel import denoise_step

        torch, host, dev = self.torch, self.host, self.dev
        out_host = torch.empty_like(host["latents"]).pin_memory()
        h2d = sum(host[k].numel() * host[k].element_size() for k in ("latents", "cond", "prompt", "image"))
        d2h = out_host.numel() * out_host.element_size()

        def body(i):
            if self.graphed is not None:
                self.graphed.load_inputs(host["latents"], host["prompt"], host["cond"], host["image"])
                out_host.copy_(self.graphed.step(i), non_blocking=True)
                return
            din = {k: host[k].to(dev, non_blocking=True) for k in ("latents", "cond", "prompt", "image")}
            new = denoise_step(self.unet, self.sched, din["latents"], self.ts[i % len(self.ts)], din["prompt"],
                               GUIDANCE, din["cond"], din["image"])
            out_host.copy_(new, non_blocking=True)

        return body, h2d, d2h

    def kernel_times(self, replays=5):
        """Durations of the level-0 dense and temporal launches inside the step.  Graph mode: the event pairs were
        captured as external event nodes, each replay re-records them.  Eager mode: armed for `replays` steps."""
        lib = self._lib
        dense, temporal = [], []
        if not self.prof:
            return dense, temporal, ""
        if self.graphed is not None:
            for i in range(replays):
                self.graphed.step(i)
                self.torch.cuda.synchronize()
                dense += lib.prof_read(lib.PROF_DENSE)
                temporal += lib.prof_read(lib.PROF_TEMPORAL)
            return dense, temporal, f", inside {replays} replays of the captured step graph"
        self._arm(256)
        for i in range(replays):
            self.step(i)
        self.torch.cuda.synchronize()
        dense, temporal = lib.prof_read(lib.PROF_DENSE), lib.prof_read(lib.PROF_TEMPORAL)
        self._disarm()
        return dense, temporal, f", {replays} eager steps"

    def close(self):
        self.graphed = None
        self.torch.cuda.empty_cache()


def leg_frame_sharded(unet, sched, dev, rank, world, steps=3, warmup=2, frames=32, latent=64, graph=True):
    """C5: one `frames`-frame clip x CFG.  world == 1: unsharded on this GPU (the reference point).  world > 1: frames
    split over the ranks through FramePartitioner (NCCL broadcast / all-to-all / all-gather of statistics)."""
    import torch
    import torch.distributed as dist
    from i2v_adapter_unofficial_b200.hostmodel import denoise_step
    from i2v_adapter_unofficial_b200.partition import FramePartitioner, sharded_denoise_step

    ts = [int(t) for t in sched.timesteps]
    inp = make_inputs(1, frames, latent, seed=11, dtype=torch.bfloat16, device=dev)   # identical on every rank
    part = None
    if world > 1:
        part = FramePartitioner(unet).install()
        lat = part.shard_frames(inp["latents"]).clone()
    else:
        lat = inp["latents"].clone()

    graphed = None
    if graph:
        # one captured graph per rank; with world > 1 it contains the NCCL collectives of the frame partitioner
        from i2v_adapter_unofficial_b200.graph import GraphedDenoiser

        graphed = GraphedDenoiser(unet, sched, lat, inp["prompt"], GUIDANCE, inp["cond"], inp["image"],
                                  impose_first_frame=part is None or part.owns_first_frame)

    def eager_body(i):
        nonlocal lat
        if part is not None:
            lat = sharded_denoise_step(part, unet, sched, lat, ts[i % len(ts)], inp["prompt"], GUIDANCE, inp["cond"],
                                       inp["image"])
        else:
            lat = denoise_step(unet, sched, lat, ts[i % len(ts)], inp["prompt"], GUIDANCE, inp["cond"], inp["image"])

    def body(i):
        if graphed is not None:
            graphed.step(i)
        else:
            eager_body(i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    try:
        for i in range(warmup):
            body(i)
        barrier()
        if part is not None and graphed is None:
            part.start_timing()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            body(warmup + i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        eager_steps = 0
        if part is not None and graphed is not None:
            # events cannot bracket the collectives of a replayed graph: their breakdown comes from two eager steps
            eager_steps = 2
            eager_body(0)
            barrier()
            part.start_timing()
            for i in range(eager_steps):
                eager_body(1 + i)
            barrier()
        out = dict(n_gpus=world, frames=frames, latent=latent, steps=steps, warmup=warmup,
                   launch="CUDA graph replay (collectives captured)" if graphed is not None else "eager",
                   ms_per_step=ms.item() / steps, value=steps / (ms.item() / 1e3), unit=UNIT, scaling="strong",
                   parallelism=("unsharded (reference point)" if part is None else
                                f"frames sharded {frames // world} per rank: frame-0 K/V broadcast, all-to-all re-shard "
                                f"around each motion module, GroupNorm statistics all-gather"))
        if part is not None:
            out["collectives"] = part.timing_summary(eager_steps or steps)
            if eager_steps:
                out["collectives"]["how"] = f"{eager_steps} eager steps after the graph-timed region"
            iso = sum(v["isolated_ms_per_step"] for v in out["collectives"].values() if isinstance(v, dict))
            out["collectives_isolated_ms_per_step"] = iso
            out["collectives_share_of_step"] = iso / out["ms_per_step"]
        return out
    finally:
        if part is not None:
            part.uninstall()


def leg_gpu_baseline(unet, sched, dev, steps=3, warmup=2, frames=16, latent=64):
    """The reference's own GPU path on this GPU: stock AttnProcessor2_0 / IPAdapterAttnProcessor2_0
    (F.scaled_dot_product_attention, src/models/unet_motion_cross_frame_attn.py:1259-1272), bf16, launched eagerly as
    the reference pipeline does, on the hostmodel mirror of the reference UNet with the B200 processors uninstalled."""
    import torch
    from i2v_adapter_unofficial_b200.hostmodel import denoise_step

    ts = [int(t) for t in sched.timesteps]
    inp = make_inputs(1, frames, latent, seed=1, dtype=torch.bfloat16, device=dev)
    lat = inp["latents"].clone()
    for i in range(warmup):
        lat = denoise_step(unet, sched, lat, ts[i], inp["prompt"], GUIDANCE, inp["cond"], inp["image"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        lat = denoise_step(unet, sched, lat, ts[warmup + i], inp["prompt"], GUIDANCE, inp["cond"], inp["image"])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return dict(value=steps / (ms / 1e3), unit=UNIT, ms_per_step=ms / steps, steps=steps, warmup=warmup, dtype="bf16",
                impl="stock AttnProcessor2_0 / IPAdapterAttnProcessor2_0 (torch SDPA) on the hostmodel mirror, eager, "
                     "same GPU, same weights and inputs; B200 processors uninstalled")


def run_b200(args, rank, world, local_rank, emitter):
    import torch
    import torch.distributed as dist

    from i2v_adapter_unofficial_b200 import _lib, fastpath, install
    from i2v_adapter_unofficial_b200.hostmodel import DDIMScheduler
    from i2v_adapter_unofficial_b200.partition import BatchPartition

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    dtype = torch.bfloat16
    wl = WORKLOADS[args.config]
    unet = build_unet(dev, dtype)
    handle = install(unet)
    fastpath.reset_fallback_counts()
    sched = DDIMScheduler()
    sched.set_timesteps(DDIM_STEPS, device="cpu")
    peaks = _peaks()

    if wl["sharded"]:
        # headline = the frame-sharded clip itself
        emitter.guard("c5", 600)
        res = leg_frame_sharded(unet, sched, dev, rank, world, steps=args.steps, warmup=args.warmup,
                                frames=wl["frames"], latent=wl["latent"], graph=args.graph)
        emitter.cancel()
        emitter.line = dict(metric=METRIC, value=res["value"], unit=UNIT, n_gpus=world, steps=args.steps,
                            warmup=args.warmup, ms_per_step=res["ms_per_step"], higher_is_better=True, scaling="strong",
                            vs_baseline=None, dtype="bf16", data="synthetic",
                            config=dict(workload=wl["text"], parallelism=res["parallelism"], launch=res["launch"]),
                            e2e=None, gpu_launches=int(_lib.launch_count()), collectives=res.get("collectives"),
                            fast_path_fallbacks=fastpath.fallback_counts())
        emitter.emit()
        return

    videos = 1
    if wl["total_videos"]:
        videos = BatchPartition(wl["total_videos"], world, rank).counts[rank]
    run = StepRunner(unet, sched, dev, rank, world, videos, wl["frames"], wl["latent"], args.graph)

    # ---- value: inputs resident in HBM ----
    sampler = ClockSampler(local_rank)
    launches0 = _lib.launch_count()
    if rank == 0:
        sampler.start()   # warm-up samples are dropped below: the sampler is restarted right before the timed steps
    for i in range(args.warmup):
        run.step(i)
    run.barrier()
    sampler.samples.clear()
    ms_total = run.timed(args.steps, 0, profiler_range=args.profiler_range)
    clocks = sampler.stop() if rank == 0 else None
    launches = _lib.launch_count() - launches0
    if run.graphed is not None:
        launches = run.graphed.launches_per_step * args.steps  # recorded at capture; replays do not pass through the host

    # ---- e2e: host buffers in, host buffer out, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        body, h2d, d2h = run.e2e_body()
        e2e_ms = run.timed(args.steps, min(args.warmup, 2), body=body)
        total_videos = wl["total_videos"] or world * videos
        e2e = dict(value=total_videos * args.steps / (e2e_ms / 1e3), unit=UNIT, h2d_bytes_per_step=h2d,
                   d2h_bytes_per_step=d2h)

    dense_ms, temporal_ms, how = run.kernel_times()
    total_videos = wl["total_videos"] or world * videos
    value = total_videos * args.steps / (ms_total / 1e3)
    fallbacks = fastpath.fallback_counts()
    if fallbacks:
        raise RuntimeError(f"module-level fast path fell back to the stock PyTorch forward: {fallbacks}")
    tkey = "_c2" if (wl["latent"] == 64 and videos == 1 and wl["frames"] == 16) else None
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_total / args.steps, higher_is_better=True, scaling=wl["scaling"], vs_baseline=None,
                dtype="bf16", data="synthetic",
                config=dict(workload=wl["text"],
                            parallelism=(f"dp{world} (independent videos, no collective)"),
                            videos_per_gpu=videos,
                            processors="install(unet, fast_path=True): B200 processors + module-level fast path",
                            launch="CUDA graph replay" if run.graphed is not None else "eager",
                            l2="working set >> 126 MB L2 (2.7 GB bf16 weights, 84 MB activations per level-0 tensor)"),
                e2e=e2e, gpu_launches=int(launches),
                roofline=dense_roofline(dense_ms, run.bf, run.seq, peaks, "dense_l0" + tkey if tkey else None, how),
                roofline_temporal=temporal_roofline(temporal_ms, 2 * videos * run.seq, wl["frames"], peaks,
                                                    "temporal_l0" + tkey if tkey else None, how),
                cpu_baseline=None, clocks=clocks, fast_path_fallbacks=fallbacks, legs={})
    emitter.line = line
    run.close()

    # ---- extra legs of the default run: C4 (long sequence), C5 (frame-sharded clip), reference GPU path, CPU port ----
    legs = line["legs"]
    if args.config == "c2" and not args.no_legs:
        try:
            emitter.guard("c4", 300)
            w4 = WORKLOADS["c4"]
            r4 = StepRunner(unet, sched, dev, rank, world, 1, w4["frames"], w4["latent"], args.graph)
            ms4 = r4.timed(3, 2)
            d4, t4, how4 = r4.kernel_times(replays=3)
            legs["c4"] = dict(workload=w4["text"], n_gpus=world, steps=3, warmup=2, ms_per_step=ms4 / 3,
                              value=world * 3 / (ms4 / 1e3), unit=UNIT, scaling="weak",
                              launch="CUDA graph replay" if r4.graphed is not None else "eager",
                              roofline=dense_roofline(d4, r4.bf, r4.seq, peaks, None, how4),
                              roofline_temporal=temporal_roofline(t4, 2 * r4.seq, w4["frames"], peaks, None, how4))
            r4.close()
        except Exception as e:  # noqa: BLE001
            legs["c4"] = dict(error=f"{type(e).__name__}: {e}")
        try:
            emitter.guard("c5", 300)
            legs["c5"] = dict(workload=WORKLOADS["c5"]["text"],
                              **leg_frame_sharded(unet, sched, dev, rank, world, steps=3, warmup=2, graph=args.graph))
        except Exception as e:  # noqa: BLE001
            legs["c5"] = dict(error=f"{type(e).__name__}: {e}")
        emitter.cancel()
        fb = fastpath.fallback_counts()
        if fb:
            legs["fast_path_fallbacks"] = fb
    if world == 1 and not args.no_legs:
        try:
            emitter.guard("gpu_baseline", 300)
            handle.uninstall()
            line["gpu_baseline"] = leg_gpu_baseline(unet, sched, dev)
        except Exception as e:  # noqa: BLE001
            line["gpu_baseline"] = dict(error=f"{type(e).__name__}: {e}")
        emitter.cancel()
    if world == 1 and not args.no_cpu_baseline:
        del unet
        torch.cuda.empty_cache()
        emitter.guard("cpu_baseline", 600)
        line["cpu_baseline"], _, _ = cpu_reference_run(1, 0)
        emitter.cancel()
    emitter.emit()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the c4 / c5 / gpu_baseline legs of the default run")
    ap.add_argument("--no-e2e", action="store_true", help="developer runs under a profiler: skip the host-buffer leg")
    ap.add_argument("--cpu-budget", type=float, default=200.0,
                    help="--impl reference: wall-clock budget (s) for the CPU steps; fewer timed steps if they do not fit")
    ap.add_argument("--graph", dest="graph", action="store_true", default=True,
                    help="(default) replay a captured CUDA graph of the denoise step (i2v_adapter_unofficial_b200.graph) "
                         "in both timed regions")
    ap.add_argument("--eager", dest="graph", action="store_false",
                    help="launch the step kernel by kernel from Python instead of replaying the captured graph")
    ap.add_argument("--profiler-range", action="store_true",
                    help="bracket the timed steps with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    import torch
    import torch.distributed as dist

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    emitter = Emitter(rank)
    try:
        run_b200(args, rank, world, local_rank, emitter)
    finally:
        emitter.cancel()
        emitter.emit()   # a failure after the main measurement still leaves the line
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
