"""Frame / batch partitioner on CPU: world_size-2 (and 4) `gloo` process groups exercise the host-side logic of the
multi-GPU path — re-shard index maps, the all-to-all round trip, the sharded GroupNorm statistics, the first-frame
broadcast and a frame-sharded UNet forward that must equal the unsharded one.  The attention arithmetic runs through
the stock (PyTorch SDPA) processors here: the partitioner is independent of which processors are installed."""
import os
import socket
import sys
import traceback

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, fn_name, errq):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.set_num_threads(2)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        globals()[fn_name](rank, world)
        dist.barrier()
    except Exception:  # noqa: BLE001
        errq.put((rank, traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def _run(fn_name, world=2):
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn_name, errq)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    errors = []
    while not errq.empty():
        errors.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            errors.append((-1, "worker timed out"))
    assert not errors, "\n".join(f"[rank {r}] {e}" for r, e in errors)
    assert all(p.exitcode == 0 for p in procs)


# ---------------------------------------------------------------------------------------------------------------
# worker bodies (module-level so that spawn can import them)
# ---------------------------------------------------------------------------------------------------------------
def _w_reshard_round_trip(rank, world):
    from i2v_adapter_unofficial_b200.partition import frames_to_positions, positions_to_frames

    V, F, S, C = 2, 4 * world, 8 * world, 6
    g = torch.Generator().manual_seed(0)
    full = torch.randn(V, F, S, C, generator=g)             # the unsharded activation, identical on every rank
    f, sl = F // world, S // world
    mine = full[:, rank * f:(rank + 1) * f].contiguous()    # my frames, all positions
    y = frames_to_positions(mine, allow_torch_layout=True)
    assert y.shape == (V, F, sl, C)
    assert torch.equal(y, full[:, :, rank * sl:(rank + 1) * sl])   # all frames, my positions
    back = positions_to_frames(y, allow_torch_layout=True)
    assert torch.equal(back, mine)
    with pytest.raises(RuntimeError):                        # no silent CPU path
        frames_to_positions(mine)


def _w_sharded_group_norm(rank, world):
    from i2v_adapter_unofficial_b200.partition import sharded_group_norm

    V, F, C, h, w = 2, 2 * world, 16, 3, 5
    g = torch.Generator().manual_seed(1)
    full = torch.randn(V, F, C, h, w, generator=g) * 3 + 1.5
    norm = torch.nn.GroupNorm(4, C, eps=1e-6)
    with torch.no_grad():
        norm.weight.copy_(torch.randn(C, generator=g))
        norm.bias.copy_(torch.randn(C, generator=g))
        ref = norm(full.permute(0, 2, 1, 3, 4)).permute(0, 2, 1, 3, 4)   # stats over (C/g, F, h, w) as diffusers does
        f = F // world
        got = sharded_group_norm(full[:, rank * f:(rank + 1) * f].contiguous(), norm)
    assert torch.allclose(got, ref[:, rank * f:(rank + 1) * f], atol=2e-5, rtol=1e-5)


def _w_batch_partition(rank, world):
    from i2v_adapter_unofficial_b200.partition import BatchPartition

    n = 2 * world + 1  # uneven on purpose
    part = BatchPartition(n, world, rank)
    x = torch.arange(n * 3, dtype=torch.float32).view(n, 3)
    local = part.split(x)
    assert sum(part.counts) == n and local.shape[0] == part.counts[rank]
    assert torch.equal(part.gather(local * 2), x * 2)
    emb = torch.arange(2 * n, dtype=torch.float32).view(2 * n, 1)
    cfg = part.split_cfg(emb)
    k = part.counts[rank]
    assert torch.equal(cfg[:k], emb[:n][part.local_slice]) and torch.equal(cfg[k:], emb[n:][part.local_slice])
    even = BatchPartition(2 * world, world, rank)
    assert torch.equal(even.gather(even.split(x[: 2 * world])), x[: 2 * world])


def _w_frame_sharded_unet(rank, world):
    """Frame-sharded UNet forward (stock processors) == unsharded forward, frame by frame."""
    from helpers import make_unet, randomize_zero_init, unet_inputs
    from i2v_adapter_unofficial_b200.partition import FramePartitioner

    unet = randomize_zero_init(make_unet(ip_adapter=True))   # same seed -> identical weights on every rank
    F = 2 * world
    sample, ctx, img = unet_inputs(unet, videos=2, frames=F, size=16, tokens=5, image_embed_dim=64)
    with torch.no_grad():
        ref = unet(sample, 37, True, ctx, added_cond_kwargs={"image_embeds": img}).sample
    part = FramePartitioner(unet, allow_torch_layout=True).install().watch_unet_forward()
    with torch.no_grad():
        out = unet(part.shard_frames(sample), 37, True, ctx, added_cond_kwargs={"image_embeds": img}).sample
        full = part.gather_frames(out)
    f = F // world
    assert out.shape == ref[:, rank * f:(rank + 1) * f].shape
    err = (full - ref).abs().max().item()
    assert err <= 2e-4 * ref.abs().max().item(), err
    # one broadcast per cross-frame block, two all-to-alls and one statistics gather per motion module
    # (16 / 42 / 21 for the SD1.5 layout with two layers per block)
    n_xf = sum(1 for m in unet.modules() if hasattr(m, "i2v_adapter") and hasattr(m, "attn1"))
    n_mm = sum(1 for m in unet.modules() if type(m).__name__ == "TransformerTemporalModel")
    assert part.stats == {"broadcasts": n_xf, "all_to_alls": 2 * n_mm, "stat_gathers": n_mm} and n_xf > 0 and n_mm > 0
    part.uninstall()
    with torch.no_grad():
        again = unet(sample, 37, True, ctx, added_cond_kwargs={"image_embeds": img}).sample
    assert torch.equal(again, ref)


def _w_sharded_denoise_first_frame(rank, world):
    """Only the owner of frame 0 re-imposes the condition latents; the sharded loop equals the unsharded loop."""
    from helpers import make_unet, randomize_zero_init, unet_inputs
    from i2v_adapter_unofficial_b200.hostmodel import DDIMScheduler, denoise_step
    from i2v_adapter_unofficial_b200.partition import FramePartitioner, sharded_denoise_step

    unet = randomize_zero_init(make_unet())
    F = 2 * world
    sample, ctx, _ = unet_inputs(unet, videos=1, frames=F, size=16, tokens=5)
    cond = torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(3))
    prompt = torch.cat([torch.zeros_like(ctx), ctx])
    sched = DDIMScheduler()
    sched.set_timesteps(4)
    ref = sample.clone()
    for t in sched.timesteps[:2]:
        ref = denoise_step(unet, sched, ref, int(t), prompt, 7.5, cond)
    part = FramePartitioner(unet, allow_torch_layout=True).install().watch_unet_forward()
    lat = part.shard_frames(sample).clone()
    for t in sched.timesteps[:2]:
        lat = sharded_denoise_step(part, unet, sched, lat, int(t), prompt, 7.5, cond)
    full = part.gather_frames(lat)
    assert (full - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [2, 4])
def test_reshard_round_trip(world):
    _run("_w_reshard_round_trip", world)


def test_sharded_group_norm_statistics():
    _run("_w_sharded_group_norm", 2)


def test_batch_partition_split_and_gather():
    _run("_w_batch_partition", 2)


def test_frame_sharded_unet_equals_unsharded():
    _run("_w_frame_sharded_unet", 2)


def test_sharded_denoise_step_first_frame_owner():
    _run("_w_sharded_denoise_first_frame", 2)
