"""Frame partitioner on real GPUs (-m gpu, needs >= 2 devices; skipped otherwise): NCCL broadcast of the frame-0
K/V, all-to-all re-shard through the C-ABI layout kernels, B200 processors installed.  The frame-sharded UNet
forward must reproduce the single-GPU forward of the same bf16 model."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from test_partition_gloo import _free_port  # noqa: E402

pytestmark = pytest.mark.gpu


def _nccl_worker(rank, world, port, errq, scenario="default"):
    import traceback

    import torch.distributed as dist

    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device(f"cuda:{rank}")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from helpers import SD15_HEADDIM_CFG, make_unet, unet_inputs
        from i2v_adapter_unofficial_b200 import _lib, install
        from i2v_adapter_unofficial_b200.partition import FramePartitioner, frames_to_positions, positions_to_frames

        # 1. re-shard round trip through the CUDA layout kernels + NCCL all-to-all
        V, F, S, C = 2, 4 * world, 64, 320
        g = torch.Generator().manual_seed(0)
        full = torch.randn(V, F, S, C, generator=g).to(torch.bfloat16).to(dev)
        f, sl = F // world, S // world
        mine = full[:, rank * f:(rank + 1) * f].contiguous()
        y = frames_to_positions(mine)
        assert torch.equal(y, full[:, :, rank * sl:(rank + 1) * sl])
        assert torch.equal(positions_to_frames(y), mine)

        # 2. frame-sharded UNet (SD1.5 head dims 40/80/160, one layer per block) vs the unsharded forward
        #    scenario "one_frame_unfused": one frame per rank and separate attn1 / i2v_adapter launches -- the local
        #    rows of every rank but the owner are NOT frame 0, the K/V must come from the owner's broadcast
        unet = make_unet(SD15_HEADDIM_CFG, ip_adapter=True, dtype=torch.bfloat16, device=dev)
        with torch.no_grad():   # the adapter's zero-initialised output projection would hide the cross-frame branch
            gen = torch.Generator().manual_seed(9)
            for name, p in unet.named_parameters():
                if ".i2v_adapter.to_out.0." in name:
                    p.copy_((torch.randn(p.shape, generator=gen) * 0.02).to(p))
        one_frame = scenario == "one_frame_unfused"
        F = world if one_frame else 4 * world
        videos = 2 if one_frame else 1
        sample, ctx, img = unet_inputs(unet, videos=videos, frames=F, size=32, tokens=77, image_embed_dim=64)
        sample, ctx, img = sample.to(dev, torch.bfloat16), ctx.to(dev, torch.bfloat16), img.to(dev, torch.bfloat16)
        install(unet, fuse_cross_frame=not one_frame)
        with torch.no_grad():
            ref = unet(sample, 37, True, ctx, added_cond_kwargs={"image_embeds": img}).sample.float()
        part = FramePartitioner(unet).install()
        n0 = _lib.launch_count()
        with torch.no_grad():
            out = unet(part.shard_frames(sample), 37, True, ctx, added_cond_kwargs={"image_embeds": img}).sample
            got = part.gather_frames(out).float()
        assert _lib.launch_count() > n0
        cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
        err = (got - ref).abs().max().item()
        assert cos >= 0.999 and err <= 0.05 * ref.abs().max().item(), (cos, err)
        assert part.stats["broadcasts"] > 0 and part.stats["all_to_alls"] > 0
        dist.barrier()
    except Exception:  # noqa: BLE001
        errq.put((rank, traceback.format_exc()))
    finally:
        import torch.distributed as dist2

        if dist2.is_initialized():
            dist2.destroy_process_group()


@pytest.mark.parametrize("scenario", ["default", "one_frame_unfused"])
def test_frame_sharded_unet_nccl(scenario):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 CUDA devices (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, errq, scenario)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
    errors = []
    while not errq.empty():
        errors.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            errors.append((-1, "worker timed out"))
    assert not errors, "\n".join(f"[rank {r}] {e}" for r, e in errors)
