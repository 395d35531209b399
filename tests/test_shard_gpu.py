"""Frame-sharded VideoUNet / fused Euler step over 2 GPUs (NCCL) against the single-GPU result of the same weights:
K/V all-gather before temporal attention, one-frame halo for the (3,1,1) conv, (sum, sumsq) all-reduce for the
(T,H,W) GroupNorm.  Needs >= 2 GPUs (skipped on the 1-GPU box; run with `gpurun --gpus 2`)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from hi3d_official_b200 import configs, spec
        dev = torch.device("cuda", rank)
        T, h = 8, 16
        model = configs.build_engine(1, device=dev, unet_overrides=dict(model_channels=64), vae_overrides=dict(ch=64),
                                     num_steps=3, num_frames=T)
        spec.synth_fill_(model, seed=1, fast=False)
        g = torch.Generator().manual_seed(0)
        x = torch.randn(T, 4, h, h, generator=g).to(dev)
        c = dict(crossattn=torch.randn(1, 1, 1024, generator=g).to(dev), vector=torch.randn(1, 768, generator=g).to(dev),
                 concat=(torch.randn(T, 4, h, h, generator=g) * 0.18).to(dev))
        uc = dict(crossattn=torch.zeros_like(c["crossattn"]), vector=c["vector"], concat=torch.zeros_like(c["concat"]))
        full = model.sample_stage1(c, uc, x.clone(), decode=False)                 # unsharded, on this GPU
        Tl = T // world
        sl = slice(rank * Tl, (rank + 1) * Tl)
        cl = dict(c, concat=c["concat"][sl].contiguous())
        ucl = dict(uc, concat=uc["concat"][sl].contiguous())
        part = model.sample_stage1(cl, ucl, x[sl].clone(), decode=False, shard=(rank, world))
        torch.cuda.synchronize()
        err = float((part - full[sl]).abs().max())
        ref = float(full.abs().mean())
        q.put((rank, err, ref))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, f"FAIL {type(e).__name__}: {e}\n{traceback.format_exc()[-1500:]}", 0.0))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_frame_sharded_sampler_matches_single_gpu():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    print(res)
    for rank, err, ref in res:
        assert not isinstance(err, str), err
        assert err < 2e-2 * max(1.0, ref), (rank, err, ref)
