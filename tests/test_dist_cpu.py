"""world_size-2 gloo tests of the multi-GPU host logic (hi3d_official_b200/dist.py): sharding, the three
frame-sharded exchanges in the in-place form the launch plan calls (unet._Plan with HI3D_SHARD_EXCHANGE=nccl: K/V
all-gather, temporal-conv halo, GroupNorm partial all-reduce) reproduce the unsharded oracle ops bit-for-bit / to fp32
round-off, and the max-over-ranks timing reduce."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from hi3d_official_b200 import dist as D


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        torch.manual_seed(0)
        b, T, hw, c = 2, 8, 6, 64
        x = torch.randn(b, T, hw, c)                       # full clip, identical on every rank
        mine = D.shard_range(T, rank, ws)
        loc = x[:, mine.start:mine.stop].contiguous()
        # (i) all-gather for temporal attention: the plan's in-place form on [B * rows, X] token matrices
        tl = len(mine)
        full2 = torch.empty(b * T * hw, c)
        D.gather_frames_(loc.reshape(b * tl * hw, c), full2, b, tl * hw, ws)
        full = full2.view(b, T, hw, c)
        assert torch.equal(full, x)
        q_ = loc.permute(0, 2, 1, 3).reshape(b * hw, len(mine), c)
        kv = full.permute(0, 2, 1, 3).reshape(b * hw, T, c)
        att = torch.softmax(q_ @ kv.transpose(1, 2) / 8.0, -1) @ kv
        ref = torch.softmax(kv @ kv.transpose(1, 2) / 8.0, -1) @ kv
        assert torch.allclose(att, ref[:, mine.start:mine.stop], atol=1e-6)
        # (ii) halo for the (3,1,1) temporal conv
        w = torch.randn(c, c, 3, 1, 1) * 0.05
        gh = torch.full((b, tl + 2, hw * c), float("nan"))                        # the plan's haloed buffer [B, T_local + 2, X]
        gh[:, 1:tl + 1] = loc.reshape(b, tl, hw * c)
        D.halo_exchange_(gh, rank, ws)
        assert not torch.isnan(gh).any()
        ext = gh.view(b, tl + 2, hw, c).permute(0, 3, 1, 2)[..., None]            # b c t hw 1
        y_loc = F.conv3d(ext, w)                                                 # valid conv over the halo
        y_ref = F.conv3d(x.permute(0, 3, 1, 2)[..., None], w, padding=(1, 0, 0))
        assert torch.allclose(y_loc, y_ref[:, :, mine.start:mine.stop], atol=1e-5)
        # (iii) GroupNorm over (C/32, T, H, W) from all-reduced partial sums
        g = loc.reshape(b, -1, 32, c // 32)
        part = torch.stack([g.sum((1, 3)), (g * g).sum((1, 3))], -1)             # b 32 2
        tot = D.allreduce_sum_(part.clone())
        cnt = T * hw * (c // 32)
        mean = tot[..., 0] / cnt
        var = tot[..., 1] / cnt - mean ** 2
        ref_gn = F.group_norm(x.permute(0, 3, 1, 2), 32, eps=1e-5)               # b c T hw
        mine_gn = (loc.reshape(b, len(mine), hw, 32, c // 32) - mean[:, None, None, :, None]) / \
            torch.sqrt(var + 1e-5)[:, None, None, :, None]
        assert torch.allclose(mine_gn.reshape(b, len(mine), hw, c).permute(0, 3, 1, 2), ref_gn[:, :, mine.start:mine.stop],
                              atol=1e-4)
        # timing reduce + gather
        assert D.max_over_ranks_ms(10.0 + rank, "cpu") == 10.0 + ws - 1
        gathered = D.gather_to_rank0(torch.full((2, 3), float(rank)))
        if rank == 0:
            assert gathered.shape == (2 * ws, 3) and float(gathered[-1, 0]) == ws - 1
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"FAIL {type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_shard_range():
    assert [list(D.shard_range(16, r, 8)) for r in range(8)][3] == [6, 7]
    assert sum(len(D.shard_range(10, r, 4)) for r in range(4)) == 10
    assert list(D.shard_range(10, 0, 4)) == [0, 1, 2] and list(D.shard_range(10, 3, 4)) == [8, 9]
    assert D.frame_owner(7, 16, 8) == 3


def test_two_rank_gloo_exchanges():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
