"""Multi-GPU plumbing: one process per GPU (`torchrun`), `torch.distributed` (NCCL on B200, gloo in CPU tests).

The hot path shards across GPUs in two ways:
  * videos: every rank denoises its own orbital video (BASELINE configs[4]); no data-path collective, only the
    final gather of decoded frames / a max-over-ranks of the device time.  This is what bench.py --gpus N runs.
  * frames of one video (BASELINE configs[3]): rank r owns frames [r*T/R, (r+1)*T/R) of both CFG halves.
    Per-frame ops need nothing; the temporal ops need (i) an all-gather of the temporal-attention K/V rows,
    (ii) a one-frame halo for the (3,1,1) temporal conv and (iii) an all-reduce of the (sum, sumsq) GroupNorm
    partials.  The helpers below implement exactly those three exchanges on plain tensors so the same code is
    exercised under gloo on CPU (tests/test_dist_cpu.py) and NCCL on GPUs.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, world_size: int) -> range:
    """Contiguous, balanced split of n items (videos or frames): the first n % world ranks get one extra."""
    base, extra = divmod(n, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def frame_owner(t: int, T: int, world_size: int) -> int:
    for r in range(world_size):
        if t in shard_range(T, r, world_size):
            return r
    raise ValueError(t)


def max_over_ranks_ms(ms: float, device) -> float:
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_to_rank0(x: torch.Tensor) -> Optional[torch.Tensor]:
    """Concatenate equally-shaped per-rank tensors along dim 0 on rank 0 (decoded frames / final latents)."""
    rank, ws = world()
    if ws == 1:
        return x
    bufs = [torch.empty_like(x) for _ in range(ws)]
    dist.all_gather(bufs, x.contiguous())
    return torch.cat(bufs, 0) if rank == 0 else None


# ---- frame-sharded exchanges ------------------------------------------------------------------------------------------
def allgather_frames(local: torch.Tensor, T: int) -> torch.Tensor:
    """local: [b, t_local, ...] rows of this rank's frames -> [b, T, ...] with every rank's frames in frame order
    (the K/V all-gather before each temporal-attention block)."""
    rank, ws = world()
    if ws == 1:
        return local
    if T % ws:
        raise ValueError("frame sharding needs T divisible by the world size")
    bufs = [torch.empty_like(local) for _ in range(ws)]
    dist.all_gather(bufs, local.contiguous())
    return torch.cat(bufs, dim=1)


def halo_exchange(local: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """local: [b, t_local, ...].  Returns (prev, next): the last frame of rank-1 and the first frame of rank+1
    (zeros at the clip boundaries = the Conv3d zero padding at t = -1 and t = T, openaimodel.py:252-261)."""
    rank, ws = world()
    first, last = local[:, :1].contiguous(), local[:, -1:].contiguous()
    prev, nxt = torch.zeros_like(first), torch.zeros_like(last)
    if ws == 1:
        return prev, nxt
    ops = []
    if rank + 1 < ws:
        ops += [dist.P2POp(dist.isend, last, rank + 1), dist.P2POp(dist.irecv, nxt, rank + 1)]
    if rank > 0:
        ops += [dist.P2POp(dist.isend, first, rank - 1), dist.P2POp(dist.irecv, prev, rank - 1)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    return prev, nxt


def allreduce_gn_partials(partials: torch.Tensor) -> torch.Tensor:
    """partials: [b, 32, 2] (sum, sumsq) of this rank's frames -> totals over all frames of the clip."""
    if world()[1] > 1:
        dist.all_reduce(partials, op=dist.ReduceOp.SUM)
    return partials
