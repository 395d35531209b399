"""Peer memory for the frame-sharded step: symmetric device buffers shared between the one-process-per-GPU
ranks, and the single-kernel exchange point (flag barrier + optional small all-reduce) that orders the peer loads / stores
the sharded kernels perform over NVLink (csrc/peer.cu, csrc/attn.cu `tattn_d64_kernel`, csrc/norm.cu `gn_apply_kernel`).

`torch.distributed` is plumbing only: it carries the 64-byte CUDA IPC handles between the ranks once per allocation
(`all_gather_object`) -- never tensor data on the step's path.  Every rank must call `alloc` in the same order with the same
size (the launch plans of all ranks are identical, so they do).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch

from . import _native as N

_DT = {torch.float16: "<f2", torch.float32: "<f4", torch.uint8: "|u1", torch.int32: "<i4"}


class _RawMem:
    """A raw device allocation presented through __cuda_array_interface__ so torch can view it without owning it."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self.ptr, self.nbytes, self.owner = ptr, nbytes, owner
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


class SymmBuffer:
    """One symmetric allocation: `ptrs[r]` is rank r's copy as seen from THIS process (ptrs[rank] = the local one)."""

    def __init__(self, group: "PeerGroup", nbytes: int):
        lib = N.load()
        self.group, self.nbytes = group, nbytes
        p = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        N.check(lib.hi3d_symm_alloc(nbytes, C.byref(p), handle), "hi3d_symm_alloc")
        self.local = int(p.value)
        handles: List[Optional[bytes]] = [None] * group.world
        import torch.distributed as dist
        dist.all_gather_object(handles, bytes(handle), group=group.pg)
        self.ptrs: List[int] = []
        self._opened: List[int] = []
        for r, hb in enumerate(handles):
            if r == group.rank:
                self.ptrs.append(self.local)
                continue
            q = C.c_void_p()
            buf = (C.c_ubyte * 64).from_buffer_copy(hb)
            N.check(lib.hi3d_symm_open(buf, C.byref(q)), f"hi3d_symm_open (rank {r}'s buffer)")
            self.ptrs.append(int(q.value))
            self._opened.append(int(q.value))
        self._raw = _RawMem(self.local, nbytes, self)
        self._u8 = torch.as_tensor(self._raw, device=group.device)
        self.ptr_array = (C.c_void_p * group.world)(*self.ptrs)

    def view(self, dtype=torch.float16) -> torch.Tensor:
        """The LOCAL copy as a flat tensor of `dtype` (no ownership: keep this SymmBuffer alive)."""
        return self._u8.view(dtype)

    def peer(self, r: int) -> Optional[int]:
        return self.ptrs[r] if 0 <= r < self.group.world else None

    def close(self):
        lib = N.load()
        for q in self._opened:
            lib.hi3d_symm_close(q)
        self._opened = []
        if self.local:
            lib.hi3d_symm_free(self.local)
            self.local = 0


class PeerGroup:
    """The ranks sharing one frame-sharded video.  Holds the exchange area (flags, payload slots, epoch)."""

    def __init__(self, rank: int, world: int, device, pg=None):
        if world < 2 or world > 16:
            raise ValueError(f"peer group of {world} ranks (2..16 supported)")
        self.rank, self.world, self.device, self.pg = rank, world, torch.device(device), pg
        self.buffers: List[SymmBuffer] = []
        nbytes = int(N.load().hi3d_peer_xchg_bytes(world))
        self.xchg = self.alloc(nbytes)
        self.n_exchanges = 0
        import torch.distributed as dist
        dist.barrier(group=pg)             # every rank has mapped every exchange area before the first exchange kernel

    def alloc(self, nbytes: int) -> SymmBuffer:
        nbytes = (int(nbytes) + 255) // 256 * 256
        b = SymmBuffer(self, nbytes)
        self.buffers.append(b)
        return b

    def exchange(self, payload: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
        """Flag barrier over NVLink on the current stream; with `payload` (fp32, <= 1024 values) also out = sum over ranks."""
        n = 0
        if payload is not None:
            if payload.dtype != torch.float32 or out is None or out.dtype != torch.float32 or out.numel() < payload.numel() \
                    or not payload.is_contiguous() or not out.is_contiguous():
                raise ValueError("peer exchange payload / out: contiguous fp32 tensors, out at least as large as payload")
            n = payload.numel()
        N.check(N.load().hi3d_peer_exchange(self.xchg.ptr_array, self.rank, self.world,
                                            None if payload is None else payload.data_ptr(), n,
                                            None if out is None else out.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "hi3d_peer_exchange")
        self.n_exchanges += 1


_groups: Dict[Tuple[int, int, int], PeerGroup] = {}


def get_group(rank: int, world: int, device) -> PeerGroup:
    """Process-wide peer group for (rank, world) on `device` (created on first use; needs torch.distributed initialised)."""
    dev = torch.device(device)
    key = (rank, world, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _groups:
        _groups[key] = PeerGroup(rank, world, dev)
    return _groups[key]
