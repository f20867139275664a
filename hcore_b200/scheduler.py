"""Tile scheduler: shard a batch of independent tile updates over the GPUs of one box.

Tiles are independent: rank ``r`` owns the contiguous tile range
``shard_range(ntiles, world, r)`` (computed by the C library so C callers get the same
partition), runs the batched kernels on its own shard with no data-path collective, and
only the per-tile ``{rank, status}`` pairs (8 bytes per tile) are gathered -- NCCL on GPUs,
gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np

from .api import shard_range

__all__ = ["TileScheduler"]


class TileScheduler:
    def __init__(self, ntiles, world=None, rank=None):
        import torch.distributed as dist
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.world = world if world is not None else (self.dist.get_world_size() if self.dist else 1)
        self.rank = rank if rank is not None else (self.dist.get_rank() if self.dist else 0)
        self.ntiles = int(ntiles)
        self.first, self.count = shard_range(self.ntiles, self.world, self.rank)

    def owner(self, tile):
        """Rank that owns global tile index ``tile``."""
        for r in range(self.world):
            f, c = shard_range(self.ntiles, self.world, r)
            if f <= tile < f + c:
                return r
        raise IndexError(tile)

    def all_ranges(self):
        return [shard_range(self.ntiles, self.world, r) for r in range(self.world)]

    def gather_results(self, local_rank_i32, local_status_i32):
        """All-gather the per-tile new ranks and status codes of every shard.

        ``local_*`` are (count,) int32 tensors on this rank's device (CUDA with NCCL, CPU with
        gloo).  Returns two (ntiles,) int32 tensors in global tile order."""
        import torch
        if self.dist is None or self.world == 1:
            return local_rank_i32.clone(), local_status_i32.clone()
        ranges = self.all_ranges()
        maxc = max(c for _, c in ranges)
        dev = local_rank_i32.device
        pack = torch.zeros(2 * maxc, dtype=torch.int32, device=dev)
        pack[: self.count] = local_rank_i32
        pack[maxc: maxc + self.count] = local_status_i32
        out = torch.empty(self.world * 2 * maxc, dtype=torch.int32, device=dev)
        self.dist.all_gather_into_tensor(out, pack)
        out = out.view(self.world, 2, maxc)
        ranks = torch.cat([out[r, 0, :c] for r, (_, c) in enumerate(ranges)])
        stats = torch.cat([out[r, 1, :c] for r, (_, c) in enumerate(ranges)])
        return ranks, stats

    # ---- the same gather through the C ABI (hcore_b200_gather_status: ncclAllGather on a raw ncclComm_t) ------------
    def nccl_comm(self):
        """A raw ``ncclComm_t`` over the ranks of this job, created with NCCL's own C API through ctypes (the unique id
        travels over the existing torch.distributed group).  What a C caller would pass to ``hcore_b200_gather_status``."""
        import ctypes as C
        import torch
        if getattr(self, "_comm", None):
            return self._comm
        nccl = C.CDLL("libnccl.so.2")

        class UniqueId(C.Structure):
            _fields_ = [("internal", C.c_char * 128)]

        nccl.ncclGetUniqueId.argtypes = [C.POINTER(UniqueId)]
        nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
        uid = UniqueId()
        if self.rank == 0:
            assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
        buf = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).clone().cuda()
        self.dist.broadcast(buf, 0)
        C.memmove(C.byref(uid), bytes(buf.cpu().numpy().tobytes()), 128)
        comm = C.c_void_p()
        rc = nccl.ncclCommInitRank(C.byref(comm), self.world, uid, self.rank)
        if rc != 0:
            raise RuntimeError(f"ncclCommInitRank failed ({rc})")
        self._nccl, self._comm = nccl, comm.value
        return self._comm

    def gather_results_c(self, ctx, local_rank_i32, local_status_i32):
        """``gather_results`` through ``hcore_b200_gather_status`` (C ABI, one ncclAllGather per array)."""
        import torch
        dev = local_rank_i32.device
        ranks = torch.empty(self.ntiles, dtype=torch.int32, device=dev)
        stats = torch.empty(self.ntiles, dtype=torch.int32, device=dev)
        comm = self.nccl_comm() if (self.dist is not None and self.world > 1) else None
        ctx.gather_status(comm, self.world, self.rank, self.ntiles, local_rank_i32, local_status_i32, ranks, stats)
        torch.cuda.synchronize()
        return ranks, stats
