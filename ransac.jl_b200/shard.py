"""Point-range sharding over the GPUs of one box: one process per GPU (torch.distributed, NCCL).

Every rank holds the whole cloud (the sampler needs random access and all ranks draw the same
Philox minimal sets), scores and refits only its own point range, and the per-candidate counts and
inlier-mask words are summed with an NCCL all-reduce that the C library triggers through a callback
(`rsc_ctx_set_allreduce`).  The same partition is used by bench.py for the scoring microbenchmark.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

from . import _lib
from ._lib import lib

ALIGN = 2048  # refit CTAs cover 2048 points; score tiles 512


def partition(n: int, world: int, align: int = ALIGN) -> List[Tuple[int, int]]:
    """Contiguous point ranges [lo, hi) per rank: aligned to `align`, covering [0, n) exactly once.
    Trailing ranks may be empty when n is small."""
    blocks = (n + align - 1) // align
    out = []
    for r in range(world):
        lo = (blocks * r // world) * align
        hi = (blocks * (r + 1) // world) * align
        out.append((min(lo, n), min(hi, n)))
    return out


class _DevInt32:
    """__cuda_array_interface__ view of a raw device pointer (int32[n])."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 3}


class ShardedContext:
    """Installs the NCCL all-reduce callback on a context and the rank's point range on a cloud."""

    def __init__(self, pc, group=None):
        import torch
        import torch.distributed as dist

        self.pc = pc
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.range = partition(pc.size, self.world)[self.rank]
        ext = torch.cuda.ExternalStream(pc.ctx.stream, device=torch.device("cuda", pc.ctx.device))

        def _allreduce(user, ptr, count, stream):
            try:
                t = torch.as_tensor(_DevInt32(ptr, count), device=torch.device("cuda", pc.ctx.device))
                with torch.cuda.stream(ext):
                    dist.all_reduce(t, group=group)
                return 0
            except Exception as e:  # never let an exception cross the C boundary
                print(f"[rsc] all-reduce callback failed: {e!r}")
                return 1

        self._cb = _lib.ALLREDUCE_FN(_allreduce)  # keep alive
        pc.ctx.check(lib.rsc_ctx_set_allreduce(pc.ctx.h, C.cast(self._cb, C.c_void_p), None))
        lo, hi = self.range
        if hi > lo:
            pc.ctx.check(lib.rsc_cloud_set_range(pc.handle, lo, hi))

    def close(self):
        lib.rsc_ctx_set_allreduce(self.pc.ctx.h, None, None)
