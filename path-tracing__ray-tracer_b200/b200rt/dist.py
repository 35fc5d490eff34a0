"""Multi-GPU plumbing: samples-per-pixel split + one sum-reduce of the accumulation buffer.

The path shards with no exchange during rendering (samples of a pixel are i.i.d.): rank g of G
renders global sample indices [offset_g, offset_g + spp_g) of EVERY pixel into its own float
buffer; the counter-based RNG is keyed by (pixel, global sample index) so the union of samples does
not depend on G.  The only collective is one ``reduce(SUM)`` to rank 0 (NCCL over NVLink on GPUs,
gloo in the CPU tests), followed by the resolve kernel on rank 0.  The reference is single-GPU
(``cuda.select_device(0)``, cuda_path_tracer.py:743); its only multi-pass hook is the
``frame_count`` reseed (:28).
"""
from __future__ import annotations

from typing import Tuple

import torch


def rank_world() -> Tuple[int, int]:
    import torch.distributed as td
    if td.is_available() and td.is_initialized():
        return td.get_rank(), td.get_world_size()
    return 0, 1


def split_samples(spp: int, rank: int, world: int) -> Tuple[int, int]:
    """-> (spp_local, sample_offset): contiguous, exhaustive, sizes differ by at most one."""
    base, rem = divmod(int(spp), int(world))
    local = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return local, offset


def reduce_to_root(buf: torch.Tensor) -> None:
    """In-place SUM onto rank 0 (no-op for a single process)."""
    import torch.distributed as td
    if td.is_available() and td.is_initialized() and td.get_world_size() > 1:
        td.reduce(buf, dst=0, op=td.ReduceOp.SUM)


class SymmetricImage:
    """Peer-accessible accumulation (float32[4*H*W]) and image (uint8[3*H*W]) buffers of ONE image size.

    Allocated in symmetric memory (``torch.distributed._symmetric_memory``: every rank's buffer is mapped into every
    process over NVLink) once per renderer and reused, so that ``libb200rt``'s fused reduce + resolve kernel
    (``b2rt_reduce_resolve``) can read all ranks' sums and write the root's image directly: reduce-scatter, resolve and
    gather in one kernel per rank, with a stream-ordered cross-rank barrier on either side.  ``create`` is collective and
    returns ``None`` on EVERY rank if any rank cannot set it up (the caller then uses one NCCL reduce + resolve on the
    root)."""

    def __init__(self, accum, u8, hdl, hdl_u8, rank, world):
        import ctypes as C
        self.accum, self.u8, self.hdl, self.hdl_u8, self.rank, self.world = accum, u8, hdl, hdl_u8, rank, world
        self.peer_ptrs = (C.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
        self.root_u8_ptr = int(hdl_u8.buffer_ptrs[0])

    @classmethod
    def create(cls, n_pixels: int, device):
        import os
        import torch.distributed as td
        rank, world = rank_world()
        if world < 2 or not (td.is_available() and td.is_initialized()):
            return None
        ok, obj = 1, None
        try:
            if os.environ.get("B200RT_NO_SYMM", "0") == "1" or world > 16:
                raise RuntimeError("symmetric memory disabled")
            import torch.distributed._symmetric_memory as symm
            accum = symm.empty(4 * n_pixels, dtype=torch.float32, device=device)
            u8 = symm.empty(3 * n_pixels, dtype=torch.uint8, device=device)
            hdl = symm.rendezvous(accum, td.group.WORLD)
            hdl_u8 = symm.rendezvous(u8, td.group.WORLD)
            obj = cls(accum, u8, hdl, hdl_u8, rank, world)
        except Exception as e:                                      # noqa: BLE001 — any failure means "use NCCL"
            import sys
            print(f"[b200rt] rank {rank}: symmetric memory unavailable ({type(e).__name__}: {e}); using ncclReduce",
                  file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        td.all_reduce(flag, op=td.ReduceOp.MIN)                     # all ranks take the same path
        return obj if int(flag.item()) == 1 else None

    def rows(self, height: int) -> Tuple[int, int]:
        per = -(-height // self.world)
        lo = min(height, self.rank * per)
        return lo, min(height, lo + per)

    def barrier(self) -> None:
        """Stream-ordered barrier across the ranks (signal pads in symmetric memory; no host synchronisation)."""
        self.hdl.barrier(channel=0)
