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
