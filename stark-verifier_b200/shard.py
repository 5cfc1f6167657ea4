"""Sharding of a proof batch over the GPUs of one box.

Proofs are independent, so the batch is cut into contiguous per-rank ranges whose boundaries are
multiples of 32 proofs (one accept-bitmap word never straddles two ranks); every rank verifies its
own range and the only exchange is one all-gather of the packed accept bitmaps
(``torch.distributed`` -- NCCL over NVLink on GPUs, gloo in the CPU tests).  The reference is
single-process; this is new.
"""
from __future__ import annotations

from typing import Tuple


def words_per_rank(n_proofs: int, world: int) -> int:
    """bitmap words owned by each rank (the last ranks may own padding words)."""
    total_words = (n_proofs + 31) // 32
    return (total_words + world - 1) // world


def shard_range(n_proofs: int, rank: int, world: int) -> Tuple[int, int]:
    """[first, last) proofs of `rank`; every boundary is a multiple of 32."""
    w = words_per_rank(n_proofs, world)
    first = min(n_proofs, rank * w * 32)
    last = min(n_proofs, (rank + 1) * w * 32)
    return first, last


def gather_bitmap(local_words, world: int, out=None):
    """all-gather the per-rank bitmap words (a 1-D int32 tensor of words_per_rank entries, padded
    with zeros) into the full bitmap, identical on every rank."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local_words
    if out is None:
        out = torch.empty(local_words.numel() * world, dtype=local_words.dtype, device=local_words.device)
    dist.all_gather_into_tensor(out, local_words)
    return out
