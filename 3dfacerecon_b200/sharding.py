"""Batch sharding across GPUs.

Every face (row of ``params``) is independent in the forward pass and in both backward passes -- the reference's batch
loop is a plain outer loop (``render_depth_op.cc:180``) -- so the path shards by contiguous batch ranges with the basis,
triangles and textures replicated per GPU and NO data-path collective.  One process per GPU (``torchrun``) calls
``shard_batch`` with its rank and runs the ordinary single-GPU API on its slice; outputs stay on the producing GPU
unless the caller gathers them.
"""
from __future__ import annotations


def shard_batch(batch: int, world_size: int, rank: int):
    """Contiguous, balanced split: returns ``(start, count)`` for ``rank``.  The first ``batch % world_size`` ranks
    get one extra face; ranks beyond ``batch`` get ``count == 0``."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    if batch < 0:
        raise ValueError("batch must be >= 0")
    base, rem = divmod(batch, world_size)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def shard_slices(batch: int, world_size: int):
    return [shard_batch(batch, world_size, r) for r in range(world_size)]
