"""Shard bookkeeping of the global sample-index space: no CUDA, no collective.

Every sample / scene is a pure function of `(seed, global index)`, so the only multi-GPU state is WHICH
indices a rank generates.  `ShardCursor` hands out, call after call, the contiguous block of `n` indices that
belongs to this rank and advances past the blocks of all ranks:

    first = cursor + rank * n ;  cursor += world_size * n

A running cursor (not `counter * n`) keeps the blocks disjoint when `n` changes between calls - the tail
batch of `create_yolo_obb_dataset` (od_datasets.py:732-791 of the reference writes train / val / test from one
generator), `Gen.random()` mixed with `random_batch(n)`, `set_batch_size`, `image_batch_by_ids`.  All ranks must
make the same sequence of calls (they do: one process per GPU running the same loop).
"""

from __future__ import annotations


class ShardCursor:
    def __init__(self, rank: int = 0, world_size: int = 1, start: int = 0):
        rank, world_size = int(rank), int(world_size)
        if world_size < 1 or not 0 <= rank < world_size:
            raise ValueError(f"rank {rank} outside world_size {world_size}")
        self.rank, self.world_size = rank, world_size
        self.cursor = int(start)

    def next_first(self, n: int) -> int:
        """First global index of this rank's block of `n`; advances past every rank's block."""
        n = int(n)
        if n < 0:
            raise ValueError("n must be >= 0")
        first = self.cursor + self.rank * n
        self.cursor += self.world_size * n
        return first

    def state_dict(self) -> dict:
        return {"cursor": self.cursor, "rank": self.rank, "world_size": self.world_size}

    def load_state_dict(self, state: dict) -> None:
        self.cursor = int(state["cursor"])
