"""Multi-GPU host logic on CPU: world_size-2 gloo processes derive their per-batch global
sample-index ranges exactly like RanMtgEncDecDataset._next_first_index and verify, after an
all_gather, that the shards are disjoint and tile the index space (no collective exists on
the data path itself: SURVEY.md section 8e)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _first_index(batch_counter, rank, world, n):
    return (batch_counter * world + rank) * n


def _worker(rank, world, port, n, batches, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = torch.tensor([_first_index(b, rank, world, n) for b in range(batches)], dtype=torch.int64)
    got = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(got, mine)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the bench's max-over-ranks timing reduction
    if rank == 0:
        starts = torch.cat(got).sort().values
        ret["ok"] = bool(torch.equal(starts, torch.arange(world * batches, dtype=torch.int64) * n)) and t.item() == world
    dist.destroy_process_group()


def test_shard_index_ranges_are_disjoint_and_cover():
    sys.path.insert(0, ROOT)
    from mtgvision_b200 import encoder_train

    src = open(encoder_train.__file__).read()
    assert "(b * self.world_size + self.rank) * n" in src  # the formula this test mirrors
    world, n, batches = 2, 512, 7
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, 29571, n, batches, ret), nprocs=world, join=True)
    assert ret.get("ok") is True
