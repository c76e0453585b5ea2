"""Multi-GPU host logic on CPU: world_size-2 gloo processes drive the PRODUCT's shard
bookkeeping - `mtgvision_b200.shards.ShardCursor`, the object behind `RanMtgEncDecDataset._next_first_index`
and `Gen.random_batch` - through the call sequence of `create_yolo_obb_dataset` (full batches, a tail batch,
single scenes), all_gather the index ranges and check that they are disjoint and tile the index space.
No collective exists on the data path itself; the all_reduce mirrors the bench's max-over-ranks timing."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _call_sizes():
    # train 600 in batches of 256 (tail 88), val 60, test 60, a few single scenes, a changed batch size
    sizes = []
    for num in (600, 60, 60):
        i = 0
        while i < num:
            sizes.append(min(256, num - i))
            i += sizes[-1]
    return sizes + [1, 1, 1, 512, 512, 7]


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    from mtgvision_b200.shards import ShardCursor

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cur = ShardCursor(rank, world)
    sizes = _call_sizes()
    mine = torch.tensor([[cur.next_first(n), n] for n in sizes], dtype=torch.int64)
    got = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(got, mine)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the bench's max-over-ranks timing reduction
    if rank == 0:
        ranges = sorted((int(f), int(f + n)) for g in got for f, n in g.tolist())
        tiled = ranges[0][0] == 0 and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        ret["ok"] = tiled and ranges[-1][1] == world * sum(sizes) and t.item() == world
        ret["cursor"] = cur.cursor
    dist.destroy_process_group()


def test_shard_index_ranges_are_disjoint_and_cover_with_varying_batch_sizes():
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, 29571, ret), nprocs=world, join=True)
    assert ret.get("ok") is True
    assert ret["cursor"] == world * sum(_call_sizes())


def test_datasets_use_the_cursor():
    """The dataset classes take their indices from ShardCursor (no private formula left behind)."""
    import inspect

    from mtgvision_b200 import encoder_train, od_datasets, shards

    ds = encoder_train.RanMtgEncDecDataset.__new__(encoder_train.RanMtgEncDecDataset)  # no CUDA context needed
    ds._shards = shards.ShardCursor(1, 2)
    assert [ds._next_first_index(n) for n in (4, 4, 1, 8)] == [4, 12, 17, 26]
    assert "self._shards.next_first(n)" in inspect.getsource(od_datasets.Gen.random_batch)


def test_single_rank_split_has_no_duplicates():
    """train/val/test of create_yolo_obb_dataset (20000/2000/2000, batch 256) must not share indices."""
    from mtgvision_b200.shards import ShardCursor

    cur = ShardCursor()
    seen = set()
    for num in (20000, 2000, 2000):
        i = 0
        while i < num:
            n = min(256, num - i)
            f = cur.next_first(n)
            block = set(range(f, f + n))
            assert not (seen & block)
            seen |= block
            i += n
    assert len(seen) == 24000
