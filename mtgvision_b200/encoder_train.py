"""Drop-in for the batch former of `mtgvision/encoder_train.py` (reference :74-249).

`RanMtgEncDecDataset` keeps the reference constructor, `__iter__`, `random_tensor_batch`,
`random_image_batch`, `image_batch_by_ids`, `set_batch_size`, `from_hparams` and the batch
dict keys (`x, x2, y, x_labels, x2_labels`).  Each batch is produced by three kernel
launches on the caller's CUDA stream (tape sampler, parameter expansion, plane
interpreter) and stays on the GPU: `x, x2[, y]` are contiguous NCHW tensors in `out_dtype`
(fp16 default; uint8 / fp32 selectable), labels are `[B, 3] int64`.

Use it with `DataLoader(ds, batch_size=None, num_workers=0)` - CUDA tensors cannot cross
the worker fork the reference relies on (encoder_train.py:517-523), and none is needed.

Multi-GPU: one process per GPU, `rank`/`world_size` select a disjoint slice of the global
Philox sample-index space; nothing is exchanged between ranks.
"""

from __future__ import annotations

import random
import uuid
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
from torch.utils.data import IterableDataset

from . import abi, synth
from .context import Context
from .shards import ShardCursor
from .encoder_datasets import IlsvrcImages, SyntheticBgFgMtgImages

_OUT = {"float16": abi.OUT_F16, "fp16": abi.OUT_F16, "uint8": abi.OUT_U8, "u8": abi.OUT_U8,
        "float32": abi.OUT_F32, "fp32": abi.OUT_F32}


class RanMtgEncDecDataset(IterableDataset):
    def __init__(
        self,
        default_batch_size: int,
        *,
        predownload: bool = False,
        paired: bool = False,
        targets: bool = True,
        x_size_hw: Tuple[int, int] = (192, 128),
        y_size_hw: Tuple[int, int] = (192, 128),
        half_upsidedown: bool = False,
        target_is_input_prob: float = 0.05,
        similar_neg_prob: float = 0.2,
        check_data: bool = False,
        # ---- B200 path ----
        mtg: Optional[SyntheticBgFgMtgImages] = None,
        ilsvrc: Optional[IlsvrcImages] = None,
        device: Optional[int] = None,
        out_dtype: str = "float16",
        seed: Optional[int] = None,
        rank: int = 0,
        world_size: int = 1,
    ):
        assert default_batch_size > 0
        self.default_batch_size = default_batch_size
        self.paired = paired
        self.targets = targets
        self.x_size_hw = tuple(x_size_hw)
        self.y_size_hw = tuple(y_size_hw)
        self.mtg = mtg if mtg is not None else SyntheticBgFgMtgImages(img_type="small", predownload=predownload)
        self.ilsvrc = ilsvrc if ilsvrc is not None else IlsvrcImages()
        self.half_upsidedown = half_upsidedown
        self.target_is_input_prob = target_is_input_prob
        self.similar_neg_prob = similar_neg_prob
        self.check_data = check_data
        self.out_dtype = _OUT[out_dtype]
        self.rank, self.world_size = int(rank), int(world_size)
        self.seed = random.getrandbits(63) if seed is None else int(seed)
        self._shards = ShardCursor(self.rank, self.world_size)
        self.ctx = Context(device)
        self.ctx.set_encoder_config(x_size_hw=self.x_size_hw, y_size_hw=self.y_size_hw,
                                    target_is_input_prob=target_is_input_prob, similar_neg_prob=similar_neg_prob,
                                    half_upsidedown=half_upsidedown, paired=paired, targets=targets)
        pool = self.mtg.pool
        self.ctx.set_card_pool(pool.images, pool.labels3, pool.grp_off, pool.grp_mem)
        self.ilsvrc.fill_pool(self.ctx)
        big = self.ctx.oversized_backgrounds(self.x_size_hw)
        if len(big):
            # refused up front instead of flagged per sample: a batch never contains a slot that failed for its size
            raise ValueError(f"{len(big)} background(s) (first: pool index {int(big[0])}, {tuple(int(v) for v in self.ctx.bg_hw[big[0]])}) need an "
                             f"INTER_AREA reduction beyond x{int(self.ctx.MAX_BG_AREA_SCALE)} at x_size_hw={self.x_size_hw} for some rotations; "
                             "downscale them before filling the pool")

    # ------------------------------------------------------------------ reference surface
    def __iter__(self):
        while True:
            yield self.random_tensor_batch()

    def set_batch_size(self, batch_size):
        self.default_batch_size = batch_size

    @classmethod
    def from_hparams(cls, hparams, **kw):
        return cls(
            default_batch_size=hparams.batch_size,
            predownload=getattr(hparams, "force_download", False),
            paired=hparams.loss_contrastive is not None or hparams.loss_set_contrastive is not None,
            targets=hparams.loss_recon is not None,
            x_size_hw=hparams.x_size_hw,
            y_size_hw=hparams.y_size_hw,
            half_upsidedown=hparams.half_upsidedown,
            target_is_input_prob=hparams.target_is_input_prob,
            similar_neg_prob=hparams.similar_neg_prob,
            check_data=getattr(hparams, "check_data", False),
            **kw,
        )

    def random_tensor_batch(self, n: int | None = None) -> dict:
        """Device tensors: x, x2, y [n,3,H,W] out_dtype; x_labels, x2_labels [n,3] int64."""
        if n is None:
            n = self.default_batch_size
        return self._generate(n, cards=None, t_prob=None, n_prob=None)

    def random_image_batch(self, n: int | None = None) -> dict:
        """numpy NHWC float32 + int labels, like the reference's BatchHintNumpy."""
        return self._to_numpy(self.random_tensor_batch(n))

    def image_batch_by_ids(self, ids, *, force_target_input: bool = False, force_similar_neg: bool = False) -> dict:
        if isinstance(ids, (str, uuid.UUID)):
            ids = [ids]
        # reference: t = 1.0 if force else (None if force is None else 0.0), then `t or default`
        # (encoder_train.py:133-134,178,217): False -> 0.0 -> falls back to the default probability
        t = 1.0 if force_target_input else None
        n = 1.0 if force_similar_neg else None
        cards = torch.tensor([self.mtg.get_card_by_id(i).index for i in ids], dtype=torch.int32)
        return self._to_numpy(self._generate(len(ids), cards=cards, t_prob=t, n_prob=n))

    def host_tensor_batch(self, card_images: torch.Tensor, bg_images: torch.Tensor) -> dict:
        """Host buffers in, host buffers out: the `_make_image_batch(cards, bg_imgs)` call of the
        reference (encoder_train.py:189-230) for callers that keep their images on the host.

        card_images (n,H,W,3) / bg_images (n,h,w,3): uint8 CPU tensors (pinned for full copy
        speed) of the pools' image sizes.  They are copied into pool slots [0, n), sample i
        uses card i and background i (x2: a random background of the batch, as in the
        reference), and the finished batch is copied back into pinned host tensors.
        Labels are those of pool slots [0, n).  Hard-negative swaps are OFF on this path (similar_neg_prob is
        forced to 0): the same-name groups describe the resident pool, not the caller's images, so a swap would
        read an unrelated (and possibly concurrently re-uploaded) slot."""
        n = card_images.shape[0]
        assert bg_images.shape[0] == n
        ctx = self.ctx
        with torch.cuda.device(ctx.device):
            if getattr(self, "_stage_n", 0) != n:
                self._stage_cards = torch.empty(card_images.shape, dtype=torch.uint8, device=ctx.device)
                self._stage_bgs = torch.empty(bg_images.shape, dtype=torch.uint8, device=ctx.device)
                self._stage_idx = torch.arange(n, dtype=torch.int32, device=ctx.device)
                self._stage_n = n
                self._host_out = {}
            self._stage_cards.copy_(card_images, non_blocking=True)
            self._stage_bgs.copy_(bg_images, non_blocking=True)
            ctx.update_card_images(self._stage_cards, 0)
            ctx.update_bg_images(self._stage_bgs, 0)
            batch = self._generate(n, cards=self._stage_idx, t_prob=None, n_prob=0.0, bgs=self._stage_idx)
            out = {}
            for k, v in batch.items():
                h = self._host_out.get(k)
                if h is None or h.shape != v.shape or h.dtype != v.dtype:
                    h = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                    self._host_out[k] = h
                h.copy_(v, non_blocking=True)
                out[k] = h
            torch.cuda.current_stream(ctx.device).synchronize()
        return out

    def prepare_jpeg_batch(self, card_files: list[bytes], bg_files: list[bytes], bg_hw: Tuple[int, int] | None = None) -> dict:
        """One batch's inputs as JPEG files (n cards of the pool's card size, n backgrounds of one size) for
        `host_tensor_batches`: the `bytes` objects are copied back to back into pinned memory by a few host threads
        (mtgv_gather_files).  `bg_hw`: the backgrounds' common frame size (parsed from the first background file when
        omitted); the device decoder checks every file's header against the sizes, so no per-file host pass is needed."""
        n = len(card_files)
        assert len(bg_files) == n and n > 0
        if bg_hw is None:
            bg_hw = self.ctx.jpeg_info(bg_files[0])
        hw = np.empty((2 * n, 2), dtype=np.int32)
        hw[:n] = self.ctx.card_hw
        hw[n:] = bg_hw
        return self._check_jpeg_item(self.ctx.prepare_jpegs(list(card_files) + list(bg_files), hw=hw), n)

    def prepare_jpeg_batch_pinned(self, blob: torch.Tensor, file_off, bg_hw: Tuple[int, int] | None = None) -> dict:
        """`prepare_jpeg_batch` without the copy: `blob` is a (pinned) uint8 CPU tensor that already holds the 2n files
        back to back - n cards then n backgrounds, file i = blob[file_off[i]:file_off[i+1]] - e.g. the arena a loader
        `readinto`s.  Cards must have the pool's card size; `bg_hw` is the backgrounds' common frame size (parsed
        from the first background file when omitted).  The device decoder checks every file's header against these sizes,
        so no per-file host pass is needed here."""
        file_off = np.ascontiguousarray(file_off, dtype=np.int64)
        n2 = len(file_off) - 1
        assert n2 > 0 and n2 % 2 == 0
        n = n2 // 2
        if bg_hw is None:
            bg_hw = tuple(int(v) for v in self.ctx.jpeg_info_batch(blob, file_off[n:n + 2])[0])
        hw = np.empty((n2, 2), dtype=np.int32)
        hw[:n] = self.ctx.card_hw
        hw[n:] = bg_hw
        return self._check_jpeg_item(self.ctx.prepare_jpeg_blob(blob, file_off, hw=hw), n)

    @staticmethod
    def lookahead(thunks):
        """Items for `host_tensor_batches` prepared one ahead on a helper thread: `thunks` yields zero-argument callables
        (e.g. `lambda: ds.prepare_jpeg_batch(cards, bgs)`); callable i+1 runs while item i is being consumed (the copy
        into pinned staging releases the GIL).  Exactly one ahead: `prepare_jpegs` rotates three staging buffers, and two
        batches are in flight in the pipeline."""
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(1) as ex:
            fut = None
            for th in thunks:
                nxt = ex.submit(th)
                if fut is not None:
                    yield fut.result()
                fut = nxt
            if fut is not None:
                yield fut.result()

    def _check_jpeg_item(self, b: dict, n: int) -> dict:
        ch, cw = (int(v) for v in b["hw"][0])
        bh, bw = (int(v) for v in b["hw"][n])
        if not (b["hw"][:n] == (ch, cw)).all() or not (b["hw"][n:] == (bh, bw)).all() or (ch, cw) != tuple(self.ctx.card_hw):
            raise ValueError("prepare_jpeg_batch: cards must have the pool's card size and backgrounds one common size")
        b.update(n_pairs=n, card_shape=(n, ch, cw, 3), bg_shape=(n, bh, bw, 3))
        return b

    def host_tensor_batches(self, source):
        """Streaming form of `host_tensor_batch`: `source` yields `(card_images, bg_images)` pairs of
        uint8 CPU tensors (pinned for full copy speed), one pair per batch, all of one batch size n;
        the generator yields one dict of pinned host tensors per pair, in order.

        An item may also be an int n: a batch of n pairs drawn from the RESIDENT pools exactly like
        `random_tensor_batch(n)` (nothing is uploaded; kernels of batch i overlap the download of batch i-1) - the
        training-loop form for callers that want the batches in host memory.

        An item may also be a dict from `prepare_jpeg_batch(card_files, bg_files)`: the batch's inputs as the JPEG
        FILES the reference's loaders read (`_load_card_image`, `IlsvrcImages._load_image` -> imread_float); only the
        compressed bytes cross PCIe and the files are decoded on the device into the staging buffers.

        Four CUDA streams overlap the upload of batch i+1 (and its conversion into the pool
        layout), the kernels of batch i and the download of batch i-1 (what the reference's DataLoader workers do with processes,
        encoder_train.py:517-523).  Batches alternate between pool slots [0, n) and [n, 2n), so
        both pools need at least 2n entries.  A yielded dict is valid until the generator is
        advanced again (its buffers are the double-buffered download targets).  Hard-negative swaps are off
        here for the same reason as in `host_tensor_batch`."""
        ctx = self.ctx
        dev = ctx.device
        with torch.cuda.device(dev):
            if getattr(self, "_pipe_streams", None) is None:
                # persistent: the caching allocator keeps one block pool per stream, new streams would cudaMalloc again
                self._pipe_streams = tuple(torch.cuda.Stream(dev) for _ in range(4))
            s_in, s_pl, s_k, s_out = self._pipe_streams  # upload | pool ingest | kernels | download
            main = torch.cuda.current_stream(dev)
            for s in (s_in, s_pl, s_k, s_out):
                s.wait_stream(main)
            slots = getattr(self, "_pipe_slots", [])  # staging + pinned buffers persist across calls
            pending = []
            i = 0
            for item in source:
                resident = isinstance(item, int)
                if resident:
                    jpeg = None
                    n, card_shape, bg_shape = item, ("resident", item), None
                elif isinstance(item, dict):
                    jpeg = item
                    n, card_shape, bg_shape = item["n_pairs"], item["card_shape"], item["bg_shape"]
                else:
                    jpeg = None
                    card_images, bg_images = item
                    n, card_shape, bg_shape = card_images.shape[0], tuple(card_images.shape), tuple(bg_images.shape)
                    assert bg_images.shape[0] == n
                if not slots or slots[0]["shapes"] != (card_shape, bg_shape):
                    slots = self._pipe_slots = []
                    if not resident and (2 * n > len(self.mtg.pool) or 2 * n > len(self.ilsvrc)):
                        raise ValueError("host_tensor_batches needs pools of at least 2 * batch entries")
                    for j in range(2):
                        slots.append({
                            "shapes": (card_shape, bg_shape), "cards": None, "bgs": None,
                            "idx": torch.arange(j * n, (j + 1) * n, dtype=torch.int32, device=dev),
                            "k_done": torch.cuda.Event(), "copy_done": torch.cuda.Event(), "in_done": torch.cuda.Event(),
                            "out_done": torch.cuda.Event(),
                            "host": {}, "dev": None,
                        })
                if jpeg is None and not resident and slots[i % 2]["cards"] is None:  # device staging for uploaded arrays (files decode straight into the pools)
                    slots[i % 2]["cards"] = torch.empty(card_shape, dtype=torch.uint8, device=dev)
                    slots[i % 2]["bgs"] = torch.empty(bg_shape, dtype=torch.uint8, device=dev)
                sl = slots[i % 2]
                if resident:
                    # nothing to upload: the kernels draw from the whole resident pools like random_tensor_batch
                    s_k.wait_event(sl["out_done"])
                    with torch.cuda.stream(s_k):
                        batch = self._generate(n, cards=None, t_prob=None, n_prob=None)
                        sl["k_done"].record(s_k)
                else:
                    # upload + pool ingest of this batch may start once the kernels that last read these slots are done
                    s_in.wait_event(sl["k_done"])
                    with torch.cuda.stream(s_in):
                        if jpeg is not None:
                            # file bytes up, Huffman / IDCT / colour kernels on this stream, pixels written in the pools' own layouts
                            ctx.decode_into_pools(jpeg, n, (i % 2) * n, n, (i % 2) * n)
                            sl["in_done"].record(s_in)
                        else:
                            sl["cards"].copy_(card_images, non_blocking=True)
                            sl["bgs"].copy_(bg_images, non_blocking=True)
                        sl["copy_done"].record(s_in)
                    if jpeg is None:
                        # the layout conversion into the pools runs on its own stream so the next upload starts right away
                        s_pl.wait_event(sl["copy_done"])
                        with torch.cuda.stream(s_pl):
                            ctx.update_card_images(sl["cards"], (i % 2) * n)
                            ctx.update_bg_images(sl["bgs"], (i % 2) * n)
                            sl["in_done"].record(s_pl)
                    s_k.wait_event(sl["in_done"])
                    s_k.wait_event(sl["out_done"])  # the previous download from this slot's device batch is finished
                    with torch.cuda.stream(s_k):
                        batch = self._generate(n, cards=sl["idx"], t_prob=None, n_prob=0.0, bgs=sl["idx"])
                        sl["k_done"].record(s_k)
                s_out.wait_event(sl["k_done"])
                with torch.cuda.stream(s_out):
                    out = {}
                    for k, v in batch.items():
                        v.record_stream(s_out)
                        h = sl["host"].get(k)
                        if h is None or h.shape != v.shape or h.dtype != v.dtype:
                            h = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                            sl["host"][k] = h
                        h.copy_(v, non_blocking=True)
                        out[k] = h
                    sl["out_done"].record(s_out)
                sl["dev"] = batch
                pending.append((sl, out))
                i += 1
                if len(pending) == 2:
                    done, res = pending.pop(0)
                    done["out_done"].synchronize()
                    yield res
            for done, res in pending:
                done["out_done"].synchronize()
                yield res
            main.wait_stream(s_out)

    # ------------------------------------------------------------------ internals
    def _next_first_index(self, n: int) -> int:
        """Disjoint global sample indices per (rank, call): a running cursor, so a changing n (set_batch_size,
        image_batch_by_ids) never revisits an index (mtgvision_b200/shards.py)."""
        return self._shards.next_first(n)

    def _generate(self, n: int, cards, t_prob, n_prob, bgs=None) -> dict:
        ctx = self.ctx
        with torch.cuda.device(ctx.device):
            first = self._next_first_index(n)
            tape = ctx.sample_encoder_tape(self.seed, first, n, cards=cards, bgs=bgs, target_is_input_prob=t_prob,
                                           similar_neg_prob=n_prob)
            params, labels = ctx.expand_params(tape)
            imgs = ctx.encoder_batch(params, self.out_dtype)
            out = {}
            if self.targets:
                # y = make_cropped of the x card (encoder_train.py:171-173, 199-201)
                base = tape.view(torch.int32)[:n, 1]  # mtgv_enc_tape.card
                out["y"] = ctx.encoder_targets(base, self.out_dtype)
            out["x"] = imgs[:n]
            out["x_labels"] = labels[:n]
            if self.paired:
                out["x2"] = imgs[n:]
                out["x2_labels"] = labels[n:]
            if self.check_data:
                st = params.view(torch.int32)[:, abi.PARAMS_DTYPE.fields["status"][1] // 4]
                if bool((st != 0).any()):
                    raise abi.MtgvError("a sample failed parameter expansion (size limits: see DESIGN.md)")
        return out

    @staticmethod
    def _to_numpy(batch: dict) -> dict:
        out = {}
        for k, v in batch.items():
            if v.ndim == 4:
                a = v.permute(0, 2, 3, 1).contiguous()
                a = a.float() / 255.0 if a.dtype == torch.uint8 else a.float()
                out[k] = a.cpu().numpy()
            else:
                out[k] = v.cpu().numpy()
        return out


class MtgDataModule:
    """Shape of the reference's LightningDataModule (encoder_train.py:504-523) without the
    pytorch_lightning dependency: `train_dataloader()` returns a DataLoader that yields the
    dataset's pre-formed device batches."""

    def __init__(self, train_dataset: RanMtgEncDecDataset, num_workers: int = 0, batch_size: int | None = None):
        if num_workers != 0:
            raise ValueError("the GPU generator runs in-process: num_workers must be 0")
        self.train_dataset = train_dataset
        if batch_size is not None:
            train_dataset.set_batch_size(batch_size)

    def train_dataloader(self):
        from torch.utils.data import DataLoader

        return DataLoader(self.train_dataset, batch_size=None, shuffle=False, num_workers=0)
