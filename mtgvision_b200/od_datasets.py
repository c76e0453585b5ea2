"""Drop-in for the detection-scene generator of `mtgvision/od_datasets.py` (reference).

`Gen(**kwargs)` keeps the reference constructor arguments (od_datasets.py:620-639),
`random()` / `random_bg()` and the sample dict keys (`image`, `keypoints`,
`keypoints_labels`, :607-611); `save_sample` / `create_yolo_obb_dataset` keep the on-disk YOLO
layout (:732-832).  Scenes are produced by libmtgv.so: a Philox tape sampler, the placement /
label kernel (rejection sampling, overlap tests and keypoint warping on device) and the pixel
kernel (background cover-warp, per-card warp of image + mask, alpha composite, photometrics).

`random_batch(n)` is the GPU-native entry: device tensors `image [n,3,S,S]` (uint8 / fp16 /
fp32 NCHW), padded `keypoints [n,Kmax,P,2] float64` in pixels, `labels [n,Kmax] int32`
(-1 padding) and `counts [n]`.
"""

from __future__ import annotations

import os
import random
import warnings
from pathlib import Path
from typing import Literal, Optional

import numpy as np
import torch

from . import abi, synth
from .context import Context
from .shards import ShardCursor
from .encoder_datasets import CocoValImages, IlsvrcImages, SyntheticBgFgMtgImages

_OUT = {"uint8": abi.OUT_U8, "u8": abi.OUT_U8, "float16": abi.OUT_F16, "fp16": abi.OUT_F16, "float32": abi.OUT_F32,
        "fp32": abi.OUT_F32}


class Gen:
    def __init__(
        self,
        *,
        bg_size_hw: tuple[int, int] | int = 640,
        num_cards_min: int = 1,
        num_cards_max: int = 10,
        card_min_visible_ratio: float = 0.5,
        card_min_visible_ratio_edges: Optional[float] = 1.0,
        card_jitter_ratio: float = 0.3,
        card_min_area_ratio: float = 0.02,
        card_max_area_ratio: float = 0.9,
        card_size_sample_mode: Literal["uniform", "log_uniform"] = "log_uniform",
        card_no_contains: bool = True,
        card_max_place_attempts: int = 10,
        ratio_bg: Optional[float] = None,
        ilsvrc_vs_coco_sample_weights: tuple[float, float] | None = (1.0, 1.0),
        kind: Literal["obb", "seg"] = "obb",
        # ---- B200 path ----
        mtg_ds: Optional[SyntheticBgFgMtgImages] = None,
        bg_ds: Optional[IlsvrcImages] = None,
        bg2_ds: Optional[CocoValImages] = None,
        device: Optional[int] = None,
        seed: Optional[int] = None,
        photometrics: bool = True,
        rank: int = 0,
        world_size: int = 1,
    ):
        if card_size_sample_mode not in ("uniform", "log_uniform"):
            raise KeyError(card_size_sample_mode)  # like place_card_on_background_get_transform (od_datasets.py:333)
        self.mtg_ds = mtg_ds if mtg_ds is not None else SyntheticBgFgMtgImages(img_type="small")
        self.bg_ds = bg_ds if bg_ds is not None else IlsvrcImages()
        self.bg2_ds = bg2_ds
        self.bg_size_hw = bg_size_hw
        self.num_cards_min, self.num_cards_max = num_cards_min, num_cards_max
        self.card_min_visible_ratio = card_min_visible_ratio
        self.card_min_visible_ratio_edges = card_min_visible_ratio_edges
        self.card_jitter_ratio = card_jitter_ratio
        self.card_min_area_ratio, self.card_max_area_ratio = card_min_area_ratio, card_max_area_ratio
        self.card_size_sample_mode = card_size_sample_mode
        self.card_no_contains = card_no_contains
        self.card_max_place_attempts = card_max_place_attempts
        self.ratio_bg = ratio_bg
        self.kind = kind
        self.seed = random.getrandbits(63) if seed is None else int(seed)
        self.rank, self.world_size = int(rank), int(world_size)
        self._shards = ShardCursor(self.rank, self.world_size)
        # both background datasets share one resident pool: bg_ds in slots [0, len(bg_ds)), bg2_ds behind it.  The sampler
        # draws the dataset with ilsvrc_vs_coco_sample_weights first and then an image uniformly inside it, like
        # Gen._get_bg_ds + ran_path (od_datasets.py:656-672); weights=None = proportional to the dataset sizes.
        sources = [self.bg_ds] + ([self.bg2_ds] if self.bg2_ds is not None else [])
        n_first = len(self.bg_ds) if self.bg2_ds is not None else 0
        if self.bg2_ds is None:
            p_first = 1.0
        elif ilsvrc_vs_coco_sample_weights is None:
            p_first = len(self.bg_ds) / float(len(self.bg_ds) + len(self.bg2_ds))
        else:
            w = np.asarray(ilsvrc_vs_coco_sample_weights, dtype=np.float64)
            if w.shape != (2,) or not (w >= 0).all() or w.sum() <= 0:
                raise ValueError("ilsvrc_vs_coco_sample_weights must be two non-negative weights")
            p_first = float((w / np.sum(w))[0])
        self._bg_p = np.asarray([p_first, 1.0 - p_first])
        self.ctx = Context(device)
        pool = self.mtg_ds.pool
        self.ctx.set_card_pool(pool.images, pool.labels3, pool.grp_off, pool.grp_mem)
        if all(getattr(d, "jpeg_files", None) is not None for d in sources):
            # file-backed sources (IlsvrcImages(root=...)): decoded on the device straight into the pool
            self.ctx.set_bg_pool_from_jpegs([f for d in sources for f in d.jpeg_files])
        else:
            self.ctx.set_bg_pool([im for d in sources for im in d.images_u8])
        # raises like the reference's random.randint when the card diagonal does not fit (od_datasets.py:322)
        try:
            self.ctx.set_det_config(bg_size_hw=bg_size_hw, num_cards_min=num_cards_min, num_cards_max=num_cards_max,
                                    card_min_visible_ratio=card_min_visible_ratio,
                                    card_min_visible_ratio_edges=card_min_visible_ratio_edges,
                                    card_jitter_ratio=card_jitter_ratio, card_min_area_ratio=card_min_area_ratio,
                                    card_max_area_ratio=card_max_area_ratio, card_no_contains=card_no_contains,
                                    card_max_place_attempts=card_max_place_attempts, ratio_bg=ratio_bg, kind=kind,
                                    photometrics=photometrics, card_size_sample_mode=card_size_sample_mode,
                                    n_bgs_first=n_first, bg_first_prob=p_first)
        except abi.MtgvError as e:
            if "empty range" in str(e):
                raise ValueError(str(e)) from e
            raise

    # ------------------------------------------------------------------ GPU-native batch
    def random_batch(self, n: int, out_dtype: str = "uint8") -> dict:
        ctx = self.ctx
        with torch.cuda.device(ctx.device):
            first = self._shards.next_first(n)  # running cursor: tail batches and random() never reuse an index
            tape = ctx.sample_det_tape(self.seed, first, n)
            params, accepted, keypoints, labels, counts = ctx.det_place(tape)
            image = ctx.det_batch(params, _OUT[out_dtype])
        return {"image": image, "keypoints": keypoints, "labels": labels, "counts": counts, "accepted": accepted}

    def host_batches(self, sizes, out_dtype: str = "uint8"):
        """`random_batch` for callers that want the scenes in host memory (the dataset writer, a DataLoader consumer):
        `sizes` yields batch sizes, the generator yields dicts of PINNED host tensors (image, keypoints, labels, counts,
        accepted), in order.  The kernels of batch i overlap the download of batch i-1 (two streams, double-buffered pinned
        targets); a yielded dict is valid until the generator is advanced again.  Same scenes as the same sequence of
        `random_batch` calls."""
        dev = self.ctx.device
        with torch.cuda.device(dev):
            if getattr(self, "_pipe_streams", None) is None:
                self._pipe_streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
            s_k, s_out = self._pipe_streams
            main = torch.cuda.current_stream(dev)
            s_k.wait_stream(main)
            s_out.wait_stream(main)
            slots = getattr(self, "_pipe_slots", None)
            if slots is None:
                slots = self._pipe_slots = [{"host": {}, "k_done": torch.cuda.Event(), "out_done": torch.cuda.Event()} for _ in range(2)]
            pending = []
            for i, n in enumerate(sizes):
                sl = slots[i % 2]
                with torch.cuda.stream(s_k):
                    batch = self.random_batch(int(n), out_dtype)
                    sl["k_done"].record(s_k)
                s_out.wait_event(sl["k_done"])
                with torch.cuda.stream(s_out):
                    out = {}
                    for k, v in batch.items():
                        v.record_stream(s_out)
                        h = sl["host"].get(k)
                        if h is None or h.shape != v.shape or h.dtype != v.dtype:
                            h = sl["host"][k] = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                        h.copy_(v, non_blocking=True)
                        out[k] = h
                    sl["out_done"].record(s_out)
                pending.append((sl, out))
                if len(pending) == 2:
                    # the slot of the batch handed out now is rewritten two iterations later, after the caller came back for more
                    done, res = pending.pop(0)
                    done["out_done"].synchronize()
                    yield res
            for done, res in pending:
                done["out_done"].synchronize()
                yield res
            main.wait_stream(s_out)

    # ------------------------------------------------------------------ reference surface
    def _one(self, force_bg: Optional[bool] = None) -> dict:
        b = self.random_batch(1, "float32")
        k = int(b["counts"][0])
        P = 4 if self.kind == "obb" else 8
        return {
            "image": b["image"][0].permute(1, 2, 0).contiguous().cpu().numpy(),
            "keypoints": b["keypoints"][0, :k, :P].cpu().numpy() if k else [],
            "keypoints_labels": b["labels"][0, :k].cpu().numpy().astype(np.int64) if k else [],
        }

    def random(self) -> dict:
        return self._one()

    def random_bg(self) -> dict:
        old = self.ctx.det_cfg.ratio_bg
        try:
            self.ctx.det_cfg.ratio_bg = 2.0  # every draw takes the background-only branch
            self.ctx._check(self.ctx.lib.mtgv_set_det_config(self.ctx._h, abi.C.byref(self.ctx.det_cfg)), "mtgv_set_det_config")
            return self._one()
        finally:
            self.ctx.det_cfg.ratio_bg = old
            self.ctx._check(self.ctx.lib.mtgv_set_det_config(self.ctx._h, abi.C.byref(self.ctx.det_cfg)), "mtgv_set_det_config")


# --------------------------------------------------------------------------------------- #
# YOLO on-disk dataset (od_datasets.py:732-832): host-side writer, same layout               #
# --------------------------------------------------------------------------------------- #


def save_sample(sample: dict, i: int, label_dir, img_dir, ext: Literal["png", "jpg"] = "jpg"):
    import cv2

    img, kps, kps_labels = sample["image"], sample["keypoints"], sample["keypoints_labels"]
    assert len(kps) == len(kps_labels)
    assert img.ndim == 3
    annotations = []
    for pts, label in zip(kps, kps_labels):
        pts = np.asarray(pts, dtype=np.float64) / img.shape[:2][::-1]  # points are (w, h)
        if np.any(pts < 0) or np.any(pts > 1):
            warnings.warn(f"points for image {i} are out of bounds, yolo will consider these invalid")
        annotations.append(f"{label} {' '.join(map(str, pts.flatten()))}")
    with open(os.path.join(label_dir, f"image_{i:04d}.txt"), "w") as f:
        for ann in annotations:
            f.write(ann + "\n")
    u8 = (img * 255).astype(np.uint8) if img.dtype != np.uint8 else img  # imwrite semantics (util/image.py:101-104)
    cv2.imwrite(os.path.join(img_dir, f"image_{i:04d}.{ext}"), cv2.cvtColor(u8, cv2.COLOR_RGB2BGR))


def save_batch(generator: Gen, batch: dict, first_i: int, label_dir, img_dir, quality: int = 95) -> int:
    """`save_sample` for a whole `Gen.random_batch(n, "uint8")`: the label text is formatted on the host exactly like
    `save_sample`, the images are JPEG-encoded on the device (`mtgv_encode_jpeg_batch`: the bytes cv2.imwrite would write)
    and only the compressed files cross PCIe.  Returns the number of samples written."""
    image = batch["image"]
    n, _, H, W = image.shape
    files = generator.ctx.encode_jpegs_host(image, quality)  # views of one pinned buffer, written out below
    counts = batch["counts"].cpu().numpy()
    kps = batch["keypoints"].cpu().numpy()
    labels = batch["labels"].cpu().numpy()
    P = 4 if generator.kind == "obb" else 8
    for k in range(n):
        i = first_i + k
        lines = []
        for pts, label in zip(kps[k, : counts[k], :P], labels[k, : counts[k]].astype(np.int64)):
            pts = np.asarray(pts, dtype=np.float64) / (W, H)  # points are (w, h)
            if np.any(pts < 0) or np.any(pts > 1):
                warnings.warn(f"points for image {i} are out of bounds, yolo will consider these invalid")
            lines.append(f"{label} {' '.join(map(str, pts.flatten()))}")
        with open(os.path.join(label_dir, f"image_{i:04d}.txt"), "w") as f:
            for ann in lines:
                f.write(ann + "\n")
        with open(os.path.join(img_dir, f"image_{i:04d}.jpg"), "wb") as f:
            f.write(files[k])
    return n


def create_yolo_obb_dataset(generator: Gen, *, output_dir: str, num_train: int = 20000, num_val_ratio: float = 0.1,
                            num_test_ratio: float = 0.1, ext: Literal["png", "jpg"] = "jpg", batch: int = 256):
    """od_datasets.py:732-791.  ext="jpg" takes the batched path: scenes are generated AND JPEG-encoded on the GPU,
    `batch` at a time; ext="png" writes one `save_sample` per scene like the reference."""
    import yaml

    output_dir = Path(output_dir)
    if output_dir.exists() and len(list(output_dir.iterdir())) > 0:
        raise FileExistsError(f"not clean... {output_dir}")
    img_dir, label_dir = output_dir / "images", output_dir / "labels"
    img_dir.mkdir(exist_ok=True, parents=True)
    label_dir.mkdir(exist_ok=True, parents=True)
    with open(output_dir / "mtg_obb.yaml", "w") as fp:
        yaml.safe_dump({"path": ".", "train": str(img_dir / "train"), "val": str(img_dir / "val"),
                        "test": str(img_dir / "test"), "names": {0: "card", 1: "card_top", 2: "card_bottom"}}, fp)
    batched = ext == "jpg" and batch > 1
    for name, num in [("train", num_train), ("val", int(num_val_ratio * num_train)), ("test", int(num_test_ratio * num_train))]:
        (img_dir / name).mkdir(exist_ok=True, parents=True)
        (label_dir / name).mkdir(exist_ok=True, parents=True)
        if batched:
            i = 0
            while i < num:
                b = generator.random_batch(min(batch, num - i), "uint8")
                i += save_batch(generator, b, i, label_dir / name, img_dir / name)
            continue
        for i in range(num):
            save_sample(generator.random(), i=i, label_dir=label_dir / name, img_dir=img_dir / name, ext=ext)
