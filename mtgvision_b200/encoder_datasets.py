"""Drop-in for `mtgvision/encoder_datasets.py` (reference) on the B200 path.

Same class and method names, argument meaning and error behaviour as the reference's
`SyntheticBgFgMtgImages` (encoder_datasets.py:515-834) and `IlsvrcImages` (:421-478);
pixels come from libmtgv.so.  What differs, by design:

  * images live in resident uint8 pools (HBM) instead of being decoded from disk per
    sample - `mtgdata`/ILSVRC downloads are out of scope, synthetic pools stand in;
  * float inputs handed to the static `make_*` helpers are quantised to uint8, the pool
    format (exact for anything that came from an 8-bit image);
  * augment magnitudes are drawn on the device from Philox streams keyed by Python's
    `random` state (`random.getrandbits`) so `random.seed(s)` still makes runs repeatable.
"""

from __future__ import annotations

import os
import random
import uuid
import warnings
from math import ceil
from typing import Iterator, Literal, Optional

import numpy as np
import torch

from . import abi, synth
from .context import Context
from .synth import CardFace, CardPool

SizeHW = tuple[int, int]
PathOrImg = "str | np.ndarray"


def _as_u8_image(img) -> np.ndarray:
    if isinstance(img, (str, os.PathLike)):
        # `path_or_img` (encoder_datasets.py:741, 759, 778, 796 -> uimg.imread_float, util/image.py:107-114): the file is read on the
        # host and decoded by the device decoder (bit-exact with cv2.imread for baseline JPEG; no CPU decode path exists here)
        path = os.fspath(img)
        try:
            with open(path, "rb") as f:
                data = f.read()
        except OSError:
            raise Exception("Image not found: {}".format(path)) from None  # imread_float's message (util/image.py:112-113)
        flat, _, hw = _StaticEngine.get().decode_jpegs([data])
        return flat.cpu().numpy().reshape(int(hw[0, 0]), int(hw[0, 1]), 3)
    a = np.asarray(img)
    if a.ndim != 3 or a.shape[2] != 3:
        raise ValueError(f"expected an (H, W, 3) image, got shape {a.shape}")
    if a.dtype == np.uint8:
        return np.ascontiguousarray(a)
    if a.dtype in (np.float16, np.float32, np.float64):
        return np.ascontiguousarray(np.rint(np.clip(a.astype(np.float32), 0, 1) * 255.0).astype(np.uint8))
    raise Exception(f"Unsupported Numpy Type: {a.dtype}")  # util/image.py:235


def _u8_to_f32(img_u8: np.ndarray) -> np.ndarray:
    """img_float32 (util/image.py:220-237) - host-side dtype conversion only."""
    return np.clip(np.divide(img_u8, 255.0, dtype=np.float32), 0, 1)


class IlsvrcImages:
    """Background image source with the reference's interface (encoder_datasets.py:421-478)
    over a resident pool of uint8 images, or - like the reference - over a directory of JPEG files
    (`root`, `subdir`), which are decoded on the device straight into the pool.  Files the
    decoder rejects (arithmetic coding, CMYK, 12-bit ...) raise `MtgvError`, or are left out with a warning
    when `skip_unsupported=True`."""

    _EXTS = (".jpeg", ".jpg")

    def __init__(self, images: Optional[list[np.ndarray]] = None, n: int = 64, *, root=None, subdir="val",
                 files: Optional[list[bytes]] = None, skip_unsupported: bool = False):
        self._images = self.jpeg_files = None
        if root is not None or files is not None:
            if files is None:
                root = os.path.join(str(root), str(subdir)) if subdir is not None else str(root)
                if subdir is not None and os.path.isabs(str(subdir)):
                    raise ValueError("subdir must be a relative path")
                paths = sorted(os.path.join(d, f) for d, _, fs in os.walk(root) for f in fs if f.lower().endswith(self._EXTS))
                assert len(paths) > 0, f"Dataset is empty. Please download the dataset. {root}"
                self._paths = paths
                files = [open(p, "rb").read() for p in paths]
            else:
                self._paths = [f"bg://{j:06d}" for j in range(len(files))]
            if skip_unsupported:
                # ILSVRC holds a few CMYK files; the decoder takes Huffman-coded YCbCr / grey files only and has no fallback:
                # leave them out of the background pool (with a warning) instead of failing the whole ingest
                keep = []
                for j, f in enumerate(files):
                    try:
                        _StaticEngine.get().jpeg_info(f)
                        keep.append(j)
                    except abi.MtgvError as e:
                        warnings.warn(f"background {self._paths[j]} skipped: {e}")
                files = [files[j] for j in keep]
                self._paths = [self._paths[j] for j in keep]
            assert len(files) > 0, "Dataset is empty."
            self.jpeg_files = list(files)
            return
        self._images = images if images is not None else synth.make_bg_pool(n)
        assert len(self._images) > 0, "Dataset is empty."
        self._paths = [f"bg://{j:06d}" for j in range(len(self._images))]

    def __len__(self):
        return len(self._paths)

    def _image_u8(self, item) -> np.ndarray:
        if self._images is not None:
            return self._images[item]
        flat, _, hw = _StaticEngine.get().decode_jpegs([self.jpeg_files[item]])  # imread_float's cv2.imread, on the device
        return flat.cpu().numpy().reshape(int(hw[0, 0]), int(hw[0, 1]), 3)

    def __getitem__(self, item):
        return _u8_to_f32(self._image_u8(item))

    def __iter__(self):
        for j in range(len(self)):
            yield self[j]

    def ran_index(self) -> int:
        return random.randrange(len(self))

    def ran_path(self) -> str:
        return random.choice(self._paths)

    def ran(self) -> np.ndarray:
        return self[self.ran_index()]

    def get(self, idx) -> np.ndarray:
        return self[idx]

    @property
    def images_u8(self) -> list[np.ndarray]:
        return self._images if self._images is not None else [self._image_u8(j) for j in range(len(self))]

    def fill_pool(self, ctx: Context) -> None:
        """Make this source the background pool of `ctx` (file-backed sources never touch a host pixel)."""
        if self.jpeg_files is not None:
            ctx.set_bg_pool_from_jpegs(self.jpeg_files)
        else:
            ctx.set_bg_pool(self._images)


class CocoValImages(IlsvrcImages):
    pass


class _StaticEngine:
    """One lazily created context per device for the static make_* helpers, with
    single-image pools re-uploaded per call."""

    _by_device: dict[int, Context] = {}

    @classmethod
    def get(cls) -> Context:
        dev = torch.cuda.current_device() if torch.cuda.is_available() else 0
        if dev not in cls._by_device:
            cls._by_device[dev] = Context(dev)
        return cls._by_device[dev]


def _single_pool(ctx: Context, card_u8: np.ndarray, bg_u8: Optional[np.ndarray], size_hw: SizeHW):
    ctx.set_encoder_config(x_size_hw=size_hw, y_size_hw=size_hw, target_is_input_prob=0.0, similar_neg_prob=0.0,
                           half_upsidedown=False, paired=False, targets=False)
    one = np.zeros((1, 3), dtype=np.int32)
    ctx.set_card_pool(card_u8[None], one, np.asarray([0, 1], dtype=np.int32), np.asarray([0], dtype=np.int32))
    if bg_u8 is not None:
        ctx.set_bg_pool([bg_u8])


class SyntheticBgFgMtgImages:
    """Card images + metadata (reference encoder_datasets.py:515-834) over a `CardPool`."""

    def __init__(self, img_type: str = "small", bulk_type: str = "default_cards", predownload: bool = False,
                 force_update: bool = False, *, pool: Optional[CardPool] = None, n_cards: int = 64):
        self.img_type, self.bulk_type = img_type, bulk_type
        self.pool = pool if pool is not None else synth.make_card_pool(n_cards)
        faces = self.pool.faces
        self._card_by_id = {f.id: f for f in faces}
        self._card_ids = sorted(f.id for f in faces)
        self._cards_by_name: dict[str, dict[str, CardFace]] = {}
        self._cards_by_set: dict[str, dict[str, CardFace]] = {}
        for f in faces:
            self._cards_by_name.setdefault(f.name, {})[f.id] = f
            self._cards_by_set.setdefault(f.set_code, {})[f.id] = f

    # ---- labels / lookup (encoder_datasets.py:586-630) ----
    def card_get_labels(self, card: CardFace) -> tuple[int, int, int]:
        a, b, c = self.pool.labels3[card.index]
        return int(a), int(b), int(c)

    def card_get_labels_by_id(self, id_) -> tuple[int, int, int]:
        return self.card_get_labels(self.get_card_by_id(id_))

    def get_card_by_id(self, id_) -> CardFace:
        return self._card_by_id[str(id_)]

    def get_image_by_id(self, id_) -> np.ndarray:
        return self._load_card_image(self.get_card_by_id(id_))

    def get_card_and_image_by_id(self, id_):
        card = self.get_card_by_id(id_)
        return card, self._load_card_image(card)

    def _get_group(self, card, mode):
        if mode == "name":
            return self._cards_by_name[card.name]
        elif mode == "set":
            return self._cards_by_name[card.set_code]  # reference quirk kept (encoder_datasets.py:615)
        raise KeyError

    def get_similar_card(self, id_, mode: Literal["name", "set"] = "name") -> Optional[CardFace]:
        card = self.get_card_by_id(id_)
        group = dict(self._get_group(card, mode=mode))
        assert card.id in group
        group.pop(card.id)
        if group:
            return random.choice(list(group.values()))
        return None

    def _load_card_image(self, card: CardFace) -> np.ndarray:
        return _u8_to_f32(self.pool.images[card.index])

    def __len__(self):
        return len(self._card_ids)

    def __getitem__(self, item):
        return self._load_card_image(self.get_card_by_id(self._card_ids[item]))

    def __iter__(self):
        for card in self.card_iter():
            yield self._load_card_image(card)

    def card_iter(self) -> Iterator[CardFace]:
        for id_ in self._card_ids:
            yield self.get_card_by_id(id_)

    def ran(self) -> np.ndarray:
        return self.ran_card_and_image()[1]

    def ran_card_and_image(self):
        card = self.ran_card()
        return card, self._load_card_image(card)

    def ran_path(self) -> str:
        return f"card://{self.ran_card().id}"

    def ran_card(self) -> CardFace:
        return self.get_card_by_id(random.choice(self._card_ids))

    def get(self, idx) -> np.ndarray:
        return self[idx]

    # ---- sample synthesis (encoder_datasets.py:733-834), numpy HWC float32 in / out ----
    @staticmethod
    def _run_single(ctx: Context, tape_np: np.ndarray) -> np.ndarray:
        tape = ctx.upload_tape(tape_np)
        params, _ = ctx.expand_params(tape, want_labels=False)
        out = ctx.encoder_batch(params, abi.OUT_F32)
        p = params.cpu().numpy().view(abi.PARAMS_DTYPE).reshape(-1)
        if int(p["status"][0]) != 0:
            raise abi.MtgvError(f"sample expansion failed with status {int(p['status'][0])} (size limits: see DESIGN.md)")
        return out[0].permute(1, 2, 0).contiguous().cpu().numpy()

    @staticmethod
    def make_cropped(path_or_img, size_hw: SizeHW | None = None, half_upsidedown: bool = False) -> np.ndarray:
        card = _as_u8_image(path_or_img)
        if size_hw is None:
            border = ceil(max(0.02 * card.shape[0], 0.02 * card.shape[1]))
            ret = _u8_to_f32(card[border : card.shape[0] - border, border : card.shape[1] - border, :])
            return np.rot90(ret, k=2) if (half_upsidedown and random.randrange(2) == 0) else ret
        ctx = _StaticEngine.get()
        _single_pool(ctx, card, None, tuple(size_hw))
        tape = np.zeros(1, dtype=abi.TAPE_DTYPE)
        tape[0]["kind"] = abi.KIND_CROPPED
        tape[0]["swap_choice"] = -1
        tape[0]["upsidedown"] = int(half_upsidedown and random.randrange(2) == 0)
        return SyntheticBgFgMtgImages._run_single(ctx, tape)

    @classmethod
    def make_masked(cls, path_or_img) -> np.ndarray:
        card = _as_u8_image(path_or_img)
        ctx = _StaticEngine.get()
        _single_pool(ctx, card, None, (192, 128))
        mask = ctx.get_mask("encoder").cpu().numpy()
        return np.concatenate([_u8_to_f32(card), mask[:, :, None]], axis=2)

    @classmethod
    def _virtual_like(cls, card_u8, bg_u8, size_hw, half_upsidedown, kind) -> np.ndarray:
        ctx = _StaticEngine.get()
        _single_pool(ctx, card_u8, bg_u8, tuple(size_hw))
        ctx.cfg.half_upsidedown = int(half_upsidedown)
        ctx.set_encoder_config(x_size_hw=size_hw, y_size_hw=size_hw, target_is_input_prob=0.0, similar_neg_prob=0.0,
                               half_upsidedown=half_upsidedown, paired=False, targets=False)
        tape = ctx.sample_encoder_tape(random.getrandbits(63), 0, 1)
        if kind != abi.KIND_VIRTUAL:
            t = tape.cpu().numpy().view(abi.TAPE_DTYPE).reshape(-1).copy()
            t["kind"] = kind
            tape = ctx.upload_tape(t)
        params, _ = ctx.expand_params(tape, want_labels=False)
        out = ctx.encoder_batch(params, abi.OUT_F32)
        p = params.cpu().numpy().view(abi.PARAMS_DTYPE).reshape(-1)
        if int(p["status"][0]) != 0:
            raise abi.MtgvError(f"sample expansion failed with status {int(p['status'][0])} (size limits: see DESIGN.md)")
        return out[0].permute(1, 2, 0).contiguous().cpu().numpy()

    @classmethod
    def make_bg(cls, bg_path_or_img, size_hw: SizeHW) -> np.ndarray:
        bg = _as_u8_image(bg_path_or_img)
        dummy = np.zeros((max(size_hw[0], 16), max(size_hw[1], 16), 3), dtype=np.uint8)
        return cls._virtual_like(dummy, bg, size_hw, False, abi.KIND_BG_ONLY)

    @classmethod
    def make_virtual(cls, card_path_or_img, bg_path_or_img, size_hw: SizeHW, half_upsidedown: bool = False) -> np.ndarray:
        card = _as_u8_image(card_path_or_img)
        bg = _as_u8_image(bg_path_or_img)
        virtual = cls._virtual_like(card, bg, size_hw, half_upsidedown, abi.KIND_VIRTUAL)
        assert virtual.shape[:2] == tuple(size_hw)
        return virtual

    @classmethod
    def make_virtual_pair(cls, card_path_or_img, bg_path_or_img, x_size_hw: SizeHW, y_size_hw: SizeHW,
                          half_upsidedown: bool = False):
        card = _as_u8_image(card_path_or_img)
        x = cls.make_virtual(card, bg_path_or_img, size_hw=x_size_hw, half_upsidedown=half_upsidedown)
        y = cls.make_cropped(card, size_hw=y_size_hw)
        return x, y
