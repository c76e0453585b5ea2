"""Synthetic card / background pools and Scryfall-style metadata.

The reference pulls ~1e5 card faces from Scryfall through `mtgdata`
(mtgvision/encoder_datasets.py:548-575) and backgrounds from ILSVRC/COCO directories
(:421-494).  Neither exists offline, and image decode is out of scope for this path
(SURVEY.md section 8f), so benchmarks and parity tests run on resident uint8 pools
generated here:

  * card k  : `default_rng(1000+k)` uniform noise blended 50/50 with a smooth RGB
              gradient, shape (680, 488, 3) uint8 ("normal" Scryfall size);
  * bg j    : low-pass noise from `default_rng(5000+j)`, shape (375, 500, 3) uint8;
  * metadata: id = pool order, names in groups cycling [1,1,2,1,3,1,1,5], set = k % 64.
"""

from __future__ import annotations

import dataclasses

import numpy as np

CARD_HW = (680, 488)
BG_HW = (375, 500)
_GROUP_CYCLE = (1, 1, 2, 1, 3, 1, 1, 5)


@dataclasses.dataclass(frozen=True)
class CardFace:
    """The three fields of mtgdata's ScryfallCardFace the path reads
    (encoder_datasets.py:569-575, 586-590)."""

    id: str
    name: str
    set_code: str
    index: int  # position in the pool


def synth_card(k: int, hw=CARD_HW) -> np.ndarray:
    h, w = hw
    rng = np.random.default_rng(1000 + k)
    noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8).astype(np.uint16)
    yy = np.linspace(0.0, 1.0, h, dtype=np.float32)[:, None]
    xx = np.linspace(0.0, 1.0, w, dtype=np.float32)[None, :]
    ph = rng.random(3, dtype=np.float32)
    grad = np.stack(
        [
            255.0 * (0.5 + 0.5 * np.sin(6.2831853 * (yy * (1 + ph[0]) + xx * ph[1]))),
            255.0 * (yy * ph[1] + xx * (1 - ph[1])) * np.ones_like(yy * xx),
            255.0 * (0.5 + 0.5 * np.cos(6.2831853 * (xx * (1 + ph[2]) - yy * ph[0]))),
        ],
        axis=-1,
    ).astype(np.uint16)
    return ((noise + grad + 1) >> 1).astype(np.uint8)


def synth_bg(j: int, hw=BG_HW) -> np.ndarray:
    h, w = hw
    rng = np.random.default_rng(5000 + j)
    # low-pass noise: coarse grid upsampled by pixel repetition then box-smoothed
    gh, gw = (h + 15) // 16 + 1, (w + 15) // 16 + 1
    coarse = rng.random((gh, gw, 3), dtype=np.float32)
    up = np.repeat(np.repeat(coarse, 16, axis=0), 16, axis=1)[: h + 8, : w + 8]
    c = np.cumsum(np.cumsum(up, axis=0), axis=1)
    c = np.pad(c, ((1, 0), (1, 0), (0, 0)))
    box = (c[8:, 8:] - c[:-8, 8:] - c[8:, :-8] + c[:-8, :-8]) / 64.0
    fine = rng.random((h, w, 3), dtype=np.float32) * 0.15
    img = np.clip(box[:h, :w] * 0.85 + fine, 0, 1)
    return (img * 255.0 + 0.5).astype(np.uint8)


class CardPool:
    """Resident card images + metadata with the reference's label/group semantics.

    labels3[k] = (id_idx, name_idx, set_idx), the rank of each key in sorted(set(keys))
    as `idx_map` computes it (encoder_datasets.py:507-512, 579-581).
    `group_of(k)` = pool indices sharing card k's name, in insertion order, which is what
    `_cards_by_name[name]` holds (:570).
    """

    def __init__(self, images: np.ndarray, faces: list[CardFace]):
        assert images.dtype == np.uint8 and images.ndim == 4 and images.shape[-1] == 3
        assert len(faces) == len(images)
        self.images = images
        self.faces = faces
        ids = sorted({f.id for f in faces})
        names = sorted({f.name for f in faces})
        sets = sorted({f.set_code for f in faces})
        id_rank = {v: i for i, v in enumerate(ids)}
        name_rank = {v: i for i, v in enumerate(names)}
        set_rank = {v: i for i, v in enumerate(sets)}
        self.labels3 = np.asarray(
            [(id_rank[f.id], name_rank[f.name], set_rank[f.set_code]) for f in faces], dtype=np.int32
        )
        groups: dict[str, list[int]] = {}
        for k, f in enumerate(faces):
            groups.setdefault(f.name, []).append(k)
        self._groups = groups
        # CSR form for the device: members of card k's group are grp_mem[grp_off[k]:grp_off[k+1]]
        off, mem = [0], []
        for f in faces:
            mem.extend(groups[f.name])
            off.append(len(mem))
        self.grp_off = np.asarray(off, dtype=np.int32)
        self.grp_mem = np.asarray(mem, dtype=np.int32)
        # `_card_ids` is sorted(ids) (:577); ran_card draws an index into it
        self.sorted_to_pool = np.asarray(
            [k for _, k in sorted((f.id, k) for k, f in enumerate(faces))], dtype=np.int32
        )
        self._by_id = {f.id: k for k, f in enumerate(faces)}

    def __len__(self):
        return len(self.faces)

    def group_of(self, k: int) -> list[int]:
        return self._groups[self.faces[k].name]

    def index_of(self, id_: str) -> int:
        return self._by_id[str(id_)]


def synth_faces(n: int) -> list[CardFace]:
    faces = []
    g, left = 0, _GROUP_CYCLE[0]
    for k in range(n):
        if left == 0:
            g += 1
            left = _GROUP_CYCLE[g % len(_GROUP_CYCLE)]
        left -= 1
        faces.append(CardFace(id=f"{k:08d}-0000-4000-8000-000000000000", name=f"card-name-{g:06d}",
                              set_code=f"s{k % 64:02d}", index=k))
    return faces


def _card_chunk(args):
    k0, k1, hw = args
    return np.stack([synth_card(k, hw) for k in range(k0, k1)])


def _bg_chunk(args):
    j0, j1, hw = args
    return [synth_bg(j, hw) for j in range(j0, j1)]


def _chunks(n: int, workers: int):
    step = max(1, (n + workers * 4 - 1) // (workers * 4))
    return [(a, min(a + step, n)) for a in range(0, n, step)]


def make_card_pool(n: int, hw=CARD_HW, workers: int = 1) -> CardPool:
    """`workers > 1` generates the images in a process pool (bench setup only)."""
    images = np.empty((n, hw[0], hw[1], 3), dtype=np.uint8)
    if workers > 1 and n >= 64:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(workers) as p:
            spans = _chunks(n, workers)
            for (a, b), part in zip(spans, p.imap(_card_chunk, [(a, b, hw) for a, b in spans])):
                images[a:b] = part
    else:
        for k in range(n):
            images[k] = synth_card(k, hw)
    return CardPool(images, synth_faces(n))


def make_bg_pool(n: int, hw=BG_HW, workers: int = 1) -> list[np.ndarray]:
    if workers > 1 and n >= 64:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(workers) as p:
            spans = _chunks(n, workers)
            out: list[np.ndarray] = []
            for part in p.imap(_bg_chunk, [(a, b, hw) for a, b in spans]):
                out.extend(part)
            return out
    return [synth_bg(j, hw) for j in range(n)]
