"""Pins the encoder oracle: (a) against the golden vectors generated from the real reference
(tests/golden/make_golden.py), everywhere; (b) against the imported reference itself when
/root/reference is present (build container only).  CPU only."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

from mtgvision_b200 import synth
from oracle import encoder_oracle as EO
from oracle import ref_import

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def pools():
    return [synth.synth_card(k) for k in range(8)], [synth.synth_bg(j) for j in range(8)]


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(GOLD, "encoder_golden.json")))


def _oracle_virtual(cards, bgs, seed, card, bg):
    random.seed(seed)
    np.random.seed(seed)
    EO.reset_shuffle_state()
    return np.ascontiguousarray(EO.make_virtual(EO.u8_to_f32(cards[card]), EO.u8_to_f32(bgs[bg]), (192, 128), True))


def test_oracle_reproduces_reference_golden_virtual(pools, golden):
    cards, bgs = pools
    imgs = np.load(os.path.join(GOLD, "encoder_golden_images.npz"))
    for g in golden["virtual"]:
        img = _oracle_virtual(cards, bgs, g["seed"], g["card"], g["bg"])
        assert hashlib.sha1(img.tobytes()).hexdigest() == g["sha1"], f"seed {g['seed']}"
        key = f"virtual_{g['seed']}"
        if key in imgs:
            assert np.array_equal(np.rint(np.clip(img, 0, 1) * 255).astype(np.uint8), imgs[key])


def test_oracle_reproduces_reference_golden_cropped(pools, golden):
    cards, _ = pools
    for g in golden["cropped"]:
        img = np.ascontiguousarray(EO.make_cropped(EO.u8_to_f32(cards[g["card"]]), (192, 128)))
        assert hashlib.sha1(img.tobytes()).hexdigest() == g["sha1"]


def test_masks_match_reference_golden(golden):
    for key, g in golden["masks"].items():
        hw, ratio = key.split("@")
        h, w = (int(v) for v in hw.split("x"))
        m = EO.round_rect_mask((h, w), float(ratio))
        assert int((m == 0).sum()) == g["zeros"]
        assert hashlib.sha1(m.tobytes()).hexdigest() == g["sha1"]


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not present on this box")
def test_oracle_equals_live_reference(pools):
    """Bit-exact against the imported reference for seeds beyond the golden set, including the
    history-dependent ApplyShuffled permutation (a sequence without resets)."""
    cards, bgs = pools
    ed, _, _ = ref_import.load_reference()
    for seed in range(200, 260):
        c, b = EO.u8_to_f32(cards[seed % 8]), EO.u8_to_f32(bgs[(seed // 8) % 8])
        random.seed(seed); np.random.seed(seed); ref_import.reset_reference_shuffles(ed)
        ref = ed.SyntheticBgFgMtgImages.make_virtual(c.copy(), b.copy(), (192, 128), True)
        random.seed(seed); np.random.seed(seed); EO.reset_shuffle_state()
        assert np.array_equal(ref, EO.make_virtual(c.copy(), b.copy(), (192, 128), True))
    random.seed(5); np.random.seed(5); ref_import.reset_reference_shuffles(ed)
    refs = [ed.SyntheticBgFgMtgImages.make_virtual(EO.u8_to_f32(cards[i % 8]), EO.u8_to_f32(bgs[i % 8]), (192, 128), False)
            for i in range(20)]
    random.seed(5); np.random.seed(5); EO.reset_shuffle_state()
    mine = [EO.make_virtual(EO.u8_to_f32(cards[i % 8]), EO.u8_to_f32(bgs[i % 8]), (192, 128), False) for i in range(20)]
    assert all(np.array_equal(a, b) for a, b in zip(refs, mine))


def test_batch_oracle_labels_and_hard_negatives():
    """_make_image_batch semantics: labels follow the (possibly swapped) card, swaps stay inside
    the same-name group and never pick the card itself (encoder_datasets.py:619-630)."""
    pool = synth.CardPool(np.stack([synth.synth_card(k, (96, 64)) for k in range(15)]), synth.synth_faces(15))
    bgs = [synth.synth_bg(j, (120, 160)) for j in range(4)]
    random.seed(3); np.random.seed(3); EO.reset_shuffle_state()
    bo = EO.BatchOracle(pool, bgs, paired=True, targets=True, x_size_hw=(48, 32), y_size_hw=(48, 32), similar_neg_prob=0.9)
    imgs, lbls, tapes = bo.random_image_batch(24)
    assert imgs["x"].shape == imgs["x2"].shape == imgs["y"].shape == (24, 48, 32, 3)
    swaps = 0
    for tx, t2, l1, l2 in zip(tapes["x"], tapes["x2"], lbls["x_labels"], lbls["x2_labels"]):
        assert tuple(l1) == tuple(pool.labels3[tx["card"]]) and tuple(l2) == tuple(pool.labels3[t2["card"]])
        if t2["swapped"]:
            swaps += 1
            assert t2["card"] != tx["card"] and t2["card"] in pool.group_of(tx["card"])
            assert l1[1] == l2[1] and l1[0] != l2[0]
        else:
            assert t2["card"] == tx["card"]
    assert swaps > 0
