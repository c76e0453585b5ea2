"""Detection path on the CPU: (a) the oracle's geometry / composite against the imported
reference's own functions (when present), (b) subsystem (4) - placement decisions, homographies
and warped keypoint labels of mtgv_det.cuh compiled for the host - against the oracle on the
oracle's recorded tapes.  Bit-exact bars."""
import ctypes as C
import os
import random

import cv2
import numpy as np
import pytest

from mtgvision_b200 import abi, synth
from oracle import det_oracle as DO
from oracle import encoder_oracle as EO
from oracle import ref_import, tape_pack

HARNESS = os.path.join(os.path.dirname(__file__), "host_harness", "libmtgv_hostharness.so")
vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
AUTHOR_RUN = dict(bg_size_hw=640, num_cards_min=1, num_cards_max=9, card_min_visible_ratio=0.5,
                  card_min_visible_ratio_edges=0.0, card_jitter_ratio=0.7)  # od_datasets.py:861-868


@pytest.fixture(scope="module")
def pools():
    return [synth.synth_card(k) for k in range(6)], [synth.synth_bg(j) for j in range(4)]


def scenes(pools, kind, seeds, photometrics=False, **kw):
    cards, bgs = pools
    out = []
    for seed in seeds:
        random.seed(seed); np.random.seed(seed)
        o = DO.DetOracle(cards, bgs, kind=kind, photometrics=photometrics, **{**AUTHOR_RUN, **kw})
        t = {}
        sample = o.generate(t)
        out.append((o, t, sample))
    return out


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not present on this box")
def test_oracle_geometry_and_composite_equal_reference(pools):
    _, od, _ = ref_import.load_reference()
    rng = np.random.default_rng(0)
    for s in range(100):
        pts = np.asarray([(0, 0), (488, 0), (488, 680), (0, 680)])
        np.random.seed(s); a = od.corner_jitter_2d(pts.copy(), 0.7)
        np.random.seed(s); b = DO.corner_jitter_2d(pts.copy(), np.random.uniform(1 - 0.7, 1 + 0.7, size=4))
        assert np.array_equal(a, b)
        deg, sc = rng.uniform(0, 360), rng.uniform(0.1, 2)
        assert np.array_equal(od.rotate_2d(a, deg, center=(244, 340), scale=sc), DO.rotate_2d(a, deg, center=(244, 340), scale=sc))
        assert np.array_equal(od.get_rotate_over_output_transform((375, 500), int(deg), 640, "cover"),
                              DO.get_rotate_over_output_transform((375, 500), int(deg), (640, 640)))
    cards, bgs = pools
    ref_card = od.make_card_with_mask(EO.u8_to_f32(cards[0]), kind="obb")
    assert np.array_equal(ref_card["keypoints"], DO.card_keypoints(680, 488, "obb"))
    assert np.array_equal(ref_card["bbox"], np.asarray(DO.box(0, 0, 488, 680)))
    for kind in ("obb", "seg"):
        for o, t, sample in scenes(pools, kind, range(4)):
            bg = od.rotate_over_output({"image": EO.u8_to_f32(bgs[t["bg"]])}, deg=t["bg_deg"], out_size_hw=640)["image"]
            placed = [(c["card"], c["attempts"][-1]["M"]) for c in t["cards"] if c["attempts"] and c["attempts"][-1]["accepted"]]
            kps = []
            for k, M in placed[::-1]:
                mask = od.apply_transform_2d_img(ref_card["mask"], M, out_size_hw=bg.shape)
                img = od.apply_transform_2d_img(EO.u8_to_f32(cards[k]), M, out_size_hw=bg.shape)
                bg = mask[:, :, None] * img + (1 - mask[:, :, None]) * bg
                kps.extend(od.apply_transform_2d(o.keypoints, M))
            assert np.array_equal(bg, sample["image"])
            assert np.array_equal(np.asarray(kps).reshape(sample["keypoints"].shape), sample["keypoints"])


def test_golden_det_geometry_kats():
    import json
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "encoder_golden.json")))["det"]
    assert np.array_equal(np.asarray(g["rotate_over_output_375x500_37_640"]), DO.get_rotate_over_output_transform((375, 500), 37, (640, 640)))
    assert np.array_equal(np.asarray(g["obb_keypoints_680x488"]), DO.card_keypoints(680, 488, "obb"))
    assert np.allclose(DO.card_keypoints(680, 488, "obb")[1], [[20.4, 20.4], [467.6, 20.4], [467.6, 329.8], [20.4, 329.8]])


def test_polygon_restatement_sanity():
    sq = [(0, 0), (4, 0), (4, 4), (0, 4)]
    assert DO.poly_area(DO.clip_convex(sq, [(2, 2), (6, 2), (6, 6), (2, 6)])) == 4.0
    assert DO.poly_area(DO.clip_convex(sq, [(6, 2), (2, 2), (2, 6), (6, 6)])) == 4.0  # orientation-agnostic
    assert DO.clip_convex(sq, [(5, 5), (6, 5), (6, 6), (5, 6)]) == []
    u = DO.CardShape(DO.card_keypoints(10, 10, "seg")[0], "seg")
    assert abs(u.area_within(None) - (100 - 2 * 5)) < 1e-12
    assert abs(u.area_within([(0, 5), (10, 5), (10, 10), (0, 10)]) - (50 - 10)) < 1e-12


@pytest.mark.parametrize("kind", ["obb", "seg"])
def test_host_placement_and_labels_match_oracle_bit_exact(pools, kind):
    hh = C.CDLL(HARNESS)
    cards, bgs = pools
    runs = scenes(pools, kind, range(40)) + scenes(pools, kind, range(40, 50), card_min_visible_ratio_edges=0.75, num_cards_max=6)
    for cfg_kw, group in ((AUTHOR_RUN, runs[:40]), ({**AUTHOR_RUN, "card_min_visible_ratio_edges": 0.75, "num_cards_max": 6}, runs[40:])):
        tapes = [t for _, t, _ in group]
        arr, _ = tape_pack.pack_det_tapes(tapes)
        n = len(tapes)
        cfg = abi.DetConfig(640, 640, cfg_kw["num_cards_min"], cfg_kw["num_cards_max"], cfg_kw["card_min_visible_ratio"],
                            cfg_kw["card_min_visible_ratio_edges"], cfg_kw["card_jitter_ratio"], 0.02, 0.9, 0.0, 1, 10,
                            0 if kind == "obb" else 1, 0)
        hh.hh_det_params_size.restype = C.c_int
        params = np.zeros((n, hh.hh_det_params_size()), dtype=np.uint8)
        nk = abi.DET_MAX_CARDS * abi.DET_MAX_KPOLY
        accepted = np.zeros((n, abi.DET_MAX_CARDS), dtype=np.int32)
        kps = np.zeros((n, nk, abi.DET_MAX_KP, 2), dtype=np.float64)
        labels = np.zeros((n, nk), dtype=np.int32)
        counts = np.zeros(n, dtype=np.int32)
        bg_hw = np.asarray([b.shape[:2] for b in bgs], dtype=np.int32)
        bad = hh.hh_det_place(vp(arr), n, C.byref(cfg), 680, 488, len(cards), len(bgs), vp(bg_hw), vp(params), vp(accepted),
                              vp(kps), vp(labels), vp(counts))
        assert bad == 0
        total_attempts = rejected = 0
        for s, (o, t, sample) in enumerate(group):
            for ci, c in enumerate(t["cards"]):
                want = len(c["attempts"]) - 1 if c["attempts"][-1]["accepted"] else -1
                assert accepted[s, ci] == want, f"scene {s} card {ci}: accept/reject decision differs"
                total_attempts += len(c["attempts"]); rejected += sum(not a["accepted"] for a in c["attempts"])
            k = sample["keypoints"].shape[0]
            assert counts[s] == k
            P = sample["keypoints"].shape[1] if k else 0
            assert np.array_equal(kps[s, :k, :P], sample["keypoints"].reshape(k, P, 2)), f"scene {s}: keypoints differ"
            assert np.array_equal(labels[s, :k], sample["keypoints_labels"]) and np.all(labels[s, k:] == -1)
        assert rejected > 10 and total_attempts > rejected  # both branches exercised


def test_keypoint_warp_matches_numpy_matmul_on_this_host():
    """apply_transform_2d relies on numpy's `pts @ M.T` rounding, an FMA chain on x86 hosts with FMA
    (host-BLAS dependent): det_apply must reproduce it bit for bit."""
    hh = C.CDLL(HARNESS)
    rng = np.random.default_rng(1)
    for _ in range(20):
        src = np.float32([[0, 0], [488, 0], [488, 680], [0, 680]])
        dst = (src * rng.uniform(0.2, 1.2) + rng.uniform(-80, 300, (4, 2))).astype(np.float32)
        M = cv2.getPerspectiveTransform(src, dst)
        pts = rng.uniform(-50, 750, (64, 2))
        out = np.zeros_like(pts)
        hh.hh_det_apply(vp(np.ascontiguousarray(M)), vp(pts), len(pts), vp(out))
        assert np.array_equal(out, DO.apply_transform_2d(pts, M))


def test_motion_blur_line_mask_equals_cv2_line():
    """The MotionBlur kernel of the production sampler (det_line_mask, mtgv_det.cuh) is cv2.line's pixel set (thickness 1,
    8-connected) for EVERY pair of cells of every kernel size of A.MotionBlur(blur_limit=(3, 11)) (od_datasets.py:453-454)."""
    hh = C.CDLL(HARNESS)
    checked = 0
    for ks in (3, 5, 7, 9, 11):
        for x1 in range(ks):
            for y1 in range(ks):
                for x2 in range(ks):
                    for y2 in range(ks):
                        k = np.zeros((ks, ks), np.uint8)
                        cv2.line(k, (x1, y1), (x2, y2), 1, thickness=1)
                        w = np.zeros(4, np.int32)
                        hh.hh_line_mask(ks, x1, y1, x2, y2, w.ctypes.data_as(C.c_void_p))
                        bits = np.unpackbits(w.view(np.uint8), bitorder="little")[: ks * ks].reshape(ks, ks)
                        assert np.array_equal(bits, k), (ks, x1, y1, x2, y2)
                        checked += 1
    assert checked == sum(ks ** 4 for ks in (3, 5, 7, 9, 11))
