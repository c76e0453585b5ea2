"""The on-disk YOLO sample format: `save_sample` of the drop-in against the reference's own
`save_sample` (od_datasets.py:794-832) on identical sample dicts - label text byte for byte, image files
pixel for pixel.  The committed known-answer keeps the check alive where the reference tree is absent."""
import os
import warnings

import cv2
import numpy as np
import pytest

from oracle import ref_import


def _sample(seed, n_poly, pts_per_poly, out_of_bounds=False):
    rng = np.random.default_rng(seed)
    img = rng.random((640, 640, 3), dtype=np.float32)
    kps = rng.uniform(0.0, 640.0, (n_poly, pts_per_poly, 2))
    if out_of_bounds:
        kps[0, 0] = (-3.25, 700.5)
    return {"image": img, "keypoints": kps, "keypoints_labels": np.arange(n_poly) % 3}


def test_label_text_known_answer(tmp_path):
    from mtgvision_b200.od_datasets import save_sample

    s = {"image": np.zeros((640, 640, 3), np.float32), "keypoints": np.array([[[64.0, 320.0], [640.0, 0.0], [1.0, 2.0], [480.0, 639.0]]]),
         "keypoints_labels": np.array([2])}
    save_sample(s, 7, tmp_path, tmp_path, ext="png")
    # str(float64) of pts / (w, h): the reference joins map(str, pts.flatten()) (od_datasets.py:819)
    assert (tmp_path / "image_0007.txt").read_text() == "2 0.1 0.5 1.0 0.0 0.0015625 0.003125 0.75 0.9984375\n"
    assert cv2.imread(str(tmp_path / "image_0007.png")).shape == (640, 640, 3)


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not present on this box")
@pytest.mark.parametrize("ext", ["png", "jpg"])
def test_save_sample_equals_reference(tmp_path, ext):
    from mtgvision_b200.od_datasets import save_sample

    _, od, _ = ref_import.load_reference()
    cases = [_sample(1, 3, 4), _sample(2, 1, 8), _sample(3, 0, 4), _sample(4, 6, 4, out_of_bounds=True)]
    for i, s in enumerate(cases):
        ours, theirs = tmp_path / f"ours{i}", tmp_path / f"ref{i}"
        ours.mkdir(); theirs.mkdir()
        ref_in = {"image": s["image"].copy(), "keypoints": np.array(s["keypoints"], copy=True), "keypoints_labels": s["keypoints_labels"].copy()}
        with warnings.catch_warnings(record=True) as w_ref:
            warnings.simplefilter("always")
            od.save_sample(ref_in, i, theirs, theirs, ext=ext)  # normalises its keypoints in place
        with warnings.catch_warnings(record=True) as w_ours:
            warnings.simplefilter("always")
            save_sample(s, i, ours, ours, ext=ext)
        assert (ours / f"image_{i:04d}.txt").read_bytes() == (theirs / f"image_{i:04d}.txt").read_bytes()
        assert bool(w_ref) == bool(w_ours)  # out-of-bounds points warn in both
        a = cv2.imread(str(ours / f"image_{i:04d}.{ext}")); b = cv2.imread(str(theirs / f"image_{i:04d}.{ext}"))
        assert np.array_equal(a, b)
        assert (ours / f"image_{i:04d}.{ext}").read_bytes() == (theirs / f"image_{i:04d}.{ext}").read_bytes()
