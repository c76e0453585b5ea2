"""Host-side pieces of bench.py that need no GPU: both arms print the same `config`, the algorithmic-bytes helper of the
detection roofline clips polygons correctly (A_k = area of the card quad inside the frame), and the reference arm's
file-backed card container decodes on access like the reference's loaders."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def _args(**kw):
    a = argparse.Namespace(gpus=1, pool_cards=2048, pool_bgs=1024)
    a.__dict__.update(kw)
    return a


def test_both_arms_state_the_same_config():
    assert bench.workload_config(_args()) == bench.workload_config(_args())
    c = bench.workload_config(_args(gpus=8))
    assert c["card_pool"] == 2048 and c["bg_pool"] == 1024 and c["pairs_per_step"] == 512 and "8 independent" in c["parallelism"]
    assert "model" not in c


def test_clip_area():
    sq = np.array([[100, 100], [300, 100], [300, 400], [100, 400]], float)
    assert abs(bench._clip_area(sq, 640) - 60000.0) < 1e-9
    assert abs(bench._clip_area(sq - 200.0, 640) - 100.0 * 200.0) < 1e-9          # clipped at the top-left corner
    assert bench._clip_area(sq + 1000.0, 640) == 0.0                              # outside
    big = np.array([[-100, -100], [900, -100], [900, 900], [-100, 900]], float)
    assert abs(bench._clip_area(big, 640) - 640.0 * 640.0) < 1e-6                 # covers the frame
    tri = np.array([[0, 0], [640, 0], [0, 640]], float)
    assert abs(bench._clip_area(tri, 640) - 0.5 * 640 * 640) < 1e-6


def test_reference_arm_decodes_files_on_access():
    import cv2

    from mtgvision_b200 import synth

    pool = synth.make_card_pool(3)
    files = [cv2.imencode(".jpg", pool.images[k][:, :, ::-1], [cv2.IMWRITE_JPEG_QUALITY, bench.JPEG_Q])[1].tobytes() for k in range(3)]
    fc = bench._FileCards(pool, files)
    assert len(fc.images) == 3 and fc.images[1].shape == (680, 488, 3) and fc.images[1].dtype == np.uint8
    ref = cv2.imdecode(np.frombuffer(files[1], np.uint8), cv2.IMREAD_COLOR_RGB)
    assert np.array_equal(fc.images[1], ref)
    assert fc.group_of(2) == pool.group_of(2) and fc.labels3 is pool.labels3
