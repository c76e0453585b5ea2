"""Pins oracle/cv2_restate.py (the written specification of the CUDA arithmetic) against the
real OpenCV calls the reference makes.  CPU only."""
import hashlib
import math

import cv2
import numpy as np
import pytest

from oracle import cv2_restate as R


@pytest.fixture(scope="module")
def rng():
    return np.random.default_rng(7)


@pytest.mark.parametrize("hw,ratio,zeros,sha", [((680, 488), 0.05, 868, "75fac4e48730e3c0"),
                                                ((680, 488), 0.046, 764, "d7923dd2973e78fe"),
                                                ((204, 146), 0.05, 68, "8fb55ccac7cc9cff")])
def test_round_rect_mask_kat(hw, ratio, zeros, sha):
    """Known answers from running the reference's round_rect_mask."""
    radius = int(math.ceil(max(hw) * ratio))
    m = R.round_rect_mask(hw, radius)
    assert int((m == 0).sum()) == zeros
    assert hashlib.sha1(m.tobytes()).hexdigest()[:16] == sha
    c = np.zeros((radius, radius), np.float32)
    cv2.circle(c, (0, 0), radius, 1, cv2.FILLED)
    assert np.array_equal(c, R.filled_quarter_circle(radius))


def test_perspective_transform_and_invert_bit_exact(rng):
    for _ in range(300):
        src = np.float32([[0, 0], [487, 0], [487, 679], [0, 679]]) + rng.normal(0, 3, (4, 2)).astype(np.float32)
        dst = (src + rng.normal(0, 60, (4, 2))).astype(np.float32)
        M = cv2.getPerspectiveTransform(src, dst)
        assert np.array_equal(M, R.get_perspective_transform(src, dst))
        assert np.array_equal(cv2.invert(M)[1], R.invert3x3(M))


def test_rotation_matrix_bit_exact(rng):
    for _ in range(500):
        c = (float(rng.integers(0, 600)), float(rng.integers(0, 600)))
        a, s = rng.uniform(0, 360), rng.uniform(0.5, 3)
        assert np.array_equal(cv2.getRotationMatrix2D(c, a, s), R.get_rotation_matrix_2d(c, a, s))


@pytest.mark.parametrize("shape", [(192, 128, 4), (192, 128), (375, 500, 3), (50, 40, 3)])
def test_warps_bit_exact(rng, shape):
    img = rng.random(shape, dtype=np.float32)
    h, w = shape[:2]
    src = np.float32([[0, 0], [w, 0], [0, h], [w, h]])
    dst = (src + rng.uniform(-0.1, 0.1, (4, 2)) * [w, h]).astype(np.float32)
    M = cv2.getPerspectiveTransform(src, dst)
    for dsize in [(w, h), (w + 70, h + 33)]:
        assert np.array_equal(cv2.warpPerspective(img, M, dsize), R.warp_perspective(img, M, dsize))
    A = cv2.getRotationMatrix2D((w / 2, h / 2), rng.uniform(0, 360), rng.uniform(0.8, 1.2))
    A[:, 2] += rng.uniform(-10, 10, 2)
    for dsize in [(w, h), (w + 77, h + 33)]:
        assert np.array_equal(cv2.warpAffine(img, A, dsize), R.warp_affine(img, A, dsize))


def test_resize_area_bit_exact_on_path_sizes(rng):
    img = rng.random((680, 488, 4), dtype=np.float32)
    assert np.array_equal(cv2.resize(img, (128, 178), interpolation=cv2.INTER_AREA), R.resize_area(img, (128, 178)))
    crop = img[14:-14, 14:-14, :3]
    assert np.array_equal(cv2.resize(crop, (128, 192), interpolation=cv2.INTER_AREA), R.resize_area(crop, (128, 192)))
    bg = rng.random((618, 603, 3), dtype=np.float32)
    assert np.array_equal(cv2.resize(bg, (187, 192), interpolation=cv2.INTER_AREA), R.resize_area(bg, (187, 192)))


def test_resize_small_kernels_within_tolerance(rng):
    """NEAREST exact; LINEAR / CUBIC / blur / sharpen within 1e-6 (cv2's SIMD contraction order is
    not restated; 1e-6 = 0.0003 uint8 LSB)."""
    img = rng.random((192, 128, 3), dtype=np.float32)
    for n in (1, 2):
        for it, tol in ((0, 0.0), (1, 1e-6), (2, 1e-6)):
            d = cv2.resize(img, (128 >> n, 192 >> n), interpolation=it)
            assert np.abs(d - R.resize(img, (128 >> n, 192 >> n), it)).max() <= tol
            u = cv2.resize(d, (128, 192), interpolation=it)
            assert np.abs(u - R.resize(d, (128, 192), it)).max() <= tol
    assert np.abs(cv2.GaussianBlur(img, (3, 3), 0) - R.gaussian_blur3(img)).max() <= 1e-6
    k = np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]])
    assert np.abs(cv2.filter2D(img, -1, k) - R.sharpen3(img)).max() <= 2e-6
    assert np.array_equal(cv2.getGaussianKernel(3, 0).ravel(), [0.25, 0.5, 0.25])


def test_warp_perspective_u8_fixed_point_bit_exact():
    """uint8 warpPerspective (serving-side dewarp, od_export.py:108): remapBilinear's 15-bit fixed point."""
    rng = np.random.default_rng(5)
    for t in range(12):
        H, W = int(rng.integers(200, 700)), int(rng.integers(200, 800))
        src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        quad = np.float32([[W * .2, H * .2], [W * .8, H * .25], [W * .75, H * .8], [W * .15, H * .7]]) + rng.uniform(-40, 40, (4, 2)).astype(np.float32)
        if t % 4 == 0:
            quad[0] = (-30, -20)
        h, w = 192, 128
        dst = ((1 + 0.05) * np.asarray([[0, 0], [w, 0], [w, h], [0, h]]) - 0.025 * np.asarray([w, h])).astype(np.float32)
        M = cv2.getPerspectiveTransform(quad, dst)
        assert np.array_equal(R.warp_perspective_u8(src, M, (w, h)), cv2.warpPerspective(src, M, (w, h)))
