"""InstanceSeg.extract_dewarped on the GPU against cv2 itself: bit-exact on uint8 frames."""
import cv2
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _ref(frame, pts, out_size_hw, expand_ratio):
    # the reference body, mtgvision/od_export.py:101-109
    h, w = out_size_hw
    dst_pts = np.asarray([[0, 0], [w, 0], [w, h], [0, h]])
    dst_pts = (1 + expand_ratio) * dst_pts - (0.5 * expand_ratio) * np.asarray([w, h])
    M = cv2.getPerspectiveTransform(np.asarray(pts).astype(np.float32), np.asarray(dst_pts).astype(np.float32))
    return cv2.warpPerspective(frame, M, out_size_hw[::-1])


@pytest.mark.parametrize("frame_hw,channels,out_hw", [((720, 1280), 3, (192, 128)), ((480, 640), 3, (192, 128)), ((1080, 1920), 4, (96, 64)),
                                                       ((333, 517), 1, (224, 160))])
def test_extract_dewarped_bit_exact(frame_hw, channels, out_hw):
    from mtgvision_b200.od_export import Dewarper

    rng = np.random.default_rng(frame_hw[0])
    H, W = frame_hw
    frame = rng.integers(0, 256, (H, W, channels), dtype=np.uint8)
    quads = []
    for k in range(12):
        cx, cy, s = rng.uniform(0.2, 0.8) * W, rng.uniform(0.2, 0.8) * H, rng.uniform(0.08, 0.3) * min(H, W)
        a = rng.uniform(0, 2 * np.pi)
        base = np.array([[-0.7, -1.0], [0.7, -1.0], [0.7, 1.0], [-0.7, 1.0]]) * s
        rot = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        q = base @ rot.T + (cx, cy) + rng.uniform(-0.08, 0.08, (4, 2)) * s
        if k % 4 == 0:
            q = q.astype(int).astype(np.float64)  # the reference's `_xyxyxyxy` are ints (:92)
        if k == 5:
            q += (W * 0.6, 0)  # partly outside the frame: constant 0 border
        quads.append(q)
    dw = Dewarper(device=0)
    f = frame if channels > 1 else frame  # HWC either way
    got = dw.extract_dewarped(frame, np.stack(quads), out_hw, 0.05).cpu().numpy()
    for k, q in enumerate(quads):
        ref = _ref(f, q, out_hw, 0.05).reshape(out_hw[0], out_hw[1], channels)
        assert np.array_equal(got[k], ref), f"card {k}: {np.abs(got[k].astype(int) - ref).max()} LSB"
    one = dw.extract_dewarped(frame, quads[1], out_hw).cpu().numpy()
    assert np.array_equal(one, got[1])
    dw.ctx.close()
