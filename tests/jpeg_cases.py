"""Synthetic JPEG files for the decode parity tests (encoded with cv2.imencode = libjpeg-turbo)."""
import cv2
import numpy as np

SAMPLING = {
    "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420,
    "444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444,
    "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
    "440": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440,
}


def image(rng, h, w, kind):
    noise = rng.integers(0, 256, (h, w, 3), np.uint8)
    if kind == "noise":
        return noise
    yy, xx = np.mgrid[0:h, 0:w]
    ramp = np.stack([xx * 255 // max(w - 1, 1), yy * 255 // max(h - 1, 1), (xx + yy) * 255 // max(h + w - 2, 1)], -1)
    if kind == "smooth":
        return ((ramp + noise // 8) * 8 // 9).astype(np.uint8)
    return ((ramp + noise) // 2).astype(np.uint8)  # "mixed"


def encode(img_rgb, quality=90, sampling="420", rst=0, optimize=0, progressive=0):
    src = img_rgb if img_rgb.ndim == 2 else img_rgb[:, :, ::-1]
    ok, buf = cv2.imencode(".jpg", src, [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SAMPLING[sampling],
                                          cv2.IMWRITE_JPEG_RST_INTERVAL, rst, cv2.IMWRITE_JPEG_OPTIMIZE, optimize,
                                          cv2.IMWRITE_JPEG_PROGRESSIVE, progressive])
    assert ok
    return buf.tobytes()


def small_suite(seed=0):
    """(name, file bytes) over sampling modes, ragged sizes down to 1x1, qualities, restart intervals, optimised tables."""
    rng = np.random.default_rng(seed)
    out = []
    for h, w in [(16, 16), (17, 23), (33, 31), (48, 64), (8, 8), (1, 1), (3, 5), (40, 2), (2, 40), (5, 4), (37, 50)]:
        for s in SAMPLING:
            for q, kind, rst, opt in [(30, "noise", 0, 0), (75, "mixed", 3, 0), (95, "smooth", 0, 1), (100, "noise", 1, 1)]:
                out.append((f"{h}x{w}_{s}_q{q}_{kind}_r{rst}_o{opt}", encode(image(rng, h, w, kind), q, s, rst, opt)))
    for h, w in [(17, 23), (32, 32)]:
        out.append((f"{h}x{w}_gray", encode(image(rng, h, w, "mixed")[:, :, 0], 80)))
    return out


def truncated_suite(seed=1):
    """Files cut short inside the scan ("Premature end of JPEG file"): cv2.imread decodes what is there and leaves the rest
    grey; cv2.imdecode refuses them, so the reference for these is `imread_ref`."""
    rng = np.random.default_rng(seed)
    out = []
    for (h, w), s, q in [((64, 80), "420", 90), ((48, 48), "444", 75), ((100, 37), "422", 95), ((33, 70), "440", 60), ((120, 160), "420", 85)]:
        data = encode(image(rng, h, w, "mixed"), q, s)
        scan = data.index(b"\xff\xda")
        for frac in (0.05, 0.31, 0.5, 0.77, 0.98):
            cut = scan + 14 + int((len(data) - scan - 16) * frac)
            out.append((f"{h}x{w}_{s}_q{q}_cut{frac}", data[:cut]))
        out.append((f"{h}x{w}_{s}_q{q}_noEOI", data[:-2]))
    return out


def imread_ref(data: bytes, tmpdir) -> np.ndarray:
    """cv2.imread(path, IMREAD_COLOR_RGB) - the reference's own call (imread_float, util/image.py:107-114)."""
    import os

    p = os.path.join(str(tmpdir), "case.jpg")
    with open(p, "wb") as f:
        f.write(data)
    return cv2.imread(p, cv2.IMREAD_COLOR_RGB)
