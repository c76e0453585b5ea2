"""Soak beyond the suite: random JPEG files (decode) and random images (encode) against cv2, bit / byte exact.
    python tests/soak_gpu_jpeg.py [n_decode] [n_encode] > profiles/rNN_jpeg_soak.json"""
import json
import sys

import cv2
import numpy as np
import torch

sys.path.insert(0, ".")
from mtgvision_b200.context import Context  # noqa: E402
from tests import jpeg_cases  # noqa: E402

n_dec = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
n_enc = int(sys.argv[2]) if len(sys.argv) > 2 else 300
rng = np.random.default_rng(2026)
ctx = Context(0)
files, meta = [], []
for i in range(n_dec):
    h, w = (int(rng.integers(1, 400)), int(rng.integers(1, 400))) if i % 5 else (int(rng.integers(400, 1300)), int(rng.integers(400, 1300)))
    kind = ["noise", "mixed", "smooth"][int(rng.integers(0, 3))]
    q, s = int(rng.integers(1, 101)), ["420", "444", "422", "440"][int(rng.integers(0, 4))]
    rst, opt, gray = int(rng.choice([0, 0, 0, 1, 7])), int(rng.integers(0, 2)), i % 17 == 0
    img = jpeg_cases.image(rng, h, w, kind)
    files.append(jpeg_cases.encode(img[:, :, 0] if gray else img, q, s, rst, opt))
    meta.append((h, w, kind, q, s, rst, opt, gray))
bad_dec, px = [], 0
for b0 in range(0, n_dec, 250):
    chunk = files[b0:b0 + 250]
    flat, off, hw = ctx.decode_jpegs(chunk)
    flat = flat.cpu().numpy()
    for i, data in enumerate(chunk):
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR_RGB)
        h, w = hw[i]
        got = flat[off[i]: off[i] + 3 * h * w].reshape(h, w, 3)
        px += int(h) * int(w)
        if not np.array_equal(got, ref):
            bad_dec.append(meta[b0 + i])
bad_enc, nbytes = [], 0
for i in range(n_enc):
    h, w = (16 * int(rng.integers(1, 50)), 16 * int(rng.integers(1, 50))) if i % 2 else (int(rng.integers(1, 500)), int(rng.integers(1, 500)))
    q = int(rng.integers(1, 101))
    imgs = np.stack([jpeg_cases.image(rng, h, w, kind) for kind in ("noise", "mixed", "smooth")])
    out = ctx.encode_jpegs(torch.from_numpy(imgs).cuda(), q, cap=((h + 16) * (w + 16) * 4 + 8192) // 4 * 4)
    for k in range(3):
        ref = cv2.imencode(".jpg", cv2.cvtColor(imgs[k], cv2.COLOR_RGB2BGR), [cv2.IMWRITE_JPEG_QUALITY, q])[1].tobytes()
        nbytes += len(ref)
        if out[k] != ref:
            bad_enc.append((h, w, q, k))
print(json.dumps({"decode": {"files": n_dec, "pixels": px, "mismatching_files": len(bad_dec), "first": bad_dec[:5]},
                  "encode": {"images": 3 * n_enc, "file_bytes": nbytes, "mismatching_files": len(bad_enc), "first": bad_enc[:5]}}))
