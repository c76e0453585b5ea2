"""JPEG encode (save_sample -> imwrite -> cv2.imwrite, od_datasets.py:794-832, util/image.py:95-104):
the oracle restatement and the arithmetic shared with the device kernels (mtgvision_b200/csrc/mtgv_jpegenc.cuh,
compiled for the host by tests/host_harness) must produce the FILE BYTES of cv2.imencode."""
import ctypes as C
import os

import cv2
import numpy as np
import pytest

from oracle import jpeg_encode_oracle as E
from tests import jpeg_cases

HARNESS = os.path.join(os.path.dirname(__file__), "host_harness", "libmtgv_hostharness.so")


def cv2_bytes(rgb, quality=None):
    params = [] if quality is None else [cv2.IMWRITE_JPEG_QUALITY, quality]
    return cv2.imencode(".jpg", cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR), params)[1].tobytes()  # imwrite's default: quality 95


def cases():
    rng = np.random.default_rng(0)
    out = []
    for h, w in [(16, 16), (32, 48), (64, 64), (128, 96)]:
        for kind in ["noise", "mixed", "smooth"]:
            for q in [None, 90, 75, 50, 30, 100, 1]:
                out.append((jpeg_cases.image(rng, h, w, kind), q))
    for h, w in [(1, 1), (7, 5), (9, 17), (17, 16), (23, 40), (31, 33), (40, 57), (49, 15), (100, 101), (15, 130)]:  # not whole MCUs
        out.append((jpeg_cases.image(rng, h, w, "mixed"), None))
        out.append((jpeg_cases.image(rng, h, w, "noise"), 60))
    out.append((np.full((32, 32, 3), 255, np.uint8), None))
    out.append((np.zeros((48, 16, 3), np.uint8), None))
    return out


def test_oracle_bytes_equal_cv2():
    for img, q in cases():
        assert E.encode(img, 95 if q is None else q) == cv2_bytes(img, q), (img.shape, q)


def host_encode(hh, img, quality=95):
    img = np.ascontiguousarray(img)
    out = np.zeros(img.size * 4 + 4096, np.uint8)
    hh.hh_jpeg_encode.restype = C.c_int64
    n = hh.hh_jpeg_encode(img.ctypes.data_as(C.c_void_p), img.shape[0], img.shape[1], quality, out.ctypes.data_as(C.c_void_p),
                          C.c_int64(out.size))
    assert n > 0
    return out[:n].tobytes()


def test_shared_arithmetic_bytes_equal_cv2():
    hh = C.CDLL(HARNESS)
    for img, q in cases():
        assert host_encode(hh, img, 95 if q is None else q) == cv2_bytes(img, q), (img.shape, q)
    rng = np.random.default_rng(4)
    scene = jpeg_cases.image(rng, 640, 640, "mixed")  # save_sample's size
    assert host_encode(hh, scene) == cv2_bytes(scene)
