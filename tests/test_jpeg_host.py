"""The JPEG decode arithmetic shared by the device kernels (mtgvision_b200/csrc/mtgv_jpeg.cuh), compiled for the host
by tests/host_harness and compared bit for bit with cv2.imdecode - the call behind the reference's imread_float
(mtgvision/util/image.py:107-114) - and with the oracle restatement."""
import ctypes as C
import os

import cv2
import numpy as np
import pytest

from oracle import jpeg_oracle
from tests import jpeg_cases

HARNESS = os.path.join(os.path.dirname(__file__), "host_harness", "libmtgv_hostharness.so")


@pytest.fixture(scope="module")
def hh():
    return C.CDLL(HARNESS)


def host_decode(hh, data: bytes):
    buf = np.frombuffer(data, np.uint8)
    hw = np.zeros(2, np.int32)
    msg = C.create_string_buffer(256)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    if hh.hh_jpeg_decode(vp(buf), C.c_int64(len(buf)), vp(hw), None, msg, 256) != 0:
        raise ValueError(msg.value.decode())
    out = np.zeros((hw[0], hw[1], 3), np.uint8)
    assert hh.hh_jpeg_decode(vp(buf), C.c_int64(len(buf)), vp(hw), vp(out), msg, 256) == 0
    return out


def test_shared_arithmetic_equals_cv2(hh):
    for name, data in jpeg_cases.small_suite():
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR_RGB)
        got = host_decode(hh, data)
        assert got.shape == ref.shape and np.array_equal(got, ref), name


def test_background_sized_files(hh):
    rng = np.random.default_rng(11)
    for s, rst in [("420", 0), ("422", 8), ("444", 0), ("440", 5)]:
        data = jpeg_cases.encode(jpeg_cases.image(rng, 375, 500, "mixed"), 90, s, rst)
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR_RGB)
        assert np.array_equal(host_decode(hh, data), ref), (s, rst)
    data = jpeg_cases.encode(jpeg_cases.image(rng, 75, 100, "noise"), 97, "420")
    assert np.array_equal(host_decode(hh, data), jpeg_oracle.decode(data))


def test_unsupported_files_are_rejected_with_a_message(hh):
    rng = np.random.default_rng(3)
    img = jpeg_cases.image(rng, 24, 24, "mixed")
    arith = jpeg_cases.encode(img).replace(b"\xff\xc0", b"\xff\xc9", 1)  # SOF9: arithmetic coding
    with pytest.raises(ValueError, match="Huffman-coded"):
        host_decode(hh, arith)
    with pytest.raises(ValueError, match="SOI"):
        host_decode(hh, b"\x89PNG....")
    good = jpeg_cases.encode(img)
    with pytest.raises(ValueError, match="truncated"):
        host_decode(hh, good[:100])


def test_truncated_files_decode_like_cv2_imread(hh, tmp_path):
    for name, data in jpeg_cases.truncated_suite():
        assert np.array_equal(host_decode(hh, data), jpeg_cases.imread_ref(data, tmp_path)), name


def test_progressive_files_decode_like_cv2(hh):
    """SOF2 files (ILSVRC and Scryfall hold some): the host entropy stage (mtgv_jpeg_prog.h: DC / AC first and refinement scans,
    end-of-band runs, restart intervals) feeding the shared IDCT / upsampling / colour arithmetic equals cv2.imdecode."""
    rng = np.random.default_rng(21)
    n = 0
    for h, w in [(24, 24), (375, 500), (17, 33), (1, 1), (100, 7)]:
        for kind in ("mixed", "noise", "smooth"):
            for samp, q, rst, opt in [("420", 90, 0, 0), ("444", 35, 3, 1), ("422", 100, 0, 1), ("440", 75, 5, 0)]:
                data = jpeg_cases.encode(jpeg_cases.image(rng, h, w, kind), q, samp, rst, opt, progressive=1)
                assert b"\xff\xc2" in data
                ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR_RGB)
                assert np.array_equal(host_decode(hh, data), ref), (h, w, kind, samp, q, rst, opt)
                n += 1
    gray = rng.integers(0, 256, (40, 56), dtype=np.uint8)
    data = jpeg_cases.encode(gray, 90, "420", progressive=1)
    assert np.array_equal(host_decode(hh, data), cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR_RGB))
    assert n == 60
