"""Pins oracle/jpeg_oracle.py (restated libjpeg-turbo baseline decode) bit for bit against cv2.imdecode,
the call behind the reference's imread_float (mtgvision/util/image.py:107-114)."""
import cv2
import numpy as np
import pytest

from oracle import jpeg_oracle
from tests import jpeg_cases


def test_oracle_decode_equals_cv2_imdecode():
    cases = jpeg_cases.small_suite()
    assert len(cases) > 150
    for name, data in cases:
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR_RGB)
        got = jpeg_oracle.decode(data)
        assert got.shape == ref.shape and np.array_equal(got, ref), name


def test_background_sized_file():
    rng = np.random.default_rng(7)
    data = jpeg_cases.encode(jpeg_cases.image(rng, 75, 100, "mixed"), 85, "420")
    assert np.array_equal(jpeg_oracle.decode(data), cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR_RGB))


def test_unsupported_files_are_rejected():
    rng = np.random.default_rng(3)
    img = jpeg_cases.image(rng, 24, 24, "mixed")
    with pytest.raises(jpeg_oracle.JpegUnsupported):
        jpeg_oracle.decode(jpeg_cases.encode(img, progressive=1))
    with pytest.raises(jpeg_oracle.JpegUnsupported):
        jpeg_oracle.decode(b"\x89PNG....")


def test_truncated_files_decode_like_cv2_imread(tmp_path):
    """A premature end of the data: the MCU in progress is finished from zero bits, later MCUs stay grey (jdhuff.c)."""
    for name, data in jpeg_cases.truncated_suite():
        ref = jpeg_cases.imread_ref(data, tmp_path)
        assert ref is not None, name
        assert np.array_equal(jpeg_oracle.decode(data), ref), name
