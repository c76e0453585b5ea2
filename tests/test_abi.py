"""The C-ABI shared library loads without a GPU, exports every symbol include/mtgv.h
declares, and the ctypes mirrors have the C struct sizes.  No compute calls here."""
import ctypes as C
import os

import numpy as np
import pytest

from mtgvision_b200 import abi


def test_library_exports_every_declared_symbol():
    lib = abi.load_library()
    names = abi.declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mtgv.h but not exported"
    assert lib.mtgv_abi_version() == 1


def test_struct_sizes_match_c():
    lib = abi.load_library()
    lib.mtgv_sizeof.restype = C.c_int
    for which, t in enumerate([abi.TapeOp, abi.EncTape, abi.XOp, abi.EncParams, abi.EncConfig, abi.PhotoOp,
                               abi.DetAttempt, abi.DetCard, abi.DetTape, abi.DetConfig]):
        assert lib.mtgv_sizeof(which) == C.sizeof(t), t.__name__
    assert abi.TAPE_DTYPE.itemsize == C.sizeof(abi.EncTape)
    assert abi.PARAMS_DTYPE.itemsize == C.sizeof(abi.EncParams)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(abi.MtgvError):
        abi.load_library(str(tmp_path / "nope.so"))


def test_null_context_is_rejected_without_a_gpu():
    lib = abi.load_library()
    assert lib.mtgv_launch_count(None) == 0
    assert lib.mtgv_last_error(None) == b"null context"
    assert lib.mtgv_expand_params(None, None, 0, None, None, None) < 0


def test_no_product_module_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "mtgvision_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
