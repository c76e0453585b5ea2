"""TEST INFRASTRUCTURE ONLY - import the real reference (when present) behind stubs.

`/root/reference` exists only in the build container, never on the GPU box.  Tests that
pin the oracle against the live reference, and `tests/golden/make_golden.py`, call
`load_reference()`; everything else uses the committed golden vectors.
"""

from __future__ import annotations

import contextlib
import io
import os
import sys

REFERENCE_ROOT = os.environ.get("MTGV_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refstubs")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "mtgvision", "encoder_datasets.py"))


def load_reference():
    """Returns (encoder_datasets, od_datasets, util.image) modules of the real reference."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for p in (_STUBS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    with contextlib.redirect_stdout(io.StringIO()):  # encoder_datasets prints DATASETS_ROOT on import
        import mtgvision.encoder_datasets as ed
        import mtgvision.od_datasets as od
        import mtgvision.util.image as uimg
    return ed, od, uimg


def reset_reference_shuffles(ed):
    """ApplyShuffled keeps its permutation between calls (util/random.py:88-97)."""
    ed.SyntheticBgFgMtgImages._RAN_BG.indices = [0, 1, 2]
    ed.SyntheticBgFgMtgImages._RAN_VRTL.indices = list(range(7))
