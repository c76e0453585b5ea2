"""Host-side helpers of the file-fed path that need no GPU: mtgv_gather_files (C ABI, context-free) and the one-ahead
item preparation of RanMtgEncDecDataset.lookahead."""
import ctypes as C
import threading
import time

import numpy as np

from mtgvision_b200 import abi


def test_gather_files_without_a_context_equals_join():
    lib = abi.load_library()
    rng = np.random.default_rng(5)
    for sizes in ([0, 1, 0, 7, 300, 0], [5 << 20, 0, 3, 1 << 20, 0, 0, 6 << 20, 17], [1 << 15] * 300, []):
        files = [rng.integers(0, 256, size=s, dtype=np.uint8).tobytes() for s in sizes]
        n = len(files)
        lens = np.asarray(sizes, dtype=np.int64).reshape(n)
        want = b"".join(files)
        dst = np.full(len(want) + 8, 0xAB, dtype=np.uint8)
        off = np.zeros(n + 1, dtype=np.int64)
        srcs = (C.c_char_p * max(n, 1))(*files)
        rc = lib.mtgv_gather_files(None, C.cast(srcs, C.c_void_p), lens.ctypes.data_as(C.c_void_p), n, dst.ctypes.data_as(C.c_void_p),
                                   len(want), off.ctypes.data_as(C.c_void_p))
        assert rc == 0
        assert dst[: len(want)].tobytes() == want and np.all(dst[len(want):] == 0xAB)  # nothing written past the end
        assert np.array_equal(off, np.concatenate([[0], np.cumsum(lens)]).astype(np.int64))
        if n:
            assert lib.mtgv_gather_files(None, C.cast(srcs, C.c_void_p), lens.ctypes.data_as(C.c_void_p), n, dst.ctypes.data_as(C.c_void_p),
                                         len(want) - 1, off.ctypes.data_as(C.c_void_p)) != 0 or len(want) == 0  # destination too small


def test_lookahead_prepares_exactly_one_ahead_in_order():
    from mtgvision_b200.encoder_train import RanMtgEncDecDataset

    started, main = [], threading.get_ident()
    tids = set()

    def thunk(i):
        def run():
            started.append(i)
            tids.add(threading.get_ident())
            time.sleep(0.01)
            return i
        return run

    seen = []
    for v in RanMtgEncDecDataset.lookahead(thunk(i) for i in range(6)):
        time.sleep(0.03)  # the consumer is slower: the helper must not run further ahead than the next item
        seen.append(v)
        assert max(started) <= v + 1
    assert seen == list(range(6)) and started == list(range(6))
    assert main not in tids  # prepared on the helper thread
    assert list(RanMtgEncDecDataset.lookahead(iter(()))) == []
