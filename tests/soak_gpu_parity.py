"""One-off soak (not collected by pytest): many more oracle seeds through the CUDA encoder path than the
regular suite, reporting the LSB error histogram.  python tests/soak_gpu_parity.py [n_seeds]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mtgvision_b200 import abi
from tests import parity_util as PU

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
pool, bgs = PU.small_pools(8, 8)
rng = np.random.default_rng(0)
bgs = list(bgs) + [rng.integers(0, 256, (375, 500, 3), dtype=np.uint8) for _ in range(4)]  # white noise: worst case for coordinates
ctx = PU.make_context(pool, bgs, half_upsidedown=True)
hist = np.zeros(4, dtype=np.int64)
worst_f = 0.0
for base in range(0, n, 100):
    refs, tapes = [], []
    for seed in range(base, min(n, base + 100)):
        img, t = PU.oracle_virtual(pool, bgs, 10000 + seed, seed % 8, seed % len(bgs))
        refs.append(img); tapes.append(t)
    out, _, p = PU.gpu_run_tapes(ctx, tapes, abi.OUT_F32)
    assert np.all(p["status"] == 0)
    for o, r in zip(out, refs):
        g8 = np.rint(np.clip(o, 0, 1) * 255).astype(np.int32); r8 = np.rint(np.clip(r, 0, 1) * 255).astype(np.int32)
        d = np.abs(g8 - r8)
        hist += np.bincount(np.minimum(d.ravel(), 3), minlength=4)
        worst_f = max(worst_f, float(np.abs(o - r).max()))
tot = hist.sum()
print({"samples": n, "values": int(tot), "lsb0": float(hist[0] / tot), "lsb1": float(hist[1] / tot), "lsb2": int(hist[2]), "lsb3plus": int(hist[3]),
       "max_abs_float_err": worst_f, "max_abs_float_err_lsb": worst_f * 255})
