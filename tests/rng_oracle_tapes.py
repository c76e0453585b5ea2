"""Oracle side of tests/test_gpu_rng.py, run as its OWN process (`python tests/rng_oracle_tapes.py enc|det N OUT.npz`):
the reference's draws are recorded by running the oracle on small images in a pool of forked workers.  Forking from inside
the pytest process is not safe once CUDA, cv2 and the decoder's host threads are alive (a child can inherit a held lock and
hang), so the test starts this script with `subprocess` and reads the packed tapes back."""

from __future__ import annotations

import multiprocessing as mp
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mtgvision_b200 import synth  # noqa: E402

X_HW = (48, 32)   # small output: the sampler only sees H, W through the erase / cutout ranges and the noise point counts
CARD_HW, BG_HW = (68, 48), (30, 40)
N_CARDS, N_BGS = 24, 8
DET_KW = dict(bg_size_hw=320, num_cards_min=1, num_cards_max=6, card_min_visible_ratio=0.5, card_min_visible_ratio_edges=0.0,
              card_jitter_ratio=0.7, ratio_bg=0.1, kind="obb")


def _enc_worker(args):
    wid, n_batches, batch = args
    import cv2

    from oracle import encoder_oracle as EO
    from oracle import tape_pack

    cv2.setNumThreads(1)
    random.seed(9000 + wid)
    np.random.seed(9000 + wid)
    EO.reset_shuffle_state()
    cards, bgs = synth.make_card_pool(N_CARDS, CARD_HW), synth.make_bg_pool(N_BGS, BG_HW)
    bo = EO.BatchOracle(cards, bgs, paired=True, targets=False, x_size_hw=X_HW, half_upsidedown=True,
                        target_is_input_prob=0.05, similar_neg_prob=0.2)
    xs, x2s = [], []
    for _ in range(n_batches):
        _, _, tp = bo.random_image_batch(batch)
        xs += tp["x"]
        x2s += tp["x2"]
    ax, _ = tape_pack.pack_tapes(xs, host_transcendentals=False)
    ax2, _ = tape_pack.pack_tapes(x2s, host_transcendentals=False)
    return ax, ax2


def encoder_tapes(n_pairs, batch=32, seed_shift=0):
    workers = min(os.cpu_count() or 1, 16)
    n_batches = (n_pairs + batch * workers - 1) // (batch * workers)
    with mp.get_context("fork").Pool(workers) as pool:
        parts = pool.map(_enc_worker, [(w + seed_shift, n_batches, batch) for w in range(workers)])
    return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])


def _det_worker(args):
    wid, n = args
    import cv2

    from oracle import det_oracle as DO
    from oracle import tape_pack

    cv2.setNumThreads(1)
    random.seed(500 + wid)
    np.random.seed(500 + wid)
    cards = synth.make_card_pool(N_CARDS, CARD_HW)
    bgs = synth.make_bg_pool(N_BGS, BG_HW)
    o = DO.DetOracle([cards.images[k] for k in range(N_CARDS)], bgs, **DET_KW)
    tapes = []
    for _ in range(n):
        t = {}
        o.random(t)
        tapes.append(t)
    arr, _ = tape_pack.pack_det_tapes(tapes, host_transcendentals=False)
    return arr


def det_tapes(n):
    workers = min(os.cpu_count() or 1, 16)
    per = (n + workers - 1) // workers
    with mp.get_context("fork").Pool(workers) as pool:
        return np.concatenate(pool.map(_det_worker, [(w, per) for w in range(workers)]))


if __name__ == "__main__":
    kind, n, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    if kind == "enc":
        x, x2 = encoder_tapes(n)
        np.savez(out, x=x.view(np.uint8), x2=x2.view(np.uint8))
    else:
        np.savez(out, t=det_tapes(n).view(np.uint8))
