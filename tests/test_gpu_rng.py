"""The PRODUCTION randomness (device Philox sampler + device noise fields) against the reference's distributions.

The parity suite feeds the CUDA path the oracle's recorded draws; what `bench.py` and every user run instead is
`k_sample_tape` / `k_det_sample` (tapes drawn on the device) and `MTGV_FIELD_PHILOX` (Box-Muller / Poisson-by-inversion /
point sets drawn inside the kernels).  Bit parity is impossible there - the reference consumes Mersenne-Twister streams - so
these tests pin the DISTRIBUTIONS: every field of >= 50 000 device-sampled tapes is compared with the same field of as many
tapes recorded from the oracle (the reference's own `random` / `np.random` calls: encoder_datasets.py:62-351, 669-731,
encoder_train.py:149-230; od_datasets.py:322-332, 420-512, 558) with two-sample Kolmogorov-Smirnov / chi-square tests, and
every device noise generator is compared with numpy's on a constant image (util/image.py:434-488,
encoder_datasets.py:118-134, 222-240, 259-351).

Significance: all seeds are fixed, so outcomes are deterministic; the thresholds correspond to alpha ~ 1e-5 per comparison
(a few hundred comparisons), i.e. a real distribution error of a few percent in any field fails.
"""

from __future__ import annotations

import math
import os
import random
import subprocess
import sys

import numpy as np
import pytest

from mtgvision_b200 import abi, synth

pytestmark = pytest.mark.gpu

KS_C = 2.5        # D < KS_C * sqrt((n+m)/(n*m))  <=>  alpha ~ 7e-6
CHI2_ALPHA = 1e-6
from tests.rng_oracle_tapes import BG_HW, CARD_HW, DET_KW, N_BGS, N_CARDS, X_HW  # the oracle side's geometry  # noqa: E402


def ks_ok(a, b):
    a, b = np.sort(np.asarray(a, np.float64)), np.sort(np.asarray(b, np.float64))
    if len(a) == 0 or len(b) == 0:
        return len(a) == len(b), 0.0, 0.0
    allv = np.concatenate([a, b])
    d = np.abs(np.searchsorted(a, allv, side="right") / len(a) - np.searchsorted(b, allv, side="right") / len(b)).max()
    lim = KS_C * math.sqrt((len(a) + len(b)) / (len(a) * len(b)))
    return d < lim, float(d), lim


def chi2_ok(a, b):
    """Two-sample chi-square on the category counts of integer samples a and b."""
    from scipy.stats import chi2

    a, b = np.asarray(a).astype(np.int64), np.asarray(b).astype(np.int64)
    cats = np.union1d(a, b)
    ca = np.array([(a == c).sum() for c in cats], np.float64)
    cb = np.array([(b == c).sum() for c in cats], np.float64)
    if len(cats) <= 1:
        return True, 0.0, 1.0
    # pool rare categories so that expected counts stay >= 8
    order = np.argsort(ca + cb)
    ca, cb = ca[order], cb[order]
    while len(ca) > 2 and ca[0] + cb[0] < 16:
        ca[1] += ca[0]; cb[1] += cb[0]
        ca, cb = ca[1:], cb[1:]
        o = np.argsort(ca + cb)
        ca, cb = ca[o], cb[o]
    n, m = ca.sum(), cb.sum()
    stat = (((ca * math.sqrt(m / n) - cb * math.sqrt(n / m)) ** 2) / (ca + cb)).sum()
    p = float(chi2.sf(stat, len(ca) - 1))
    return p > CHI2_ALPHA, float(stat), p


def compare_columns(name, a, b, failures, checked):
    """a, b: 1-D samples of one tape field from the device and from the oracle."""
    a, b = np.asarray(a), np.asarray(b)
    if len(a) == 0 and len(b) == 0:
        return
    if len(a) == 0 or len(b) == 0:
        failures.append(f"{name}: present on one side only ({len(a)} vs {len(b)})")
        return
    if not a.any() and not b.any():
        return
    integer = np.issubdtype(a.dtype, np.integer)
    if integer and len(np.union1d(a, b)) <= 512:
        ok, stat, p = chi2_ok(a, b)
        checked.append(name)
        if not ok:
            failures.append(f"{name}: chi2 {stat:.1f}, p {p:.2e} (n {len(a)} / {len(b)})")
    else:
        ok, d, lim = ks_ok(a, b)
        checked.append(name)
        if not ok:
            failures.append(f"{name}: KS D {d:.4f} >= {lim:.4f} (n {len(a)} / {len(b)})")


# --------------------------------------------------------------------------------------- #
# encoder tapes                                                                            #
# --------------------------------------------------------------------------------------- #


def _oracle_npz(kind, n, tmp_path):
    """Run tests/rng_oracle_tapes.py in a fresh process (forking from this one is unsafe once CUDA and cv2 threads exist)."""
    out = os.path.join(str(tmp_path), f"oracle_{kind}.npz")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rng_oracle_tapes.py")
    subprocess.run([sys.executable, script, kind, str(n), out], check=True, timeout=600)
    return np.load(out)


def oracle_tapes(n_pairs, tmp_path):
    z = _oracle_npz("enc", n_pairs, tmp_path)
    return z["x"].view(abi.TAPE_DTYPE).reshape(-1), z["x2"].view(abi.TAPE_DTYPE).reshape(-1)


_SKIP_OP_FIELDS = {"n_field", "n_field2", "_pad", "field", "field2"}  # injected-field bookkeeping: not part of the draw


def compare_tapes(tag, dev, ora, failures, checked):
    for f in ("kind", "card", "swap_choice", "bg", "upsidedown", "n_fg", "n_bg", "n_vrtl"):
        compare_columns(f"{tag}.{f}", dev[f], ora[f], failures, checked)
    vd, vo = dev[dev["kind"] == abi.KIND_VIRTUAL], ora[ora["kind"] == abi.KIND_VIRTUAL]
    sections = {"fg": (lambda t: np.zeros(len(t), np.int64), "n_fg"),
                "bg": (lambda t: t["n_fg"].astype(np.int64), "n_bg"),
                "vrtl": (lambda t: (t["n_fg"] + t["n_bg"]).astype(np.int64), "n_vrtl")}
    for sec, (start_of, nkey) in sections.items():
        recs = {}
        for side, t in (("dev", vd), ("ora", vo)):
            start, cnt = start_of(t), t[nkey].astype(np.int64)
            codes_at = []
            for p in range(int(max(cnt.max(), 1))):
                has = cnt > p
                code = np.where(has, t["ops"]["code"][np.arange(len(t)), np.minimum(start + p, abi.MTGV_TAPE_MAX_OPS - 1)], -1)
                codes_at.append(code)
            recs[side] = (start, cnt, codes_at)
        # which op sits at position p of the section (pins the shuffles and the choice probabilities)
        for p in range(min(len(recs["dev"][2]), len(recs["ora"][2]))):
            compare_columns(f"{tag}.{sec}[{p}].code", recs["dev"][2][p], recs["ora"][2][p], failures, checked)
        # per opcode: every integer and double parameter
        codes = np.union1d(np.concatenate(recs["dev"][2]), np.concatenate(recs["ora"][2]))
        for code in codes:
            if code < 0:
                continue
            rows = {}
            for side, t in (("dev", vd), ("ora", vo)):
                start, cnt, _ = recs[side]
                sel = []
                for p in range(int(cnt.max())):
                    idx = np.nonzero((cnt > p) & (t["ops"]["code"][np.arange(len(t)), np.minimum(start + p, abi.MTGV_TAPE_MAX_OPS - 1)] == code))[0]
                    sel.append(t["ops"][idx, start[idx] + p])
                rows[side] = np.concatenate(sel) if sel else np.zeros(0, t["ops"].dtype)
            for k in range(rows["dev"]["i"].shape[1]):
                compare_columns(f"{tag}.{sec}.op{code}.i[{k}]", rows["dev"]["i"][:, k], rows["ora"]["i"][:, k], failures, checked)
            for k in range(rows["dev"]["d"].shape[1]):
                compare_columns(f"{tag}.{sec}.op{code}.d[{k}]", rows["dev"]["d"][:, k], rows["ora"]["d"][:, k], failures, checked)


def test_production_sampler_matches_oracle_distributions(tmp_path):
    import torch

    from mtgvision_b200.context import Context

    n_pairs = 25600  # 51 200 x-samples per side
    ora_x, ora_x2 = oracle_tapes(n_pairs, tmp_path)
    cards, bgs = synth.make_card_pool(N_CARDS, CARD_HW), synth.make_bg_pool(N_BGS, BG_HW)
    ctx = Context(0)
    ctx.set_encoder_config(x_size_hw=X_HW, y_size_hw=X_HW, target_is_input_prob=0.05, similar_neg_prob=0.2, half_upsidedown=True,
                           paired=True, targets=False)
    ctx.set_card_pool(cards.images, cards.labels3, cards.grp_off, cards.grp_mem)
    ctx.set_bg_pool(bgs)
    dev_x, dev_x2 = [], []
    batch = 32  # the x2 background is drawn from the batch (encoder_train.py:224): same batch size as the oracle run
    for b in range(n_pairs // batch):
        t = ctx.sample_encoder_tape(20261018, b * batch, batch)
        a = t.cpu().numpy().view(abi.TAPE_DTYPE).reshape(-1)
        dev_x.append(a[:batch].copy())
        dev_x2.append(a[batch:].copy())
        if b % 64 == 0:
            torch.cuda.synchronize()
    dev_x, dev_x2 = np.concatenate(dev_x), np.concatenate(dev_x2)
    assert len(dev_x) + len(dev_x2) >= 50000 and len(ora_x) + len(ora_x2) >= 50000
    failures, checked = [], []
    compare_tapes("x", dev_x, ora_x, failures, checked)
    compare_tapes("x2", dev_x2, ora_x2, failures, checked)
    # the hard-negative swap only exists where the card has same-name siblings: rate among those (similar_neg_prob = 0.2)
    has_sib = np.diff(cards.grp_off)[dev_x2["card"]] > 1
    rate_dev = (dev_x2["swap_choice"][has_sib] >= 0).mean()
    has_sib_o = np.diff(cards.grp_off)[ora_x2["card"]] > 1
    rate_ora = (ora_x2["swap_choice"][has_sib_o] >= 0).mean()
    se = math.sqrt(0.2 * 0.8 * (1 / has_sib.sum() + 1 / has_sib_o.sum()))
    assert abs(rate_dev - rate_ora) < 4.5 * se and abs(rate_dev - 0.2) < 4.5 * math.sqrt(0.16 / has_sib.sum()), (rate_dev, rate_ora)
    assert len(checked) > 150, len(checked)  # every opcode's parameters were really compared
    assert not failures, f"{len(failures)} of {len(checked)} distributions differ:\n" + "\n".join(failures[:40])
    ctx.close()


# --------------------------------------------------------------------------------------- #
# device noise fields                                                                      #
# --------------------------------------------------------------------------------------- #


def _xop(code, i=(), f=(), n_field=0, n_field2=0):
    a = np.zeros(1, dtype=abi.XOP_DTYPE)
    a["code"] = code
    a["field"] = a["field2"] = abi.MTGV_FIELD_PHILOX
    a["n_field"], a["n_field2"] = n_field, n_field2
    a["i"][0, : len(i)] = i
    a["f"][0, : len(f)] = f
    return a


def _run_device(ctx, level, op, n_img=6, H=192, W=128, seed=77):
    import torch

    img = torch.full((n_img, H, W, 3), float(level), dtype=torch.float32)
    out = ctx.run_plane_ops(img, op, fields=None, seed=seed)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_philox_noise_fields_match_numpy_distributions():
    from mtgvision_b200.context import Context
    from oracle import encoder_oracle as EO

    ctx = Context(0)
    H, W, n_img = 192, 128, 6
    np.random.seed(4242)
    random.seed(4242)
    failures = []

    def ref_images(fn, level):
        return np.stack([fn(np.full((H, W, 3), level, np.float32))[0] for _ in range(n_img)])

    def check(name, dev, ref, tol_mean=None):
        ok, d, lim = ks_ok(dev.reshape(-1)[::3], ref.reshape(-1)[::3])
        if not ok:
            failures.append(f"{name}: KS D {d:.4f} >= {lim:.4f}")
        sd = max(float(ref.std()), 1e-6) / math.sqrt(ref.size)
        if abs(float(dev.mean()) - float(ref.mean())) > 6 * sd * math.sqrt(2) + 1e-6:
            failures.append(f"{name}: mean {dev.mean():.6f} vs {ref.mean():.6f}")
        if abs(float(dev.var()) - float(ref.var())) > 0.03 * float(ref.var()) + 1e-7:
            failures.append(f"{name}: var {dev.var():.6f} vs {ref.var():.6f}")

    # Mutate.gaussian_noise: clip(img + N(0, 0.25)) (encoder_datasets.py:222-226)
    for level in (0.5, 0.1):
        dev = _run_device(ctx, level, _xop(abi.X_GAUSS_NOISE))
        check(f"gaussian_noise@{level}", dev, ref_images(EO.op_gaussian_noise, level))
        # three channels get independent fields
        assert abs(np.corrcoef(dev[..., 0].ravel(), dev[..., 1].ravel())[0, 1]) < 0.01

    # Mutate.noise, 4 kinds blended with ratio u * 0.5 (encoder_datasets.py:118-134, util/image.py:434-488)
    u = 0.6
    ra = np.float32(u * 0.5)
    rb = np.float32(1 - u * 0.5)
    n_sp = int(np.ceil(0.1 * (H * W * 3) * 0.5))
    for kind, nm in enumerate(("speckle", "gaussian", "pepper", "poisson")):
        for level in (0.5, 0.9):
            def ref_fn(img, kind=kind):
                # op_noise with the kind and ratio pinned (its own draws: field, then ratio)
                import unittest.mock as um

                with um.patch.object(EO, "_choice", lambda n: kind), um.patch.object(np.random, "random", lambda: u):
                    return EO.op_noise(img)
            dev = _run_device(ctx, level, _xop(abi.X_NOISE, i=(kind,), f=(ra, rb), n_field=n_sp, n_field2=n_sp))
            ref = ref_images(ref_fn, level)
            if kind == 2:
                # point sets: fractions of salted / peppered values (a value is hit with p ~ n / (H W 3), collisions included)
                for val in (level * rb + ra * 1.0, level * rb):
                    fd, fr = np.isclose(dev, val, atol=1e-6).mean(), np.isclose(ref, val, atol=1e-6).mean()
                    if abs(fd - fr) > 0.02 * fr + 5e-4:
                        failures.append(f"noise.pepper@{level}: fraction at {val:.3f} {fd:.5f} vs {fr:.5f}")
            else:
                check(f"noise.{nm}@{level}", dev, ref)

    # Mutate.salt_pepper_noise: ceil(1 % of img.size) salt points then as many pepper points, all channels (:228-240)
    n_pts = int(np.ceil(0.01 * H * W * 3))
    dev = _run_device(ctx, 0.5, _xop(abi.X_SALT_PEPPER, n_field=n_pts, n_field2=n_pts))
    ref = ref_images(EO.op_salt_pepper_noise, 0.5)
    for val in (0.0, 1.0):
        fd, fr = (dev == val).mean(), (ref == val).mean()
        if abs(fd - fr) > 0.03 * fr:
            failures.append(f"salt_pepper: fraction of {val} {fd:.5f} vs {fr:.5f}")
    assert (dev[..., 0] == dev[..., 1]).all() and (dev[..., 1] == dev[..., 2]).all()  # a point hits all channels
    # np.random.randint(0, i - 1): the last row and column are never hit
    assert (dev[:, H - 1] == 0.5).all() and (dev[:, :, W - 1] == 0.5).all()

    # Mutate.random_erasing, colour "random": uniform field inside the block only (:291-339)
    y0, y1, x0, x1 = 40, 120, 16, 100
    dev = _run_device(ctx, 0.5, _xop(abi.X_ERASE, i=(y0, y1, x0, x1, 0)))
    blk = dev[:, y0:y1, x0:x1]
    ok, d, lim = ks_ok(blk.ravel()[::2], np.random.uniform(0, 1, blk.size // 2))
    if not ok:
        failures.append(f"erase.random: KS D {d:.4f} >= {lim:.4f}")
    outside = dev.copy()
    outside[:, y0:y1, x0:x1] = 0.5
    assert (outside == 0.5).all()
    assert not failures, "\n".join(failures)
    ctx.close()


# --------------------------------------------------------------------------------------- #
# detection tapes                                                                          #
# --------------------------------------------------------------------------------------- #

def test_det_sampler_matches_oracle_distributions(tmp_path):
    from mtgvision_b200.context import Context

    ora = _oracle_npz("det", 12000, tmp_path)["t"].view(abi.DET_TAPE_DTYPE).reshape(-1)
    cards, bgs = synth.make_card_pool(N_CARDS, CARD_HW), synth.make_bg_pool(N_BGS, BG_HW)
    ctx = Context(0)
    ctx.set_card_pool(cards.images, cards.labels3, cards.grp_off, cards.grp_mem)
    ctx.set_bg_pool(bgs)
    kw = dict(DET_KW)
    ctx.set_det_config(card_min_area_ratio=0.02, card_max_area_ratio=0.9, card_no_contains=True, card_max_place_attempts=10,
                       photometrics=True, **kw)
    dev = ctx.sample_det_tape(99, 0, len(ora)).cpu().numpy().view(abi.DET_TAPE_DTYPE).reshape(-1)
    failures, checked = [], []
    for f in ("bg_only", "bg", "bg_deg", "n_pre", "n_post"):
        compare_columns(f"det.{f}", dev[f], ora[f], failures, checked)
    sd, so = dev[dev["bg_only"] == 0], ora[ora["bg_only"] == 0]
    compare_columns("det.n_cards", sd["n_cards"], so["n_cards"], failures, checked)
    # first placement attempt of the first card: (cx, cy), rotation, area, the four jitter draws (od_datasets.py:322-332)
    for f in ("cx", "cy", "deg", "area"):
        compare_columns(f"det.att0.{f}", sd["cards"]["att"][f][:, 0, 0], so["cards"]["att"][f][:, 0, 0], failures, checked)
    for k in range(4):
        compare_columns(f"det.att0.jitter[{k}]", sd["cards"]["att"]["jitter"][:, 0, 0, k], so["cards"]["att"]["jitter"][:, 0, 0, k], failures, checked)
    compare_columns("det.card0", sd["cards"]["card"][:, 0], so["cards"]["card"][:, 0], failures, checked)
    compare_columns("det.card0.n_photo", sd["cards"]["n_photo"][:, 0], so["cards"]["n_photo"][:, 0], failures, checked)
    # photometric programs: op code at each position + parameters per code, for the pre / post / per-card lists
    for lst, nkey, sel_d, sel_o in (("pre", "n_pre", dev, ora), ("post", "n_post", dev, ora)):
        for p in range(sel_d[lst].shape[1]):
            cd = np.where(sel_d[nkey] > p, sel_d[lst]["code"][:, p], -1)
            co = np.where(sel_o[nkey] > p, sel_o[lst]["code"][:, p], -1)
            compare_columns(f"det.{lst}[{p}].code", cd, co, failures, checked)
        for code in np.union1d(sel_d[lst]["code"], sel_o[lst]["code"]):
            rd = np.concatenate([sel_d[lst][sel_d[nkey] > p, p][sel_d[lst]["code"][sel_d[nkey] > p, p] == code] for p in range(sel_d[lst].shape[1])])
            ro = np.concatenate([sel_o[lst][sel_o[nkey] > p, p][sel_o[lst]["code"][sel_o[nkey] > p, p] == code] for p in range(sel_o[lst].shape[1])])
            if len(rd) == 0 and len(ro) == 0:
                continue
            for k in range(rd["d"].shape[1]):
                compare_columns(f"det.{lst}.op{code}.d[{k}]", rd["d"][:, k], ro["d"][:, k], failures, checked)
            for k in range(rd["i"].shape[1]):
                compare_columns(f"det.{lst}.op{code}.i[{k}]", rd["i"][:, k], ro["i"][:, k], failures, checked)
    assert len(checked) > 40, len(checked)
    assert not failures, f"{len(failures)} of {len(checked)} distributions differ:\n" + "\n".join(failures[:40])
    ctx.close()


def test_det_dataset_weights_and_uniform_size_mode():
    """Gen knobs (od_datasets.py:329-332, 656-672): the background DATASET is drawn with ilsvrc_vs_coco_sample_weights first,
    then an image uniformly inside it; card_size_sample_mode='uniform' draws the target area uniformly."""
    from mtgvision_b200.encoder_datasets import CocoValImages, IlsvrcImages, SyntheticBgFgMtgImages
    from mtgvision_b200.od_datasets import Gen

    cards = synth.make_card_pool(N_CARDS, CARD_HW)
    bgs = synth.make_bg_pool(16, BG_HW)
    n = 20000
    for weights, want in (((1.0, 1.0), 0.5), ((3.0, 1.0), 0.75), (None, 12 / 16)):
        gen = Gen(**DET_KW, ilsvrc_vs_coco_sample_weights=weights, card_size_sample_mode="uniform",
                  mtg_ds=SyntheticBgFgMtgImages(pool=cards), bg_ds=IlsvrcImages(images=bgs[:12]), bg2_ds=CocoValImages(images=bgs[12:]), seed=3)
        t = gen.ctx.sample_det_tape(gen.seed, 0, n).cpu().numpy().view(abi.DET_TAPE_DTYPE).reshape(-1)
        first = (t["bg"] < 12).mean()
        assert abs(first - want) < 4.5 * math.sqrt(want * (1 - want) / n), (weights, first)
        for lo, hi in ((0, 12), (12, 16)):  # uniform inside each dataset
            sel = t["bg"][(t["bg"] >= lo) & (t["bg"] < hi)]
            ok, stat, p = chi2_ok(sel, np.random.default_rng(0).integers(lo, hi, len(sel)))
            assert ok, (weights, lo, hi, stat, p)
        # uniform target area over [min_area_ratio, max_area_ratio] * S^2
        area = t["cards"]["att"]["area"][t["bg_only"] == 0][:, 0, 0]
        S2 = 320.0 * 320.0
        ok, d, lim = ks_ok(area, np.random.default_rng(1).uniform(0.02 * S2, 0.9 * S2, len(area)))
        assert ok, (d, lim)
        gen.ctx.close()
    with pytest.raises(KeyError):
        Gen(**DET_KW, card_size_sample_mode="cubic", mtg_ds=SyntheticBgFgMtgImages(pool=cards), bg_ds=IlsvrcImages(images=bgs))
