"""Detection path parity on the GPU, through the C ABI, against oracle/det_oracle.py on the
oracle's recorded tapes (identical sampled parameters, identical inputs).

Bars: accept/reject decisions, homographies-derived keypoints and labels bit-exact; pixels within
+-1 LSB of uint8 (without photometrics the float32 image is expected to be exact: same warp
arithmetic, same composite order)."""
import ctypes as C
import os
import random

import numpy as np
import pytest

from mtgvision_b200 import abi, synth
from oracle import det_oracle as DO
from oracle import tape_pack
from tests import parity_util as PU

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

AUTHOR_RUN = dict(bg_size_hw=640, num_cards_min=1, num_cards_max=9, card_min_visible_ratio=0.5,
                  card_min_visible_ratio_edges=0.0, card_jitter_ratio=0.7)  # od_datasets.py:861-868


@pytest.fixture(scope="module")
def env():
    pool, bgs = PU.small_pools(8, 8)
    from mtgvision_b200.context import Context

    ctx = Context(0)
    ctx.set_card_pool(pool.images, pool.labels3, pool.grp_off, pool.grp_mem)
    ctx.set_bg_pool(bgs)
    yield pool, bgs, ctx
    ctx.close()


def oracle_scenes(pool, bgs, kind, seeds, photometrics, **kw):
    out = []
    for seed in seeds:
        random.seed(seed); np.random.seed(seed)
        o = DO.DetOracle(list(pool.images), bgs, kind=kind, photometrics=photometrics, **{**AUTHOR_RUN, **kw})
        t = {}
        out.append((t, o.generate(t)))
    return out


def run_gpu(ctx, tapes, kind, photometrics, out_dtype=abi.OUT_F32, host_transcendentals=True, **kw):
    cfgkw = {**AUTHOR_RUN, **kw}
    ctx.set_det_config(kind=kind, photometrics=photometrics, **cfgkw)
    arr, fields = tape_pack.pack_det_tapes(tapes, host_transcendentals=host_transcendentals)
    f = torch.from_numpy(fields.to_array().view(np.int32)).to(ctx.device)
    params, accepted, kps, labels, counts = ctx.det_place(ctx.upload_det_tape(arr))
    img = ctx.det_batch(params, out_dtype, fields=f)
    torch.cuda.synchronize()
    return (img.permute(0, 2, 3, 1).contiguous().cpu().numpy(), accepted.cpu().numpy(), kps.cpu().numpy(),
            labels.cpu().numpy(), counts.cpu().numpy())


@pytest.mark.parametrize("kind", ["obb", "seg"])
def test_placement_labels_and_pixels_without_photometrics(env, kind):
    pool, bgs, ctx = env
    scenes = oracle_scenes(pool, bgs, kind, range(24), photometrics=False)
    img, accepted, kps, labels, counts = run_gpu(ctx, [t for t, _ in scenes], kind, False)
    worst = 0.0
    for s, (t, sample) in enumerate(scenes):
        for ci, c in enumerate(t["cards"]):
            want = len(c["attempts"]) - 1 if c["attempts"][-1]["accepted"] else -1
            assert accepted[s, ci] == want, f"scene {s} card {ci}: accept/reject differs"
        k = sample["keypoints"].shape[0]
        P = sample["keypoints"].shape[1] if k else 0
        assert counts[s] == k
        assert np.array_equal(kps[s, :k, :P], sample["keypoints"].reshape(k, P, 2))
        assert np.array_equal(labels[s, :k], sample["keypoints_labels"])
        worst = max(worst, float(np.abs(img[s] - sample["image"]).max()))
    assert worst <= 1e-6, f"max abs pixel error {worst}"  # expected 0: identical arithmetic


def test_pixels_with_photometrics_within_one_lsb(env):
    pool, bgs, ctx = env
    scenes = oracle_scenes(pool, bgs, "obb", range(100, 132), photometrics=True)
    seen = set()
    for t, _ in scenes:
        for rec in t["pre"] + t["post"] + [r for c in t["cards"] for r in c.get("photo", [])]:
            seen.add(rec["ph"])
    assert {DO.PH_RBC, DO.PH_HSV, DO.PH_GAUSS_NOISE, DO.PH_GAUSS_BLUR, DO.PH_ERASE} <= seen
    img, accepted, kps, labels, counts = run_gpu(ctx, [t for t, _ in scenes], "obb", True)
    worst = [PU.lsb_diff(img[s], sample["image"])[0] for s, (t, sample) in enumerate(scenes)]
    assert max(worst) <= 1, f"max uint8 LSB error {max(worst)} at scene {int(np.argmax(worst))}"
    u8, *_ = run_gpu(ctx, [t for t, _ in scenes], "obb", True, out_dtype=abi.OUT_U8)
    for s, (t, sample) in enumerate(scenes):  # uint8 output follows imwrite: (img*255).astype(uint8) (util/image.py:101-104)
        ref = (np.clip(sample["image"], 0, 1) * 255).astype(np.uint8).astype(np.int32)
        assert np.abs(u8[s].astype(np.int32) - ref).max() <= 1


def test_device_transcendentals_change_nothing_discrete(env):
    """Production mode derives the corner targets with CUDA libm instead of taking the host's float32
    values: decisions and keypoints may move by float32 rounding of dst only, never structurally."""
    pool, bgs, ctx = env
    scenes = oracle_scenes(pool, bgs, "obb", range(300, 316), photometrics=False)
    tapes = [t for t, _ in scenes]
    a = run_gpu(ctx, tapes, "obb", False, host_transcendentals=True)
    b = run_gpu(ctx, tapes, "obb", False, host_transcendentals=False)
    same = (a[1] == b[1]).mean()
    assert same > 0.97
    both = (a[4] == b[4])
    assert np.abs(a[2][both] - b[2][both]).max() < 0.05  # keypoints agree to a few hundredths of a pixel


def _scene_with_post(pool, bgs, seed, post_makers, **kw):
    """A card scene of the oracle without photometrics, then a hand-picked post-augment list applied like
    post_transform_bg (od_datasets.py:603-606); returns (tape with that list, final image)."""
    random.seed(seed); np.random.seed(seed)
    o = DO.DetOracle(list(pool.images), bgs, kind="obb", photometrics=False, **{**AUTHOR_RUN, **kw})
    t = {}
    img = o.generate(t)["image"]
    post = [m() for m in post_makers]
    for rec in post:
        img = DO.apply_photo(img, rec)
    t["post"] = post
    return t, img


def test_remaining_albumentations_ops_within_one_lsb(env):
    """ISONoise, ShotNoise, MedianBlur, MotionBlur, GlassBlur (od_datasets.py:443-457) - each alone and chained (up to
    eight pass boundaries + pointwise ops), noise fields injected from the oracle.  The oracle's inner kernels are the real cv2
    calls (cvtColor RGB<->HLS, meanStdDev, pow, medianBlur, line + filter2D, GaussianBlur); GlassBlur's shuffle is numpy's."""
    pool, bgs, ctx = env
    cases = {
        "iso": [lambda: DO.draw_iso_noise((0.01, 0.4))],
        "shot": [lambda: DO.draw_shot_noise((0.1, 0.3))],
        "median3": [lambda: {"ph": DO.PH_MEDIAN_BLUR, "ksize": 3}],
        "median5": [lambda: {"ph": DO.PH_MEDIAN_BLUR, "ksize": 5}],
        "median7": [lambda: {"ph": DO.PH_MEDIAN_BLUR, "ksize": 7}],
        "motion": [lambda: DO.draw_motion_blur((3, 11))],
        "motion11": [lambda: {"ph": DO.PH_MOTION_BLUR, "ksize": 11, "pts": (0, 10, 10, 3), "mask": DO.motion_kernel_mask(11, 0, 10, 10, 3)}],
        "glass": [lambda: DO.draw_glass_blur((640, 640))],
        "glass_chain": [lambda: DO.draw_glass_blur((640, 640)), lambda: DO.draw_hsv((-30, 30), (-40, 40), (0, 0)), lambda: DO.draw_glass_blur((640, 640))],
        "chain": [lambda: DO.draw_motion_blur((3, 11)), lambda: DO.draw_rbc((-0.4, 0.4), (-0.5, 0.5)), lambda: DO.draw_iso_noise((0.01, 0.4)),
                  lambda: {"ph": DO.PH_MEDIAN_BLUR, "ksize": 5}, lambda: DO.draw_shot_noise((0.1, 0.3))],
        "chain2": [lambda: DO.draw_iso_noise((0.01, 0.4)), lambda: DO.draw_gauss_blur((1.0, 3.0)), lambda: DO.draw_iso_noise((0.01, 0.4)),
                   lambda: DO.draw_motion_blur((3, 11))],
    }
    tapes, refs, names = [], [], []
    for k, (name, makers) in enumerate(cases.items()):
        for rep in range(2):
            t, img = _scene_with_post(pool, bgs, 4000 + 10 * k + rep, makers)
            tapes.append(t); refs.append(img); names.append(f"{name}/{rep}")
    img, *_ = run_gpu(ctx, tapes, "obb", True)
    for s, (name, ref) in enumerate(zip(names, refs)):
        # (median: a 1e-7 difference in front of rint() moves a quantised neighbour by at most one byte value, and the median
        # of values that each move by at most 1 moves by at most 1)
        lsb, _ = PU.lsb_diff(img[s], ref)
        assert lsb <= 1, f"{name}: max uint8 LSB error {lsb}"


def test_remaining_albumentations_ops_production_fields(env):
    """The same ops with DEVICE-generated fields (Philox): ISONoise's Poisson rate comes from the device's own cv2.meanStdDev
    (k_det_stats), ShotNoise's from the pixel; the output distributions must match the oracle's numpy draws on the same scene."""
    from tests.test_gpu_rng import ks_ok

    pool, bgs, ctx = env
    for name, maker in (("iso", lambda: DO.draw_iso_noise((0.2, 0.4), (0.3, 0.5))), ("shot", lambda: DO.draw_shot_noise((0.1, 0.3)))):
        t, ref = _scene_with_post(pool, bgs, 4100, [maker])
        for rec in t["post"]:  # forget the oracle's fields: the device draws its own
            for k in ("lum", "col", "field"):
                rec.pop(k, None)
        img, *_ = run_gpu(ctx, [t], "obb", True)
        for c in range(3):
            # rounded: ShotNoise outputs are atoms (k * scale)^(1/2.2) that powf and cv2.pow place 1e-7 apart
            ok, d, lim = ks_ok(np.round(img[0][::2, ::2, c].ravel(), 4), np.round(ref[::2, ::2, c].ravel(), 4))
            assert ok, f"{name} channel {c}: KS D {d:.4f} >= {lim:.4f}"
        assert abs(float(img[0].mean()) - float(ref.mean())) < 2e-3, name


def test_config4_1280_up_to_32_cards_matches_oracle(env):
    """BASELINE config 4 against the oracle (od_datasets.py:520-611): 1280x1280 scenes, num_cards_max=33 (up to 32 cards),
    photometrics on - the <= 32-card placement loop, the 1280^2 tile lists and the multi-pass blur scratch that 640^2 / 8
    cards does not stress.  Decisions, keypoints and labels exact; pixels within 1 uint8 LSB."""
    pool, bgs, ctx = env
    kw = dict(bg_size_hw=1280, num_cards_max=33)
    for kind, seeds in (("obb", (900, 901, 902)), ("seg", (903, 904, 905))):
        scenes = oracle_scenes(pool, bgs, kind, seeds, photometrics=True, **kw)
        assert max(t["n_cards"] for t, _ in scenes) >= 12  # the seeds do reach many-card scenes
        img, accepted, kps, labels, counts = run_gpu(ctx, [t for t, _ in scenes], kind, True, **kw)
        for s, (t, sample) in enumerate(scenes):
            for ci, c in enumerate(t["cards"]):
                want = len(c["attempts"]) - 1 if c["attempts"][-1]["accepted"] else -1
                assert accepted[s, ci] == want, f"{kind} scene {s} card {ci}: accept/reject differs"
            k = sample["keypoints"].shape[0]
            P = sample["keypoints"].shape[1] if k else 0
            assert counts[s] == k
            assert np.array_equal(kps[s, :k, :P], sample["keypoints"].reshape(k, P, 2))
            assert np.array_equal(labels[s, :k], sample["keypoints_labels"])
            lsb, _ = PU.lsb_diff(img[s], sample["image"])
            assert lsb <= 1, f"{kind} scene {s}: max uint8 LSB error {lsb}"
    # without photometrics the float32 image is the reference's own arithmetic: exact
    scenes = oracle_scenes(pool, bgs, "obb", (906, 907), photometrics=False, **kw)
    img, accepted, *_ = run_gpu(ctx, [t for t, _ in scenes], "obb", False, **kw)
    assert max(float(np.abs(img[s] - sample["image"]).max()) for s, (t, sample) in enumerate(scenes)) <= 1e-6


def test_full_size_1280_32_cards_properties(env):
    """BASELINE config 4 shape: 1280x1280, up to 32 cards, photometrics on, production sampler."""
    pool, bgs, ctx = env
    ctx.set_det_config(bg_size_hw=1280, num_cards_min=1, num_cards_max=33, card_min_visible_ratio=0.5,
                       card_min_visible_ratio_edges=0.0, card_jitter_ratio=0.7, ratio_bg=0.1, kind="seg", photometrics=True)
    tape = ctx.sample_det_tape(99, 0, 16)
    params, accepted, kps, labels, counts = ctx.det_place(tape)
    img = ctx.det_batch(params, abi.OUT_U8)
    img2 = ctx.det_batch(params, abi.OUT_U8)
    torch.cuda.synchronize()
    assert img.shape == (16, 3, 1280, 1280) and img.dtype == torch.uint8
    assert torch.equal(img, img2)  # idempotent
    t = tape.cpu().numpy().view(abi.DET_TAPE_DTYPE).reshape(-1)
    acc = accepted.cpu().numpy(); cnt = counts.cpu().numpy(); lab = labels.cpu().numpy()
    for s in range(16):
        placed = int((acc[s] >= 0).sum())
        assert placed <= t["n_cards"][s] <= 32
        assert cnt[s] == placed  # seg: one polygon per card
        assert np.all(lab[s, :cnt[s]] == 0) and np.all(lab[s, cnt[s]:] == -1)
        if t["bg_only"][s]:
            assert placed == 0
    assert (acc >= 0).sum() > 16
    again = ctx.sample_det_tape(99, 0, 16)
    assert torch.equal(tape, again)  # deterministic per (seed, index)
    shard = ctx.sample_det_tape(99, 8, 8).cpu().numpy().view(abi.DET_TAPE_DTYPE).reshape(-1)
    assert np.array_equal(shard["bg"], t["bg"][8:]) and np.array_equal(shard["n_cards"], t["n_cards"][8:])


def test_gen_dropin_surface():
    from mtgvision_b200.encoder_datasets import IlsvrcImages, SyntheticBgFgMtgImages
    from mtgvision_b200.od_datasets import Gen

    pool, bgs = PU.small_pools(8, 8)
    gen = Gen(card_min_visible_ratio=0.5, card_min_visible_ratio_edges=0.0, card_jitter_ratio=0.7, ratio_bg=0.1, kind="seg",
              mtg_ds=SyntheticBgFgMtgImages(pool=pool), bg_ds=IlsvrcImages(images=bgs), seed=5)
    s = gen.random()
    assert s["image"].shape == (640, 640, 3) and s["image"].dtype == np.float32
    assert len(s["keypoints"]) == len(s["keypoints_labels"])
    b = gen.random_batch(8)
    assert b["image"].shape == (8, 3, 640, 640) and b["image"].dtype == torch.uint8 and b["image"].is_cuda
    bg = gen.random_bg()
    assert len(bg["keypoints"]) == 0 and bg["image"].shape == (640, 640, 3)
    # pipelined host form: the same scenes as the same sequence of random_batch calls (ragged batch sizes), in pinned host memory
    mk = lambda: Gen(card_min_visible_ratio=0.5, card_min_visible_ratio_edges=0.0, card_jitter_ratio=0.7, ratio_bg=0.1, kind="seg",  # noqa: E731
                     mtg_ds=SyntheticBgFgMtgImages(pool=pool), bg_ds=IlsvrcImages(images=bgs), seed=6)
    g_a, g_b = mk(), mk()
    sizes = [5, 5, 3, 5]
    got = [{k: v.clone() for k, v in r.items()} for r in g_a.host_batches(sizes)]
    assert len(got) == len(sizes)
    for n, r in zip(sizes, got):
        want = g_b.random_batch(n)
        assert set(r) == set(want)
        for k in want:
            assert not r[k].is_cuda
            assert torch.equal(r[k], want[k].cpu()), k
    with pytest.raises(ValueError):  # randrange empty range, like the reference with the default edge ratio
        Gen(mtg_ds=SyntheticBgFgMtgImages(pool=pool), bg_ds=IlsvrcImages(images=bgs))
