"""Parity of the CUDA encoder path against the oracle away from the benchmark geometry: ragged
background pools, other card / output sizes (wider cards, more INTER_AREA taps, canvases beyond the
kernels' fixed-point tables, photograph-sized backgrounds), empty batches and the documented size limits.  Same bar as
test_gpu_encoder.py: labels exact, pixels within 1 uint8 LSB."""
import numpy as np
import pytest

from mtgvision_b200 import abi, synth
from oracle import encoder_oracle as EO
from tests import parity_util as PU

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

MAX_LSB = 1


def _white_noise_bg(seed, hw):
    return np.random.default_rng(seed).integers(0, 256, (*hw, 3), dtype=np.uint8)


CASES = {
    # name: (card_hw, [bg_hw, ...], x_size_hw)
    "ragged_backgrounds": ((680, 488), [(375, 500), (480, 640), (333, 500), (500, 375), (256, 256), (427, 640)], (192, 128)),
    "scryfall_small_cards": ((204, 146), [(375, 500), (300, 400)], (192, 128)),
    "large_backgrounds_no_tables": ((680, 488), [(600, 800), (640, 854)], (192, 128)),
    "small_output_many_taps": ((680, 488), [(375, 500), (300, 400)], (128, 96)),  # INTER_AREA scales up to 5.3 (limit 6)
    "wide_cards_two_chunks": ((936, 672), [(375, 500), (500, 375)], (192, 128)),
    "square_output": ((680, 488), [(375, 500), (400, 400)], (160, 160)),
    # backgrounds smaller than the output: crop_to_size enlarges with cv2's 2-tap "area-linear" INTER_AREA kernel
    "tiny_backgrounds_enlarged": ((680, 488), [(90, 120), (60, 50), (128, 100), (150, 127), (40, 333)], (192, 128)),
    # full-size photographs (ILSVRC holds them): INTER_AREA reductions of x7 .. x16, past the register-resident tap lists
    "photograph_sized_backgrounds": ((680, 488), [(1000, 1500), (1536, 2048)], (192, 128)),
    "vga_background_small_output": ((340, 244), [(480, 640), (768, 1024)], (96, 64)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_virtual_samples_other_geometries(name):
    card_hw, bg_hws, x_hw = CASES[name]
    pool = synth.make_card_pool(4, card_hw)
    # one low-pass and one white-noise background per size: the 1/32-px coordinate grid matters most on noise
    bgs = []
    for j, hw in enumerate(bg_hws):
        bgs.append(synth.synth_bg(j, hw))
        bgs.append(_white_noise_bg(100 + j, hw))
    ctx = PU.make_context(pool, bgs, half_upsidedown=True, x_size_hw=x_hw, y_size_hw=x_hw)
    try:
        refs, tapes = [], []
        for seed in range((2 if "photograph" in name else 3) * len(bgs)):
            img, t = PU.oracle_virtual(pool, bgs, 7000 + seed, seed % 4, seed % len(bgs), size_hw=x_hw)
            refs.append(img)
            tapes.append(t)
        for k in range(2):  # make_cropped at this geometry
            t = {"card": k, "bg": 0}
            ref = EO.make_cropped(EO.u8_to_f32(pool.images[k]), x_hw, tape=t)
            if k:  # rot180 of the crop (make_cropped(half_upsidedown=True) drawing upsidedown)
                t["upsidedown"] = True
                ref = ref[::-1, ::-1]
            refs.append(ref)
            tapes.append(t)
        out, labels, p = PU.gpu_run_tapes(ctx, tapes, abi.OUT_F32)
        assert np.all(p["status"] == 0), p["status"]
        assert out.shape == (len(tapes), x_hw[0], x_hw[1], 3)
        worst = [PU.lsb_diff(o, r)[0] for o, r in zip(out, refs)]
        assert max(worst) <= MAX_LSB, f"{name}: max uint8 LSB error {max(worst)} at sample {int(np.argmax(worst))}"
        want = np.asarray([pool.labels3[t["card"]] for t in tapes], dtype=np.int64)
        assert np.array_equal(labels, want)
        out16, _, _ = PU.gpu_run_tapes(ctx, tapes[:4], abi.OUT_F16)
        assert max(PU.lsb_diff(o, r)[0] for o, r in zip(out16, refs[:4])) <= MAX_LSB
    finally:
        ctx.close()


def test_empty_batch_and_limits():
    pool, bgs = PU.small_pools(4, 4)
    ctx = PU.make_context(pool, bgs)
    try:
        # n = 0: every entry point is a no-op that succeeds
        tape = ctx.sample_encoder_tape(1, 0, 0)
        params, labels = ctx.expand_params(tape)
        out = ctx.encoder_batch(params, abi.OUT_F16)
        assert out.shape[0] == 0 and labels.shape[0] == 0
        # odd batch sizes (no pair partner to alias), single sample
        for n in (1, 3):
            tape = ctx.sample_encoder_tape(5, 10, n)
            params, labels = ctx.expand_params(tape)
            x = ctx.encoder_batch(params, abi.OUT_U8)
            torch.cuda.synchronize()
            assert x.shape[0] == 2 * n and bool((x.float().mean(dim=(1, 2, 3)) > 1).all())
    finally:
        ctx.close()
    # x_size_hw whose two float32 planes do not fit shared memory: a clean error, not a crash
    ctx = PU.make_context(pool, bgs, x_size_hw=(256, 192), y_size_hw=(256, 192))
    try:
        tape = ctx.sample_encoder_tape(1, 0, 2)
        params, _ = ctx.expand_params(tape)
        with pytest.raises(abi.MtgvError):
            ctx.encoder_batch(params, abi.OUT_F16)
    finally:
        ctx.close()


def test_pair_partner_plane_reuse_is_invisible():
    """x2 reuses x's area-resized card planes when the inputs match (k_foreground skips them): the second half
    of a batch must equal the same samples generated without a partner in front of them."""
    pool, bgs = PU.small_pools(8, 8)
    ctx = PU.make_context(pool, bgs, similar_neg_prob=0.5)
    try:
        tape = ctx.sample_encoder_tape(99, 0, 16)
        params, _ = ctx.expand_params(tape)
        both = ctx.encoder_batch(params, abi.OUT_F32)
        second_alone = ctx.encoder_batch(params[16:].contiguous(), abi.OUT_F32)  # 16 samples: partners are x2 samples, cards differ
        torch.cuda.synchronize()
        assert torch.equal(both[16:], second_alone)
    finally:
        ctx.close()


def test_oversized_background_is_refused_at_ingest():
    """A background whose rotated canvas can exceed the x24 INTER_AREA limit is refused when the dataset is built (no sample
    can then fail for its size at run time); full-size photographs below the limit are accepted without a warning."""
    import warnings

    from mtgvision_b200.encoder_datasets import IlsvrcImages, SyntheticBgFgMtgImages
    from mtgvision_b200.encoder_train import RanMtgEncDecDataset

    pool, bgs = PU.small_pools(4, 4)
    huge = np.zeros((3000, 4000, 3), np.uint8)  # diagonal 5000 px: x26 at 192 rows
    with pytest.raises(ValueError, match="INTER_AREA"):
        RanMtgEncDecDataset(2, paired=True, targets=False, mtg=SyntheticBgFgMtgImages(pool=pool), ilsvrc=IlsvrcImages(images=list(bgs) + [huge]), seed=1)
    photo = np.random.default_rng(0).integers(0, 256, (1500, 2000, 3), dtype=np.uint8)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        ds = RanMtgEncDecDataset(8, paired=True, targets=False, mtg=SyntheticBgFgMtgImages(pool=pool), ilsvrc=IlsvrcImages(images=[photo] * 2), seed=1,
                                 check_data=True)
    b = ds.random_tensor_batch()  # check_data=True raises on any failed sample
    assert b["x"].shape == (8, 3, 192, 128)
    ds.ctx.close()
