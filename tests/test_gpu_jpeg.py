"""Device JPEG decode through the C ABI against cv2.imdecode itself: bit-exact."""
import cv2
import numpy as np
import pytest

from tests import jpeg_cases

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _ref(data):
    return cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR_RGB)


def _check(ctx, files, names=None):
    flat, off, hw = ctx.decode_jpegs(files)
    flat = flat.cpu().numpy()
    for i, data in enumerate(files):
        ref = _ref(data)
        h, w = hw[i]
        assert (h, w) == ref.shape[:2]
        got = flat[off[i]: off[i] + 3 * h * w].reshape(h, w, 3)
        assert np.array_equal(got, ref), (names[i] if names else i, int(np.abs(got.astype(int) - ref).max()))


def test_small_suite_bit_exact():
    from mtgvision_b200.context import Context

    ctx = Context(0)
    cases = jpeg_cases.small_suite()
    _check(ctx, [d for _, d in cases], [n for n, _ in cases])
    ctx.close()


def test_pool_sized_batches_bit_exact():
    """375x500 backgrounds and 680x488 cards, with and without restart intervals, all lanes-per-warp regimes."""
    from mtgvision_b200.context import Context

    ctx = Context(0)
    rng = np.random.default_rng(5)
    files = []
    for k in range(24):
        hw = (375, 500) if k % 3 else (680, 488)
        files.append(jpeg_cases.encode(jpeg_cases.image(rng, *hw, ["mixed", "smooth", "noise"][k % 3]), [85, 92, 75][k % 3],
                                       ["420", "444", "422", "440"][k % 4], rst=[0, 0, 4, 1][k % 4], optimize=k % 2))
    _check(ctx, files)
    _check(ctx, files[:1])
    # long scans: many subsequences per thread of the parallel entropy decoder; flat image: very short scan
    big = [jpeg_cases.encode(jpeg_cases.image(rng, 680, 488, "noise"), 100, "444", optimize=1),
           jpeg_cases.encode(jpeg_cases.image(rng, 1080, 1920, "mixed"), 95, "420"),
           jpeg_cases.encode(np.full((375, 500, 3), 128, np.uint8), 90, "420")]
    _check(ctx, big)
    # many restart intervals: decoders share warps
    many = [jpeg_cases.encode(jpeg_cases.image(rng, 375, 500, "mixed"), 80, "420", rst=1) for _ in range(64)]
    _check(ctx, many)
    assert ctx.decode_jpegs([])[0].numel() == 0
    ctx.close()


def test_decoded_files_feed_the_background_pool():
    """IlsvrcImages over JPEG files (decoded on the device into the pool) == over the cv2-decoded arrays."""
    from mtgvision_b200 import synth
    from mtgvision_b200.encoder_datasets import IlsvrcImages, SyntheticBgFgMtgImages
    from mtgvision_b200.encoder_train import RanMtgEncDecDataset
    from tests import parity_util as PU

    pool, _ = PU.small_pools(8, 1)
    files = [jpeg_cases.encode(synth.synth_bg(j), 90, "420") for j in range(8)]
    outs = []
    for src in (IlsvrcImages(images=[_ref(f) for f in files]), IlsvrcImages(files=files)):
        ds = RanMtgEncDecDataset(8, paired=True, targets=False, mtg=SyntheticBgFgMtgImages(pool=pool), ilsvrc=src, seed=3)
        b = next(iter(ds))
        outs.append((b["x"].cpu().numpy(), b["x2"].cpu().numpy()))
        ds.ctx.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(IlsvrcImages(files=files)[2], _ref(files[2]).astype(np.float32) / 255.0)


def test_truncated_files_decode_like_cv2_imread(tmp_path):
    """Premature end of the data (ILSVRC has such files): finished MCU from zero bits, the rest grey - cv2.imread's result."""
    from mtgvision_b200.context import Context

    ctx = Context(0)
    cases = jpeg_cases.truncated_suite()
    rng = np.random.default_rng(8)
    big = jpeg_cases.encode(jpeg_cases.image(rng, 375, 500, "mixed"), 90, "420")
    cases += [(f"375x500_cut{c}", big[:c]) for c in (len(big) // 3, len(big) - 1000, 700)]
    flat, off, hw = ctx.decode_jpegs([d for _, d in cases])
    flat = flat.cpu().numpy()
    for i, (name, data) in enumerate(cases):
        ref = jpeg_cases.imread_ref(data, tmp_path)
        h, w = hw[i]
        got = flat[off[i]: off[i] + 3 * h * w].reshape(h, w, 3)
        assert np.array_equal(got, ref), (name, int((got != ref).any(axis=2).sum()))
    ctx.close()


def test_host_pipeline_fed_with_jpeg_files_equals_decoded_arrays():
    """host_tensor_batches with prepare_jpeg_batch items == the same call with the cv2-decoded images as tensors."""
    from mtgvision_b200 import synth
    from mtgvision_b200.encoder_datasets import IlsvrcImages, SyntheticBgFgMtgImages
    from mtgvision_b200.encoder_train import RanMtgEncDecDataset
    from tests import parity_util as PU

    pool, bgs = PU.small_pools(8, 8)
    n = 4
    card_files = [jpeg_cases.encode(pool.images[k], 90, "420") for k in range(n)]
    bg_files = [jpeg_cases.encode(bgs[j], 90, "420") for j in range(n)]
    hc = torch.from_numpy(np.stack([_ref(f) for f in card_files]))
    hb = torch.from_numpy(np.stack([_ref(f) for f in bg_files]))
    outs = []
    for mode in ("arrays", "files"):
        ds = RanMtgEncDecDataset(n, paired=True, targets=False, mtg=SyntheticBgFgMtgImages(pool=pool), ilsvrc=IlsvrcImages(images=bgs), seed=5)
        item = (hc, hb) if mode == "arrays" else ds.prepare_jpeg_batch(card_files, bg_files)
        got = [{k: v.clone() for k, v in res.items()} for res in ds.host_tensor_batches(item for _ in range(3))]
        outs.append(got)
        ds.ctx.close()
    for a, b in zip(*outs):
        for k in a:
            assert torch.equal(a[k], b[k]), k


def test_card_pool_from_jpeg_files():
    """mtgv_set_card_pool fed from device-decoded files == fed from the cv2-decoded arrays (same generated batch)."""
    from mtgvision_b200 import abi, synth
    from tests import parity_util as PU

    pool, bgs = PU.small_pools(4, 4)
    files = [jpeg_cases.encode(pool.images[k], 92, "420") for k in range(4)]
    decoded = synth.CardPool(np.stack([_ref(f) for f in files]), pool.faces)
    outs = []
    for mode in ("arrays", "files"):
        ctx = PU.make_context(decoded, bgs)
        if mode == "files":
            ctx.set_card_pool_from_jpegs(files, decoded.labels3, decoded.grp_off, decoded.grp_mem)
        tape = ctx.sample_encoder_tape(7, 0, 4)
        params, labels = ctx.expand_params(tape)
        outs.append((ctx.encoder_batch(params, abi.OUT_U8).cpu().numpy(), labels.cpu().numpy()))
        ctx.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_decode_straight_into_the_pools():
    """mtgv_decode_jpeg_to_pools (files -> the pools' own planar / RGBX layouts) == set_*_pool from the cv2-decoded arrays:
    same generated batch, for a slot range in the middle of the pools, ragged widths (not multiples of 4) included."""
    from mtgvision_b200 import abi, synth
    from mtgvision_b200.abi import MtgvError
    from tests import parity_util as PU

    for card_hw, bg_hw, mode in (((680, 488), (375, 500), "420"), ((121, 86), (77, 101), "444"), ((90, 62), (50, 67), "422")):
        pool, bgs = PU.small_pools(6, 6, card_hw, bg_hw)
        cfiles = [jpeg_cases.encode(pool.images[k], 92, mode) for k in range(6)]
        bfiles = [jpeg_cases.encode(bgs[j], 88, mode) for j in range(6)]
        dec_cards = synth.CardPool(np.stack([_ref(f) for f in cfiles]), pool.faces)
        dec_bgs = [_ref(f) for f in bfiles]
        outs = []
        for via_files in (False, True):
            if via_files:
                ctx = PU.make_context(synth.CardPool(np.zeros_like(dec_cards.images), pool.faces), [np.zeros_like(b) for b in dec_bgs])
                # slots [0,2) from arrays, [2,6) decoded from files into the pool layouts
                ctx.update_card_images(torch.from_numpy(dec_cards.images[:2]).cuda(), 0)
                ctx.update_bg_images(torch.from_numpy(np.stack(dec_bgs[:2])).cuda(), 0)
                batch = ctx.prepare_jpegs(cfiles[2:] + bfiles[2:])
                ctx.decode_into_pools(batch, 4, 2, 4, 2)
            else:
                ctx = PU.make_context(dec_cards, dec_bgs)
            tape = ctx.sample_encoder_tape(11, 0, 24)
            params, labels = ctx.expand_params(tape)
            outs.append(ctx.encoder_batch(params, abi.OUT_U8).cpu().numpy())
            if via_files:
                with pytest.raises(MtgvError, match="destination"):  # a background file into a card slot: sizes differ
                    ctx.decode_into_pools(ctx.prepare_jpegs(bfiles[:1]), 1, 0, 0, 0)
                with pytest.raises(MtgvError, match="outside the pool"):
                    ctx.decode_into_pools(ctx.prepare_jpegs(cfiles[:2]), 2, 5, 0, 0)
            ctx.close()
        assert np.array_equal(outs[0], outs[1]), (card_hw, bg_hw, mode)


def test_static_helpers_accept_file_paths(tmp_path):
    """`path_or_img: str` of make_cropped / make_masked / make_bg / make_virtual (encoder_datasets.py:741, 759, 778, 796): the file
    goes through the device decoder and gives what the cv2-decoded array gives."""
    import random

    from mtgvision_b200 import synth
    from mtgvision_b200.encoder_datasets import SyntheticBgFgMtgImages as S

    card, bg = synth.synth_card(3), synth.synth_bg(2)
    cpath, bpath = str(tmp_path / "card.jpg"), str(tmp_path / "bg.jpg")
    open(cpath, "wb").write(jpeg_cases.encode(card, 92, "420"))
    open(bpath, "wb").write(jpeg_cases.encode(bg, 90, "420"))
    dcard, dbg = _ref(open(cpath, "rb").read()), _ref(open(bpath, "rb").read())
    assert np.array_equal(S.make_cropped(cpath, (192, 128)), S.make_cropped(dcard, (192, 128)))
    assert np.array_equal(S.make_masked(cpath), S.make_masked(dcard))
    random.seed(5); a = S.make_bg(bpath, (192, 128))
    random.seed(5); b = S.make_bg(dbg, (192, 128))
    assert np.array_equal(a, b)
    random.seed(6); a = S.make_virtual(cpath, bpath, (192, 128))
    random.seed(6); b = S.make_virtual(dcard, dbg, (192, 128))
    assert np.array_equal(a, b)
    with pytest.raises(Exception, match="Image not found"):
        S.make_cropped(str(tmp_path / "missing.jpg"), (192, 128))


def test_detection_generator_over_jpeg_backgrounds():
    """Gen over a file-backed background source == Gen over the cv2-decoded arrays."""
    from mtgvision_b200 import synth
    from mtgvision_b200.encoder_datasets import IlsvrcImages, SyntheticBgFgMtgImages
    from mtgvision_b200.od_datasets import Gen
    from tests import parity_util as PU

    pool, _ = PU.small_pools(8, 1)
    files = [jpeg_cases.encode(synth.synth_bg(j), 90, "420") for j in range(6)]
    outs = []
    for src in (IlsvrcImages(images=[_ref(f) for f in files]), IlsvrcImages(files=files)):
        gen = Gen(card_min_visible_ratio=0.5, card_min_visible_ratio_edges=0.0, card_jitter_ratio=0.7, ratio_bg=0.1, kind="seg",
                  mtg_ds=SyntheticBgFgMtgImages(pool=pool), bg_ds=src, seed=21)
        b = gen.random_batch(6)
        outs.append((b["image"].cpu().numpy(), b["keypoints"].cpu().numpy(), b["counts"].cpu().numpy()))
        gen.ctx.close()
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def test_progressive_files_in_a_mixed_batch():
    """Progressive files (host entropy stage, coefficients uploaded, IDCT / upsampling / colour on the device) next to baseline
    files in one call: bit-exact with cv2.imdecode, into plain HWC images and straight into the pools."""
    from mtgvision_b200 import abi, synth
    from mtgvision_b200.context import Context
    from tests import parity_util as PU

    rng = np.random.default_rng(8)
    files = []
    for k, (h, w) in enumerate([(375, 500), (24, 31), (375, 500), (1, 1), (97, 64), (680, 488)]):
        img = jpeg_cases.image(rng, h, w, ("mixed", "noise", "smooth")[k % 3])
        files.append(jpeg_cases.encode(img, (90, 35, 100)[k % 3], ("420", "444", "422", "440")[k % 4], rst=(0, 3)[k % 2], progressive=int(k % 3 != 2)))
    ctx = Context(0)
    flat, off, hw = ctx.decode_jpegs(files)
    torch.cuda.synchronize()
    host = flat.cpu().numpy()
    for k, f in enumerate(files):
        h, w = hw[k]
        assert np.array_equal(host[off[k]: off[k] + 3 * h * w].reshape(h, w, 3), _ref(f)), k
    ctx.close()
    # pools: slots decoded from progressive files == slots set from the cv2-decoded arrays
    pool, bgs = PU.small_pools(4, 4)
    cfiles = [jpeg_cases.encode(pool.images[k], 92, "420", progressive=k % 2) for k in range(4)]
    bfiles = [jpeg_cases.encode(bgs[j], 88, "420", progressive=(j + 1) % 2) for j in range(4)]
    dec_cards = synth.CardPool(np.stack([_ref(f) for f in cfiles]), pool.faces)
    dec_bgs = [_ref(f) for f in bfiles]
    outs = []
    for via_files in (False, True):
        if via_files:
            ctx = PU.make_context(synth.CardPool(np.zeros_like(dec_cards.images), pool.faces), [np.zeros_like(b) for b in dec_bgs])
            ctx.decode_into_pools(ctx.prepare_jpegs(cfiles + bfiles), 4, 0, 4, 0)
        else:
            ctx = PU.make_context(dec_cards, dec_bgs)
        params, _ = ctx.expand_params(ctx.sample_encoder_tape(3, 0, 16))
        outs.append(ctx.encoder_batch(params, abi.OUT_U8).cpu().numpy())
        ctx.close()
    assert np.array_equal(outs[0], outs[1])


def test_rejects_unsupported_and_mismatched():
    from mtgvision_b200.abi import MtgvError
    from mtgvision_b200.context import Context

    ctx = Context(0)
    rng = np.random.default_rng(3)
    img = jpeg_cases.image(rng, 24, 24, "mixed")
    arith = jpeg_cases.encode(img).replace(b"\xff\xc0", b"\xff\xc9", 1)  # SOF9: arithmetic coding
    with pytest.raises(MtgvError, match="Huffman-coded"):
        ctx.decode_jpegs([arith])
    with pytest.raises(MtgvError, match="SOI"):
        ctx.jpeg_info(b"\x89PNG....")
    ctx.close()
    # a file-backed background source can leave such files out instead of failing
    from mtgvision_b200.encoder_datasets import IlsvrcImages

    files = [jpeg_cases.encode(img), arith, jpeg_cases.encode(img, 80, "444")]
    with pytest.warns(UserWarning, match="Huffman-coded"):
        src = IlsvrcImages(files=files, skip_unsupported=True)
    assert len(src) == 2 and np.array_equal(src.images_u8[1], _ref(files[2]))


def test_gather_files_equals_join():
    """mtgv_gather_files (the copy of a list of `bytes` into pinned staging, split across host threads inside files):
    byte-identical to b"".join for empty, tiny and multi-megabyte buffers, one thread and several."""
    from mtgvision_b200.context import Context

    rng = np.random.default_rng(3)
    ctx = Context(0)
    for sizes in ([0, 1, 0, 7, 300, 0], [5 << 20, 0, 3, 1 << 20, 0, 0, 9 << 20, 17], [1 << 16] * 200, []):
        files = [rng.integers(0, 256, size=s, dtype=np.uint8).tobytes() for s in sizes]
        b = ctx.prepare_jpegs(files, hw=np.zeros((len(files), 2), np.int32))
        want = b"".join(files)
        assert b["blob"].numpy()[: len(want)].tobytes() == want
        assert np.array_equal(b["file_off"], np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64))
    ctx.close()
