"""Device JPEG encode through the C ABI against cv2.imencode itself: identical FILE BYTES."""
import os

import cv2
import numpy as np
import pytest

from tests import jpeg_cases

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def cv2_bytes(rgb, quality=None):
    params = [] if quality is None else [cv2.IMWRITE_JPEG_QUALITY, quality]
    return cv2.imencode(".jpg", cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR), params)[1].tobytes()


@pytest.mark.parametrize("hw,quality", [((640, 640), None), ((1280, 1280), None), ((64, 48), 75), ((16, 16), 100), ((128, 256), 30), ((320, 208), 1),
                                        ((1, 1), None), ((7, 5), 90), ((9, 17), None), ((17, 16), 60), ((31, 33), None), ((100, 101), 80), ((375, 500), None),
                                        ((49, 15), 100), ((15, 130), None)])
def test_file_bytes_equal_cv2(hw, quality):
    from mtgvision_b200.context import Context

    ctx = Context(0)
    rng = np.random.default_rng(hw[0] + hw[1])
    imgs = [jpeg_cases.image(rng, *hw, kind) for kind in ("mixed", "smooth", "noise")]
    imgs += [np.zeros((*hw, 3), np.uint8), np.full((*hw, 3), 255, np.uint8)]
    batch = np.stack(imgs)
    q = 95 if quality is None else quality
    cap = (hw[0] + 16) * (hw[1] + 16) * 3 + 4096  # noise at quality 100 expands
    nhwc = ctx.encode_jpegs(torch.from_numpy(batch).cuda(), q, cap=cap // 4 * 4)
    nchw = ctx.encode_jpegs(torch.from_numpy(batch).cuda().permute(0, 3, 1, 2).contiguous(), q, layout="nchw", cap=cap // 4 * 4)
    for k, img in enumerate(imgs):
        ref = cv2_bytes(img, quality)
        assert nhwc[k] == ref, (k, len(nhwc[k]), len(ref))
        assert nchw[k] == ref, (k, "nchw")
    ctx.close()


def test_capacity_and_size_errors():
    from mtgvision_b200.abi import MtgvError
    from mtgvision_b200.context import Context

    ctx = Context(0)
    rng = np.random.default_rng(1)
    noise = torch.from_numpy(jpeg_cases.image(rng, 64, 64, "noise")[None]).cuda()
    with pytest.raises(MtgvError, match="do not fit"):
        ctx.encode_jpegs(noise, 100, cap=2048)
    with pytest.raises(MtgvError, match="16384"):
        ctx.encode_jpegs(torch.zeros((1, 1, 16400, 3), dtype=torch.uint8, device="cuda"))
    assert ctx.encode_jpegs(torch.zeros((0, 64, 64, 3), dtype=torch.uint8, device="cuda")) == []
    ctx.close()


def test_batched_dataset_writer_equals_cv2_imwrite(tmp_path):
    """create_yolo_obb_dataset's batched path: every image file is what cv2.imwrite writes for that scene, every label
    file what save_sample writes."""
    from mtgvision_b200 import od_datasets as OD
    from mtgvision_b200.encoder_datasets import IlsvrcImages, SyntheticBgFgMtgImages
    from tests import parity_util as PU

    pool, bgs = PU.small_pools(8, 8)
    gen = OD.Gen(card_min_visible_ratio=0.5, card_min_visible_ratio_edges=0.0, card_jitter_ratio=0.7, ratio_bg=0.1, kind="seg",
                 mtg_ds=SyntheticBgFgMtgImages(pool=pool), bg_ds=IlsvrcImages(images=bgs), seed=11)
    b = gen.random_batch(6, "uint8")
    OD.save_batch(gen, b, 0, tmp_path, tmp_path)
    imgs = b["image"].permute(0, 2, 3, 1).contiguous().cpu().numpy()
    for k in range(6):
        ref_path = tmp_path / f"ref_{k}.jpg"
        cv2.imwrite(str(ref_path), cv2.cvtColor(imgs[k], cv2.COLOR_RGB2BGR))
        assert (tmp_path / f"image_{k:04d}.jpg").read_bytes() == ref_path.read_bytes()
        cnt = int(b["counts"][k])
        sample = {"image": imgs[k], "keypoints": b["keypoints"][k, :cnt, :8].cpu().numpy(), "keypoints_labels": b["labels"][k, :cnt].cpu().numpy().astype(np.int64)}
        one = tmp_path / "one"
        one.mkdir(exist_ok=True)
        OD.save_sample(sample, k, one, one, ext="jpg")
        assert (one / f"image_{k:04d}.txt").read_text() == (tmp_path / f"image_{k:04d}.txt").read_text()
        assert (one / f"image_{k:04d}.jpg").read_bytes() == (tmp_path / f"image_{k:04d}.jpg").read_bytes()


def test_create_yolo_obb_dataset_batched_layout(tmp_path):
    """od_datasets.py:732-791 through the batched GPU path: same directory layout, yaml and file naming as the reference."""
    import yaml

    from mtgvision_b200 import od_datasets as OD
    from mtgvision_b200.encoder_datasets import IlsvrcImages, SyntheticBgFgMtgImages
    from tests import parity_util as PU

    pool, bgs = PU.small_pools(8, 8)
    gen = OD.Gen(card_min_visible_ratio=0.5, card_min_visible_ratio_edges=0.0, card_jitter_ratio=0.7, ratio_bg=0.1, kind="seg",
                 mtg_ds=SyntheticBgFgMtgImages(pool=pool), bg_ds=IlsvrcImages(images=bgs), seed=3)
    out = tmp_path / "ds"
    OD.create_yolo_obb_dataset(gen, output_dir=str(out), num_train=5, num_val_ratio=0.4, num_test_ratio=0.2, batch=4)
    cfg = yaml.safe_load((out / "mtg_obb.yaml").read_text())
    assert cfg["names"] == {0: "card", 1: "card_top", 2: "card_bottom"}
    for name, num in (("train", 5), ("val", 2), ("test", 1)):
        imgs = sorted(os.listdir(out / "images" / name))
        assert imgs == [f"image_{i:04d}.jpg" for i in range(num)]
        assert sorted(os.listdir(out / "labels" / name)) == [f"image_{i:04d}.txt" for i in range(num)]
        im = cv2.imread(str(out / "images" / name / imgs[0]))
        assert im is not None and im.shape == (640, 640, 3)
    with pytest.raises(FileExistsError):
        OD.create_yolo_obb_dataset(gen, output_dir=str(out), num_train=1)
