"""Generate the golden vectors from the REAL reference (needs /root/reference; run in the
build container only):   python tests/golden/make_golden.py

For each seed: random.seed(s); np.random.seed(s); fresh ApplyShuffled state; then the
reference's own SyntheticBgFgMtgImages.make_virtual / make_cropped on the synthetic pools
of mtgvision_b200.synth.  Stored: sha1 of the float32 output bytes for every seed (exact
pin) and the uint8-rounded image for a few seeds (debugging aid).  Also geometry KATs
obtained by running reference code (od_datasets geometry helpers, round_rect_mask)."""
import hashlib
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from mtgvision_b200 import synth  # noqa: E402
from oracle import ref_import  # noqa: E402

N_SEEDS = 96
N_IMAGES = 6


def main():
    ed, od, uimg = ref_import.load_reference()
    cards = [synth.synth_card(k) for k in range(8)]
    bgs = [synth.synth_bg(j) for j in range(8)]
    f32 = lambda u8: uimg.img_float32(u8)
    out = {"virtual": [], "cropped": [], "masks": {}, "det": {}}
    keep = {}
    for seed in range(N_SEEDS):
        card, bg = seed % 8, (seed // 8) % 8
        random.seed(seed)
        np.random.seed(seed)
        ref_import.reset_reference_shuffles(ed)
        img = ed.SyntheticBgFgMtgImages.make_virtual(f32(cards[card]), f32(bgs[bg]).copy(), (192, 128), True)
        img = np.ascontiguousarray(img, dtype=np.float32)
        out["virtual"].append({"seed": seed, "card": card, "bg": bg, "sha1": hashlib.sha1(img.tobytes()).hexdigest(),
                               "sum": float(img.astype(np.float64).sum())})
        if seed < N_IMAGES:
            keep[f"virtual_{seed}"] = np.rint(np.clip(img, 0, 1) * 255).astype(np.uint8)
    for card in range(8):
        img = np.ascontiguousarray(ed.SyntheticBgFgMtgImages.make_cropped(f32(cards[card]), (192, 128)))
        out["cropped"].append({"card": card, "sha1": hashlib.sha1(img.tobytes()).hexdigest()})
    for hw, ratio in (((680, 488), 0.05), ((680, 488), 0.046), ((204, 146), 0.05)):
        m = uimg.round_rect_mask(hw, radius_ratio=ratio)
        out["masks"][f"{hw[0]}x{hw[1]}@{ratio}"] = {"zeros": int((m == 0).sum()), "sha1": hashlib.sha1(m.tobytes()).hexdigest()}
    # detection geometry KATs (od_datasets.py:85-118, 218-279)
    M = od.get_rotate_over_output_transform((375, 500), 37, 640, "cover")
    out["det"]["rotate_over_output_375x500_37_640"] = M.tolist()
    s = od.make_card_with_mask(f32(cards[0]), kind="obb")
    out["det"]["obb_keypoints_680x488"] = s["keypoints"].tolist()
    out["det"]["bbox_680x488"] = s["bbox"].tolist()
    json.dump(out, open(os.path.join(HERE, "encoder_golden.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(HERE, "encoder_golden_images.npz"), **keep)
    print("wrote", len(out["virtual"]), "virtual,", len(out["cropped"]), "cropped")


if __name__ == "__main__":
    main()
