"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference encoder-sample generator.

This is the parity oracle (and the timed CPU baseline) for the encoder half of the hot
path.  It restates, function by function, what nmichlo/mtg-vision does in

    mtgvision/encoder_datasets.py   Mutate.* (:68-404), SyntheticBgFgMtgImages.make_* (:733-834),
                                    pipelines _RAN_BG/_RAN_FG/_RAN_VRTL (:669-731)
    mtgvision/util/image.py         img_float32 (:220), rgba_over_rgb (:246-290), resize (:322),
                                    remove_border_resized (:338), crop_to_size (:350),
                                    rotate_bounded (:381), round_rect_mask (:407), noise_* (:434-488)
    mtgvision/util/random.py        ApplyOrdered/ApplyShuffled/ApplyChoice (:82-102)
    mtgvision/encoder_train.py      RanMtgEncDecDataset._make_image_batch (:189-230)

with the same third-party calls (opencv-python 4.13.0, numpy legacy RNG, Python `random`)
in the same order, so that with identical seeds it reproduces the reference bit for bit
(pinned by tests/test_oracle_vs_reference.py against the imported reference, and by the
golden vectors in tests/golden/ generated from the reference itself).

Unlike the reference it *records* every sampled choice and magnitude in a per-sample
"tape" (plain dicts), which is what the CUDA path is fed for parity tests: identical
sampled parameters, identical inputs.

Nothing under `mtgvision_b200/` imports this module.
"""

from __future__ import annotations

import math
import random
from typing import Any

import cv2
import numpy as np

# persistent shuffle state, mirrors ApplyShuffled.indices (util/random.py:88-97)
_BG_INDICES = [0, 1, 2]
_VRTL_INDICES = [0, 1, 2, 3, 4, 5, 6]

_INTERPS = (cv2.INTER_NEAREST, cv2.INTER_LINEAR, cv2.INTER_CUBIC)


def reset_shuffle_state():
    _BG_INDICES[:] = [0, 1, 2]
    _VRTL_INDICES[:] = [0, 1, 2, 3, 4, 5, 6]


def _choice(n: int) -> int:
    """random.choice(seq) consumes exactly random._randbelow(len(seq)) (util/random.py:102)."""
    return random.randrange(n)


def u8_to_f32(img_u8: np.ndarray) -> np.ndarray:
    """img_float32 for uint8 input (util/image.py:232-233)."""
    return np.clip(np.divide(img_u8, 255.0, dtype=np.float32), 0, 1)


def _clip(img):
    return np.clip(img, 0, 1)


# --------------------------------------------------------------------------- #
# static pieces                                                                #
# --------------------------------------------------------------------------- #


def round_rect_mask(size_hw, radius_ratio: float) -> np.ndarray:
    """util/image.py:406-425 (uses the real cv2.circle)."""
    radius = int(math.ceil(max(size_hw) * radius_ratio))
    img = np.ones(size_hw[:2], dtype=np.float32)
    corner = np.zeros((radius, radius), dtype=np.float32)
    cv2.circle(corner, (0, 0), radius, 1, cv2.FILLED)
    y1, x1 = size_hw[:2]
    img[y1 - radius :, x1 - radius :] = np.rot90(corner, 0)
    img[:radius, x1 - radius :] = np.rot90(corner, 1)
    img[:radius, :radius] = np.rot90(corner, 2)
    img[y1 - radius :, :radius] = np.rot90(corner, 3)
    return img


def resize_area_clip(img, size_hw):
    """util/image.py:321-334 with shrink=True: INTER_AREA always, then clip."""
    h, w = size_hw
    return _clip(cv2.resize(img, (w, h), interpolation=cv2.INTER_AREA))


def crop_to_size_geometry(in_hw, size_hw, pad: bool):
    """The integer geometry of crop_to_size (util/image.py:359-376):
    returns (rh, rw, y0, x0) or None when the size already matches."""
    (ih, iw), (sh, sw) = in_hw, size_hw
    if ih == sh and iw == sw:
        return None
    rh, rw = ih / sh, iw / sw
    r = max(rh, rw) if pad else min(rh, rw)
    rh, rw = int(ih / r), int(iw / r)
    if pad:
        return rh, rw, (sh - rh) // 2, (sw - rw) // 2
    return rh, rw, (rh - sh) // 2, (rw - sw) // 2


def crop_to_size(img, size_hw, pad=False):
    g = crop_to_size_geometry(img.shape[:2], size_hw, pad)
    if g is None:
        return img
    rh, rw, y0, x0 = g
    sh, sw = size_hw
    resized = resize_area_clip(img, (rh, rw))
    if pad:
        out = np.zeros((sh, sw, img.shape[2]), dtype=img.dtype)
        out[y0 : y0 + rh, x0 : x0 + rw, :] = resized
        return out
    return resized[y0 : y0 + sh, x0 : x0 + sw, :]


def rgba_over_rgb(fg_rgba, bg_rgb):
    """util/image.py:246-290: per channel fg*a and bg*(1-a), summed, clipped."""
    a = fg_rgba[:, :, 3]
    fg = cv2.merge([fg_rgba[:, :, c] * a for c in range(3)])
    inv = 1 - a
    bg = cv2.merge([bg_rgb[:, :, c] * inv for c in range(3)])
    return _clip(bg + fg)


# --------------------------------------------------------------------------- #
# Mutate.* restated; every function returns (image, op-record)                 #
# --------------------------------------------------------------------------- #


def op_downscale_upscale(img):
    """encoder_datasets.py:142-163."""
    n = int(np.random.randint(0, 3))
    h, w = img.shape[:2]
    nh, nw = h // (2**n), w // (2**n)
    down = int(np.random.choice(_INTERPS))
    up = int(np.random.choice(_INTERPS))
    img = cv2.resize(img, (nw, nh), interpolation=down)
    img = cv2.resize(img, (w, h), interpolation=up)
    return img, {"op": "downup", "n": n, "down": down, "up": up}


def _warp_from_rand(img, ran_u, ratio, ratio_min):
    """encoder_datasets.py:94-111 (note the (row, col) points handed to cv2 as (x, y))."""
    h, w = img.shape[0] - 1, img.shape[1] - 1
    src = np.asarray([(0, 0), (0, w), (h, 0), (h, w)], dtype=np.float32)
    ran = ratio_min + ran_u * (abs(ratio - ratio_min) * 0.5)
    dst = ran * np.asarray([(h, w), (h, -w), (-h, w), (-h, -w)], dtype=np.float32) + src
    dst = np.asarray(dst, dtype=np.float32)
    M = cv2.getPerspectiveTransform(src, dst)
    out = cv2.warpPerspective(img, M, (img.shape[1], img.shape[0]))
    return out, src, dst, M


def op_warp(img):
    u = np.random.rand(4, 2)
    out, src, dst, M = _warp_from_rand(img, u, 0.3, -0.25)
    return out, {"op": "warp", "u": u, "src": src, "dst": dst, "M": M}


def op_warp_inv(img):
    """encoder_datasets.py:113-116: warp(ratio=-0.5, min=-0.25)."""
    u = np.random.rand(4, 2)
    out, src, dst, M = _warp_from_rand(img, u, -0.5, -0.25)
    return out, {"op": "warp_inv", "u": u, "src": src, "dst": dst, "M": M}


def op_affine(img):
    """encoder_datasets.py:353-375."""
    angle = np.random.uniform(-5, 5)
    tx = np.random.uniform(-10, 10)
    ty = np.random.uniform(-10, 10)
    s = min(1.0 + 0.1, 1.0 / (1.0 + 0.1))
    scale = np.random.uniform(s, 1 / s)
    shear = np.random.uniform(-0.3, 0.3)
    rows, cols, _ = img.shape
    center = (cols / 2, rows / 2)
    R = cv2.getRotationMatrix2D(center, angle, scale)
    M = np.vstack([R, [0, 0, 1]])
    M = np.dot(np.array([[1, shear, 0], [0, 1, 0], [0, 0, 1]]), M)
    M[0, 2] += tx
    M[1, 2] += ty
    out = cv2.warpAffine(img, M[:2, :], (cols, rows))
    rec = {
        "op": "affine", "angle": angle, "tx": tx, "ty": ty, "scale": scale, "shear": shear,
        "alpha": float(R[0, 0]), "beta": float(R[0, 1]), "M": M[:2, :].copy(),
    }
    return out, rec


def op_perspective(img):
    """encoder_datasets.py:377-403."""
    rows, cols, _ = img.shape
    u = [np.random.uniform(-0.1, 0.1) for _ in range(8)]
    src = np.float32([[0, 0], [cols, 0], [0, rows], [cols, rows]])
    dst = np.float32(
        [
            [u[0] * cols, u[1] * rows],
            [cols + u[2] * cols, u[3] * rows],
            [u[4] * cols, rows + u[5] * rows],
            [cols + u[6] * cols, rows + u[7] * rows],
        ]
    )
    M = cv2.getPerspectiveTransform(src, dst)
    out = cv2.warpPerspective(img, M, (cols, rows))
    return out, {"op": "perspective", "u": np.asarray(u), "src": src, "dst": dst, "M": M}


def op_tint(img):
    """encoder_datasets.py:165-171 (in place on channels :3, clip per channel)."""
    u = []
    for i in range(3):
        ui = np.random.random()
        u.append(ui)
        r = 1 + 0.15 * (2 * ui - 1)
        img[:, :, i] = _clip(r * img[:, :, i])
    return img, {"op": "tint", "u": np.asarray(u)}


def op_fade_white(img):
    u = np.random.random()
    ratio = u * 0.33
    img[:, :, :3] = ratio * 1 + (1 - ratio) * img[:, :, :3]
    return img, {"op": "fade_white", "u": u}


def op_fade_black(img):
    u = np.random.random()
    ratio = u * 0.5
    img[:, :, :3] = ratio * 0 + (1 - ratio) * img[:, :, :3]
    return img, {"op": "fade_black", "u": u}


def op_brightness_contrast(img):
    """encoder_datasets.py:187-193: whole array (alpha channel included), then clip."""
    uc = np.random.uniform(-0.2, 0.2)
    ub = np.random.uniform(-0.2, 0.2)
    alpha = 1.0 + uc
    img = alpha * img + ub
    return _clip(img), {"op": "bc", "uc": uc, "ub": ub}


_FADES = (op_fade_black, op_fade_white, op_brightness_contrast, None)


def op_flip(img):
    """encoder_datasets.py:73-80 -> util/image.py:312-318."""
    horr = random.random() >= 0.5
    vert = random.random() >= 0.5
    if vert:
        img = cv2.flip(img, 0)
    if horr:
        img = cv2.flip(img, 1)
    return img, {"op": "flip", "horr": bool(horr), "vert": bool(vert)}


def op_rotate_bounded(img):
    """encoder_datasets.py:82-87 -> util/image.py:380-398."""
    u = np.random.random()
    deg = 0 + u * (360 - 0)
    h, w = img.shape[:2]
    cy, cx = h // 2, w // 2
    M = cv2.getRotationMatrix2D(center=(cx, cy), angle=deg, scale=1.0)
    alpha, beta = float(M[0, 0]), float(M[0, 1])
    cos, sin = np.abs(M[0, 0]), np.abs(M[0, 1])
    nw, nh = int((h * sin) + (w * cos)), int((h * cos) + (w * sin))
    M[0, 2] += (nw / 2) - cx
    M[1, 2] += (nh / 2) - cy
    out = cv2.warpAffine(img, M, (nw, nh))
    return out, {"op": "rotate", "u": u, "deg": deg, "alpha": alpha, "beta": beta, "M": M.copy(), "nh": nh, "nw": nw}


def op_blur(img):
    """encoder_datasets.py:136-140 with n_max=3: ksize 1 or 3, sigma 0."""
    k = int(np.random.randint(0, 2))
    n = k * 2 + 1
    return cv2.GaussianBlur(img, (n, n), 0), {"op": "blur", "n": n}


def op_sharpen(img):
    kernel = np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]])
    return _clip(cv2.filter2D(img, -1, kernel)), {"op": "sharpen"}


def op_noise(img):
    """encoder_datasets.py:118-134 + util/image.py:434-488."""
    kind = _choice(4)  # speckle, gaussian, pepper, poisson
    h, w = img.shape[:2]
    rec: dict[str, Any] = {"op": "noise", "kind": kind}
    if kind == 0:
        gauss = np.random.randn(h, w, 3)
        noisy = np.copy(img)
        noisy[:, :, :3] = img[:, :, :3] * (1 + gauss * 0.3)
        noisy = _clip(noisy)
        rec["field"] = gauss.astype(np.float32)
    elif kind == 1:
        gauss = np.random.normal(0, 0.05**0.5, (h, w, 3))
        noisy = np.copy(img)
        noisy[:, :, :3] = img[:, :, :3] + gauss
        noisy = _clip(noisy)
        rec["field"] = gauss.astype(np.float32)
    elif kind == 2:
        noisy = np.copy(img)
        num_salt = int(np.ceil(0.1 * img.size * 0.5))
        sy, sx, sc = [np.random.randint(0, i - 1, num_salt) for i in img.shape]
        noisy[sy, sx, sc] = 1
        num_pepper = int(np.ceil(0.1 * img.size * (1 - 0.5)))
        py, px, pc = [np.random.randint(0, i - 1, num_pepper) for i in img.shape]
        noisy[py, px, pc] = 0
        noisy = _clip(noisy)
        rec["salt"] = np.stack([sy, sx, sc], 1).astype(np.int32)
        rec["pepper"] = np.stack([py, px, pc], 1).astype(np.int32)
    else:
        src = _clip(img)
        noise = np.zeros_like(src)
        counts = np.random.poisson(src[:, :, :3] * 0.8)
        noise[:, :, :3] = counts / 0.8
        noisy = _clip((1 - 0.5) * src + 0.5 * noise).astype(np.float32)
        rec["field"] = counts.astype(np.float32)
    u = np.random.random()
    ratio = u * 0.5
    img[:, :, :3] = ratio * noisy[:, :, :3] + (1 - ratio) * img[:, :, :3]
    rec["u"] = u
    return img, rec


def op_gaussian_noise(img):
    noise = np.random.normal(0, 0.25, img.shape).astype(np.float32)
    return _clip(img + noise), {"op": "gaussian_noise", "field": noise}


def op_salt_pepper_noise(img):
    """encoder_datasets.py:228-240: 1 % salt then 1 % pepper, all channels."""
    noisy = img.copy()
    n_s = int(np.ceil(0.01 * img.size))
    n_p = int(np.ceil(0.01 * img.size))
    c = [np.random.randint(0, i - 1, n_s) for i in img.shape]
    noisy[c[0], c[1], :] = 1
    salt = np.stack([c[0], c[1]], 1).astype(np.int32)
    c = [np.random.randint(0, i - 1, n_p) for i in img.shape]
    noisy[c[0], c[1], :] = 0
    pepper = np.stack([c[0], c[1]], 1).astype(np.int32)
    return noisy, {"op": "salt_pepper", "salt": salt, "pepper": pepper}


_ERASE_COLORS = ("random", "uniform_random", "zeros", "ones", "mean")


def op_random_erasing(img):
    """encoder_datasets.py:273-351 with defaults (inside=False)."""
    h, w = img.shape[:2]
    u_scale = np.random.uniform(0.2, 0.4)
    target_area = u_scale * (h * w)
    u_aspect = np.random.uniform(1, 3)
    aspect = u_aspect
    u_flip = np.random.random()
    if u_flip < 0.5:
        aspect = 1 / aspect
    bw = int((target_area / aspect) ** 0.5)
    bh = int((target_area * aspect) ** 0.5)
    rec: dict[str, Any] = {"op": "erase", "u_scale": u_scale, "u_aspect": u_aspect, "u_flip": u_flip,
                           "bw": bw, "bh": bh, "active": False}
    mx, Mx = 0 - bw // 2, w + bw // 2
    my, My = 0 - bh // 2, h + bh // 2
    if Mx <= mx or My <= my:
        return img, rec
    cx = int(np.random.randint(mx, Mx))
    cy = int(np.random.randint(my, My))
    rec["cx"], rec["cy"] = cx, cy
    x0, y0 = max(0, cx - bw // 2), max(0, cy - bh // 2)
    x1, y1 = min(w, cx + bw // 2), min(h, cy + bh // 2)
    rec["rect"] = (y0, y1, x0, x1)
    if y1 <= y0 or x1 <= x0:
        return img, rec
    color = _choice(5)
    rec["color"] = color
    rec["active"] = True
    name = _ERASE_COLORS[color]
    if name == "uniform_random":
        c = np.random.uniform(0, 1, (img.shape[-1],))
        rec["c"] = c.astype(np.float32)
    elif name == "random":
        c = np.random.uniform(0, 1, (y1 - y0, x1 - x0, img.shape[-1]))
        rec["field"] = c.astype(np.float32)
    elif name == "zeros":
        c = np.zeros((img.shape[-1],))
    elif name == "mean":
        c = img[y0:y1, x0:x1, :].mean(axis=(0, 1))
    else:
        c = np.ones((img.shape[-1],))
    img[y0:y1, x0:x1, :] = c
    return img, rec


def op_cutout(img):
    """encoder_datasets.py:259-271: 8 holes of 8x8 set to 0."""
    h, w, _ = img.shape
    holes = []
    for _ in range(8):
        y = int(np.random.randint(h))
        x = int(np.random.randint(w))
        y1, y2 = int(np.clip(y - 4, 0, h)), int(np.clip(y + 4, 0, h))
        x1, x2 = int(np.clip(x - 4, 0, w)), int(np.clip(x + 4, 0, w))
        img[y1:y2, x1:x2, :] = 0
        holes.append((y, x))
    return img, {"op": "cutout", "holes": np.asarray(holes, dtype=np.int32)}


_NOISE6 = (op_noise, op_gaussian_noise, op_salt_pepper_noise, op_random_erasing, op_cutout, None)


def _apply(fn, img, ops):
    if fn is None:
        return img
    img, rec = fn(img)
    ops.append(rec)
    return img


# --------------------------------------------------------------------------- #
# pipelines                                                                    #
# --------------------------------------------------------------------------- #


def ran_fg(fg, ops):
    """_RAN_FG (encoder_datasets.py:684-699): four ordered choices."""
    fg = _apply((op_downscale_upscale, None, None, None)[_choice(4)], fg, ops)
    fg = _apply((op_warp, op_affine, op_perspective, None)[_choice(4)], fg, ops)
    fg = _apply((op_tint, None)[_choice(2)], fg, ops)
    fg = _apply(_FADES[_choice(4)], fg, ops)
    return fg


def ran_bg(bg, ops):
    """_RAN_BG (encoder_datasets.py:669-683): persistent in-place shuffle of 3 groups."""
    random.shuffle(_BG_INDICES)
    for g in list(_BG_INDICES):
        if g == 0:
            bg = _apply(op_flip, bg, ops)
            bg = _apply(op_rotate_bounded, bg, ops)
            bg = _apply(op_warp_inv, bg, ops)
        elif g == 1:
            bg = _apply((op_tint, None)[_choice(2)], bg, ops)
        else:
            bg = _apply(_FADES[_choice(4)], bg, ops)
    return bg


def ran_vrtl(img, ops):
    """_RAN_VRTL (encoder_datasets.py:700-731): persistent in-place shuffle of 7 slots."""
    random.shuffle(_VRTL_INDICES)
    for s in list(_VRTL_INDICES):
        if s == 0:
            img = _apply((op_downscale_upscale, None, None, None)[_choice(4)], img, ops)
        elif s == 1:
            img = _apply((op_blur, None, None)[_choice(3)], img, ops)
        elif s == 2:
            img = _apply((op_sharpen, None, None)[_choice(3)], img, ops)
        elif s == 3:
            img = _apply(_NOISE6[_choice(6)], img, ops)
        elif s == 4:
            if _choice(2) == 0:
                img = _apply(_NOISE6[_choice(6)], img, ops)
        elif s == 5:
            img = _apply((op_tint, None)[_choice(2)], img, ops)
        else:
            img = _apply(_FADES[_choice(4)], img, ops)
    return img


def make_cropped(card_f32, size_hw=None, half_upsidedown=False, tape=None):
    """SyntheticBgFgMtgImages.make_cropped (encoder_datasets.py:733-753)."""
    ih, iw = card_f32.shape[:2]
    border = math.ceil(max(0.02 * ih, 0.02 * iw))
    ret = card_f32[border : ih - border, border : iw - border, :]
    if size_hw is not None:
        ret = resize_area_clip(ret, size_hw)
    ud = False
    if half_upsidedown:
        ud = _choice(2) == 0
        if ud:
            ret = np.rot90(ret, k=2)
    if tape is not None:
        tape.update(kind="cropped", upsidedown=bool(ud), border=border, size_hw=tuple(size_hw) if size_hw else None)
    return ret


def make_masked(card_f32):
    """make_masked (encoder_datasets.py:755-772): RGB + rounded-rect alpha (ratio 0.05)."""
    mask = round_rect_mask(card_f32.shape[:2], radius_ratio=0.05)
    return cv2.merge((card_f32[:, :, 0], card_f32[:, :, 1], card_f32[:, :, 2], mask))


def make_bg(bg_f32, size_hw, ops):
    """make_bg (encoder_datasets.py:774-784)."""
    bg = ran_bg(bg_f32, ops)
    return crop_to_size(bg, size_hw)


def make_virtual(card_f32, bg_f32, size_hw, half_upsidedown=False, tape=None):
    """make_virtual (encoder_datasets.py:786-813).  `bg_f32` is modified in place by the
    reference when tint/fade run before the geometric group; pass a copy."""
    t: dict[str, Any] = {} if tape is None else tape
    ud = False
    card = card_f32
    if half_upsidedown:
        ud = _choice(2) == 0
        if ud:
            card = np.rot90(card, k=2)
    fg_ops: list = []
    bg_ops: list = []
    v_ops: list = []
    fg = make_masked(card)
    fg = crop_to_size(fg, size_hw, pad=True)
    fg = ran_fg(fg, fg_ops)
    bg = make_bg(bg_f32, size_hw, bg_ops)
    virtual = rgba_over_rgb(fg, bg)
    virtual = ran_vrtl(virtual, v_ops)
    assert virtual.shape[:2] == tuple(size_hw)
    t.update(kind="virtual", upsidedown=bool(ud), size_hw=tuple(size_hw), fg_ops=fg_ops, bg_ops=bg_ops, vrtl_ops=v_ops,
             card_hw=tuple(card_f32.shape[:2]), bg_hw=tuple(bg_f32.shape[:2]))
    return virtual


# --------------------------------------------------------------------------- #
# batch former                                                                 #
# --------------------------------------------------------------------------- #


class BatchOracle:
    """RanMtgEncDecDataset (encoder_train.py:90-249) over in-memory uint8 pools.

    `cards`  : mtgvision_b200.synth.CardPool-like object with `.images` (N,H,W,3) uint8,
               `.labels3` (N,3), `.group_of(k)` -> list of pool indices with the same name
               in insertion order (mirrors `_cards_by_name`, encoder_datasets.py:570).
    `bgs`    : list/array of uint8 background images.
    Card/background *selection* uses Python `random` exactly like `ran_card`
    (encoder_datasets.py:662-664, choice over the sorted id list == pool order here)
    and `IlsvrcImages.ran` (:470-474).
    """

    def __init__(self, cards, bgs, *, paired=True, targets=False, x_size_hw=(192, 128), y_size_hw=(192, 128),
                 half_upsidedown=False, target_is_input_prob=0.05, similar_neg_prob=0.2):
        self.cards, self.bgs = cards, bgs
        self.paired, self.targets = paired, targets
        self.x_size_hw, self.y_size_hw = tuple(x_size_hw), tuple(y_size_hw)
        self.half_upsidedown = half_upsidedown
        self.target_is_input_prob = target_is_input_prob
        self.similar_neg_prob = similar_neg_prob

    def _make_x(self, card_f32, bg_f32, target_is_input_prob, tape):
        """__make_x__ (encoder_train.py:175-187); `p or default` quirk kept."""
        u = random.random()
        tape["u_target_is_input"] = u
        if u < (target_is_input_prob or self.target_is_input_prob):
            return make_cropped(card_f32, size_hw=self.x_size_hw, tape=tape)
        return make_virtual(card_f32, bg_f32.copy(), self.x_size_hw, self.half_upsidedown, tape=tape)

    def similar_card(self, k: int):
        """get_similar_card (encoder_datasets.py:619-630): same-name group minus self."""
        group = [j for j in self.cards.group_of(k) if j != k]
        if group:
            c = _choice(len(group))
            return group[c], c
        return None

    def make_image_batch(self, card_idx, bg_idx, *, target_in_prob=None, similar_neg_prob=None):
        """_make_image_batch (encoder_train.py:189-230)."""
        assert len(card_idx) == len(bg_idx)
        bg_imgs = [u8_to_f32(self.bgs[j]) for j in bg_idx]
        imgs = {"x": [], "x2": [], "y": []}
        lbls = {"x_labels": [], "x2_labels": []}
        tapes = {"x": [], "x2": []}
        for i, k in enumerate(card_idx):
            card_img = u8_to_f32(self.cards.images[k])
            if self.targets:
                imgs["y"].append(make_cropped(card_img, size_hw=self.y_size_hw))
            t = {"card": int(k), "bg": int(bg_idx[i])}
            imgs["x"].append(self._make_x(card_img, bg_imgs[i], target_in_prob, t))
            tapes["x"].append(t)
            lbls["x_labels"].append(tuple(int(v) for v in self.cards.labels3[k]))
            if self.paired:
                pk, pair_img = k, card_img
                u = random.random()
                swapped, swap_choice = False, -1
                if u < (similar_neg_prob or self.similar_neg_prob):
                    hit = self.similar_card(k)
                    if hit is not None:
                        pk, swap_choice = hit
                        pair_img, swapped = u8_to_f32(self.cards.images[pk]), True
                b = _choice(len(bg_imgs))  # bg1 = random.choice(bg_imgs)
                t2 = {"card": int(pk), "bg": int(bg_idx[b]), "u_similar_neg": u, "swapped": swapped,
                      "device_card": int(k), "swap_choice": int(swap_choice), "bg_slot": int(b)}
                imgs["x2"].append(self._make_x(pair_img, bg_imgs[b], target_in_prob, t2))
                tapes["x2"].append(t2)
                lbls["x2_labels"].append(tuple(int(v) for v in self.cards.labels3[pk]))
        out_imgs = {k: np.stack(v, 0) for k, v in imgs.items() if v}
        out_lbls = {k: np.asarray(v, dtype=np.int64) for k, v in lbls.items() if v}
        return out_imgs, out_lbls, tapes

    def random_image_batch(self, n: int):
        """_random_image_batch (encoder_train.py:149-156): all cards first, then all bgs."""
        card_idx = [_choice(len(self.cards.images)) for _ in range(n)]
        bg_idx = [_choice(len(self.bgs)) for _ in range(n)]
        return self.make_image_batch(card_idx, bg_idx)
