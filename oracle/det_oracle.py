"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference detection-scene generator.

Follows mtgvision/od_datasets.py function by function:

    corner_jitter_2d (:29-47)  rotate_2d (:50-56)  translate_2d (:59-61)  apply_transform_2d (:64-70)
    apply_transform_2d_img (:73-82)  get_rotate_over_output_transform (:85-118)  make_background (:195-203)
    make_card_with_mask (:218-279)  place_card_on_background_get_transform (:287-377)
    get_bg_transform_light / get_bg_transform / get_card_transform (:420-512)
    generate_synthetic_image (:520-611)  Gen.random / random_bg (:674-704)

What is pinned and what is not:
  * geometry (jitter / rotate / translate / homography / keypoint warp / cover transform),
    image warps and the alpha composite use the same numpy / cv2 calls as the reference and
    are pinned bit-exactly against the imported reference in tests/test_det_oracle.py;
  * the placement test uses shapely (GEOS) in the reference, which is NOT installed here:
    `PARITY UNPINNED` - restated with fp64 Sutherland-Hodgman clipping (all clip polygons are
    convex) following od_datasets.py:354-372 including the operator-precedence
    quirk at :369.  GEOS areas may differ in the last ulp; decisions can only flip within
    ~1e-12 of a threshold;
  * the photometric transforms use albumentations 2.x in the reference, which is NOT installed
    here and keeps private RNGs: `PARITY UNPINNED` - the parameter ranges and probabilities are
    the ones written at od_datasets.py:420-512, the per-op arithmetic follows albumentations'
    documented formulas on top of the cv2 kernels it delegates to, and the sampling order
    defined here is OUR specification.  ISONoise, ShotNoise, MedianBlur and MotionBlur
    are restated the same way - their inner kernels (cv2.cvtColor RGB<->HLS, cv2.meanStdDev, cv2.pow,
    cv2.medianBlur, cv2.line + cv2.filter2D, cv2.GaussianBlur) are the real cv2 calls, GlassBlur's
    pixel shuffle is albumentations' numpy statement verbatim.
  * the seg-kind keypoint polygon comes from a GEOS difference whose vertex order is
    GEOS-defined: `PARITY UNPINNED` - we start at (0,0) in the card box's own orientation.
"""

from __future__ import annotations

import math
import random
from typing import Any

import cv2
import numpy as np

from oracle import encoder_oracle as EO

# --------------------------------------------------------------------------- #
# geometry helpers (pinned against the reference)                             #
# --------------------------------------------------------------------------- #


def apply_transform_2d(pts: np.ndarray, M: np.ndarray) -> np.ndarray:
    pts = np.concatenate([pts, np.ones((*pts.shape[:-1], 1))], axis=-1)
    pts = pts @ M.T
    return pts[..., :2] / pts[..., 2:3]


def corner_jitter_2d(pts: np.ndarray, jitter_u: np.ndarray) -> np.ndarray:
    """od_datasets.py:29-47 with the uniform draw passed in (`jitter_u` = the factors)."""
    center = np.mean(pts, axis=0)
    deltas = np.linalg.norm((pts - center), axis=-1)
    deltas *= jitter_u
    angles = np.arctan2(pts[:, 1] - center[1], pts[:, 0] - center[0])
    return np.stack([center[0] + deltas * np.cos(angles), center[1] + deltas * np.sin(angles)], axis=-1)


def rotate_2d(pts, deg, center=(0, 0), scale=1):
    M = np.identity(3)
    M[:2, :] = cv2.getRotationMatrix2D(center, deg, scale)
    return apply_transform_2d(pts, M)


def get_rotate_over_output_transform(in_hw, deg, out_hw, scale_factor=1.0):
    """od_datasets.py:85-118, mode='cover'."""
    h, w = in_hw
    oh, ow = out_hw
    scale = math.hypot(oh / max(ow, oh), ow / max(ow, oh)) * max(oh, ow) / min(h, w)
    M0 = cv2.getRotationMatrix2D((w // 2, h // 2), deg, scale * scale_factor)
    M0 = np.concatenate([M0, np.array([[0, 0, 1]])], axis=0)
    M1 = np.array([[1, 0, (ow - w) // 2], [0, 1, (oh - h) // 2], [0, 0, 1]])
    return M1 @ M0


def box(lft, top, rht, bot, margin=0.0, mlr=1.0, mrr=1.0, mtr=1.0, mbr=1.0):
    return [
        (lft + margin * mlr, top + margin * mtr), (rht - margin * mrr, top + margin * mtr),
        (rht - margin * mrr, bot - margin * mbr), (lft + margin * mlr, bot - margin * mbr),
    ]


def card_keypoints(h: int, w: int, kind: str) -> np.ndarray:
    """make_card_with_mask keypoints (od_datasets.py:244-270)."""
    if kind == "obb":
        r, m = 0.5, 0.03 * max(w, h)
        return np.asarray([box(0, 0, w, h, margin=0), box(0, 0, w, r * h, margin=m, mbr=0.5),
                           box(0, (1 - r) * h, w, h, margin=m, mtr=0.5)])
    if kind == "seg":
        # Polygon(card_box).difference(Polygon(bottom_indent)).exterior minus the closing point.
        # Vertex order is GEOS-defined in the reference (PARITY UNPINNED); ours: card-box order from (0,0).
        x0, x1, y0 = w * 0.4, w * 0.6, h * 0.5
        return np.asarray([[(0, 0), (w, 0), (w, h), (x1, h), (x1, y0), (x0, y0), (x0, h), (0, h)]], dtype=np.float64)
    raise KeyError(f"invalid: {kind}")


# --------------------------------------------------------------------------- #
# convex clipping in fp64 (restates shapely's intersection/difference/contains) #
# same operation order as mtgvision_b200/csrc/mtgv_poly.cuh                     #
# --------------------------------------------------------------------------- #


def poly_signed2(p) -> float:
    s = 0.0
    n = len(p)
    for i in range(n):
        j = 0 if i + 1 == n else i + 1
        s = s + (p[i][0] * p[j][1] - p[j][0] * p[i][1])
    return s


def poly_area(p) -> float:
    if len(p) < 3:
        return 0.0
    return abs(poly_signed2(p) * 0.5)


def clip_convex(subj, clip):
    """Sutherland-Hodgman: `subj` (any simple polygon) clipped by the CONVEX polygon `clip`."""
    a = [(float(x), float(y)) for x, y in subj]
    orient = 1.0 if poly_signed2(clip) >= 0.0 else -1.0
    nc = len(clip)
    for e in range(nc):
        if not a:
            break
        e2 = 0 if e + 1 == nc else e + 1
        ex, ey = float(clip[e][0]), float(clip[e][1])
        dx, dy = float(clip[e2][0]) - ex, float(clip[e2][1]) - ey
        b = []
        na = len(a)
        for i in range(na):
            j = 0 if i + 1 == na else i + 1
            px, py = a[i]
            qx, qy = a[j]
            sp = orient * (dx * (py - ey) - dy * (px - ex))
            sq = orient * (dx * (qy - ey) - dy * (qx - ex))
            pin, qin = sp >= 0.0, sq >= 0.0
            if pin and len(b) < 16:
                b.append((px, py))
            if pin != qin and len(b) < 16:
                t = sp / (sp - sq)
                b.append((px + t * (qx - px), py + t * (qy - py)))
        a = b
    return a


def inside_convex(pt, poly) -> bool:
    orient = 1.0 if poly_signed2(poly) >= 0.0 else -1.0
    n = len(poly)
    for e in range(n):
        e2 = 0 if e + 1 == n else e + 1
        ex, ey = poly[e]
        dx, dy = poly[e2][0] - ex, poly[e2][1] - ey
        if orient * (dx * (pt[1] - ey) - dy * (pt[0] - ex)) < 0.0:
            return False
    return True


class CardShape:
    """The region tested by place_card_on_background_get_transform: Polygon(keypoints[0]).
    obb: the warped card quad Q.  seg: Q minus the warped bottom indent I (od_datasets.py:258-266),
    handled as area(Q n C) - area(Q n I n C) for convex C."""

    def __init__(self, kp0: np.ndarray, kind: str):
        pts = [tuple(map(float, p)) for p in kp0]
        if kind == "obb":
            self.quad, self.indent = pts, None
        else:  # vertices 0,1,2,7 are the card corners; 3..6 the indent's corners inside the card
            self.quad = [pts[0], pts[1], pts[2], pts[7]]
            self.indent = [pts[5], pts[4], pts[3], pts[6]]  # (x0,y0),(x1,y0),(x1,h),(x0,h)
        self.pts = pts

    def area_within(self, conv) -> float:
        """area(shape n conv) for a convex polygon `conv` (None = the whole plane)."""
        q = self.quad if conv is None else clip_convex(self.quad, conv)
        a = poly_area(q)
        if self.indent is not None and len(q) >= 3:
            a = a - poly_area(clip_convex(q, self.indent))
        return a


def placement_visible(shape: CardShape, S_hw, existing: list, min_visible: float, min_visible_edge: float,
                      no_contains: bool = True):
    """The accept/reject tests of od_datasets.py:353-372.  `existing`: previously accepted card
    bboxes (convex quads, od_datasets.py:585).  Returns (accepted, reason)."""
    bh, bw = S_hw
    img = [(0.0, 0.0), (float(bw), 0.0), (float(bw), float(bh)), (0.0, float(bh))]
    card_area = shape.area_within(None)
    vis_area = shape.area_within(img)                    # card_visible = img_bound n card_bound
    if vis_area / card_area < min_visible_edge:
        return False, "edge"
    visible = True
    for p in existing:
        pq = [tuple(map(float, v)) for v in p]
        pc = clip_convex(pq, img)                        # p n img  (convex)
        inter = shape.area_within(pc) if len(pc) >= 3 else 0.0   # area(card_visible n p)
        if (vis_area - inter) / card_area < min_visible:  # card_visible.difference(p).area / card_bound.area
            visible = False
            break
        p_area = poly_area(pq)
        if (p_area - inter) / p_area < min_visible:        # p.difference(card_visible).area / p.area
            visible = False
            break
        # `no_contains and p.contains(card_visible) or card_visible.contains(p)` (precedence as written)
        vis_pts = clip_convex(shape.quad, img)
        p_contains_vis = len(vis_pts) >= 3 and all(inside_convex(v, pq) for v in vis_pts)
        vis_contains_p = (all(inside_convex(v, img) and inside_convex(v, shape.quad) for v in pq)
                          and (shape.indent is None or poly_area(clip_convex(pq, shape.indent)) == 0.0))
        if (no_contains and p_contains_vis) or vis_contains_p:
            visible = False
    return visible, ("ok" if visible else "overlap")


# --------------------------------------------------------------------------- #
# photometrics (our specification of the albumentations graphs, see header)     #
# --------------------------------------------------------------------------- #

PH_RBC, PH_HSV, PH_GAUSS_NOISE, PH_GAUSS_BLUR, PH_ERASE, PH_UNSUPPORTED = 1, 2, 3, 4, 5, 6
PH_ISO_NOISE, PH_SHOT_NOISE, PH_MEDIAN_BLUR, PH_MOTION_BLUR, PH_GLASS_BLUR = 7, 8, 9, 10, 11
FILL_RANDOM, FILL_RANDOM_UNIFORM, FILL_ONE, FILL_ZERO = 0, 1, 2, 3


def gaussian_kernel_1d(sigma: float) -> np.ndarray:
    ksize = max(3, int(sigma * 6 + 1) | 1)
    r = ksize // 2
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-(x * x) / (2.0 * sigma * sigma))
    return (k / k.sum()).astype(np.float32)


def draw_rbc(brightness, contrast):
    alpha = 1.0 + np.random.uniform(*contrast)
    beta = np.random.uniform(*brightness)
    return {"ph": PH_RBC, "alpha": alpha, "beta": beta}


def draw_hsv(hue, sat, val):
    return {"ph": PH_HSV, "hue": np.random.uniform(*hue), "sat": np.random.uniform(*sat), "val": np.random.uniform(*val)}


def draw_gauss_noise(std_range, hw):
    sigma = np.random.uniform(*std_range)
    field = np.random.normal(0.0, 1.0, (hw[0], hw[1], 3)).astype(np.float32)  # unit normals, scaled by sigma on apply
    return {"ph": PH_GAUSS_NOISE, "sigma": sigma, "field": field}


def draw_gauss_blur(sigma_limit):
    return {"ph": PH_GAUSS_BLUR, "sigma": np.random.uniform(*sigma_limit)}


def draw_iso_noise(color_shift, intensity=(0.1, 0.5)):
    """A.ISONoise(color_shift=(0.01, 0.4)) with its default intensity=(0.1, 0.5).  The noise fields depend on the image
    (Poisson rate from the luminance spread), so they are drawn when the op is applied and recorded into this dict."""
    return {"ph": PH_ISO_NOISE, "color_shift": np.random.uniform(*color_shift), "intensity": np.random.uniform(*intensity)}


def draw_shot_noise(scale_range):
    """A.ShotNoise(scale_range=(0.1, 0.3)); the Poisson counts are drawn on apply (rate = linearised image / scale)."""
    return {"ph": PH_SHOT_NOISE, "scale": np.random.uniform(*scale_range)}


def draw_median_blur(blur_limit):
    ks = list(range(blur_limit[0] | 1, blur_limit[1] + 1, 2))
    return {"ph": PH_MEDIAN_BLUR, "ksize": int(ks[int(np.random.randint(len(ks)))])}


def motion_kernel_mask(ksize, x1, y1, x2, y2):
    k = np.zeros((ksize, ksize), dtype=np.uint8)
    cv2.line(k, (int(x1), int(y1)), (int(x2), int(y2)), 1, thickness=1)
    return k


def draw_motion_blur(blur_limit):
    """A.MotionBlur(blur_limit=(3, 11)), allow_shifted: a one-pixel line between two random cells of a ksize x ksize
    kernel (cv2.line), normalised.  When both x coordinates coincide the y pair is drawn distinct, so the line never
    degenerates to a point."""
    ks = list(range(blur_limit[0] | 1, blur_limit[1] + 1, 2))
    ksize = int(ks[int(np.random.randint(len(ks)))])
    x1, x2 = int(np.random.randint(ksize)), int(np.random.randint(ksize))
    if x1 == x2:
        y1 = int(np.random.randint(ksize))
        y2 = int(np.random.randint(ksize - 1))
        y2 += y2 >= y1
    else:
        y1, y2 = int(np.random.randint(ksize)), int(np.random.randint(ksize))
    return {"ph": PH_MOTION_BLUR, "ksize": ksize, "pts": (x1, y1, x2, y2), "mask": motion_kernel_mask(ksize, x1, y1, x2, y2)}


def draw_glass_blur(hw, sigma=0.5, max_delta=4, iterations=2):
    """A.GlassBlur(sigma=0.5, max_delta=4, iterations=2), mode "fast": one (dy, dx) in [-max_delta, max_delta) per interior
    pixel and iteration (get_params_dependent_on_data)."""
    n = (hw[0] - 2 * max_delta) * (hw[1] - 2 * max_delta)
    dxy = np.random.randint(-max_delta, max_delta, size=(n, iterations, 2)).astype(np.int32)
    return {"ph": PH_GLASS_BLUR, "sigma": sigma, "max_delta": max_delta, "iterations": iterations, "dxy": dxy}


def draw_erase(scale, fill, hw):
    h, w = hw
    area = np.random.uniform(*scale) * h * w
    aspect = math.exp(np.random.uniform(math.log(0.3), math.log(3.3)))
    eh, ew = int(round(math.sqrt(area * aspect))), int(round(math.sqrt(area / aspect)))
    rec: dict[str, Any] = {"ph": PH_ERASE, "fill": int(fill), "eh": eh, "ew": ew, "active": False}
    if eh < 1 or ew < 1 or eh >= h or ew >= w:
        return rec
    rec["top"] = int(np.random.randint(0, h - eh + 1))
    rec["left"] = int(np.random.randint(0, w - ew + 1))
    rec["active"] = True
    if fill == FILL_RANDOM:
        rec["field"] = np.random.uniform(0, 1, (eh, ew, 3)).astype(np.float32)
    elif fill == FILL_RANDOM_UNIFORM:
        rec["color"] = np.random.uniform(0, 1, 3).astype(np.float32)
    return rec


def apply_photo(img: np.ndarray, rec: dict) -> np.ndarray:
    ph = rec["ph"]
    if ph == PH_RBC:  # RandomBrightnessContrast, brightness_by_max=True on float32: img*alpha + beta*1.0, clip
        return np.clip(img * np.float32(rec["alpha"]) + np.float32(rec["beta"]), 0, 1).astype(np.float32)
    if ph == PH_HSV:
        if rec["hue"] == 0 and rec["sat"] == 0 and rec["val"] == 0:
            return img
        hsv = cv2.cvtColor(np.ascontiguousarray(img, dtype=np.float32), cv2.COLOR_RGB2HSV)
        h, s, v = hsv[:, :, 0], hsv[:, :, 1], hsv[:, :, 2]
        h = np.mod(h + np.float32(rec["hue"]), np.float32(360))
        s = np.clip(s + np.float32(rec["sat"] / 255.0), 0, 1)
        v = np.clip(v + np.float32(rec["val"] / 255.0), 0, 1)
        return cv2.cvtColor(cv2.merge([h, s, v]), cv2.COLOR_HSV2RGB)
    if ph == PH_GAUSS_NOISE:
        return np.clip(img + rec["field"] * np.float32(rec["sigma"]), 0, 1).astype(np.float32)
    if ph == PH_GAUSS_BLUR:
        k = gaussian_kernel_1d(rec["sigma"])
        return cv2.sepFilter2D(np.ascontiguousarray(img, dtype=np.float32), -1, k, k, borderType=cv2.BORDER_REFLECT_101)
    if ph == PH_ERASE:
        if not rec["active"]:
            return img
        out = img.copy()
        t, l, eh, ew = rec["top"], rec["left"], rec["eh"], rec["ew"]
        if rec["fill"] == FILL_RANDOM:
            out[t : t + eh, l : l + ew] = rec["field"]
        elif rec["fill"] == FILL_RANDOM_UNIFORM:
            out[t : t + eh, l : l + ew] = rec["color"]
        else:
            out[t : t + eh, l : l + ew] = 1.0 if rec["fill"] == FILL_ONE else 0.0
        return out
    if ph == PH_ISO_NOISE:
        # camera-sensor noise in HLS space: hue jitter ~ N(0, color_shift * 360 * intensity), luminance lifted towards 1 by
        # Poisson(std(L) * intensity * 255) / 255 of the remaining headroom (cv2.cvtColor RGB<->HLS on float32, cv2.meanStdDev)
        img = np.ascontiguousarray(img, dtype=np.float32)
        hls = cv2.cvtColor(img, cv2.COLOR_RGB2HLS)
        _, std = cv2.meanStdDev(hls)
        if "lum" not in rec:
            rec["lam"] = float(std[1, 0]) * rec["intensity"] * 255.0
            rec["lum"] = np.random.poisson(rec["lam"], size=hls.shape[:2]).astype(np.float32)
            rec["col"] = np.random.normal(0.0, 1.0, hls.shape[:2]).astype(np.float32)
        hue = np.mod(hls[:, :, 0] + rec["col"] * np.float32(rec["color_shift"] * 360.0 * rec["intensity"]), np.float32(360))
        lum = hls[:, :, 1] + (rec["lum"] / np.float32(255)) * (np.float32(1) - hls[:, :, 1])
        out = cv2.cvtColor(cv2.merge([hue.astype(np.float32), lum.astype(np.float32), hls[:, :, 2]]), cv2.COLOR_HLS2RGB)
        return np.clip(out, 0, 1).astype(np.float32)
    if ph == PH_SHOT_NOISE:
        # photon noise in linear light: gamma 2.2 -> Poisson((x + scale * 1e-6) / scale) * scale -> clip -> gamma 1 / 2.2
        img = np.ascontiguousarray(img, dtype=np.float32)
        scale = np.float32(rec["scale"])
        lin = cv2.pow(np.clip(img, 0, 1), 2.2)
        lam = (lin + scale * np.float32(1e-6)) / scale
        if "field" not in rec:
            rec["field"] = np.random.poisson(lam).astype(np.float32)
        return cv2.pow(np.clip(rec["field"] * scale, 0, 1), 1.0 / 2.2).astype(np.float32)
    if ph == PH_MEDIAN_BLUR:
        # cv2.medianBlur supports float32 only up to ksize 5: the image goes through uint8 (albumentations' uint8_io)
        u8 = np.rint(np.clip(img, 0, 1) * np.float32(255)).astype(np.uint8)
        return np.divide(cv2.medianBlur(np.ascontiguousarray(u8), rec["ksize"]), 255.0, dtype=np.float32)
    if ph == PH_MOTION_BLUR:
        k = rec["mask"].astype(np.float32) / np.float32(rec["mask"].sum())
        return cv2.filter2D(np.ascontiguousarray(img, dtype=np.float32), -1, k)  # BORDER_REFLECT_101
    if ph == PH_GLASS_BLUR:
        # albumentations' glass_blur, mode "fast": blur, `iterations` rounds of pixel swaps with a random neighbour (numpy's
        # simultaneous fancy assignment: the gathers are taken first, then x[h,w] is written, then x[h+dy,w+dx] - where two
        # sources target the same pixel the later index wins), blur again
        md, sigma = rec["max_delta"], rec["sigma"]
        x = cv2.GaussianBlur(np.array(img, dtype=np.float32), sigmaX=sigma, ksize=(0, 0))
        hs = np.arange(img.shape[0] - md, md, -1)
        ws = np.arange(img.shape[1] - md, md, -1)
        h = np.tile(hs, ws.shape[0])
        w = np.repeat(ws, hs.shape[0])
        for i in range(rec["iterations"]):
            dy, dx = rec["dxy"][:, i, 0], rec["dxy"][:, i, 1]
            x[h, w], x[h + dy, w + dx] = x[h + dy, w + dx], x[h, w]
        return cv2.GaussianBlur(x, sigmaX=sigma, ksize=(0, 0))
    return img  # PH_UNSUPPORTED: drawn, applied as identity


def _maybe(p: float, fn):
    """albumentations applies a transform with probability p; parameters are drawn only when it fires."""
    return fn() if np.random.random() < p else None


def _random_order(makers: list, n: int) -> list:
    """A.RandomOrder(n of k, replace=False): a random n-subset in random order, each wrapped so its own p applies."""
    order = np.random.permutation(len(makers))[:n]
    out = []
    for i in order:
        rec = makers[int(i)]()
        if rec is not None:
            out.append(rec)
    return out


def _one_of(makers: list):
    return makers[int(np.random.randint(len(makers)))]()


def draw_bg_light(fill, hw):
    """get_bg_transform_light (od_datasets.py:420-438)."""
    return _random_order([
        lambda: _maybe(0.5, lambda: draw_rbc((-0.4, 0.4), (-0.4, 0.4))),
        lambda: _maybe(0.2, lambda: draw_gauss_blur((0, 2))),
        lambda: _maybe(0.2, lambda: draw_gauss_noise((0.0, 0.1), hw)),
        lambda: _maybe(0.4, lambda: draw_erase((0.02, 0.2), fill, hw)),
    ], 3)


def draw_card_transform(fill, hw):
    """get_card_transform (od_datasets.py:490-512); contrast_limit=(-0.4,-0.4) is degenerate."""
    return _random_order([
        lambda: _maybe(0.8, lambda: draw_rbc((-0.2, 0.2), (-0.4, -0.4))),
        lambda: _maybe(0.8, lambda: draw_hsv((-30, 30), (-40, 40), (0, 0))),
        lambda: _maybe(0.3, lambda: draw_erase((0.02, 0.2), fill, hw)),
    ], 2)


def draw_bg_transform(hw, extra=False, fill=None):
    """get_bg_transform (od_datasets.py:441-487)."""
    unsupported = {"ph": PH_UNSUPPORTED}

    def noise(p):
        return lambda: _one_of([lambda: _maybe(p, lambda: draw_gauss_noise((0.0, 0.2), hw)),
                                lambda: _maybe(p, lambda: draw_iso_noise((0.01, 0.4))),
                                lambda: _maybe(p, lambda: draw_shot_noise((0.1, 0.3)))])

    def blur(p):
        return lambda: _one_of([lambda: _maybe(p, lambda: draw_gauss_blur((0, 3))),
                                lambda: _maybe(p, lambda: draw_median_blur((3, 7))),
                                lambda: _maybe(p, lambda: draw_motion_blur((3, 11))),
                                lambda: _maybe(p, lambda: draw_motion_blur((3, 11))),
                                lambda: _maybe(p / 3 * 2, lambda: draw_glass_blur(hw))])

    makers = [
        lambda: _maybe(0.5, lambda: draw_rbc((-0.4, 0.4) if not extra else (-0.7, 0.7), (-0.5, 0.5))),
        lambda: _maybe(0.5, lambda: draw_hsv((-30, 30), (-40, 40), (0, 0) if not extra else (-30, 30))),
        noise(0.5), blur(0.5), noise(0.1), blur(0.1),
    ]
    if extra:
        makers.append(lambda: _maybe(0.4, lambda: draw_erase((0.02, 0.4), fill, hw)))
    return [r for r in _random_order(makers, 4) if r["ph"] != PH_UNSUPPORTED]


def draw_fill(p):
    return int(np.random.choice(4, p=p))  # ["random", "random_uniform", 1, 0]


# --------------------------------------------------------------------------- #
# scene generation                                                             #
# --------------------------------------------------------------------------- #


class DetOracle:
    """Gen (od_datasets.py:619-704) over resident uint8 pools."""

    def __init__(self, cards_u8, bgs_u8, *, bg_size_hw=640, num_cards_min=1, num_cards_max=10, card_min_visible_ratio=0.5,
                 card_min_visible_ratio_edges=1.0, card_jitter_ratio=0.3, card_min_area_ratio=0.02, card_max_area_ratio=0.9,
                 card_no_contains=True, card_max_place_attempts=10, ratio_bg=None, kind="obb", photometrics=True):
        self.cards, self.bgs = cards_u8, bgs_u8
        self.S = (bg_size_hw, bg_size_hw) if isinstance(bg_size_hw, int) else tuple(bg_size_hw)
        self.num_cards_min, self.num_cards_max = num_cards_min, num_cards_max
        self.min_visible = card_min_visible_ratio
        self.min_visible_edges = card_min_visible_ratio_edges
        self.jitter = card_jitter_ratio
        self.min_area, self.max_area = card_min_area_ratio, card_max_area_ratio
        self.no_contains, self.max_attempts = card_no_contains, card_max_place_attempts
        self.ratio_bg, self.kind, self.photometrics = ratio_bg, kind, photometrics
        ch, cw = cards_u8[0].shape[:2]
        self.card_hw = (ch, cw)
        self.mask = EO.round_rect_mask((ch, cw), radius_ratio=0.046)
        self.keypoints = card_keypoints(ch, cw, kind)
        self.bbox = np.asarray(box(0, 0, cw, ch))

    # -- place_card_on_background_get_transform (od_datasets.py:287-377)
    def place(self, existing: list, attempts_out: list):
        bh, bw = self.S
        ch, cw = self.card_hw
        card_diag = int(math.hypot(ch, cw))
        edge = self.min_visible if self.min_visible_edges is None else self.min_visible_edges
        edge = max(self.min_visible, edge)
        for _ in range(self.max_attempts):
            edge_pad = card_diag // 2
            edge_ovr = int(card_diag * (1 - edge))
            cx = random.randint(0 + edge_pad - edge_ovr, bw - edge_pad + edge_ovr)
            cy = random.randint(0 + edge_pad - edge_ovr, bh - edge_pad + edge_ovr)
            deg = np.random.uniform(0, 360)
            min_area, max_area = bh * bw * self.min_area, bh * bw * self.max_area
            target_area = np.exp(np.random.uniform(np.log(min_area), np.log(max_area)))
            scale = target_area / (ch * cw)
            src = np.asarray([(0, 0), (cw, 0), (cw, ch), (0, ch)])
            jitter_u = np.random.uniform(1 - self.jitter, 1 + self.jitter, size=4)
            dst = corner_jitter_2d(src.copy(), jitter_u)
            dst = rotate_2d(dst, deg=deg, center=(cw / 2, ch / 2), scale=scale)
            dst = dst + np.array([cx - (cw / 2) * scale, cy - (ch / 2) * scale])
            dst32 = dst.astype(np.float32)
            M = cv2.getPerspectiveTransform(src.astype(np.float32), dst32)
            kp = apply_transform_2d(self.keypoints, M)
            ok, why = placement_visible(CardShape(kp[0], self.kind), (bh, bw), existing, self.min_visible, edge, self.no_contains)
            attempts_out.append({"cx": cx, "cy": cy, "deg": float(deg), "area": float(target_area), "jitter": jitter_u,
                                 "dst": dst32, "M": M, "accepted": ok, "why": why})
            if ok:
                return M
        return None

    def make_background(self, bg_f32, tape):
        deg = int(np.random.randint(0, 360))
        M = get_rotate_over_output_transform(bg_f32.shape[:2], deg, self.S)
        tape.update(bg_deg=deg, bg_M=M, bg_hw=tuple(bg_f32.shape[:2]), size_hw=tuple(self.S))
        return cv2.warpPerspective(bg_f32, M, self.S[::-1], flags=cv2.INTER_LINEAR)

    def generate(self, tape: dict | None = None):
        """generate_synthetic_image (od_datasets.py:520-611)."""
        t = {} if tape is None else tape
        S = self.S
        fill_light = draw_fill([0.1, 0.5, 0.2, 0.2])   # get_bg_transform_light(): Erasing fill drawn at construction
        fill_card = draw_fill([0.1, 0.5, 0.2, 0.2])    # get_card_transform()
        bg_idx = EO._choice(len(self.bgs))             # bg_ds.ran_path()
        bg = self.make_background(EO.u8_to_f32(self.bgs[bg_idx]), t)
        pre = draw_bg_light(fill_light, S) if self.photometrics else []
        for rec in pre:
            bg = apply_photo(bg, rec)
        n_cards = int(np.random.randint(self.num_cards_min, self.num_cards_max))
        placed, collide, cards_tape = [], [], []
        for _ in range(n_cards):
            k = EO._choice(len(self.cards))            # mtg_ds.ran_path()
            ct: dict[str, Any] = {"card": k, "attempts": []}
            M = self.place(collide, ct["attempts"])
            cards_tape.append(ct)
            if M is None:
                continue
            ops = draw_card_transform(fill_card, self.card_hw) if self.photometrics else []
            img = EO.u8_to_f32(self.cards[k])
            for rec in ops:
                img = apply_photo(img, rec)
            ct["photo"] = ops
            placed.append((img, M, k))
            collide.append(apply_transform_2d(self.bbox, M))
        keypoints, labels, order = [], [], []
        for img, M, k in placed[::-1]:
            mask = cv2.warpPerspective(self.mask, M, S[::-1], flags=cv2.INTER_LINEAR)
            wimg = cv2.warpPerspective(img, M, S[::-1], flags=cv2.INTER_LINEAR)
            pts = apply_transform_2d(self.keypoints, M)
            bg = mask[:, :, None] * wimg + (1 - mask[:, :, None]) * bg
            keypoints.extend(pts)
            labels.extend(np.arange(len(self.keypoints)))
            order.append(k)
        post = draw_bg_transform(S) if self.photometrics else []
        for rec in post:
            bg = apply_photo(bg, rec)
        t.update(bg_only=False, bg=bg_idx, pre=pre, post=post, n_cards=n_cards, cards=cards_tape, composite_order=order)
        return {"image": bg, "keypoints": np.asarray(keypoints), "keypoints_labels": np.asarray(labels)}

    def random_bg(self, tape: dict | None = None):
        """Gen.random_bg / make_aug_background (od_datasets.py:206-210, 674-683)."""
        t = {} if tape is None else tape
        bg_idx = EO._choice(len(self.bgs))
        bg = self.make_background(EO.u8_to_f32(self.bgs[bg_idx]), t)
        fill_light = draw_fill([0.1, 0.5, 0.2, 0.2])
        pre = draw_bg_light(fill_light, self.S) if self.photometrics else []
        fill_extra = draw_fill([0.1, 0.1, 0.4, 0.4])
        post = draw_bg_transform(self.S, extra=True, fill=fill_extra) if self.photometrics else []
        for rec in pre + post:
            bg = apply_photo(bg, rec)
        t.update(bg_only=True, bg=bg_idx, pre=pre, post=post, n_cards=0, cards=[], composite_order=[])
        return {"image": bg, "keypoints": np.zeros((0, self.keypoints.shape[1], 2)), "keypoints_labels": np.zeros((0,), dtype=np.int64)}

    def random(self, tape: dict | None = None):
        """Gen.random (od_datasets.py:685-704)."""
        if self.ratio_bg and np.random.uniform(0, 1) < self.ratio_bg:
            return self.random_bg(tape)
        return self.generate(tape)
