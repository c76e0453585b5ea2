"""The placement accept/reject tests (place_card_on_background_get_transform, od_datasets.py:353-372) are shapely calls in
the reference; shapely/GEOS is absent here, so the oracle restates them with fp64 convex clipping ("parity unpinned").
This test bounds what that can cost: every decision of the oracle on recorded scenes is re-evaluated in EXACT rational
arithmetic on the same float64 vertices (what any correct polygon library converges to), must agree, and the smallest
distance of a tested ratio from its threshold is orders of magnitude above fp64 polygon-area error (~1e-13) - so GEOS
cannot decide differently on these scenes either."""
import random
from fractions import Fraction as Fr

import numpy as np
import pytest

from mtgvision_b200 import synth
from oracle import det_oracle as DO

AUTHOR_RUN = dict(bg_size_hw=640, num_cards_min=1, num_cards_max=9, card_min_visible_ratio=0.5,
                  card_min_visible_ratio_edges=0.0, card_jitter_ratio=0.7)  # od_datasets.py:861-868


def fr_poly(p):
    return [(Fr(float(x)), Fr(float(y))) for x, y in p]


def signed2(p):
    return sum(p[i][0] * p[(i + 1) % len(p)][1] - p[(i + 1) % len(p)][0] * p[i][1] for i in range(len(p)))


def area(p):
    return abs(signed2(p)) / 2 if len(p) >= 3 else Fr(0)


def clip(subj, conv):
    a = list(subj)
    o = 1 if signed2(conv) >= 0 else -1
    for e in range(len(conv)):
        if not a:
            break
        (ex, ey), (fx, fy) = conv[e], conv[(e + 1) % len(conv)]
        dx, dy = fx - ex, fy - ey
        b = []
        for i in range(len(a)):
            (px, py), (qx, qy) = a[i], a[(i + 1) % len(a)]
            sp = o * (dx * (py - ey) - dy * (px - ex))
            sq = o * (dx * (qy - ey) - dy * (qx - ex))
            if sp >= 0:
                b.append((px, py))
            if (sp >= 0) != (sq >= 0):
                t = sp / (sp - sq)
                b.append((px + t * (qx - px), py + t * (qy - py)))
        a = b
    return a


def inside(pt, conv):
    o = 1 if signed2(conv) >= 0 else -1
    return all(o * ((conv[(e + 1) % len(conv)][0] - conv[e][0]) * (pt[1] - conv[e][1])
                    - (conv[(e + 1) % len(conv)][1] - conv[e][1]) * (pt[0] - conv[e][0])) >= 0 for e in range(len(conv)))


def exact_decision(shape, S_hw, existing, min_visible, min_visible_edge, no_contains, margins):
    quad, indent = fr_poly(shape.quad), (fr_poly(shape.indent) if shape.indent is not None else None)

    def area_within(conv):
        q = quad if conv is None else clip(quad, conv)
        a = area(q)
        if indent is not None and len(q) >= 3:
            a -= area(clip(q, indent))
        return a

    bh, bw = S_hw
    img = fr_poly([(0, 0), (bw, 0), (bw, bh), (0, bh)])
    card_area, vis_area = area_within(None), area_within(img)
    margins.append(abs(vis_area / card_area - Fr(float(min_visible_edge))))
    if vis_area / card_area < Fr(float(min_visible_edge)):
        return False
    visible = True
    for p in existing:
        pq = fr_poly(p)
        pc = clip(pq, img)
        inter = area_within(pc) if len(pc) >= 3 else Fr(0)
        margins.append(abs((vis_area - inter) / card_area - Fr(float(min_visible))))
        if (vis_area - inter) / card_area < Fr(float(min_visible)):
            visible = False
            break
        p_area = area(pq)
        margins.append(abs((p_area - inter) / p_area - Fr(float(min_visible))))
        if (p_area - inter) / p_area < Fr(float(min_visible)):
            visible = False
            break
        vis_pts = clip(quad, img)
        p_contains_vis = len(vis_pts) >= 3 and all(inside(v, pq) for v in vis_pts)
        vis_contains_p = all(inside(v, img) and inside(v, quad) for v in pq) and (indent is None or area(clip(pq, indent)) == 0)
        if (no_contains and p_contains_vis) or vis_contains_p:
            visible = False
    return visible


@pytest.mark.parametrize("kind", ["obb", "seg"])
def test_oracle_decisions_equal_exact_rational_arithmetic(kind, monkeypatch):
    cards, bgs = [synth.synth_card(k) for k in range(4)], [synth.synth_bg(j) for j in range(2)]
    calls = []
    real = DO.placement_visible

    def recording(shape, S_hw, existing, min_visible, min_visible_edge, no_contains=True):
        res = real(shape, S_hw, existing, min_visible, min_visible_edge, no_contains)
        calls.append((shape, S_hw, [np.array(p, dtype=np.float64) for p in existing], min_visible, min_visible_edge, no_contains, res[0]))
        return res

    monkeypatch.setattr(DO, "placement_visible", recording)
    for seed, kw in [(s, {}) for s in range(16)] + [(s, {"card_min_visible_ratio_edges": 0.75, "num_cards_max": 6}) for s in range(16, 24)]:
        random.seed(seed); np.random.seed(seed)
        DO.DetOracle(cards, bgs, kind=kind, photometrics=False, **{**AUTHOR_RUN, **kw}).generate({})
    assert len(calls) > 100
    margins = []
    accepted = 0
    for shape, S_hw, existing, mv, mve, nc, got in calls:
        assert exact_decision(shape, S_hw, existing, mv, mve, nc, margins) == got
        accepted += got
    assert 0 < accepted < len(calls)  # both outcomes exercised
    nonzero = [float(m) for m in margins if m != 0]  # ratio == threshold exactly only for min_visible_edge = 0 (0/.. vs 0)
    assert min(nonzero) > 1e-9, min(nonzero)
