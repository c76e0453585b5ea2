"""Subsystem (1) on the CPU: the host/device parameter-expansion headers, compiled by g++
into tests/host_harness, must reproduce bit for bit the matrices/geometry the reference's
cv2 calls produced while the oracle recorded its tapes."""
import ctypes as C
import os
import random

import cv2
import numpy as np
import pytest

from mtgvision_b200 import abi, synth
from oracle import cv2_restate as R
from oracle import encoder_oracle as EO
from oracle import tape_pack

HARNESS = os.path.join(os.path.dirname(__file__), "host_harness", "libmtgv_hostharness.so")
vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731


@pytest.fixture(scope="module")
def hh():
    return C.CDLL(HARNESS)


@pytest.fixture(scope="module")
def recorded():
    pool = synth.CardPool(np.stack([synth.synth_card(k) for k in range(4)]), synth.synth_faces(4))
    bgs = [synth.synth_bg(j) for j in range(4)]
    tapes = []
    for seed in range(160):
        random.seed(seed); np.random.seed(seed); EO.reset_shuffle_state()
        t = {"card": seed % 4, "bg": (seed // 4) % 4}
        EO.make_virtual(EO.u8_to_f32(pool.images[seed % 4]), EO.u8_to_f32(bgs[(seed // 4) % 4]), (192, 128), True, tape=t)
        tapes.append(t)
    return pool, bgs, tapes


def _expand(hh, pool, bgs, tapes, **kw):
    arr, _ = tape_pack.pack_tapes(tapes, **kw)
    cfg = abi.EncConfig(192, 128, 192, 128, 0.05, 0.2, 1, 1, 0, 0)
    params = np.zeros(len(tapes), dtype=abi.PARAMS_DTYPE)
    bg_hw = np.asarray([b.shape[:2] for b in bgs], dtype=np.int32)
    bad = hh.hh_expand_encoder(vp(arr), len(tapes), C.byref(cfg), 680, 488, len(pool), vp(pool.labels3), vp(pool.grp_off),
                               vp(pool.grp_mem), len(bgs), vp(bg_hw), vp(params))
    assert bad == 0
    return params


def test_expansion_matches_cv2_matrices_bit_exact(hh, recorded):
    pool, bgs, tapes = recorded
    params = _expand(hh, pool, bgs, tapes)
    n = {"persp": 0, "affine": 0, "rot": 0}
    for t, p in zip(tapes, params):
        assert (p["fg_rh"], p["fg_rw"], p["fg_y0"], p["fg_x0"]) == (178, 128, 7, 0)
        k = 0
        for rec in t["fg_ops"]:
            if rec["op"] == "downup" and rec["n"] == 0:
                continue
            x = p["ops"][k]; k += 1
            if rec["op"] in ("warp", "perspective"):
                n["persp"] += 1
                assert np.array_equal(cv2.invert(rec["M"])[1].reshape(-1), x["d"])
            elif rec["op"] == "affine":
                n["affine"] += 1
                assert np.array_equal(R.invert_affine(rec["M"]).reshape(-1), x["d"][:6])
        for rec in t["bg_ops"]:
            if rec["op"] == "rotate":
                n["rot"] += 1
                assert np.array_equal(R.invert_affine(rec["M"]).reshape(-1), p["rot_inv"])
                assert (rec["nh"], rec["nw"]) == (p["rot_nh"], p["rot_nw"])
                g = EO.crop_to_size_geometry((rec["nh"], rec["nw"]), (192, 128), False)
                assert g == (p["bg_rh"], p["bg_rw"], p["bg_y0"], p["bg_x0"])
            elif rec["op"] == "warp_inv":
                assert np.array_equal(cv2.invert(rec["M"])[1].reshape(-1), p["winv"])
    assert min(n.values()) > 10


def test_expansion_with_device_transcendentals_is_close(hh, recorded):
    """Without the host-libm alpha/beta override (production mode) the matrices agree to a few ulp."""
    pool, bgs, tapes = recorded
    a = _expand(hh, pool, bgs, tapes, host_transcendentals=True)
    b = _expand(hh, pool, bgs, tapes, host_transcendentals=False)
    assert np.allclose(a["rot_inv"], b["rot_inv"], rtol=1e-13, atol=1e-10)
    assert np.array_equal(a["rot_nh"], b["rot_nh"]) and np.array_equal(a["rot_nw"], b["rot_nw"])


def test_coordinate_generators_match_restatement(hh):
    rng = np.random.default_rng(3)
    for (dh, dw) in [(192, 128), (640, 640), (50, 40), (618, 603)]:
        src = np.float32([[0, 0], [dw, 0], [0, dh], [dw, dh]])
        dst = (src + rng.uniform(-0.2, 0.2, (4, 2)) * [dw, dh]).astype(np.float32)
        Mi = cv2.invert(cv2.getPerspectiveTransform(src, dst))[1]
        X = np.zeros((dh, dw), np.int32); Y = np.zeros((dh, dw), np.int32)
        hh.hh_persp_coords(vp(np.ascontiguousarray(Mi)), dh, dw, vp(X), vp(Y))
        Xr, Yr = R.warp_perspective_coords(Mi, (dw, dh))
        assert np.array_equal(X, Xr) and np.array_equal(Y, Yr)
        A = cv2.getRotationMatrix2D((dw / 2, dh / 2), rng.uniform(0, 360), 1.0)
        Ai = R.invert_affine(A)
        hh.hh_affine_coords(vp(np.ascontiguousarray(Ai)), dh, dw, vp(X), vp(Y))
        Xr, Yr = R.warp_affine_coords(Ai, (dw, dh))
        assert np.array_equal(X, Xr) and np.array_equal(Y, Yr)


def test_area_taps_match_restatement(hh):
    for ssize, dsize in [(488, 128), (680, 178), (652, 192), (460, 128), (618, 191), (500, 256), (375, 192), (192, 192)]:
        ref = {}
        for d, s, a in R.area_tab(ssize, dsize):
            ref.setdefault(d, []).append((s, a))
        for d in range(dsize):
            start = C.c_int(0)
            w = (C.c_float * 8)()
            n = hh.hh_area_taps(ssize, dsize, d, C.byref(start), w)
            assert n == len(ref[d]) and start.value == ref[d][0][0]
            assert [np.float32(w[k]) for k in range(n)] == [a for _, a in ref[d]]


def test_area_linear_taps_reproduce_cv2_enlarging_inter_area(hh):
    """cv2.resize(INTER_AREA) with either axis enlarged (crop_to_size of a small background, util/image.py:321-334)
    is a separable 2-tap filter with the 'area' coefficient rule: rebuilt from the harness tables it must equal cv2."""
    rng = np.random.default_rng(0)
    for (sh, sw), (dh, dw) in [((90, 120), (192, 256)), ((60, 50), (230, 192)), ((128, 100), (192, 150)), ((191, 127), (192, 128)),
                               ((100, 300), (192, 576)), ((7, 5), (192, 137))]:
        src = rng.random((sh, sw, 3), dtype=np.float32)
        ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA)

        def table(ssize, dsize):
            T = np.zeros((dsize, ssize), np.float64)
            for d in range(dsize):
                start = C.c_int(0)
                w = (C.c_float * 2)()
                n = hh.hh_area_linear_taps(ssize, dsize, d, C.byref(start), w)
                for k in range(n):
                    T[d, start.value + k] += w[k]
            return T

        got = np.einsum("ys,sxc->yxc", table(sh, dh), np.einsum("xs,ysc->yxc", table(sw, dw), src.astype(np.float64)))
        assert np.abs(got - ref).max() < 1e-6, ((sh, sw), (dh, dw), np.abs(got - ref).max())


def test_host_mask_matches_reference_kat(hh):
    import hashlib
    for hw, rad, sha in [((680, 488), 34, "75fac4e48730e3c0"), ((680, 488), 32, "d7923dd2973e78fe")]:
        m = np.zeros(hw, np.float32)
        hh.hh_round_rect_mask(hw[0], hw[1], rad, vp(m))
        assert hashlib.sha1(m.tobytes()).hexdigest()[:16] == sha
