"""TEST INFRASTRUCTURE ONLY - numpy restatements of the OpenCV arithmetic on the hot path.

The reference (nmichlo/mtg-vision) has no native code: every pixel on its sample
generator path is produced by an OpenCV call (pinned here: opencv-python 4.13.0,
the version installed in this image and on the GPU box).  This module restates, in
plain numpy, the arithmetic of exactly those calls so that

  * the CUDA kernels have a written specification (each function below is what one
    device routine in `mtgvision_b200/csrc/` implements), and
  * the restatement itself can be pinned: `tests/test_cv2_restate.py` compares every
    function here against the real `cv2` call (bit-exact where stated).

Reference call sites (file:line under /root/reference):
  cv2.getPerspectiveTransform  mtgvision/encoder_datasets.py:109,402  od_datasets.py:347
  cv2.getRotationMatrix2D      encoder_datasets.py:369  util/image.py:390  od_datasets.py:54,106
  cv2.warpPerspective          encoder_datasets.py:111,403  od_datasets.py:82
  cv2.warpAffine               encoder_datasets.py:375  util/image.py:398
  cv2.resize                   util/image.py:328-332  encoder_datasets.py:161-162
  cv2.GaussianBlur/filter2D    encoder_datasets.py:140,246
  cv2.circle                   util/image.py:418

Nothing under `mtgvision_b200/` may import this module; only tests/, bench.py's
cpu_baseline leg and __graft_entry__.smoke() do.
"""

from __future__ import annotations

import math

import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS  # 32
AB_BITS = 10
AB_SCALE = 1 << AB_BITS  # 1024

INTER_NEAREST, INTER_LINEAR, INTER_CUBIC, INTER_AREA = 0, 1, 2, 3


# --------------------------------------------------------------------------- #
# small fp64 helpers                                                           #
# --------------------------------------------------------------------------- #


def invert3x3(M: np.ndarray) -> np.ndarray:
    """cv::invert for a 3x3 CV_64F matrix (DECOMP_LU small-matrix path):
    cofactor formula scaled by 1/det, every product/sum a separate fp64 rounding."""
    m = np.asarray(M, dtype=np.float64)
    a, b, c = m[0]
    d, e, f = m[1]
    g, h, i = m[2]
    det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g)
    if det == 0.0:
        return np.zeros((3, 3))
    s = 1.0 / det
    t = np.empty((3, 3))
    t[0, 0] = (e * i - f * h) * s
    t[0, 1] = (c * h - b * i) * s
    t[0, 2] = (b * f - c * e) * s
    t[1, 0] = (f * g - d * i) * s
    t[1, 1] = (a * i - c * g) * s
    t[1, 2] = (c * d - a * f) * s
    t[2, 0] = (d * h - e * g) * s
    t[2, 1] = (b * g - a * h) * s
    t[2, 2] = (a * e - b * d) * s
    return t


def invert_affine(A: np.ndarray) -> np.ndarray:
    """cv::invertAffineTransform as inlined in cv::warpAffine (imgwarp.cpp):
    D = 1/(a00 a11 - a01 a10); the translation is pushed through the new 2x2."""
    M = np.asarray(A, dtype=np.float64).reshape(2, 3).copy()
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11 = M[1, 1] * D
    A22 = M[0, 0] * D
    M[0, 0] = A11
    M[0, 1] *= -D
    M[1, 0] *= -D
    M[1, 1] = A22
    b1 = -M[0, 0] * M[0, 2] - M[0, 1] * M[1, 2]
    b2 = -M[1, 0] * M[0, 2] - M[1, 1] * M[1, 2]
    M[0, 2] = b1
    M[1, 2] = b2
    return M


def get_rotation_matrix_2d(center, angle_deg: float, scale: float) -> np.ndarray:
    """cv::getRotationMatrix2D: centre is a Point2f (rounded to float32), the angle is
    converted with CV_PI/180 and alpha/beta come from the host libm cos/sin."""
    cx = float(np.float32(center[0]))
    cy = float(np.float32(center[1]))
    a = angle_deg * (math.pi / 180.0)
    alpha = math.cos(a) * scale
    beta = math.sin(a) * scale
    return rotation_matrix_from_ab(cx, cy, alpha, beta)


def rotation_matrix_from_ab(cx: float, cy: float, alpha: float, beta: float) -> np.ndarray:
    """The part of getRotationMatrix2D after the transcendental functions."""
    M = np.empty((2, 3))
    M[0, 0] = alpha
    M[0, 1] = beta
    M[0, 2] = (1 - alpha) * cx - beta * cy
    M[1, 0] = -beta
    M[1, 1] = alpha
    M[1, 2] = beta * cx + (1 - alpha) * cy
    return M


def get_perspective_transform(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """cv::getPerspectiveTransform(src f32[4,2], dst f32[4,2]) with DECOMP_LU.

    The 8x8 system is built in fp64 but the four products x*u, y*u, x*v, y*v are formed
    in float32 (Point2f arithmetic) and only then widened.  The solve is cv::LU
    (partial pivoting, first maximum, strict '>'), multiply and add as separate fp64
    roundings, then back substitution.
    """
    src = np.asarray(src, dtype=np.float32).reshape(4, 2)
    dst = np.asarray(dst, dtype=np.float32).reshape(4, 2)
    A = np.zeros((8, 8), dtype=np.float64)
    b = np.zeros(8, dtype=np.float64)
    for i in range(4):
        x, y = src[i]
        u, v = dst[i]
        A[i, 0] = A[i + 4, 3] = x
        A[i, 1] = A[i + 4, 4] = y
        A[i, 2] = A[i + 4, 5] = 1.0
        A[i, 6] = np.float32(-x) * u  # float32 product (numpy scalar arithmetic)
        A[i, 7] = np.float32(-y) * u
        A[i + 4, 6] = np.float32(-x) * v
        A[i + 4, 7] = np.float32(-y) * v
        b[i] = u
        b[i + 4] = v
    x = lu_solve8(A, b)
    if x is None:
        return np.zeros((3, 3))
    M = np.empty(9)
    M[:8] = x
    M[8] = 1.0
    return M.reshape(3, 3)


def lu_solve8(A: np.ndarray, b: np.ndarray):
    """cv::hal::LU64f + back substitution for one right-hand side (n = 8)."""
    A = A.copy()
    b = b.copy()
    n = A.shape[0]
    eps = np.finfo(np.float64).eps * 100
    for i in range(n):
        k = i
        for j in range(i + 1, n):
            if abs(A[j, i]) > abs(A[k, i]):
                k = j
        if abs(A[k, i]) < eps:
            return None
        if k != i:
            A[[i, k], i:] = A[[k, i], i:]
            b[[i, k]] = b[[k, i]]
        d = -1.0 / A[i, i]
        for j in range(i + 1, n):
            alpha = A[j, i] * d
            for kk in range(i + 1, n):
                A[j, kk] = A[j, kk] + alpha * A[i, kk]
            b[j] = b[j] + alpha * b[i]
    for i in range(n - 1, -1, -1):
        s = b[i]
        for k in range(i + 1, n):
            s = s - A[i, k] * b[k]
        b[i] = s / A[i, i]
    return b


# --------------------------------------------------------------------------- #
# bilinear remap core shared by warpPerspective / warpAffine                   #
# --------------------------------------------------------------------------- #

_F32 = np.float32


def _bilinear_tab() -> np.ndarray:
    """initInterTab2D(INTER_LINEAR): 32x32x4 float32 weights, w = vy * vx in float."""
    t1 = np.empty((INTER_TAB_SIZE, 2), dtype=np.float32)
    scale = _F32(1.0) / _F32(INTER_TAB_SIZE)
    for i in range(INTER_TAB_SIZE):
        x = _F32(i) * scale
        t1[i, 0] = _F32(1.0) - x
        t1[i, 1] = x
    tab = np.empty((INTER_TAB_SIZE, INTER_TAB_SIZE, 4), dtype=np.float32)
    for iy in range(INTER_TAB_SIZE):
        for ix in range(INTER_TAB_SIZE):
            tab[iy, ix, 0] = t1[iy, 0] * t1[ix, 0]
            tab[iy, ix, 1] = t1[iy, 0] * t1[ix, 1]
            tab[iy, ix, 2] = t1[iy, 1] * t1[ix, 0]
            tab[iy, ix, 3] = t1[iy, 1] * t1[ix, 1]
    return tab


BILINEAR_TAB = _bilinear_tab()


def remap_bilinear_fixed(src: np.ndarray, X: np.ndarray, Y: np.ndarray) -> np.ndarray:
    """remapBilinear<float> with BORDER_CONSTANT(0): X, Y are integer coordinates in
    1/32 px.  out = ((S00*w0 + S01*w1) + S10*w2) + S11*w3 in float32, taps outside the
    source read as 0.  (cv2 saturates the integer part to int16; sizes here are far
    below that.)"""
    src = np.asarray(src, dtype=np.float32)
    squeeze = src.ndim == 2
    if squeeze:
        src = src[:, :, None]
    h, w = src.shape[:2]
    sx = np.clip(X >> INTER_BITS, -32768, 32767).astype(np.int64)
    sy = np.clip(Y >> INTER_BITS, -32768, 32767).astype(np.int64)
    ax = (X & (INTER_TAB_SIZE - 1)).astype(np.int64)
    ay = (Y & (INTER_TAB_SIZE - 1)).astype(np.int64)
    wts = BILINEAR_TAB[ay, ax]  # (..., 4)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        v = src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]
        return np.where(ok[..., None], v, _F32(0))

    out = tap(sy, sx) * wts[..., 0:1]
    out = out + tap(sy, sx + 1) * wts[..., 1:2]
    out = out + tap(sy + 1, sx) * wts[..., 2:3]
    out = out + tap(sy + 1, sx + 1) * wts[..., 3:4]
    out = out.astype(np.float32)
    return out[:, :, 0] if squeeze else out


def warp_perspective_coords(Minv: np.ndarray, dsize_wh) -> tuple[np.ndarray, np.ndarray]:
    """WarpPerspectiveInvoker coordinate generation (fp64), per destination pixel:
    block origin bx = 64*floor(x/64) (BLOCK_SZ^2/16 columns when width >= 64),
    X0 = M0*bx + M1*y + M2 etc., then per column x1 = x - bx:
      W = W0 + M6*x1;  W = W ? 32/W : 0;  X = rint(clamp((X0 + M0*x1)*W))."""
    dw, dh = dsize_wh
    M = np.asarray(Minv, dtype=np.float64).reshape(9)
    bh0 = min(16, dh)
    bw0 = min(1024 // bh0, dw)
    x = np.arange(dw, dtype=np.int64)
    bx = (x // bw0) * bw0
    x1 = (x - bx).astype(np.float64)
    bx = bx.astype(np.float64)
    y = np.arange(dh, dtype=np.float64)[:, None]
    X0 = M[0] * bx[None, :] + M[1] * y + M[2]
    Y0 = M[3] * bx[None, :] + M[4] * y + M[5]
    W0 = M[6] * bx[None, :] + M[7] * y + M[8]
    W = W0 + M[6] * x1[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        W = np.where(W != 0, INTER_TAB_SIZE / W, 0.0)
    fX = np.clip((X0 + M[0] * x1[None, :]) * W, -2147483648.0, 2147483647.0)
    fY = np.clip((Y0 + M[3] * x1[None, :]) * W, -2147483648.0, 2147483647.0)
    X = np.rint(fX).astype(np.int64)
    Y = np.rint(fY).astype(np.int64)
    return X, Y


def warp_perspective(src: np.ndarray, M: np.ndarray, dsize_wh) -> np.ndarray:
    """cv2.warpPerspective(src f32, M, dsize, INTER_LINEAR, BORDER_CONSTANT 0)."""
    Minv = invert3x3(M)
    X, Y = warp_perspective_coords(Minv, dsize_wh)
    return remap_bilinear_fixed(src, X, Y)


def warp_perspective_u8(src: np.ndarray, M: np.ndarray, dsize_wh) -> np.ndarray:
    """cv2.warpPerspective on uint8 images (od_export.py:108, the serving-side dewarp): same
    coordinates as the float path, but remapBilinear's fixed-point blend: the four weights are
    saturate_cast<short>(w * 32768) = 32*(32-ax|ax)*(32-ay|ay) (exact, they sum to 32768) and the
    pixel is (sum(tap * weight) + 2^14) >> 15; taps outside the source are 0."""
    assert src.dtype == np.uint8 and src.ndim == 3
    X, Y = warp_perspective_coords(invert3x3(M), dsize_wh)
    sh, sw = src.shape[:2]
    sx = np.clip(X >> INTER_BITS, -32768, 32767)
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    ax, ay = X & (INTER_TAB_SIZE - 1), Y & (INTER_TAB_SIZE - 1)
    acc = np.zeros((*X.shape, src.shape[2]), dtype=np.int64)
    weights = (32 * (32 - ax) * (32 - ay), 32 * ax * (32 - ay), 32 * (32 - ax) * ay, 32 * ax * ay)
    for wgt, (dy, dx) in zip(weights, ((0, 0), (0, 1), (1, 0), (1, 1))):
        yy, xx = sy + dy, sx + dx
        ok = (yy >= 0) & (yy < sh) & (xx >= 0) & (xx < sw)
        tap = np.where(ok[..., None], src[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)].astype(np.int64), 0)
        acc += tap * wgt[..., None]
    return ((acc + (1 << 14)) >> 15).astype(np.uint8)


def warp_affine_coords(Ainv: np.ndarray, dsize_wh) -> tuple[np.ndarray, np.ndarray]:
    """WarpAffineInvoker coordinate generation: 10-bit fixed point per row/column
    tables, result in 1/32 px.  saturate_cast<int>(double) is round-half-even."""
    dw, dh = dsize_wh
    M = np.asarray(Ainv, dtype=np.float64).reshape(6)
    x = np.arange(dw, dtype=np.float64)
    adelta = np.rint(M[0] * x * AB_SCALE).astype(np.int64)
    bdelta = np.rint(M[3] * x * AB_SCALE).astype(np.int64)
    y = np.arange(dh, dtype=np.float64)
    rd = AB_SCALE // INTER_TAB_SIZE // 2
    X0 = np.rint((M[1] * y + M[2]) * AB_SCALE).astype(np.int64) + rd
    Y0 = np.rint((M[4] * y + M[5]) * AB_SCALE).astype(np.int64) + rd
    X = (X0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    return X, Y


def warp_affine(src: np.ndarray, A: np.ndarray, dsize_wh) -> np.ndarray:
    """cv2.warpAffine(src f32, A(2x3), dsize) with default INTER_LINEAR / constant 0."""
    Ainv = invert_affine(A)
    X, Y = warp_affine_coords(Ainv, dsize_wh)
    return remap_bilinear_fixed(src, X, Y)


# --------------------------------------------------------------------------- #
# resize                                                                       #
# --------------------------------------------------------------------------- #


def area_tab(ssize: int, dsize: int):
    """computeResizeAreaTab: list of (dst index, src index, float32 weight)."""
    scale = 1.0 / (dsize / ssize)  # cv::resize passes scale = 1./inv_scale
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1 = math.ceil(fsx1)
        sx2 = math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def resize_area(src: np.ndarray, dsize_wh) -> np.ndarray:
    """cv2.resize(f32, INTER_AREA) for a down-scale in both axes (general, non-integer
    path ResizeArea_): horizontal pass accumulates S*alpha per source row in table
    order, vertical pass accumulates beta*row in source-row order, all float32."""
    src = np.asarray(src, dtype=np.float32)
    squeeze = src.ndim == 2
    if squeeze:
        src = src[:, :, None]
    sh, sw, cn = src.shape
    dw, dh = dsize_wh
    assert sw >= dw and sh >= dh, "INTER_AREA restatement covers down-scaling only"
    xtab = area_tab(sw, dw)
    ytab = area_tab(sh, dh)
    hbuf = np.zeros((sh, dw, cn), dtype=np.float32)
    for dx, sx, a in xtab:
        hbuf[:, dx] = hbuf[:, dx] + src[:, sx] * a
    out = np.zeros((dh, dw, cn), dtype=np.float32)
    first = np.ones(dh, dtype=bool)
    for dy, sy, b in ytab:
        if first[dy]:
            out[dy] = hbuf[sy] * b
            first[dy] = False
        else:
            out[dy] = out[dy] + hbuf[sy] * b
    return out[:, :, 0] if squeeze else out


def resize_nearest(src: np.ndarray, dsize_wh) -> np.ndarray:
    """cv2.resize(INTER_NEAREST): sx = min(floor(dx * (1/(dsize/ssize))), ssize-1)."""
    sh, sw = src.shape[:2]
    dw, dh = dsize_wh
    ifx = 1.0 / (dw / sw)
    ify = 1.0 / (dh / sh)
    xs = np.minimum(np.floor(np.arange(dw) * ifx).astype(np.int64), sw - 1)
    ys = np.minimum(np.floor(np.arange(dh) * ify).astype(np.int64), sh - 1)
    return src[ys][:, xs]


def _linear_ofs(ssize: int, dsize: int):
    scale = 1.0 / (dsize / ssize)  # double, cv::resize: 1./inv_scale
    sx = np.empty(dsize, dtype=np.int64)
    fx = np.empty(dsize, dtype=np.float32)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)  # cv2 computes this in float
        s = math.floor(f)
        f = np.float32(f - s)
        if s < 0:
            s, f = 0, np.float32(0)
        if s >= ssize - 1:
            s, f = ssize - 1, np.float32(0)
        sx[d], fx[d] = s, f
    return sx, fx


def resize_linear(src: np.ndarray, dsize_wh) -> np.ndarray:
    """cv2.resize(f32, INTER_LINEAR): half-pixel centres, separable, float32."""
    src = np.asarray(src, dtype=np.float32)
    sh, sw = src.shape[:2]
    dw, dh = dsize_wh
    sx, fx = _linear_ofs(sw, dw)
    sy, fy = _linear_ofs(sh, dh)
    sx1 = np.minimum(sx + 1, sw - 1)
    sy1 = np.minimum(sy + 1, sh - 1)
    shp = (1, dw) + (1,) * (src.ndim - 2)
    fxr = fx.reshape(shp)
    rows = src[:, sx] * (np.float32(1) - fxr) + src[:, sx1] * fxr
    shp = (dh, 1) + (1,) * (src.ndim - 2)
    fyr = fy.reshape(shp)
    out = rows[sy] * (np.float32(1) - fyr) + rows[sy1] * fyr
    return out.astype(np.float32)


def _cubic_coeffs(x: np.float32) -> np.ndarray:
    A = np.float32(-0.75)
    one = np.float32(1)
    c = np.empty(4, dtype=np.float32)
    c[0] = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
    c[1] = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
    c[2] = ((A + np.float32(2)) * (one - x) - (A + np.float32(3))) * (one - x) * (one - x) + one
    c[3] = one - c[0] - c[1] - c[2]
    return c


def _cubic_ofs(ssize: int, dsize: int):
    scale = 1.0 / (dsize / ssize)
    idx = np.empty((dsize, 4), dtype=np.int64)
    cf = np.empty((dsize, 4), dtype=np.float32)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = math.floor(f)
        f = np.float32(f - s)
        cf[d] = _cubic_coeffs(f)
        idx[d] = np.clip(np.arange(s - 1, s + 3), 0, ssize - 1)
    return idx, cf


def resize_cubic(src: np.ndarray, dsize_wh) -> np.ndarray:
    """cv2.resize(f32, INTER_CUBIC): Keys cubic A=-0.75, taps sx-1..sx+2 replicated at
    the border, separable, float32."""
    src = np.asarray(src, dtype=np.float32)
    sh, sw = src.shape[:2]
    dw, dh = dsize_wh
    ix, cx = _cubic_ofs(sw, dw)
    iy, cy = _cubic_ofs(sh, dh)
    tail = (1,) * (src.ndim - 2)
    rows = np.zeros((sh, dw) + src.shape[2:], dtype=np.float32)
    for k in range(4):
        rows = rows + src[:, ix[:, k]] * cx[:, k].reshape((1, dw) + tail)
    out = np.zeros((dh, dw) + src.shape[2:], dtype=np.float32)
    for k in range(4):
        out = out + rows[iy[:, k]] * cy[:, k].reshape((dh, 1) + tail)
    return out.astype(np.float32)


def resize(src: np.ndarray, dsize_wh, interp: int) -> np.ndarray:
    if tuple(dsize_wh) == (src.shape[1], src.shape[0]):
        return np.array(src, copy=True)  # cv2 copies when the size is unchanged
    if interp == INTER_NEAREST:
        return resize_nearest(src, dsize_wh)
    if interp == INTER_LINEAR:
        return resize_linear(src, dsize_wh)
    if interp == INTER_CUBIC:
        return resize_cubic(src, dsize_wh)
    if interp == INTER_AREA:
        return resize_area(src, dsize_wh)
    raise ValueError(interp)


# --------------------------------------------------------------------------- #
# 3x3 filters                                                                  #
# --------------------------------------------------------------------------- #


def _reflect101_pad(img: np.ndarray) -> np.ndarray:
    pad = [(1, 1), (1, 1)] + [(0, 0)] * (img.ndim - 2)
    return np.pad(img, pad, mode="reflect")


def gaussian_blur3(img: np.ndarray) -> np.ndarray:
    """cv2.GaussianBlur(img, (3,3), 0): separable [1/4, 1/2, 1/4], BORDER_REFLECT_101."""
    p = _reflect101_pad(np.asarray(img, dtype=np.float32))
    q, h = np.float32(0.25), np.float32(0.5)
    rows = p[:, :-2] * q + p[:, 1:-1] * h + p[:, 2:] * q
    return (rows[:-2] * q + rows[1:-1] * h + rows[2:] * q).astype(np.float32)


def sharpen3(img: np.ndarray) -> np.ndarray:
    """cv2.filter2D(img, -1, [[0,-1,0],[-1,5,-1],[0,-1,0]]) (no clip inside cv2)."""
    p = _reflect101_pad(np.asarray(img, dtype=np.float32))
    c = p[1:-1, 1:-1]
    out = np.float32(5) * c - p[:-2, 1:-1] - p[1:-1, :-2] - p[1:-1, 2:] - p[2:, 1:-1]
    return out.astype(np.float32)


# --------------------------------------------------------------------------- #
# rounded-rectangle mask                                                       #
# --------------------------------------------------------------------------- #


def filled_quarter_circle(radius: int) -> np.ndarray:
    """cv2.circle(zeros(r,r), (0,0), r, 1, FILLED): Bresenham-style midpoint circle
    (cv::Circle in drawing.cpp) rasterised as horizontal spans, clipped to the r x r
    corner tile.  Returned array is float32 {0,1}."""
    r = radius
    img = np.zeros((r, r), dtype=np.float32)
    err, dx, dy, plus, minus = 0, r, 0, 1, (r << 1) - 1
    while dx >= dy:
        # centre (0,0): rows y = +-dy span x in [-dx, dx]; rows y = +-dx span [-dy, dy]
        for yy, half in ((dy, dx), (dx, dy)):
            if 0 <= yy < r:
                img[yy, 0 : min(half, r - 1) + 1] = 1
        dy += 1
        err += plus
        plus += 2
        mask = (err <= 0) - 1  # 0 if err <= 0 else -1
        err -= minus & mask
        dx += mask
        minus -= mask & 2
    return img


def round_rect_mask(size_hw, radius: int) -> np.ndarray:
    """mtgvision/util/image.py:406-425 with the cv2.circle call restated."""
    h, w = size_hw
    img = np.ones((h, w), dtype=np.float32)
    corner = filled_quarter_circle(radius)
    img[h - radius :, w - radius :] = np.rot90(corner, 0)
    img[:radius, w - radius :] = np.rot90(corner, 1)
    img[:radius, :radius] = np.rot90(corner, 2)
    img[h - radius :, :radius] = np.rot90(corner, 3)
    return img
