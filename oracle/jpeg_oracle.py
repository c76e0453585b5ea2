"""TEST INFRASTRUCTURE ONLY - restatement of the baseline JPEG decode behind `cv2.imread`.

SURVEY 8f.1: the step in front of the generator path is `imread_float`
(/root/reference/mtgvision/util/image.py:107-114: `cv2.imread(path, IMREAD_COLOR_RGB)` then
`img_float32`), reached from `IlsvrcImages._load_image/ran` (encoder_datasets.py:458-474) for
every background.  The arithmetic lives in a third-party dependency that is not under
/root/reference: **libjpeg-turbo 3.1.2 as bundled in opencv-python 4.13.0** (cv2 build info:
"JPEG: build-libjpeg-turbo (ver 3.1.2-70)"), with OpenCV's defaults: `dct_method = JDCT_ISLOW`,
`do_fancy_upsampling = TRUE`, output colour space RGB.  This module restates its published
algorithm for baseline (SOF0/SOF1, 8-bit, Huffman, one interleaved scan) files:

  * entropy decode            jdhuff.c      (decode_mcu_slow semantics, ITU T.81 F.2.2)
  * dequantise + inverse DCT  jidctint.c    (jpeg_idct_islow: CONST_BITS 13, PASS1_BITS 2)
  * chroma upsampling         jdsample.c    (h2v1/h2v2/h1v2 "fancy" triangle filters, plain
                                             replication when the chroma plane is <= 2 wide;
                                             context rows replicated at the image edge: jdmainct.c)
  * YCbCr -> RGB              jdcolor.c     (16-bit fixed-point tables, SCALEBITS 16)

Pinned: `tests/test_jpeg_oracle.py` compares `decode()` bit for bit with `cv2.imdecode` over
sampling modes, odd sizes, qualities, restart intervals and optimised Huffman tables.

Nothing under `mtgvision_b200/` may import this module.
"""

from __future__ import annotations

import numpy as np

ZIGZAG = np.array(
    [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
     28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54,
     47, 55, 62, 63], dtype=np.int64)


class JpegUnsupported(ValueError):
    pass


# --------------------------------------------------------------------------- #
# container                                                                    #
# --------------------------------------------------------------------------- #


def parse(data: bytes) -> dict:
    """Marker walk (ITU T.81 B.2).  Returns the frame header, tables and the entropy segment."""
    if data[:2] != b"\xff\xd8":
        raise JpegUnsupported("not a JPEG (no SOI)")
    pos = 2
    qt = {}
    ht = {}
    frame = None
    dri = 0
    adobe_transform = None
    while True:
        if pos + 4 > len(data):
            raise JpegUnsupported("truncated before SOS")
        if data[pos] != 0xFF:
            raise JpegUnsupported("marker expected")
        while data[pos + 1] == 0xFF:
            pos += 1
        m = data[pos + 1]
        pos += 2
        if m == 0xD8 or (0xD0 <= m <= 0xD7) or m == 0x01:
            continue
        seglen = (data[pos] << 8) | data[pos + 1]
        seg = data[pos + 2 : pos + seglen]
        if m == 0xDB:
            q = 0
            while q < len(seg):
                pq, tq = seg[q] >> 4, seg[q] & 15
                q += 1
                if pq:
                    tab = [(seg[q + 2 * i] << 8) | seg[q + 2 * i + 1] for i in range(64)]
                    q += 128
                else:
                    tab = list(seg[q : q + 64])
                    q += 64
                nat = np.zeros(64, np.int64)
                nat[ZIGZAG] = tab  # tables are stored in zig-zag order
                qt[tq] = nat
        elif m == 0xC4:
            q = 0
            while q < len(seg):
                tc, th = seg[q] >> 4, seg[q] & 15
                counts = list(seg[q + 1 : q + 17])
                n = sum(counts)
                ht[(tc, th)] = (counts, list(seg[q + 17 : q + 17 + n]))
                q += 17 + n
        elif m in (0xC0, 0xC1):
            if seg[0] != 8:
                raise JpegUnsupported("only 8-bit samples")
            h, w, nc = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4], seg[5]
            comps = [dict(id=seg[6 + 3 * i], h=seg[7 + 3 * i] >> 4, v=seg[7 + 3 * i] & 15, tq=seg[8 + 3 * i]) for i in range(nc)]
            frame = dict(h=h, w=w, comps=comps)
        elif 0xC2 <= m <= 0xCF and m not in (0xC4, 0xC8, 0xCC):
            raise JpegUnsupported("only baseline/extended sequential Huffman JPEG (SOF0/SOF1), got SOF%d" % (m - 0xC0))
        elif m == 0xDD:
            dri = (seg[0] << 8) | seg[1]
        elif m == 0xEE and seg[:5] == b"Adobe":
            adobe_transform = seg[11]
        elif m == 0xDA:
            if frame is None:
                raise JpegUnsupported("SOS before SOF")
            ns = seg[0]
            if ns != len(frame["comps"]):
                raise JpegUnsupported("only one interleaved scan")
            for i in range(ns):
                cs, t = seg[1 + 2 * i], seg[2 + 2 * i]
                c = next(c for c in frame["comps"] if c["id"] == cs)
                c["td"], c["ta"] = t >> 4, t & 15
            pos += seglen
            break
        pos += seglen
    nc = len(frame["comps"])
    if nc not in (1, 3):
        raise JpegUnsupported("only grayscale or YCbCr")
    if nc == 3 and adobe_transform == 0:
        raise JpegUnsupported("Adobe RGB JPEG")
    return dict(frame=frame, qt=qt, ht=ht, dri=dri, scan=pos)


# --------------------------------------------------------------------------- #
# entropy decode                                                               #
# --------------------------------------------------------------------------- #


class _Bits:
    def __init__(self, data: bytes, pos: int):
        self.d, self.p, self.buf, self.n = data, pos, 0, 0
        self.hit_marker = False
        self.fake = 0  # zero bits at the tail of the buffer that are not file data (fed after a marker / the end of the file)

    def _fill(self):
        while self.n <= 24:
            b = 0
            self.fake += 8
            if not self.hit_marker and self.p < len(self.d):
                self.fake -= 8
                b = self.d[self.p]
                if b == 0xFF:
                    nx = self.d[self.p + 1] if self.p + 1 < len(self.d) else 0xD9
                    if nx == 0:
                        self.p += 2
                    else:
                        self.hit_marker = True  # jdhuff.c feeds zero bits once a marker is reached
                        b = 0
                        self.fake += 8
                else:
                    self.p += 1
            self.buf = ((self.buf << 8) | b) & 0xFFFFFFFFFFFF
            self.n += 8

    def get(self, k: int) -> int:
        if k == 0:
            return 0
        self._fill()
        self.n -= k
        return (self.buf >> self.n) & ((1 << k) - 1)

    def restart(self):
        """Discard the partial byte and step over the RSTn marker."""
        self.buf = self.n = self.fake = 0
        self.hit_marker = False
        while self.p + 1 < len(self.d) and not (self.d[self.p] == 0xFF and 0xD0 <= self.d[self.p + 1] <= 0xD7):
            self.p += 1
        self.p += 2


def _huff_tables(counts, vals):
    """Canonical code assignment (T.81 C.2): code -> (length, symbol)."""
    lut = {}
    code = 0
    k = 0
    for ln in range(1, 17):
        for _ in range(counts[ln - 1]):
            lut[(ln, code)] = vals[k]
            code += 1
            k += 1
        code <<= 1
    return lut


def _decode_symbol(bits: _Bits, lut) -> int:
    code = 0
    for ln in range(1, 17):
        code = (code << 1) | bits.get(1)
        s = lut.get((ln, code))
        if s is not None:
            return s
    return 0  # corrupt code: libjpeg warns and returns 0


def _extend(v: int, s: int) -> int:
    return v - ((1 << s) - 1) if s and v < (1 << (s - 1)) else v


def entropy_decode(data: bytes, info: dict):
    """Per component an int16 array [block_rows, block_cols, 64] (natural order, not dequantised),
    block counts padded to whole MCUs."""
    fr = info["frame"]
    comps = fr["comps"]
    hmax = max(c["h"] for c in comps)
    vmax = max(c["v"] for c in comps)
    if len(comps) == 1:
        hmax = vmax = comps[0]["h"] = comps[0]["v"] = 1  # a single-component scan is never interleaved (T.81 A.2.2)
    mcux = -(-fr["w"] // (8 * hmax))
    mcuy = -(-fr["h"] // (8 * vmax))
    out = [np.zeros((mcuy * c["v"], mcux * c["h"], 64), np.int16) for c in comps]
    luts = {k: _huff_tables(*v) for k, v in info["ht"].items()}
    bits = _Bits(data, info["scan"])
    pred = [0] * len(comps)
    dri = info["dri"]
    n = 0
    for my in range(mcuy):
        for mx in range(mcux):
            if dri and n and n % dri == 0:
                bits.restart()
                pred = [0] * len(comps)
            n += 1
            # jdhuff.c decode_mcu: once a request for bits ran past the data (premature end of the file) the MCU in progress
            # is finished from zero bits and every later MCU is skipped - its coefficients stay zero (grey) - until a restart
            if bits.n < bits.fake:
                continue
            for ci, c in enumerate(comps):
                dc_lut, ac_lut = luts[(0, c["td"])], luts[(1, c["ta"])]
                for by in range(c["v"]):
                    for bx in range(c["h"]):
                        blk = out[ci][my * c["v"] + by, mx * c["h"] + bx]
                        s = _decode_symbol(bits, dc_lut)
                        pred[ci] += _extend(bits.get(s), s)
                        blk[0] = np.int16(pred[ci])
                        k = 1
                        while k < 64:
                            rs = _decode_symbol(bits, ac_lut)
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            if k > 63:
                                break
                            blk[ZIGZAG[k]] = np.int16(_extend(bits.get(s), s))
                            k += 1
    return out, (hmax, vmax)


# --------------------------------------------------------------------------- #
# jidctint.c: jpeg_idct_islow                                                  #
# --------------------------------------------------------------------------- #

_C = dict(f0298=2446, f0390=3196, f0541=4433, f0765=6270, f0899=7373, f1175=9633, f1501=12299, f1847=15137, f1961=16069,
          f2053=16819, f2562=20995, f3072=25172)


def _idct_1d(x, shift):
    """One pass over the LAST axis' 8 entries held as x[..., 0..7]; int64 arithmetic, DESCALE by `shift`."""
    C = _C
    z2, z3 = x[..., 2], x[..., 6]
    z1 = (z2 + z3) * C["f0541"]
    tmp2 = z1 - z3 * C["f1847"]
    tmp3 = z1 + z2 * C["f0765"]
    z2, z3 = x[..., 0], x[..., 4]
    tmp0 = (z2 + z3) << 13
    tmp1 = (z2 - z3) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = x[..., 7], x[..., 5], x[..., 3], x[..., 1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * C["f1175"]
    tmp0 = tmp0 * C["f0298"]
    tmp1 = tmp1 * C["f2053"]
    tmp2 = tmp2 * C["f3072"]
    tmp3 = tmp3 * C["f1501"]
    z1 = -z1 * C["f0899"]
    z2 = -z2 * C["f2562"]
    z3 = -z3 * C["f1961"] + z5
    z4 = -z4 * C["f0390"] + z5
    tmp0 = tmp0 + z1 + z3
    tmp1 = tmp1 + z2 + z4
    tmp2 = tmp2 + z2 + z3
    tmp3 = tmp3 + z1 + z4
    r = 1 << (shift - 1)
    o = np.stack([tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3], axis=-1)
    return (o + r) >> shift


def idct_islow(coefs: np.ndarray, quant: np.ndarray) -> np.ndarray:
    """[..., 64] int16 coefficients (natural order) -> [..., 8, 8] uint8 samples."""
    x = (coefs.astype(np.int64) * quant.astype(np.int64)).reshape(coefs.shape[:-1] + (8, 8))
    ws = _idct_1d(np.swapaxes(x, -1, -2), 13 - 2)  # pass 1: columns; ws[..., col, row]
    ws = np.swapaxes(ws, -1, -2)
    o = _idct_1d(ws, 13 + 2 + 3)  # pass 2: rows
    return np.clip(o + 128, 0, 255).astype(np.uint8)  # range_limit table, centred


def blocks_to_plane(samples: np.ndarray) -> np.ndarray:
    br, bc = samples.shape[:2]
    return samples.transpose(0, 2, 1, 3).reshape(br * 8, bc * 8)


# --------------------------------------------------------------------------- #
# jdsample.c                                                                   #
# --------------------------------------------------------------------------- #


def _other_rows(c: np.ndarray, out_h: int, vfac: int):
    """For every output row: (this input row, neighbouring input row), edges replicated (jdmainct.c)."""
    y = np.arange(out_h)
    if vfac == 1:
        return c[y], None
    dh = c.shape[0]
    inrow = y >> 1
    other = np.where(y & 1, inrow + 1, inrow - 1).clip(0, dh - 1)
    return c[inrow], c[other]


def upsample(c: np.ndarray, out_h: int, out_w: int, hfac: int, vfac: int) -> np.ndarray:
    """c: the component's sample plane cut to (downsampled_height, downsampled_width)."""
    c = c.astype(np.int64)
    dw = c.shape[1]
    if hfac == 1 and vfac == 1:
        return c[:out_h, :out_w]
    if hfac == 1 and vfac == 2:  # h1v2_fancy_upsample
        a, b = _other_rows(c, out_h, 2)
        bias = np.where(np.arange(out_h) & 1, 2, 1)[:, None]
        return ((3 * a + b + bias) >> 2)[:, :out_w]
    if hfac == 2 and dw <= 2:  # too narrow for the triangle filter: h2v1_upsample / h2v2_upsample replicate
        y = np.arange(out_h) >> (vfac - 1)
        return c[y][:, np.arange(out_w) >> 1]
    x = np.arange(out_w)
    j = x >> 1
    odd = (x & 1).astype(bool)
    if hfac == 2 and vfac == 1:  # h2v1_fancy_upsample
        rows = c[:out_h]
        nb = np.where(odd, j + 1, j - 1)
        edge = (nb < 0) | (nb >= dw)
        nbc = nb.clip(0, dw - 1)
        v = (3 * rows[:, j] + rows[:, nbc] + np.where(odd, 2, 1)) >> 2
        return np.where(edge[None, :], rows[:, j], v)
    if hfac == 2 and vfac == 2:  # h2v2_fancy_upsample
        a, b = _other_rows(c, out_h, 2)
        cs = 3 * a + b
        nb = np.where(odd, j + 1, j - 1)
        edge = (nb < 0) | (nb >= dw)
        nbc = nb.clip(0, dw - 1)
        bias = np.where(odd, 7, 8)
        v = (3 * cs[:, j] + cs[:, nbc] + bias) >> 4
        e = (4 * cs[:, j] + bias) >> 4
        return np.where(edge[None, :], e, v)
    raise JpegUnsupported("sampling factors %dx%d" % (hfac, vfac))


# --------------------------------------------------------------------------- #
# jdcolor.c                                                                    #
# --------------------------------------------------------------------------- #


def ycc_to_rgb(y, cb, cr):
    def fix(v):
        return int(v * 65536 + 0.5)

    half = 1 << 15
    cbx, crx = cb - 128, cr - 128
    r = y + ((fix(1.40200) * crx + half) >> 16)
    b = y + ((fix(1.77200) * cbx + half) >> 16)
    g = y + ((-fix(0.34414) * cbx + half - fix(0.71414) * crx) >> 16)
    return np.stack([r, g, b], axis=-1).clip(0, 255).astype(np.uint8)


# --------------------------------------------------------------------------- #


def decode(data: bytes) -> np.ndarray:
    """== cv2.imdecode(data, cv2.IMREAD_COLOR_RGB) for supported files."""
    info = parse(data)
    fr = info["frame"]
    H, W = fr["h"], fr["w"]
    coefs, (hmax, vmax) = entropy_decode(data, info)
    planes = []
    for c, cf in zip(fr["comps"], coefs):
        p = blocks_to_plane(idct_islow(cf, info["qt"][c["tq"]]))
        dh, dw = -(-H * c["v"] // vmax), -(-W * c["h"] // hmax)
        planes.append(upsample(p[:dh, :dw], H, W, hmax // c["h"], vmax // c["v"]))
    if len(planes) == 1:
        g = planes[0].astype(np.uint8)
        return np.stack([g, g, g], axis=-1)
    return ycc_to_rgb(*planes)
