"""TEST INFRASTRUCTURE ONLY - restatement of the baseline JPEG encode behind `cv2.imwrite`.

SURVEY 8f.2: `save_sample` (/root/reference/mtgvision/od_datasets.py:794-832) writes every generated
scene with `imwrite` (/root/reference/mtgvision/util/image.py:95-104): `(img * 255).astype(uint8)`,
`cv2.cvtColor(RGB2BGR)`, `cv2.imwrite(path, img)` with OpenCV's defaults.  The arithmetic lives in a
third-party dependency that is not under /root/reference: **libjpeg-turbo 3.1.2 as bundled in
opencv-python 4.13.0**, driven by OpenCV's defaults: quality 95, 4:2:0 chroma subsampling, baseline
sequential Huffman with the standard (Annex K) tables, no restart markers, JFIF 1.01 header with
density 1:1.  This module restates its published algorithm:

  * RGB -> YCbCr               jccolor.c    (rgb_ycc_convert: 16-bit fixed-point tables)
  * chroma 2x2 downsampling    jcsample.c   (h2v2_downsample: box filter, alternating bias 1,2)
  * forward DCT                jfdctint.c   (jpeg_fdct_islow: CONST_BITS 13, PASS1_BITS 2, output x8)
  * quantisation               jcdctmgr.c   (round-half-away division by 8*Q; tables scaled by
                                             jpeg_quality_scaling / jpeg_add_quant_table, jcparam.c)
  * Huffman coding             jchuff.c     (encode_one_block, standard tables of jstdhuff.c, 0xFF stuffing,
                                             final byte padded with one bits)
  * markers                    jcmarker.c   (SOI, APP0, DQT x2, SOF0, DHT x4, SOS, EOI)

Image sizes that are not whole 16x16 MCUs (`save_sample` itself asserts 640x640) follow the edge-replication and
dummy-block rules of jcsample.c / jcprepct.c / jccoefct.c, restated in `encode`.

Pinned: `tests/test_jpeg_encode.py` compares `encode()` byte for byte with `cv2.imencode`.

Nothing under `mtgvision_b200/` may import this module.
"""

from __future__ import annotations

import numpy as np

from .jpeg_oracle import ZIGZAG

# Annex K.1 / K.2 (jcparam.c std_luminance_quant_tbl, std_chrominance_quant_tbl), natural order
STD_LUMA_Q = np.array(
    [16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
     18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100,
     103, 99], dtype=np.int64)
STD_CHROMA_Q = np.array(
    [17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99],
    dtype=np.int64)

# Annex K.3 (jstdhuff.c): code-length counts for lengths 1..16 and the symbols in code order
DC_LUMA_BITS = [0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0]
DC_LUMA_VALS = list(range(12))
DC_CHROMA_BITS = [0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0]
DC_CHROMA_VALS = list(range(12))
AC_LUMA_BITS = [0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7D]
AC_LUMA_VALS = [
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91,
    0xA1, 0x08, 0x23, 0x42, 0xB1, 0xC1, 0x15, 0x52, 0xD1, 0xF0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0A, 0x16, 0x17, 0x18, 0x19, 0x1A,
    0x25, 0x26, 0x27, 0x28, 0x29, 0x2A, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4A, 0x53,
    0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79,
    0x7A, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A, 0xA2, 0xA3, 0xA4, 0xA5,
    0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA, 0xC2, 0xC3, 0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9,
    0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA, 0xE1, 0xE2, 0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF1, 0xF2,
    0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9, 0xFA]
AC_CHROMA_BITS = [0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77]
AC_CHROMA_VALS = [
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14,
    0x42, 0x91, 0xA1, 0xB1, 0xC1, 0x09, 0x23, 0x33, 0x52, 0xF0, 0x15, 0x62, 0x72, 0xD1, 0x0A, 0x16, 0x24, 0x34, 0xE1, 0x25, 0xF1, 0x17,
    0x18, 0x19, 0x1A, 0x26, 0x27, 0x28, 0x29, 0x2A, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4A,
    0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78,
    0x79, 0x7A, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A, 0xA2, 0xA3,
    0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA, 0xC2, 0xC3, 0xC4, 0xC5, 0xC6, 0xC7,
    0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA, 0xE2, 0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF2,
    0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9, 0xFA]


def quant_table(std: np.ndarray, quality: int) -> np.ndarray:
    """jpeg_quality_scaling + jpeg_add_quant_table(force_baseline=TRUE), natural order."""
    quality = min(max(int(quality), 1), 100)
    scale = 5000 // quality if quality < 50 else 200 - quality * 2
    return np.clip((std * scale + 50) // 100, 1, 255)


def huff_codes(bits, vals):
    """jchuff.c jpeg_make_c_derived_tbl: symbol -> (code, length)."""
    table = {}
    code = 0
    k = 0
    for ln in range(1, 17):
        for _ in range(bits[ln - 1]):
            table[vals[k]] = (code, ln)
            code += 1
            k += 1
        code <<= 1
    return table


def rgb_to_ycc(rgb: np.ndarray):
    """jccolor.c rgb_ycc_convert (SCALEBITS 16)."""
    def fix(v):
        return int(v * 65536 + 0.5)

    r, g, b = (rgb[..., k].astype(np.int64) for k in range(3))
    half, off = 1 << 15, 128 << 16
    y = (fix(0.29900) * r + fix(0.58700) * g + fix(0.11400) * b + half) >> 16
    cb = (-fix(0.16874) * r - fix(0.33126) * g + fix(0.50000) * b + off + half - 1) >> 16
    cr = (fix(0.50000) * r - fix(0.41869) * g - fix(0.08131) * b + off + half - 1) >> 16
    return y, cb, cr


def downsample_h2v2(c: np.ndarray) -> np.ndarray:
    """jcsample.c h2v2_downsample: bias 1, 2, 1, 2, ... along each output row."""
    s = c[0::2, 0::2] + c[0::2, 1::2] + c[1::2, 0::2] + c[1::2, 1::2]
    bias = np.where(np.arange(s.shape[1]) & 1, 2, 1)[None, :]
    return (s + bias) >> 2


_F = dict(f0298=2446, f0390=3196, f0541=4433, f0765=6270, f0899=7373, f1175=9633, f1501=12299, f1847=15137, f1961=16069,
          f2053=16819, f2562=20995, f3072=25172)


def _fdct_1d(d, first: bool):
    """One pass of jpeg_fdct_islow over the last axis (8 entries)."""
    F = _F

    def descale(x, n):
        return (x + (1 << (n - 1))) >> n

    tmp0, tmp7 = d[..., 0] + d[..., 7], d[..., 0] - d[..., 7]
    tmp1, tmp6 = d[..., 1] + d[..., 6], d[..., 1] - d[..., 6]
    tmp2, tmp5 = d[..., 2] + d[..., 5], d[..., 2] - d[..., 5]
    tmp3, tmp4 = d[..., 3] + d[..., 4], d[..., 3] - d[..., 4]
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    sh = 13 - 2 if first else 13 + 2
    if first:
        o0, o4 = (tmp10 + tmp11) << 2, (tmp10 - tmp11) << 2
    else:
        o0, o4 = descale(tmp10 + tmp11, 2), descale(tmp10 - tmp11, 2)
    z1 = (tmp12 + tmp13) * F["f0541"]
    o2 = descale(z1 + tmp13 * F["f0765"], sh)
    o6 = descale(z1 - tmp12 * F["f1847"], sh)
    z1, z2, z3, z4 = tmp4 + tmp7, tmp5 + tmp6, tmp4 + tmp6, tmp5 + tmp7
    z5 = (z3 + z4) * F["f1175"]
    tmp4 = tmp4 * F["f0298"]
    tmp5 = tmp5 * F["f2053"]
    tmp6 = tmp6 * F["f3072"]
    tmp7 = tmp7 * F["f1501"]
    z1 = -z1 * F["f0899"]
    z2 = -z2 * F["f2562"]
    z3 = -z3 * F["f1961"] + z5
    z4 = -z4 * F["f0390"] + z5
    o7 = descale(tmp4 + z1 + z3, sh)
    o5 = descale(tmp5 + z2 + z4, sh)
    o3 = descale(tmp6 + z2 + z3, sh)
    o1 = descale(tmp7 + z1 + z4, sh)
    return np.stack([o0, o1, o2, o3, o4, o5, o6, o7], axis=-1)


def fdct_quant(plane: np.ndarray, q: np.ndarray) -> np.ndarray:
    """Sample plane (H, W multiples of 8) -> quantised coefficients [H/8, W/8, 64] in natural order."""
    h, w = plane.shape
    blocks = plane.reshape(h // 8, 8, w // 8, 8).transpose(0, 2, 1, 3).astype(np.int64) - 128
    ws = _fdct_1d(blocks, True)  # rows
    out = _fdct_1d(np.swapaxes(ws, -1, -2), False)  # columns; out[..., col, k]
    coef = np.swapaxes(out, -1, -2).reshape(h // 8, w // 8, 64)
    q8 = (q << 3)[None, None, :]
    mag = (np.abs(coef) + (q8 >> 1)) // q8  # jcdctmgr.c quantize(): round half away from zero
    return np.where(coef < 0, -mag, mag)


class _BitWriter:
    def __init__(self):
        self.out = bytearray()
        self.acc = 0
        self.n = 0

    def put(self, code: int, size: int):
        self.acc = (self.acc << size) | (code & ((1 << size) - 1))
        self.n += size
        while self.n >= 8:
            byte = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(byte)
            if byte == 0xFF:
                self.out.append(0)
            self.n -= 8
        self.acc &= (1 << self.n) - 1

    def flush(self):
        if self.n:
            self.put((1 << (8 - self.n)) - 1, 8 - self.n)  # fill the last byte with one bits


def _encode_block(bw: _BitWriter, blk: np.ndarray, last_dc: int, dc_tbl, ac_tbl) -> int:
    """jchuff.c encode_one_block; blk in natural order."""
    diff = int(blk[0]) - last_dc
    t = -diff if diff < 0 else diff
    t2 = diff - 1 if diff < 0 else diff
    nbits = t.bit_length()
    bw.put(*dc_tbl[nbits])
    if nbits:
        bw.put(t2, nbits)
    r = 0
    for k in range(1, 64):
        v = int(blk[ZIGZAG[k]])
        if v == 0:
            r += 1
            continue
        while r > 15:
            bw.put(*ac_tbl[0xF0])
            r -= 16
        t = -v if v < 0 else v
        t2 = v - 1 if v < 0 else v
        nbits = t.bit_length()
        bw.put(*ac_tbl[(r << 4) + nbits])
        bw.put(t2, nbits)
        r = 0
    if r > 0:
        bw.put(*ac_tbl[0x00])
    return int(blk[0])


def _marker(code: int, payload: bytes) -> bytes:
    return bytes([0xFF, code]) + (len(payload) + 2).to_bytes(2, "big") + payload


def header(h: int, w: int, ql: np.ndarray, qc: np.ndarray) -> bytes:
    """jcmarker.c: write_file_header + write_frame_header + write_scan_header for a 4:2:0 baseline YCbCr image."""
    out = b"\xff\xd8" + _marker(0xE0, b"JFIF\x00" + bytes([1, 1, 0, 0, 1, 0, 1, 0, 0]))
    for i, q in enumerate((ql, qc)):
        out += _marker(0xDB, bytes([i]) + bytes(int(q[ZIGZAG[k]]) for k in range(64)))
    out += _marker(0xC0, bytes([8]) + h.to_bytes(2, "big") + w.to_bytes(2, "big") + bytes([3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1]))
    for tc_th, bits, vals in ((0x00, DC_LUMA_BITS, DC_LUMA_VALS), (0x10, AC_LUMA_BITS, AC_LUMA_VALS),
                              (0x01, DC_CHROMA_BITS, DC_CHROMA_VALS), (0x11, AC_CHROMA_BITS, AC_CHROMA_VALS)):
        out += _marker(0xC4, bytes([tc_th]) + bytes(bits) + bytes(vals))
    out += _marker(0xDA, bytes([3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0]))
    return out


def _pad_edge(a: np.ndarray, h: int, w: int) -> np.ndarray:
    return np.pad(a, ((0, h - a.shape[0]), (0, w - a.shape[1])), mode="edge")


def encode(rgb: np.ndarray, quality: int = 95) -> bytes:
    """== cv2.imencode('.jpg', rgb[:, :, ::-1], [cv2.IMWRITE_JPEG_QUALITY, quality]).tobytes().

    Sizes that are not whole 16x16 MCUs follow libjpeg's edge rules: columns are replicated at full resolution before
    the chroma box filter (jcsample.c expand_right_edge), rows are replicated to an even count before it and the
    DOWNSAMPLED rows are replicated below that (jcprepct.c), and luma blocks that lie wholly outside the image are
    "dummy" blocks - zero AC terms, DC term copied from the block before them in the MCU (jccoefct.c compress_data)."""
    h, w, _ = rgb.shape
    mcuy, mcux = -(-h // 16), -(-w // 16)
    ql, qc = quant_table(STD_LUMA_Q, quality), quant_table(STD_CHROMA_Q, quality)
    y, cb, cr = rgb_to_ycc(rgb)
    cy = fdct_quant(_pad_edge(y, mcuy * 16, mcux * 16), ql)
    even_h = h + (h & 1)
    ccb = fdct_quant(_pad_edge(downsample_h2v2(_pad_edge(cb, even_h, mcux * 16)), mcuy * 8, mcux * 8), qc)
    ccr = fdct_quant(_pad_edge(downsample_h2v2(_pad_edge(cr, even_h, mcux * 16)), mcuy * 8, mcux * 8), qc)
    hib, wib = -(-h // 8), -(-w // 8)  # luma blocks that hold image samples
    dcl, acl = huff_codes(DC_LUMA_BITS, DC_LUMA_VALS), huff_codes(AC_LUMA_BITS, AC_LUMA_VALS)
    dcc, acc = huff_codes(DC_CHROMA_BITS, DC_CHROMA_VALS), huff_codes(AC_CHROMA_BITS, AC_CHROMA_VALS)
    bw = _BitWriter()
    last = [0, 0, 0]
    for my in range(mcuy):
        for mx in range(mcux):
            blocks = []  # the MCU's luma blocks in scan order, dummy rules applied in that order
            for by in range(2):
                row_real = 2 * my + by < hib
                for bx in range(2):
                    if row_real and 2 * mx + bx < wib:
                        blk = cy[2 * my + by, 2 * mx + bx]
                    else:
                        blk = np.zeros(64, dtype=cy.dtype)
                        blk[0] = blocks[-1][0] if row_real else blocks[2 * by - 1][0]
                    blocks.append(blk)
            for blk in blocks:
                last[0] = _encode_block(bw, blk, last[0], dcl, acl)
            last[1] = _encode_block(bw, ccb[my, mx], last[1], dcc, acc)
            last[2] = _encode_block(bw, ccr[my, mx], last[2], dcc, acc)
    bw.flush()
    return header(h, w, ql, qc) + bytes(bw.out) + b"\xff\xd9"
