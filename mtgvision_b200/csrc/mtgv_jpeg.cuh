// mtgv_jpeg.cuh - baseline JPEG decode arithmetic shared by the device kernels (mtgv_jpeg.cu) and the
// host-side unit-test harness (tests/host_harness.cpp).  SURVEY 8f.1: the step in front of the generator is
// imread_float (mtgvision/util/image.py:107-114, cv2.imread(path, IMREAD_COLOR_RGB)), called for every
// background by IlsvrcImages._load_image (encoder_datasets.py:457-474).  cv2 delegates to libjpeg-turbo
// (3.1.2 in opencv-python 4.13.0) with its defaults: ISLOW inverse DCT, "fancy" chroma upsampling, RGB out.
// All of it is integer arithmetic, restated here so that the decoded pool bytes equal cv2's bit for bit:
//   entropy decode   jdhuff.c   (T.81 F.2.2; lookahead table for codes <= 9 bits, canonical search above)
//   dequant + IDCT   jidctint.c (jpeg_idct_islow, CONST_BITS 13, PASS1_BITS 2)
//   upsampling       jdsample.c (h2v1 / h2v2 / h1v2 triangle filters; edge rows replicated as in jdmainct.c)
//   YCbCr -> RGB     jdcolor.c  (SCALEBITS 16 fixed-point tables)
// Supported files: SOF0/SOF1, 8-bit, Huffman, one interleaved scan, 1 (gray) or 3 (YCbCr) components with
// chroma subsampled by 1 or 2 in each direction; restart intervals are decoded in parallel.  Anything else
// (progressive, arithmetic, CMYK, 4:1:1 ...) is rejected by the parser with a message; there is no fallback.
#pragma once

#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "mtgv_geom.cuh"

namespace mtgv {

constexpr int kJpegFastBits = 9;

struct JpegTables {  // per image
  uint16_t qt[4][64];                    // quantisation tables, natural (row-major) order
  uint16_t fast[4][1 << kJpegFastBits];  // [class*2 + id][next 9 bits] = length << 8 | symbol, 0 = longer code
  int32_t maxcode[4][18];                // largest code of each length 1..16 (-1 none), [17] = sentinel
  int32_t valoff[4][18];                 // index of the first symbol of a length minus its first code
  uint8_t vals[4][256];
};

struct JpegImg {
  int64_t file_off;      // of the file in the byte buffer
  int64_t coef_blk;      // first 8x8 block of component 0 in the coefficient scratch (components follow each other)
  int64_t plane_off[3];  // component sample planes in the plane scratch (bytes)
  int64_t out_off;       // output image (bytes)
  int32_t file_len, scan_off;
  int32_t h, w, ncomp, hmax, vmax, mcux, mcuy, dri;
  int32_t ch[3], cv[3], tq[3], td[3], ta[3];
  int32_t bw[3], bh[3];  // component extent in blocks (whole MCUs)
  int32_t blk0[3];       // first block of component c relative to coef_blk
  int32_t hf[3], vf[3];  // upsampling factors hmax / ch, vmax / cv
  int32_t dh[3], dw[3];  // downsampled component size ceil(h * cv / vmax), ceil(w * ch / hmax)
  int32_t seg0, nseg;
  int32_t nblk, par;     // decoded by k_jpeg_entropy_par: 1 = one interval (self-synchronising runs, DC terms integrated by
                         // k_jpeg_dc), 2 = restart intervals (a thread per interval, DC terms written directly)
  int64_t clean_off;     // byte-unstuffed copy of the scan in the clean scratch (bytes, multiple of 16)
  int64_t sync_off;      // checkpoint states of the subsequences (uint64 units)
  int64_t rst_off;       // par == 2: clean byte position of every restart interval but the first (uint32 units)
  int32_t out_layout;    // 0: [h][w][3] RGB bytes; 1: planar [3][h][out_pitch] bytes (card pool); 2: RGBX words [h][out_pitch] (background pool)
  int32_t out_pitch;     // layout 1: bytes per plane row; layout 2: words per row
  int32_t progressive;   // SOF2: the entropy stage runs on the host (mtgv_jpeg_prog.h), par = 0, coefficients are uploaded
  int32_t comp_id[3];    // component identifiers of the frame header (scan headers of progressive files refer to them)
};

struct JpegSeg {  // one restart interval (or the whole scan): decoded by one thread
  int32_t img, mcu0, nmcu, byte_off;
};

static const uint8_t kJpegZigzag[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// ------------------------------------------------------------------------------------ //
// container: marker walk on the host (T.81 B.2)                                          //
// ------------------------------------------------------------------------------------ //

inline int jpeg_fail(std::string* err, const char* msg) {
  if (err) *err = std::string("jpeg: ") + msg;
  return -1;
}

// Fills im (geometry, table selectors; offsets other than scan_off are the caller's), tb and appends the
// image's segments (segs may be null: the interval count is then derived from DRI alone).  Returns 0 or -1 with *err set.
inline int jpeg_parse(const uint8_t* d, int64_t len, int img_index, JpegImg* im, JpegTables* tb, std::vector<JpegSeg>* segs,
                      std::string* err) {
  memset(im, 0, sizeof(*im));
  memset(tb, 0, sizeof(*tb));
  if (len < 4 || d[0] != 0xFF || d[1] != 0xD8) return jpeg_fail(err, "not a JPEG file (no SOI marker)");
  if (len > 0x7fffffff) return jpeg_fail(err, "file larger than 2 GiB");
  int64_t pos = 2;
  bool have_frame = false, have_q[4] = {false, false, false, false}, have_h[4] = {false, false, false, false};
  bool progressive = false, ext_tables = false;
  int comp_id[3] = {0, 0, 0}, adobe_transform = -1;
  for (;;) {
    if (pos + 4 > len) return jpeg_fail(err, "truncated before the scan");
    if (d[pos] != 0xFF) return jpeg_fail(err, "marker expected");
    while (pos + 1 < len && d[pos + 1] == 0xFF) pos++;
    if (pos + 1 >= len) return jpeg_fail(err, "truncated before the scan");  // header ends in 0xFF fill bytes
    const int m = d[pos + 1];
    pos += 2;
    if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
    if (pos + 2 > len) return jpeg_fail(err, "truncated segment");
    const int seglen = (d[pos] << 8) | d[pos + 1];
    if (seglen < 2 || pos + seglen > len) return jpeg_fail(err, "truncated segment");
    const uint8_t* s = d + pos + 2;
    const int n = seglen - 2;
    if (m == 0xDB) {
      int q = 0;
      while (q < n) {
        const int pq = s[q] >> 4, tq = s[q] & 15;
        q++;
        if (tq > 3 || q + (pq ? 128 : 64) > n) return jpeg_fail(err, "bad quantisation table");
        for (int i = 0; i < 64; i++) tb->qt[tq][kJpegZigzag[i]] = pq ? (uint16_t)((s[q + 2 * i] << 8) | s[q + 2 * i + 1]) : s[q + i];
        q += pq ? 128 : 64;
        have_q[tq] = true;
      }
    } else if (m == 0xC4) {
      int q = 0;
      while (q < n) {
        if (q + 17 > n) return jpeg_fail(err, "bad Huffman table");
        const int tc = s[q] >> 4, th = s[q] & 15;
        if (tc > 1 || th > 3) return jpeg_fail(err, "bad Huffman table id");
        if (th > 1) {  // ids 2, 3: legal in extended / progressive files only; the host scan decoder reads its own tables
          int total2 = 0;
          for (int l = 0; l < 16; l++) total2 += s[q + 1 + l];
          if (total2 > 256 || q + 17 + total2 > n) return jpeg_fail(err, "bad Huffman table");
          ext_tables = true;
          q += 17 + total2;
          continue;
        }
        const int ti = tc * 2 + th;
        const uint8_t* counts = s + q + 1;
        int total = 0;
        for (int l = 0; l < 16; l++) total += counts[l];
        if (total > 256 || q + 17 + total > n) return jpeg_fail(err, "bad Huffman table");
        memcpy(tb->vals[ti], s + q + 17, total);
        memset(tb->fast[ti], 0, sizeof(tb->fast[ti]));
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
          tb->valoff[ti][l] = k - code;
          for (int i = 0; i < counts[l - 1]; i++, k++, code++) {
            if (l <= kJpegFastBits) {
              const int lo = code << (kJpegFastBits - l), cnt = 1 << (kJpegFastBits - l);
              if (lo + cnt > (1 << kJpegFastBits)) return jpeg_fail(err, "bad Huffman table (code overflow)");
              for (int j = 0; j < cnt; j++) tb->fast[ti][lo + j] = (uint16_t)((l << 8) | tb->vals[ti][k]);
            }
          }
          tb->maxcode[ti][l] = counts[l - 1] ? code - 1 : -1;
          if (code > (1 << l)) return jpeg_fail(err, "bad Huffman table (code overflow)");
          code <<= 1;
        }
        tb->maxcode[ti][17] = 0x7fffffff;
        have_h[ti] = true;
        q += 17 + total;
      }
    } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {
      progressive = m == 0xC2;
      if (n < 6 || s[0] != 8) return jpeg_fail(err, "only 8-bit samples are supported");
      im->h = (s[1] << 8) | s[2];
      im->w = (s[3] << 8) | s[4];
      im->ncomp = s[5];
      if (im->ncomp != 1 && im->ncomp != 3) return jpeg_fail(err, "only grayscale or YCbCr files are supported");
      if (n < 6 + 3 * im->ncomp || im->h <= 0 || im->w <= 0) return jpeg_fail(err, "bad frame header");
      if (im->h > 16384 || im->w > 16384) return jpeg_fail(err, "image larger than 16384 pixels per side");
      for (int c = 0; c < im->ncomp; c++) {
        comp_id[c] = s[6 + 3 * c];
        im->ch[c] = s[7 + 3 * c] >> 4;
        im->cv[c] = s[7 + 3 * c] & 15;
        im->tq[c] = s[8 + 3 * c];
        if (im->tq[c] > 3) return jpeg_fail(err, "bad frame header");
      }
      have_frame = true;
    } else if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      return jpeg_fail(err, "only Huffman-coded baseline, extended sequential or progressive JPEG (SOF0, SOF1, SOF2) is supported");
    } else if (m == 0xDD) {
      if (n < 2) return jpeg_fail(err, "bad DRI");
      im->dri = (s[0] << 8) | s[1];
    } else if (m == 0xEE && n >= 12 && memcmp(s, "Adobe", 5) == 0) {
      adobe_transform = s[11];
    } else if (m == 0xDA) {
      if (!have_frame) return jpeg_fail(err, "scan before frame header");
      if (progressive) {  // several scans, each with its own header: mtgv_jpeg_prog.h walks them; only the quantisation tables matter here
        for (int c = 0; c < im->ncomp; c++)
          if (!have_q[im->tq[c]]) return jpeg_fail(err, "missing quantisation table");
        pos -= 2;  // scan_off = the first SOS marker
        break;
      }
      if (ext_tables) return jpeg_fail(err, "Huffman table id > 1 (not baseline)");
      if (n < 1 || s[0] != im->ncomp || n < 1 + 2 * im->ncomp + 3) return jpeg_fail(err, "only one interleaved scan is supported");
      for (int i = 0; i < im->ncomp; i++) {
        int c = -1;
        for (int j = 0; j < im->ncomp; j++)
          if (comp_id[j] == s[1 + 2 * i]) c = j;
        if (c < 0) return jpeg_fail(err, "bad scan header");
        im->td[c] = s[2 + 2 * i] >> 4;
        im->ta[c] = s[2 + 2 * i] & 15;
        if (im->td[c] > 1 || im->ta[c] > 1 || !have_h[im->td[c]] || !have_h[2 + im->ta[c]]) return jpeg_fail(err, "missing Huffman table");
        if (!have_q[im->tq[c]]) return jpeg_fail(err, "missing quantisation table");
      }
      pos += seglen;
      break;
    }
    pos += seglen;
  }
  if (im->ncomp == 3 && adobe_transform == 0) return jpeg_fail(err, "Adobe RGB / CMYK files are not supported");
  if (im->ncomp == 1) im->ch[0] = im->cv[0] = 1;  // a single-component scan is never interleaved (T.81 A.2.2)
  im->hmax = im->vmax = 1;
  for (int c = 0; c < im->ncomp; c++) {
    if (im->ch[c] > im->hmax) im->hmax = im->ch[c];
    if (im->cv[c] > im->vmax) im->vmax = im->cv[c];
  }
  if (im->ch[0] != im->hmax || im->cv[0] != im->vmax) return jpeg_fail(err, "luma must be the full-resolution component");
  for (int c = 0; c < im->ncomp; c++) {
    const int hf = im->hmax / im->ch[c], vf = im->vmax / im->cv[c];
    if (im->ch[c] < 1 || im->cv[c] < 1 || hf * im->ch[c] != im->hmax || vf * im->cv[c] != im->vmax || hf > 2 || vf > 2 || im->hmax > 2 ||
        im->vmax > 2)
      return jpeg_fail(err, "unsupported sampling factors (chroma may be subsampled by 1 or 2 per direction)");
  }
  im->mcux = (im->w + 8 * im->hmax - 1) / (8 * im->hmax);
  im->mcuy = (im->h + 8 * im->vmax - 1) / (8 * im->vmax);
  im->nblk = 0;
  for (int c = 0; c < im->ncomp; c++) {
    im->bw[c] = im->mcux * im->ch[c];
    im->bh[c] = im->mcuy * im->cv[c];
    im->blk0[c] = im->nblk;
    im->nblk += im->bw[c] * im->bh[c];
    im->hf[c] = im->hmax / im->ch[c];
    im->vf[c] = im->vmax / im->cv[c];
    im->dh[c] = (im->h * im->cv[c] + im->vmax - 1) / im->vmax;
    im->dw[c] = (im->w * im->ch[c] + im->hmax - 1) / im->hmax;
  }
  im->scan_off = (int32_t)pos;
  im->file_len = (int32_t)len;
  im->progressive = progressive ? 1 : 0;
  for (int c = 0; c < 3; c++) im->comp_id[c] = comp_id[c];
  if (progressive) {  // no device entropy stage: one dummy segment keeps the scratch arithmetic of the caller uniform
    im->seg0 = segs ? (int32_t)segs->size() : 0;
    im->nseg = 1;
    im->dri = 0;
    return 0;
  }
  // segments: the whole scan, or one per restart interval (the RSTn markers are byte aligned: B.1.1.5).  The device
  // decoder finds the markers itself while it unstuffs the scan (segs == nullptr): only the interval count is needed.
  const int total = im->mcux * im->mcuy;
  if (!segs) {
    im->seg0 = 0;
    im->nseg = im->dri > 0 ? (total + im->dri - 1) / im->dri : 1;
    return 0;
  }
  im->seg0 = (int32_t)segs->size();
  if (im->dri <= 0) {
    segs->push_back(JpegSeg{img_index, 0, total, (int32_t)pos});
  } else {
    int64_t p = pos;
    int mcu = 0;
    for (;;) {
      const int cnt = total - mcu < im->dri ? total - mcu : im->dri;
      segs->push_back(JpegSeg{img_index, mcu, cnt, (int32_t)p});
      mcu += cnt;
      if (mcu >= total) break;
      for (;;) {  // next RSTn
        const uint8_t* f = (const uint8_t*)memchr(d + p, 0xFF, (size_t)(len - p));
        if (!f || f + 1 >= d + len) return jpeg_fail(err, "restart marker missing");
        p = f - d;
        const int nx = d[p + 1];
        if (nx >= 0xD0 && nx <= 0xD7) { p += 2; break; }
        if (nx == 0x00 || nx == 0xFF) { p += nx == 0 ? 2 : 1; continue; }
        return jpeg_fail(err, "restart marker missing");
      }
    }
  }
  im->nseg = (int32_t)segs->size() - im->seg0;
  return 0;
}

// ------------------------------------------------------------------------------------ //
// entropy decode of one segment (one thread)                                             //
// ------------------------------------------------------------------------------------ //

struct JpegBits {
  const uint8_t* p;
  const uint8_t* end;
  uint64_t buf;  // the low n bits are unread
  int n;
  int marker;    // a marker was reached: the rest of the segment reads as zero bits (jdhuff.c)
  int fake;      // zero bits at the tail of buf that are not file data
};

MTGV_HD unsigned jpeg_ldb(const uint8_t* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

MTGV_HD void jpeg_fill(JpegBits& b) {
  while (b.n <= 56) {
    unsigned c = 0;
    if (!b.marker && b.p < b.end) {
      c = jpeg_ldb(b.p);
      if (c == 0xFF) {
        const unsigned nx = b.p + 1 < b.end ? jpeg_ldb(b.p + 1) : 0xD9u;
        if (nx == 0) b.p += 2;  // stuffed zero byte
        else { b.marker = 1; c = 0; b.fake += 8; }
      } else {
        b.p++;
      }
    } else {
      b.fake += 8;
    }
    b.buf = (b.buf << 8) | c;
    b.n += 8;
  }
}

MTGV_HD unsigned jpeg_peek(const JpegBits& b, int k) { return (unsigned)(b.buf >> (b.n - k)) & ((1u << k) - 1u); }

// needs >= 16 unread bits
MTGV_HD int jpeg_symbol(JpegBits& b, const JpegTables* t, int ti) {
  const unsigned e = t->fast[ti][jpeg_peek(b, kJpegFastBits)];
  if (e) {
    b.n -= (int)(e >> 8);
    return (int)(e & 255u);
  }
  int l = kJpegFastBits + 1;
  int code = (int)jpeg_peek(b, l);
  while (l <= 16 && code > t->maxcode[ti][l]) {
    l++;
    if (l <= 16) code = (int)jpeg_peek(b, l);
  }
  if (l > 16) {  // corrupt data: libjpeg warns and returns a zero symbol
    b.n -= 16;
    return 0;
  }
  b.n -= l;
  return t->vals[ti][(code + t->valoff[ti][l]) & 255];
}

MTGV_HD int jpeg_receive_extend(JpegBits& b, int s) {
  if (s == 0) return 0;
  const int v = (int)jpeg_peek(b, s);
  b.n -= s;
  return v < (1 << (s - 1)) ? v - ((1 << s) - 1) : v;
}

// coef: the zero-initialised coefficient scratch (int16, 64 per block, natural order, not dequantised)
MTGV_HD void jpeg_decode_segment(const uint8_t* file, const JpegImg& im, const JpegTables* tb, const JpegSeg& sg, int16_t* coef,
                                 const uint8_t* zz) {
  JpegBits b;
  b.p = file + sg.byte_off;
  b.end = file + im.file_len;
  b.buf = 0; b.n = 0; b.marker = 0; b.fake = 0;
  int pred[3] = {0, 0, 0};
  int my = sg.mcu0 / im.mcux, mx = sg.mcu0 - my * im.mcux;
  for (int m = 0; m < sg.nmcu; m++) {
    // jdhuff.c decode_mcu: once a request for bits ran past the data (premature end of the file) the MCU in progress is
    // finished from zero bits and the later MCUs of the interval are skipped: their coefficients stay zero (grey)
    if (b.n < b.fake) break;
    for (int c = 0; c < im.ncomp; c++) {
      const int tdc = im.td[c], tac = 2 + im.ta[c];
      for (int by = 0; by < im.cv[c]; by++) {
        for (int bx = 0; bx < im.ch[c]; bx++) {
          int16_t* blk = coef + (im.coef_blk + im.blk0[c] + (int64_t)(my * im.cv[c] + by) * im.bw[c] + mx * im.ch[c] + bx) * 64;
          if (b.n < 32) jpeg_fill(b);
          int s = jpeg_symbol(b, tb, tdc);
          pred[c] += jpeg_receive_extend(b, s & 15);
          blk[0] = (int16_t)pred[c];
          int k = 1;
          while (k < 64) {
            if (b.n < 32) jpeg_fill(b);
            const int rs = jpeg_symbol(b, tb, tac);
            const int r = rs >> 4;
            s = rs & 15;
            if (s == 0) {
              if (r != 15) break;
              k += 16;
              continue;
            }
            k += r;
            const int v = jpeg_receive_extend(b, s);
            if (k > 63) break;
            blk[zz[k]] = (int16_t)v;
            k++;
          }
        }
      }
    }
    if (++mx == im.mcux) { mx = 0; my++; }
  }
}

// ------------------------------------------------------------------------------------ //
// jidctint.c: one 8-point pass of jpeg_idct_islow                                        //
// ------------------------------------------------------------------------------------ //

MTGV_HD void jpeg_idct8(const int* x, int* o, int shift) {
  int z2 = x[2], z3 = x[6];
  int z1 = (z2 + z3) * 4433;
  int tmp2 = z1 - z3 * 15137;
  int tmp3 = z1 + z2 * 6270;
  z2 = x[0]; z3 = x[4];
  int tmp0 = (z2 + z3) * 8192;
  int tmp1 = (z2 - z3) * 8192;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = x[7]; tmp1 = x[5]; tmp2 = x[3]; tmp3 = x[1];
  z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * 9633;
  tmp0 *= 2446; tmp1 *= 16819; tmp2 *= 25172; tmp3 *= 12299;
  z1 *= -7373; z2 *= -20995;
  z3 = z3 * -16069 + z5;
  z4 = z4 * -3196 + z5;
  tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
  const int r = 1 << (shift - 1);
  o[0] = (tmp10 + tmp3 + r) >> shift; o[7] = (tmp10 - tmp3 + r) >> shift;
  o[1] = (tmp11 + tmp2 + r) >> shift; o[6] = (tmp11 - tmp2 + r) >> shift;
  o[2] = (tmp12 + tmp1 + r) >> shift; o[5] = (tmp12 - tmp1 + r) >> shift;
  o[3] = (tmp13 + tmp0 + r) >> shift; o[4] = (tmp13 - tmp0 + r) >> shift;
}

constexpr int kJpegPass1Shift = 13 - 2, kJpegPass2Shift = 13 + 2 + 3;

MTGV_HD int jpeg_clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// ------------------------------------------------------------------------------------ //
// jdsample.c + jdcolor.c: one output pixel                                               //
// ------------------------------------------------------------------------------------ //

MTGV_HD int jpeg_chroma(const JpegImg& im, const uint8_t* planes, int c, int y, int x) {
  const uint8_t* P = planes + im.plane_off[c];
  const int pitch = im.bw[c] * 8;
  const int hf = im.hf[c], vf = im.vf[c];
  if (hf == 1 && vf == 1) return P[y * pitch + x];
  const int dh = im.dh[c], dw = im.dw[c];
  int inrow = y, other = y;
  if (vf == 2) {  // the neighbouring row of the triangle filter; edge rows replicated (jdmainct.c context rows)
    inrow = y >> 1;
    other = (y & 1) ? (inrow + 1 < dh ? inrow + 1 : dh - 1) : (inrow > 0 ? inrow - 1 : 0);
  }
  const uint8_t* r0 = P + inrow * pitch;
  const uint8_t* r1 = P + other * pitch;
  if (hf == 1) return (3 * r0[x] + r1[x] + ((y & 1) ? 2 : 1)) >> 2;  // h1v2_fancy_upsample
  const int j = x >> 1, odd = x & 1;
  if (dw <= 2) return r0[j];  // h2v1_upsample / h2v2_upsample: too narrow for the filter, replicate
  const int nb = odd ? j + 1 : j - 1;
  const bool edge = nb < 0 || nb >= dw;
  if (vf == 1) return edge ? r0[j] : (3 * r0[j] + r0[nb] + (odd ? 2 : 1)) >> 2;  // h2v1_fancy_upsample
  const int cs = 3 * r0[j] + r1[j], bias = odd ? 7 : 8;                           // h2v2_fancy_upsample
  if (edge) return (4 * cs + bias) >> 4;
  return (3 * cs + 3 * r0[nb] + r1[nb] + bias) >> 4;
}

MTGV_HD void jpeg_pixel(const JpegImg& im, const uint8_t* planes, int y, int x, int* rgb) {
  const int Y = planes[im.plane_off[0] + (int64_t)y * (im.bw[0] * 8) + x];
  if (im.ncomp == 1) {
    rgb[0] = rgb[1] = rgb[2] = Y;
    return;
  }
  const int cb = jpeg_chroma(im, planes, 1, y, x) - 128, cr = jpeg_chroma(im, planes, 2, y, x) - 128;
  rgb[0] = jpeg_clamp255(Y + ((91881 * cr + 32768) >> 16));
  rgb[1] = jpeg_clamp255(Y + ((-22554 * cb + 32768 - 46802 * cr) >> 16));
  rgb[2] = jpeg_clamp255(Y + ((116130 * cb + 32768) >> 16));
}

}  // namespace mtgv
