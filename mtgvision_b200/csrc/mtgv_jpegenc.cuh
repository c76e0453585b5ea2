// mtgv_jpegenc.cuh - baseline JPEG encode arithmetic shared by the device kernels (mtgv_jpegenc.cu) and the host-side
// unit-test harness (tests/host_harness.cpp).  SURVEY 8f.2: save_sample (mtgvision/od_datasets.py:794-832) writes
// every scene with imwrite (mtgvision/util/image.py:95-104) = cv2.imwrite with OpenCV's defaults, i.e. libjpeg-turbo
// (3.1.2 in opencv-python 4.13.0) at quality 95, 4:2:0, standard Huffman tables, JFIF header.  All of it is integer
// arithmetic, restated here so that the file bytes equal cv2's:
//   RGB -> YCbCr     jccolor.c   (rgb_ycc_convert, SCALEBITS 16)
//   downsampling     jcsample.c  (h2v2_downsample, bias 1,2,1,2.. along a row)
//   forward DCT      jfdctint.c  (jpeg_fdct_islow, CONST_BITS 13, PASS1_BITS 2, output scaled by 8)
//   quantisation     jcdctmgr.c  (round half away from zero by 8*Q; tables: jcparam.c jpeg_quality_scaling)
//   entropy coding   jchuff.c    (encode_one_block with the tables of jstdhuff.c)
//   markers          jcmarker.c  (SOI APP0 DQT DQT SOF0 DHT*4 SOS ... EOI)
// Sizes that are not whole 16x16 MCUs (save_sample itself asserts 640x640) follow libjpeg's edge replication and
// dummy-block rules (jcsample.c, jcprepct.c, jccoefct.c): jpegenc_col / _luma_row / _chroma_row / _dummy_blocks.
#pragma once

#include <stdint.h>
#include <string.h>

#include <vector>

#include "mtgv_geom.cuh"
#include "mtgv_jpeg.cuh"  // kJpegZigzag

namespace mtgv {

struct JpegEncTables {
  uint16_t q[2][64];    // quantisation tables (luma, chroma), natural order
  uint32_t dc[2][16];   // [table][category] = length << 16 | code
  uint32_t ac[2][256];  // [table][run << 4 | size]
};

// ---------------------------------------------------------------- host: tables and header
inline void jpegenc_huff(const uint8_t* bits, const uint8_t* vals, uint32_t* tab, int ntab) {
  for (int i = 0; i < ntab; i++) tab[i] = 0;
  unsigned code = 0;
  int k = 0;
  for (int l = 1; l <= 16; l++) {
    for (int i = 0; i < bits[l - 1]; i++, k++, code++) tab[vals[k]] = ((uint32_t)l << 16) | code;
    code <<= 1;
  }
}

struct JpegEncStd {  // Annex K tables (jcparam.c, jstdhuff.c)
  uint8_t ql[64], qc[64];
  uint8_t dcl_bits[16], dcc_bits[16], acl_bits[16], acc_bits[16];
  uint8_t dc_vals[12];
  uint8_t acl_vals[162], acc_vals[162];
};

inline const JpegEncStd& jpegenc_std() {
  static const JpegEncStd S = {
      {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
       18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100,
       103, 99},
      {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
       99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99},
      {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0},
      {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0},
      {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7D},
      {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77},
      {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11},
      {0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81,
       0x91, 0xA1, 0x08, 0x23, 0x42, 0xB1, 0xC1, 0x15, 0x52, 0xD1, 0xF0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0A, 0x16, 0x17, 0x18,
       0x19, 0x1A, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2A, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48,
       0x49, 0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6A, 0x73, 0x74, 0x75,
       0x76, 0x77, 0x78, 0x79, 0x7A, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99,
       0x9A, 0xA2, 0xA3, 0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA, 0xC2, 0xC3,
       0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA, 0xE1, 0xE2, 0xE3, 0xE4, 0xE5,
       0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF1, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9, 0xFA},
      {0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08,
       0x14, 0x42, 0x91, 0xA1, 0xB1, 0xC1, 0x09, 0x23, 0x33, 0x52, 0xF0, 0x15, 0x62, 0x72, 0xD1, 0x0A, 0x16, 0x24, 0x34, 0xE1, 0x25,
       0xF1, 0x17, 0x18, 0x19, 0x1A, 0x26, 0x27, 0x28, 0x29, 0x2A, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47,
       0x48, 0x49, 0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6A, 0x73, 0x74,
       0x75, 0x76, 0x77, 0x78, 0x79, 0x7A, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97,
       0x98, 0x99, 0x9A, 0xA2, 0xA3, 0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA,
       0xC2, 0xC3, 0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA, 0xE2, 0xE3, 0xE4,
       0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9, 0xFA}};
  return S;
}

// jcparam.c: jpeg_quality_scaling + jpeg_add_quant_table(force_baseline)
inline void jpegenc_tables(int quality, JpegEncTables* T) {
  const JpegEncStd& S = jpegenc_std();
  quality = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
  const int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
  for (int i = 0; i < 64; i++) {
    int l = (S.ql[i] * scale + 50) / 100, c = (S.qc[i] * scale + 50) / 100;
    T->q[0][i] = (uint16_t)(l < 1 ? 1 : (l > 255 ? 255 : l));
    T->q[1][i] = (uint16_t)(c < 1 ? 1 : (c > 255 ? 255 : c));
  }
  jpegenc_huff(S.dcl_bits, S.dc_vals, T->dc[0], 16);
  jpegenc_huff(S.dcc_bits, S.dc_vals, T->dc[1], 16);
  jpegenc_huff(S.acl_bits, S.acl_vals, T->ac[0], 256);
  jpegenc_huff(S.acc_bits, S.acc_vals, T->ac[1], 256);
}

// jcmarker.c: everything in front of the entropy-coded data
inline std::vector<uint8_t> jpegenc_header(int h, int w, const JpegEncTables& T) {
  const JpegEncStd& S = jpegenc_std();
  std::vector<uint8_t> o;
  auto marker = [&](int code, const std::vector<uint8_t>& p) {
    o.push_back(0xFF); o.push_back((uint8_t)code);
    o.push_back((uint8_t)((p.size() + 2) >> 8)); o.push_back((uint8_t)((p.size() + 2) & 255));
    o.insert(o.end(), p.begin(), p.end());
  };
  o.push_back(0xFF); o.push_back(0xD8);
  marker(0xE0, {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0});
  for (int t = 0; t < 2; t++) {
    std::vector<uint8_t> p{(uint8_t)t};
    for (int k = 0; k < 64; k++) p.push_back((uint8_t)T.q[t][kJpegZigzag[k]]);
    marker(0xDB, p);
  }
  marker(0xC0, {8, (uint8_t)(h >> 8), (uint8_t)(h & 255), (uint8_t)(w >> 8), (uint8_t)(w & 255), 3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1});
  const uint8_t* bits[4] = {S.dcl_bits, S.acl_bits, S.dcc_bits, S.acc_bits};
  const uint8_t* vals[4] = {S.dc_vals, S.acl_vals, S.dc_vals, S.acc_vals};
  const int nvals[4] = {12, 162, 12, 162}, ids[4] = {0x00, 0x10, 0x01, 0x11};
  for (int t = 0; t < 4; t++) {
    std::vector<uint8_t> p{(uint8_t)ids[t]};
    p.insert(p.end(), bits[t], bits[t] + 16);
    p.insert(p.end(), vals[t], vals[t] + nvals[t]);
    marker(0xC4, p);
  }
  marker(0xDA, {3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0});
  return o;
}

// ---------------------------------------------------------------- shared arithmetic
MTGV_HD void jpegenc_ycc(int r, int g, int b, int* y, int* cb, int* cr) {
  *y = (19595 * r + 38470 * g + 7471 * b + 32768) >> 16;
  *cb = (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16;
  *cr = (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16;
}

// one 8-point pass of jpeg_fdct_islow; first: the row pass (outputs scaled up by PASS1_BITS)
MTGV_HD void jpegenc_fdct8(const int* d, int* o, bool first) {
  int tmp0 = d[0] + d[7], tmp7 = d[0] - d[7], tmp1 = d[1] + d[6], tmp6 = d[1] - d[6];
  int tmp2 = d[2] + d[5], tmp5 = d[2] - d[5], tmp3 = d[3] + d[4], tmp4 = d[3] - d[4];
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  const int sh = first ? 13 - 2 : 13 + 2, rnd = 1 << (sh - 1);
  if (first) {
    o[0] = (tmp10 + tmp11) * 4;
    o[4] = (tmp10 - tmp11) * 4;
  } else {
    o[0] = (tmp10 + tmp11 + 2) >> 2;
    o[4] = (tmp10 - tmp11 + 2) >> 2;
  }
  int z1 = (tmp12 + tmp13) * 4433;
  o[2] = (z1 + tmp13 * 6270 + rnd) >> sh;
  o[6] = (z1 - tmp12 * 15137 + rnd) >> sh;
  z1 = tmp4 + tmp7;
  int z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
  const int z5 = (z3 + z4) * 9633;
  tmp4 *= 2446; tmp5 *= 16819; tmp6 *= 25172; tmp7 *= 12299;
  z1 *= -7373; z2 *= -20995;
  z3 = z3 * -16069 + z5;
  z4 = z4 * -3196 + z5;
  o[7] = (tmp4 + z1 + z3 + rnd) >> sh;
  o[5] = (tmp5 + z2 + z4 + rnd) >> sh;
  o[3] = (tmp6 + z2 + z3 + rnd) >> sh;
  o[1] = (tmp7 + z1 + z4 + rnd) >> sh;
}

MTGV_HD int jpegenc_quant(int coef, int q) {  // q = table entry; the DCT output carries a factor 8
  const int q8 = q << 3;
  const int mag = ((coef < 0 ? -coef : coef) + (q8 >> 1)) / q8;
  return coef < 0 ? -mag : mag;
}

MTGV_HD int jpegenc_nbits(int v) {  // bits needed for |v| (JPEG_NBITS)
  v = v < 0 ? -v : v;
  int n = 0;
  while (v) { n++; v >>= 1; }
  return n;
}

// jchuff.c encode_one_block.  zz: the block's quantised coefficients in ZIGZAG order; put(code, size) appends bits.
template <class Put>
MTGV_HD void jpegenc_block(const int16_t* zz, int last_dc, const uint32_t* dc_tab, const uint32_t* ac_tab, Put& put) {
  int diff = (int)zz[0] - last_dc;
  int nb = jpegenc_nbits(diff);
  uint32_t e = dc_tab[nb];
  put(e & 0xffffu, (int)(e >> 16));
  if (nb) put((unsigned)(diff < 0 ? diff - 1 : diff) & ((1u << nb) - 1u), nb);
  int r = 0;
  for (int k = 1; k < 64; k++) {
    const int v = zz[k];
    if (v == 0) { r++; continue; }
    while (r > 15) { e = ac_tab[0xF0]; put(e & 0xffffu, (int)(e >> 16)); r -= 16; }
    nb = jpegenc_nbits(v);
    e = ac_tab[(r << 4) + nb];
    put(e & 0xffffu, (int)(e >> 16));
    put((unsigned)(v < 0 ? v - 1 : v) & ((1u << nb) - 1u), nb);
    r = 0;
  }
  if (r > 0) { e = ac_tab[0]; put(e & 0xffffu, (int)(e >> 16)); }
}

// Edge rules for sizes that are not whole MCUs.  Columns are replicated at full resolution before the chroma box filter
// (jcsample.c expand_right_edge); luma rows are replicated; for chroma the rows are replicated to an even count, filtered,
// and the DOWNSAMPLED rows replicated below that (jcprepct.c) - so a tap of a chroma row past the last one reads the rows
// of the last one.
MTGV_HD int jpegenc_col(int x, int W) { return x < W ? x : W - 1; }
MTGV_HD int jpegenc_luma_row(int y, int H) { return y < H ? y : H - 1; }
MTGV_HD int jpegenc_chroma_row(int y, int H) {
  const int ch = (H + 1) >> 1;
  int cr = y >> 1;
  if (cr >= ch) cr = ch - 1;
  const int r = 2 * cr + (y & 1);
  return r < H ? r : H - 1;
}

// Luma blocks of an MCU that lie wholly outside the image are "dummy" blocks: zero AC terms, DC term of the block before
// them in the MCU (jccoefct.c compress_data).  blk: the MCU's four luma blocks (scan order, 64 coefficients each, DC
// first); right / bottom: the MCU's second block column / row is outside the image.
MTGV_HD void jpegenc_dummy_blocks(int16_t* blk, bool right, bool bottom) {
  if (!right && !bottom) return;
  for (int j = 1; j < 4; j++) {
    const bool dummy = (j >= 2 && bottom) || ((j & 1) && right);
    if (!dummy) continue;
    const int src = (j >= 2 && bottom) ? 1 : j - 1;  // a dummy row copies the last block of the row above
    const int16_t dcv = blk[src * 64];
    for (int k = 0; k < 64; k++) blk[j * 64 + k] = 0;
    blk[j * 64] = dcv;
  }
}

// index (within the 6 blocks of an MCU, scan order Y00 Y01 Y10 Y11 Cb Cr) of the block whose DC predicts block j of
// MCU m: returns the linear block index (m' * 6 + j') or -1 for "no predecessor" (prediction 0)
MTGV_HD int jpegenc_pred_block(int m, int j) {
  if (j >= 1 && j <= 3) return m * 6 + j - 1;
  if (m == 0) return -1;
  return (m - 1) * 6 + (j == 0 ? 3 : j);
}

}  // namespace mtgv
