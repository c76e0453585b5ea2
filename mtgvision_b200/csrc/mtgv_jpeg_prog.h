// mtgv_jpeg_prog.h - progressive JPEG (SOF2): host entropy stage.
//
// imread_float (mtgvision/util/image.py:107-114) reads whatever libjpeg reads; ILSVRC and Scryfall hold some progressive
// files.  A progressive file spreads every 8x8 block over several scans (spectral selection + successive approximation,
// T.81 Annex G) whose refinement passes depend, bit by bit, on what earlier scans left in the block, so its Huffman stage has
// no per-file parallelism worth a kernel.  It runs here, on the host threads that already walk the markers of the batch, and
// only produces the COEFFICIENTS (int16 [block][64], natural order, same block layout as the device's scratch); every pixel
// operation - dequantisation, ISLOW inverse DCT, fancy upsampling, colour conversion - stays on the device kernels of
// mtgv_jpeg.cu, which is why the result is bit-exact with cv2.imdecode like the baseline path (for a complete file libjpeg
// applies no block smoothing: jdcoefct.c smoothing_ok() needs an AC band that is still imprecise).
//
// The bit-level logic follows libjpeg's published algorithm (jdphuff.c: decode_mcu_DC_first / AC_first / DC_refine /
// AC_refine), restated; libjpeg-turbo 3.1.2 is the version inside opencv-python 4.13.0.
#pragma once

#include <cstring>
#include <string>
#include <vector>

#include "mtgv_jpeg.cuh"

namespace mtgv {

struct ProgHuff {  // canonical Huffman table, 16-bit walk
  bool set = false;
  int maxcode[18];
  int valoff[17];
  uint8_t vals[256];
};

struct ProgBits {
  const uint8_t* p;
  const uint8_t* end;
  uint32_t buf = 0;
  int n = 0;
  bool hit_marker = false;  // a marker was reached: further bits are zeros (libjpeg's behaviour at the end of a scan's data)
  int get_byte() {
    if (hit_marker || p >= end) { hit_marker = true; return 0; }
    int b = *p++;
    if (b == 0xFF) {
      if (p < end && *p == 0x00) { p++; return 0xFF; }
      // a marker (or fill bytes in front of one): leave it for the scan loop
      p--;
      hit_marker = true;
      return 0;
    }
    return b;
  }
  int bit() {
    if (n == 0) { buf = (uint32_t)get_byte(); n = 8; }
    n--;
    return (int)((buf >> n) & 1u);
  }
  int bits(int k) {
    int v = 0;
    for (int i = 0; i < k; i++) v = (v << 1) | bit();
    return v;
  }
  void align() { n = 0; }
};

inline bool prog_build_huff(const uint8_t* counts, const uint8_t* vals, int total, ProgHuff* h) {
  int code = 0, k = 0;
  for (int l = 1; l <= 16; l++) {
    h->valoff[l] = k - code;
    k += counts[l - 1];
    code += counts[l - 1];
    h->maxcode[l] = counts[l - 1] ? code - 1 : -1;
    if (code > (1 << l)) return false;
    code <<= 1;
  }
  h->maxcode[17] = 0x7fffffff;
  memcpy(h->vals, vals, (size_t)total);
  h->set = true;
  return true;
}

inline int prog_decode(ProgBits& b, const ProgHuff& h) {
  // canonical codes: after l bits the value is a code of length l exactly when it does not exceed the largest one of that length
  int code = 0;
  for (int l = 1; l <= 16; l++) {
    code = (code << 1) | b.bit();
    if (code <= h.maxcode[l]) return h.vals[(code + h.valoff[l]) & 255];
  }
  return 0;  // corrupt data: libjpeg warns and returns a zero symbol
}

inline int prog_extend(int v, int s) { return s == 0 ? 0 : (v < (1 << (s - 1)) ? v - ((1 << s) - 1) : v); }

// Decodes every scan of a progressive file into coef (zero-initialised, [im.nblk][64], natural order, component c's block
// (by, bx) at index im.blk0[c] + by * im.bw[c] + bx).  im: geometry from jpeg_parse (frame header).  Returns 0 or -1.
inline int jpeg_decode_progressive(const uint8_t* d, int64_t len, const JpegImg& im, const int* comp_id, int16_t* coef, std::string* err) {
  ProgHuff dc[4], ac[4];
  int restart_interval = 0;
  int64_t pos = 2;
  bool any_scan = false;
  auto fail = [&](const char* m) { *err = std::string("progressive: ") + m; return -1; };
  for (;;) {
    if (pos + 4 > len) break;  // ran out of segments: treat what was decoded as the image (a missing EOI is common)
    if (d[pos] != 0xFF) return fail("marker expected");
    while (pos + 1 < len && d[pos + 1] == 0xFF) pos++;
    if (pos + 1 >= len) break;
    const int m = d[pos + 1];
    pos += 2;
    if (m == 0xD9) break;  // EOI
    if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
    if (pos + 2 > len) return fail("truncated segment");
    const int seglen = (d[pos] << 8) | d[pos + 1];
    if (seglen < 2 || pos + seglen > len) return fail("truncated segment");
    const uint8_t* s = d + pos + 2;
    const int n = seglen - 2;
    if (m == 0xC4) {
      int q = 0;
      while (q < n) {
        if (q + 17 > n) return fail("bad Huffman table");
        const int tc = s[q] >> 4, th = s[q] & 15;
        if (tc > 1 || th > 3) return fail("bad Huffman table id");
        int total = 0;
        for (int l = 0; l < 16; l++) total += s[q + 1 + l];
        if (total > 256 || q + 17 + total > n) return fail("bad Huffman table");
        if (!prog_build_huff(s + q + 1, s + q + 17, total, tc ? &ac[th] : &dc[th])) return fail("bad Huffman table (code overflow)");
        q += 17 + total;
      }
    } else if (m == 0xDD) {
      if (n < 2) return fail("bad DRI");
      restart_interval = (s[0] << 8) | s[1];
    } else if (m == 0xDA) {
      // ---- one scan ----
      if (n < 1) return fail("bad scan header");
      const int ns = s[0];
      if (ns < 1 || ns > im.ncomp || n < 1 + 2 * ns + 3) return fail("bad scan header");
      int comps[3], tdc[3], tac[3];
      for (int i = 0; i < ns; i++) {
        int c = -1;
        for (int j = 0; j < im.ncomp; j++)
          if (comp_id[j] == s[1 + 2 * i]) c = j;
        if (c < 0) return fail("bad scan header");
        comps[i] = c; tdc[i] = s[2 + 2 * i] >> 4; tac[i] = s[2 + 2 * i] & 15;
        if (tdc[i] > 3 || tac[i] > 3) return fail("bad scan header");
      }
      const int Ss = s[1 + 2 * ns], Se = s[2 + 2 * ns], Ah = s[3 + 2 * ns] >> 4, Al = s[3 + 2 * ns] & 15;
      if (Ss > Se || Se > 63 || Al > 13 || (Ss == 0 && Se != 0) || (Ss > 0 && ns != 1)) return fail("bad progression parameters");
      for (int i = 0; i < ns; i++) {
        if (Ss == 0 && Ah == 0 && !dc[tdc[i]].set) return fail("missing Huffman table");
        if (Ss > 0 && !ac[tac[i]].set) return fail("missing Huffman table");
      }
      ProgBits b;
      b.p = d + pos + seglen;
      b.end = d + len;
      // MCU geometry of this scan: interleaved = the frame's MCUs; a single component = its own blocks in raster order
      const bool inter = ns > 1;
      int mcus_x, mcus_y;
      if (inter) { mcus_x = im.mcux; mcus_y = im.mcuy; }
      else { const int c = comps[0]; mcus_x = (im.dw[c] + 7) / 8; mcus_y = (im.dh[c] + 7) / 8; }
      int pred[3] = {0, 0, 0};
      int eobrun = 0, to_restart = restart_interval, next_rst = 0;
      const int p1 = 1 << Al, m1 = -(1 << Al);
      for (int my = 0; my < mcus_y; my++)
        for (int mx = 0; mx < mcus_x; mx++) {
          if (restart_interval && to_restart == 0) {
            // expect RSTn: byte-align, skip to the marker, reset the predictors and the end-of-band run
            b.align();
            const uint8_t* q = b.p;
            while (q + 1 < b.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) {
              if (q[0] == 0xFF && q[1] != 0x00 && q[1] != 0xFF) break;  // some other marker: give up on resynchronising
              q++;
            }
            if (q + 1 < b.end && q[0] == 0xFF && q[1] == 0xD0 + next_rst) { b.p = q + 2; b.hit_marker = false; }
            next_rst = (next_rst + 1) & 7;
            pred[0] = pred[1] = pred[2] = 0;
            eobrun = 0;
            to_restart = restart_interval;
          }
          if (restart_interval) to_restart--;
          for (int i = 0; i < ns; i++) {
            const int c = comps[i];
            const int nbx = inter ? im.ch[c] : 1, nby = inter ? im.cv[c] : 1;
            for (int by = 0; by < nby; by++)
              for (int bx = 0; bx < nbx; bx++) {
                const int gy = inter ? my * im.cv[c] + by : my, gx = inter ? mx * im.ch[c] + bx : mx;
                int16_t* blk = coef + ((int64_t)im.blk0[c] + (int64_t)gy * im.bw[c] + gx) * 64;
                if (Ss == 0) {
                  if (Ah == 0) {  // DC first
                    const int t = prog_decode(b, dc[tdc[i]]);
                    const int diff = prog_extend(b.bits(t), t);
                    pred[c] += diff;
                    blk[0] = (int16_t)(pred[c] * (1 << Al));
                  } else {        // DC refine
                    if (b.bit()) blk[0] = (int16_t)(blk[0] | p1);
                  }
                } else if (Ah == 0) {  // AC first
                  if (eobrun > 0) { eobrun--; continue; }
                  for (int k = Ss; k <= Se; k++) {
                    const int rs = prog_decode(b, ac[tac[i]]);
                    const int r = rs >> 4, sz = rs & 15;
                    if (sz) {
                      k += r;
                      const int v = prog_extend(b.bits(sz), sz);
                      if (k <= 63) blk[kJpegZigzag[k]] = (int16_t)(v * (1 << Al));
                    } else if (r == 15) {
                      k += 15;
                    } else {
                      eobrun = 1 << r;
                      if (r) eobrun += b.bits(r);
                      eobrun--;
                      break;
                    }
                  }
                } else {  // AC refine
                  int k = Ss;
                  if (eobrun == 0) {
                    for (; k <= Se; k++) {
                      const int rs = prog_decode(b, ac[tac[i]]);
                      int r = rs >> 4, sz = rs & 15, val = 0;
                      if (sz) {
                        val = b.bit() ? p1 : m1;  // sz must be 1
                      } else if (r != 15) {
                        eobrun = 1 << r;
                        if (r) eobrun += b.bits(r);
                        break;  // end of band
                      }
                      // skip r still-zero coefficients, feeding correction bits to the nonzero ones on the way
                      do {
                        int16_t* t = blk + kJpegZigzag[k];
                        if (*t != 0) {
                          if (b.bit() && (*t & p1) == 0) *t = (int16_t)(*t >= 0 ? *t + p1 : *t + m1);
                        } else if (--r < 0) {
                          break;
                        }
                        k++;
                      } while (k <= Se);
                      if (val && k <= 63) blk[kJpegZigzag[k]] = (int16_t)val;
                    }
                  }
                  if (eobrun > 0) {
                    // the rest of the band: correction bits for the coefficients that are already nonzero
                    for (; k <= Se; k++) {
                      int16_t* t = blk + kJpegZigzag[k];
                      if (*t != 0 && b.bit() && (*t & p1) == 0) *t = (int16_t)(*t >= 0 ? *t + p1 : *t + m1);
                    }
                    eobrun--;
                  }
                }
              }
          }
        }
      any_scan = true;
      // continue at the marker that ended the entropy-coded data
      const uint8_t* q = b.p;
      while (q + 1 < b.end && !(q[0] == 0xFF && q[1] != 0x00 && q[1] != 0xFF && !(q[1] >= 0xD0 && q[1] <= 0xD7))) q++;
      pos = q - d;
      continue;
    }
    pos += seglen;
  }
  if (!any_scan) return fail("no scan");
  return 0;
}

}  // namespace mtgv
