// TEST INFRASTRUCTURE ONLY.  Host build (g++ -ffp-contract=off) of the host/device fp64
// parameter-expansion headers so the `-m "not gpu"` suite can pin them against cv2 and the
// oracle without a GPU.  Not linked into libmtgv.so and never used by the product path.
#include "../../mtgvision_b200/csrc/mtgv_expand.cuh"
#include "../../mtgvision_b200/csrc/mtgv_poly.cuh"
#include "../../mtgvision_b200/csrc/mtgv_mask.h"
#include "../../mtgvision_b200/csrc/mtgv_det.cuh"
#include "../../mtgvision_b200/csrc/mtgv_jpeg.cuh"
#include "../../mtgvision_b200/csrc/mtgv_jpeg_prog.h"
#include "../../mtgvision_b200/csrc/mtgv_jpegenc.cuh"

using namespace mtgv;

extern "C" {

int hh_get_perspective_transform(const float* src, const float* dst, double* M) {
  return get_perspective_transform(src, dst, M) ? 0 : -1;
}
int hh_invert3x3(const double* M, double* Mi) { return invert3x3(M, Mi) ? 0 : -1; }
void hh_invert_affine(const double* A, double* Ai) { invert_affine(A, Ai); }

void hh_persp_coords(const double* Minv, int dh, int dw, int32_t* X, int32_t* Y) {
  int bw0 = persp_block_w(dh, dw);
  for (int y = 0; y < dh; y++)
    for (int x = 0; x < dw; x++) persp_coord(Minv, x, y, bw0, &X[y * dw + x], &Y[y * dw + x]);
}

void hh_affine_coords(const double* Ainv, int dh, int dw, int32_t* X, int32_t* Y) {
  for (int y = 0; y < dh; y++) {
    int X0 = affine_row_origin(Ainv[1], Ainv[2], y), Y0 = affine_row_origin(Ainv[4], Ainv[5], y);
    for (int x = 0; x < dw; x++) {
      X[y * dw + x] = (X0 + affine_col_delta(Ainv[0], x)) >> 5;
      Y[y * dw + x] = (Y0 + affine_col_delta(Ainv[3], x)) >> 5;
    }
  }
}

int hh_area_taps(int ssize, int dsize, int d, int* start, float* w) { return area_taps(ssize, dsize, d, start, w); }

// the 2-tap table cv::resize(INTER_AREA) uses when enlarging: returns the tap count, weights in w[0..1]
int hh_area_linear_taps(int ssize, int dsize, int d, int* start, float* w) {
  int n;
  float wl, wm, wr;
  area_linear_compact(ssize, dsize, d, start, &n, &wl, &wm, &wr);
  w[0] = wl; w[1] = wr;
  return n & 255;
}

int hh_expand_encoder(const mtgv_enc_tape* tape, int n, const mtgv_enc_config* cfg, int card_h, int card_w, int n_cards,
                      const int32_t* labels3, const int32_t* grp_off, const int32_t* grp_mem, int n_bgs,
                      const int32_t* bg_hw, mtgv_enc_params* out) {
  PoolMeta pm{card_h, card_w, n_cards, n_bgs, labels3, grp_off, grp_mem, bg_hw};
  int bad = 0;
  for (int s = 0; s < n; s++) bad += expand_encoder_sample(&tape[s], cfg, pm, &out[s]) != 0;
  return bad;
}

void hh_round_rect_mask(int h, int w, int radius, float* out) { host_round_rect_mask(h, w, radius, out); }

int hh_det_params_size() { return (int)sizeof(DetParams); }

void hh_line_mask(int ks, int x1, int y1, int x2, int y2, int32_t* words) { det_line_mask(ks, x1, y1, x2, y2, words); }

// runs the placement / label warping of mtgv_det.cuh for n scenes on the host
int hh_det_place(const mtgv_det_tape* tape, int n, const mtgv_det_config* cfg, int card_h, int card_w, int n_cards, int n_bgs,
                 const int32_t* bg_hw, void* params, int32_t* accepted, double* keypoints, int32_t* labels, int32_t* counts) {
  DetKeypoints kp;
  det_keypoints(card_h, card_w, cfg->kind, &kp);
  const size_t nk = (size_t)MTGV_DET_MAX_CARDS * MTGV_DET_MAX_KPOLY;
  int bad = 0;
  for (int s = 0; s < n; s++)
    bad += det_place_scene(&tape[s], cfg, &kp, card_h, card_w, n_cards, n_bgs, bg_hw, (DetParams*)params + s,
                           accepted + (size_t)s * MTGV_DET_MAX_CARDS, keypoints + (size_t)s * nk * MTGV_DET_MAX_KP * 2,
                           labels + (size_t)s * nk, counts + s) != 0;
  return bad;
}

void hh_det_keypoints(int h, int w, int kind, double* pts, double* bbox, int* n_poly, int* n_pts) {
  DetKeypoints kp;
  det_keypoints(h, w, kind, &kp);
  *n_poly = kp.n_poly; *n_pts = kp.n_pts;
  for (int q = 0; q < MTGV_DET_MAX_KPOLY; q++)
    for (int k = 0; k < MTGV_DET_MAX_KP; k++) { pts[(q * MTGV_DET_MAX_KP + k) * 2] = kp.pts[q][k][0]; pts[(q * MTGV_DET_MAX_KP + k) * 2 + 1] = kp.pts[q][k][1]; }
  for (int k = 0; k < 4; k++) { bbox[2 * k] = kp.bbox[k][0]; bbox[2 * k + 1] = kp.bbox[k][1]; }
}

void hh_det_apply(const double* M, const double* pts, int n, double* out) {
  for (int k = 0; k < n; k++) det_apply(M, pts[2 * k], pts[2 * k + 1], &out[2 * k], &out[2 * k + 1]);
}

double hh_poly_area(const double* p, int n) { return poly_area(p, n); }
int hh_clip_convex(const double* subj, int ns, const double* clip, int nc, double* out) {
  return clip_convex(subj, ns, clip, nc, out);
}

// Host run of the JPEG decode arithmetic of mtgv_jpeg.cuh (the device kernels call the same functions per
// segment / block / pixel).  hw: out [2]; out: [h,w,3] RGB or NULL to only parse.  Returns 0, or -1 with msg set.
int hh_jpeg_decode(const uint8_t* file, int64_t len, int32_t* hw, uint8_t* out, char* msg, int msg_cap) {
  JpegImg im;
  JpegTables tb;
  std::vector<JpegSeg> segs;
  std::string err;
  if (jpeg_parse(file, len, 0, &im, &tb, &segs, &err) != 0) {
    snprintf(msg, msg_cap, "%s", err.c_str());
    return -1;
  }
  hw[0] = im.h; hw[1] = im.w;
  if (!out) return 0;
  std::vector<int16_t> coef((size_t)im.nblk * 64, 0);
  int64_t plane_total = 0;
  for (int c = 0; c < im.ncomp; c++) { im.plane_off[c] = plane_total; plane_total += (int64_t)im.bw[c] * im.bh[c] * 64; }
  std::vector<uint8_t> planes((size_t)plane_total);
  if (im.progressive) {
    if (jpeg_decode_progressive(file, len, im, im.comp_id, coef.data(), &err) != 0) {
      snprintf(msg, msg_cap, "%s", err.c_str());
      return -1;
    }
  } else {
    for (const JpegSeg& sg : segs) jpeg_decode_segment(file, im, &tb, sg, coef.data(), kJpegZigzag);
  }
  for (int c = 0; c < im.ncomp; c++)
    for (int b = 0; b < im.bw[c] * im.bh[c]; b++) {
      const int16_t* blk = coef.data() + (size_t)(im.blk0[c] + b) * 64;
      int ws[8][8];
      for (int t = 0; t < 8; t++) {
        int x[8], o[8];
        for (int r = 0; r < 8; r++) x[r] = (int)blk[r * 8 + t] * (int)tb.qt[im.tq[c]][r * 8 + t];
        jpeg_idct8(x, o, kJpegPass1Shift);
        for (int r = 0; r < 8; r++) ws[r][t] = o[r];
      }
      const int by = b / im.bw[c], bx = b % im.bw[c];
      for (int t = 0; t < 8; t++) {
        int o[8];
        jpeg_idct8(ws[t], o, kJpegPass2Shift);
        for (int k = 0; k < 8; k++)
          planes[(size_t)im.plane_off[c] + (size_t)(by * 8 + t) * (im.bw[c] * 8) + bx * 8 + k] = (uint8_t)jpeg_clamp255(o[k] + 128);
      }
    }
  for (int y = 0; y < im.h; y++)
    for (int x = 0; x < im.w; x++) {
      int rgb[3];
      jpeg_pixel(im, planes.data(), y, x, rgb);
      for (int k = 0; k < 3; k++) out[((size_t)y * im.w + x) * 3 + k] = (uint8_t)rgb[k];
    }
  return 0;
}

// Host run of the JPEG encode arithmetic of mtgv_jpegenc.cuh (the device kernels call the same functions per pixel /
// block).  rgb: [h,w,3] uint8, h and w multiples of 16.  Returns the file length (<= cap) or -1.
int64_t hh_jpeg_encode(const uint8_t* rgb, int h, int w, int quality, uint8_t* out, int64_t cap) {
  if (h < 1 || w < 1) return -1;
  JpegEncTables T;
  jpegenc_tables(quality, &T);
  std::vector<uint8_t> o = jpegenc_header(h, w, T);
  const int mx = (w + 15) / 16, my = (h + 15) / 16, nmcu = mx * my;
  std::vector<int16_t> coef((size_t)nmcu * 6 * 64);
  for (int m = 0; m < nmcu; m++) {
    const int y0 = (m / mx) * 16, x0 = (m % mx) * 16;
    int Y[16][16], Cb[8][8], Cr[8][8];
    for (int qy = 0; qy < 8; qy++)
      for (int qx = 0; qx < 8; qx++) {
        int sb = 0, sr = 0;
        for (int dy = 0; dy < 2; dy++)
          for (int dx = 0; dx < 2; dx++) {
            const int gy = y0 + 2 * qy + dy, gx = jpegenc_col(x0 + 2 * qx + dx, w);
            const uint8_t* p = rgb + ((size_t)jpegenc_luma_row(gy, h) * w + gx) * 3;
            const uint8_t* pc = rgb + ((size_t)jpegenc_chroma_row(gy, h) * w + gx) * 3;
            int y, cb, cr, yc;
            jpegenc_ycc(p[0], p[1], p[2], &y, &cb, &cr);
            jpegenc_ycc(pc[0], pc[1], pc[2], &yc, &cb, &cr);
            Y[2 * qy + dy][2 * qx + dx] = y; sb += cb; sr += cr;
          }
        const int bias = ((x0 / 2 + qx) & 1) ? 2 : 1;
        Cb[qy][qx] = (sb + bias) >> 2; Cr[qy][qx] = (sr + bias) >> 2;
      }
    for (int j = 0; j < 6; j++) {
      int ws[8][8], d[8], r[8];
      for (int row = 0; row < 8; row++) {
        for (int k = 0; k < 8; k++) d[k] = (j < 4 ? Y[(j >> 1) * 8 + row][(j & 1) * 8 + k] : (j == 4 ? Cb[row][k] : Cr[row][k])) - 128;
        jpegenc_fdct8(d, r, true);
        for (int k = 0; k < 8; k++) ws[row][k] = r[k];
      }
      int nat[64];
      for (int col = 0; col < 8; col++) {
        for (int k = 0; k < 8; k++) d[k] = ws[k][col];
        jpegenc_fdct8(d, r, false);
        for (int k = 0; k < 8; k++) nat[k * 8 + col] = jpegenc_quant(r[k], T.q[j < 4 ? 0 : 1][k * 8 + col]);
      }
      for (int k = 0; k < 64; k++) coef[((size_t)m * 6 + j) * 64 + k] = (int16_t)nat[kJpegZigzag[k]];
    }
    jpegenc_dummy_blocks(&coef[(size_t)m * 6 * 64], (m % mx) == mx - 1 && (((w + 7) / 8) & 1), (m / mx) == my - 1 && (((h + 7) / 8) & 1));
  }
  struct Bits {
    std::vector<uint8_t>* o; uint64_t acc = 0; int n = 0;
    void operator()(unsigned code, int size) {
      acc = (acc << size) | code; n += size;
      while (n >= 8) { uint8_t b = (uint8_t)(acc >> (n - 8)); o->push_back(b); if (b == 0xFF) o->push_back(0); n -= 8; }
    }
  } put;
  put.o = &o;
  for (int m = 0; m < nmcu; m++)
    for (int j = 0; j < 6; j++) {
      const int pb = jpegenc_pred_block(m, j);
      const int t = j < 4 ? 0 : 1;
      jpegenc_block(&coef[((size_t)m * 6 + j) * 64], pb < 0 ? 0 : coef[(size_t)pb * 64], T.dc[t], T.ac[t], put);
    }
  if (put.n) put((1u << (8 - put.n)) - 1u, 8 - put.n);
  o.push_back(0xFF); o.push_back(0xD9);
  if ((int64_t)o.size() > cap) return -1;
  memcpy(out, o.data(), o.size());
  return (int64_t)o.size();
}
}
