// TEST INFRASTRUCTURE ONLY.  Host build (g++ -ffp-contract=off) of the host/device fp64
// parameter-expansion headers so the `-m "not gpu"` suite can pin them against cv2 and the
// oracle without a GPU.  Not linked into libmtgv.so and never used by the product path.
#include "../../mtgvision_b200/csrc/mtgv_expand.cuh"
#include "../../mtgvision_b200/csrc/mtgv_poly.cuh"
#include "../../mtgvision_b200/csrc/mtgv_mask.h"

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

int hh_expand_encoder(const mtgv_enc_tape* tape, int n, const mtgv_enc_config* cfg, int card_h, int card_w, int n_cards,
                      const int32_t* labels3, const int32_t* grp_off, const int32_t* grp_mem, int n_bgs,
                      const int32_t* bg_hw, mtgv_enc_params* out) {
  PoolMeta pm{card_h, card_w, n_cards, n_bgs, labels3, grp_off, grp_mem, bg_hw};
  int bad = 0;
  for (int s = 0; s < n; s++) bad += expand_encoder_sample(&tape[s], cfg, pm, &out[s]) != 0;
  return bad;
}

void hh_round_rect_mask(int h, int w, int radius, float* out) { host_round_rect_mask(h, w, radius, out); }

double hh_poly_area(const double* p, int n) { return poly_area(p, n); }
int hh_clip_convex(const double* subj, int ns, const double* clip, int nc, double* out) {
  return clip_convex(subj, ns, clip, nc, out);
}
}
