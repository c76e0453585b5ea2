// mtgv_det.cuh - detection path, subsystem (4): placement rejection sampling, label warping and
// scene-program expansion for one scene.  Host/device: unit-tested on the CPU through
// tests/host_harness against oracle/det_oracle.py, runs in k_det_place on the GPU.
//
// Reference: mtgvision/od_datasets.py  place_card_on_background_get_transform (:287-377),
// apply_transform_2d (:64-70), get_rotate_over_output_transform (:85-118),
// make_card_with_mask keypoints (:244-279), generate_synthetic_image bookkeeping (:555-611).
#pragma once

#include "../../include/mtgv.h"
#include "mtgv_geom.cuh"
#include "mtgv_poly.cuh"

namespace mtgv {

// a GlassBlur expands to blur, <= 2 swap rounds, blur (3 extra entries); the graphs hold at most two (one per blur family)
constexpr int kDetProgMax = MTGV_DET_MAX_PRE + MTGV_DET_MAX_POST + 1 + 6;
constexpr int kPhCards = 100;      // marker in the scene program: composite the placed cards here
constexpr int kPhGlassSwap = 101;  // one swap round of a GlassBlur: i[0] = max_delta, i[1] = round, i[2] = rounds
constexpr int kBlurHalfMax = 10;  // sigma <= 3 -> ksize <= 19
// "Boundary" ops need the whole image of the previous step (neighbourhoods: Gaussian / median / motion blur; image
// statistic: ISONoise): the scene program is cut there and continues in another pixel pass over a float32 scratch image.
// The reference graphs hold at most 11 of them: one blur in bg_light, two noise families (ISONoise) and two blur families
// (a GlassBlur = blur + two swap rounds + blur = 4).
constexpr int kDetMaxBlur = 11;

MTGV_HD bool det_is_boundary(int code) {
  return code == MTGV_PH_GAUSS_BLUR || code == MTGV_PH_MEDIAN_BLUR || code == MTGV_PH_MOTION_BLUR || code == MTGV_PH_ISO_NOISE ||
         code == kPhGlassSwap;
}

// cv2.line(kernel, (x1,y1), (x2,y2), 1, thickness=1) on a ks x ks kernel as a bit mask (bit y*ks + x): cv::LineIterator,
// 8-connected, left to right (drawing.cpp Line()): the endpoints are swapped when x2 < x1, the longer axis advances one cell
// per point and the other whenever the error term err = major - 2*minor has gone negative (checked BEFORE it is updated).
MTGV_HD void det_line_mask(int ks, int x1, int y1, int x2, int y2, int32_t* words) {
  for (int k = 0; k < 4; k++) words[k] = 0;
  if (x2 < x1) { int t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
  int dx = x2 - x1, dy = y2 - y1;
  const int sy = dy < 0 ? -1 : 1;
  dy = dy < 0 ? -dy : dy;
  const bool steep = dy > dx;
  const int major = steep ? dy : dx, minor = steep ? dx : dy;
  int err = major - 2 * minor, x = x1, y = y1;
  for (int i = 0; i <= major; i++) {
    const int b = y * ks + x;
    words[b >> 5] |= (int32_t)(1u << (b & 31));
    const bool both = err < 0;
    err += both ? 2 * major - 2 * minor : -2 * minor;
    if (steep) { y += sy; if (both) x += 1; }
    else { x += 1; if (both) y += sy; }
  }
}

struct DetPhotoX {
  int32_t code;
  int32_t i[5];   // ERASE: top,left,h,w,fill   BLUR / MEDIAN: i[0] = ksize   MOTION: i[0] = ksize, i[1..4] = kernel bit mask
  float f[4];     // RBC: alpha,beta   HSV: hue, sat/255, val/255   NOISE: sigma   ERASE: colour   ISO: hue sigma, intensity * 255
                  // SHOT: scale, scale * 1e-6   MOTION: 1 / popcount(mask)
  float k[kBlurHalfMax];  // BLUR: k[0] centre weight, k[j] weight at +-j
  int32_t slot, _pad;
  int64_t field;
};

struct DetCardX {
  double Minv[9];       // cv::invert(M): destination -> card coordinates (cv::warpPerspective)
  int32_t x0, y0, x1, y1;  // destination pixels that can receive a non-zero mask tap
  int32_t card, n_ops;
  DetPhotoX ops[MTGV_DET_MAX_CARD_OPS];
};

struct DetParams {
  int32_t status, bg, n_placed, n_prog;
  int32_t size_h, size_w, n_blur, _pad;
  double bg_Minv[9];
  uint64_t seed;
  DetPhotoX prog[kDetProgMax];
  DetCardX cards[MTGV_DET_MAX_CARDS];  // composite order = reverse placement order (od_datasets.py:594)
};

struct DetKeypoints {  // make_card_with_mask keypoints for the pool's card size
  int n_poly, n_pts;
  double pts[MTGV_DET_MAX_KPOLY][MTGV_DET_MAX_KP][2];
  double bbox[4][2];
};

MTGV_HD void det_box(double lft, double top, double rht, double bot, double margin, double mtr, double mbr, double out[4][2]) {
  // _box (od_datasets.py:236-242) with mlr = mrr = 1
  out[0][0] = MTGV_DADD(lft, MTGV_DMUL(margin, 1.0)); out[0][1] = MTGV_DADD(top, MTGV_DMUL(margin, mtr));
  out[1][0] = MTGV_DSUB(rht, MTGV_DMUL(margin, 1.0)); out[1][1] = MTGV_DADD(top, MTGV_DMUL(margin, mtr));
  out[2][0] = MTGV_DSUB(rht, MTGV_DMUL(margin, 1.0)); out[2][1] = MTGV_DSUB(bot, MTGV_DMUL(margin, mbr));
  out[3][0] = MTGV_DADD(lft, MTGV_DMUL(margin, 1.0)); out[3][1] = MTGV_DSUB(bot, MTGV_DMUL(margin, mbr));
}

MTGV_HDN void det_keypoints(int h, int w, int kind, DetKeypoints* kp) {
  const double W = (double)w, H = (double)h;
  det_box(0.0, 0.0, W, H, 0.0, 1.0, 1.0, kp->bbox);
  if (kind == 0) {  // obb: card, top, bottom boxes
    kp->n_poly = 3;
    kp->n_pts = 4;
    const double r = 0.5, m = MTGV_DMUL(0.03, (double)(w > h ? w : h));
    det_box(0.0, 0.0, W, H, 0.0, 1.0, 1.0, kp->pts[0]);
    double t[4][2];
    det_box(0.0, 0.0, W, MTGV_DMUL(r, H), m, 1.0, 0.5, t);
    for (int k = 0; k < 4; k++) { kp->pts[1][k][0] = t[k][0]; kp->pts[1][k][1] = t[k][1]; }
    det_box(0.0, MTGV_DMUL(MTGV_DSUB(1.0, r), H), W, H, m, 0.5, 1.0, t);
    for (int k = 0; k < 4; k++) { kp->pts[2][k][0] = t[k][0]; kp->pts[2][k][1] = t[k][1]; }
  } else {  // seg: card box minus the bottom indent [0.4w,0.6w] x [0.5h,1.1h]; vertex order: ours (unpinned)
    kp->n_poly = 1;
    kp->n_pts = 8;
    const double x0 = MTGV_DMUL(W, 0.4), x1 = MTGV_DMUL(W, 0.6), y0 = MTGV_DMUL(H, 0.5);
    const double v[8][2] = {{0, 0}, {W, 0}, {W, H}, {x1, H}, {x1, y0}, {x0, y0}, {x0, H}, {0, H}};
    for (int k = 0; k < 8; k++) { kp->pts[0][k][0] = v[k][0]; kp->pts[0][k][1] = v[k][1]; }
  }
}

// apply_transform_2d (od_datasets.py:64-70): [x,y,1] @ M.T then divide.  numpy's matmul is a
// left-to-right FMA chain on x86 hosts with FMA.
MTGV_HD void det_apply(const double* M, double x, double y, double* ox, double* oy) {
  double a = MTGV_DADD(MTGV_DFMA(y, M[1], MTGV_DMUL(x, M[0])), M[2]);
  double b = MTGV_DADD(MTGV_DFMA(y, M[4], MTGV_DMUL(x, M[3])), M[5]);
  double c = MTGV_DADD(MTGV_DFMA(y, M[7], MTGV_DMUL(x, M[6])), M[8]);
  *ox = MTGV_DDIV(a, c);
  *oy = MTGV_DDIV(b, c);
}

MTGV_HD bool poly_inside_convex(double px, double py, const double* poly, int n) {
  double orient = poly_signed2(poly, n) >= 0.0 ? 1.0 : -1.0;
  for (int e = 0; e < n; e++) {
    int e2 = e + 1 == n ? 0 : e + 1;
    double ex = poly[2 * e], ey = poly[2 * e + 1];
    double dx = MTGV_DSUB(poly[2 * e2], ex), dy = MTGV_DSUB(poly[2 * e2 + 1], ey);
    if (MTGV_DMUL(orient, MTGV_DSUB(MTGV_DMUL(dx, MTGV_DSUB(py, ey)), MTGV_DMUL(dy, MTGV_DSUB(px, ex)))) < 0.0) return false;
  }
  return true;
}

// area(shape n conv): shape = quad (obb) or quad minus indent (seg); conv == nullptr: whole plane
MTGV_HDN double det_shape_area(const double* quad, const double* indent, const double* conv, int nconv) {
  double q[2 * kPolyMax];
  int nq;
  if (conv) nq = clip_convex(quad, 4, conv, nconv, q);
  else { nq = 4; for (int k = 0; k < 8; k++) q[k] = quad[k]; }
  double a = poly_area(q, nq);
  if (indent && nq >= 3) {
    double t[2 * kPolyMax];
    int nt = clip_convex(q, nq, indent, 4, t);
    a = MTGV_DSUB(a, poly_area(t, nt));
  }
  return a;
}

// The accept/reject tests of od_datasets.py:353-372 (shapely restated, see oracle/det_oracle.py).
MTGV_HDN bool det_visible(const double* kp0, int kind, int size_h, int size_w, const double* existing, int n_existing,
                         double min_visible, double min_visible_edge, bool no_contains) {
  double quad[8], indent_buf[8];
  const double* indent = nullptr;
  if (kind == 0) {
    for (int k = 0; k < 8; k++) quad[k] = kp0[k];
  } else {
    const int qi[4] = {0, 1, 2, 7}, ii[4] = {5, 4, 3, 6};
    for (int k = 0; k < 4; k++) {
      quad[2 * k] = kp0[2 * qi[k]]; quad[2 * k + 1] = kp0[2 * qi[k] + 1];
      indent_buf[2 * k] = kp0[2 * ii[k]]; indent_buf[2 * k + 1] = kp0[2 * ii[k] + 1];
    }
    indent = indent_buf;
  }
  const double bw = (double)size_w, bh = (double)size_h;
  const double img[8] = {0.0, 0.0, bw, 0.0, bw, bh, 0.0, bh};
  const double card_area = det_shape_area(quad, indent, nullptr, 0);
  const double vis_area = det_shape_area(quad, indent, img, 4);
  if (MTGV_DDIV(vis_area, card_area) < min_visible_edge) return false;
  bool visible = true;
  // loop invariants: the candidate clipped to the frame, and its bounding box
  double vis_pts[2 * kPolyMax];
  const int nv = clip_convex(quad, 4, img, 4, vis_pts);
  double qx0 = quad[0], qx1 = quad[0], qy0 = quad[1], qy1 = quad[1];
  for (int k = 1; k < 4; k++) {
    qx0 = fmin(qx0, quad[2 * k]); qx1 = fmax(qx1, quad[2 * k]);
    qy0 = fmin(qy0, quad[2 * k + 1]); qy1 = fmax(qy1, quad[2 * k + 1]);
  }
  for (int e = 0; e < n_existing; e++) {
    const double* pq = existing + 8 * e;
    {
      // Strictly separated bounding boxes: the polygons are disjoint, so the intersection is empty (inter = 0: both
      // area tests reduce to ones that already passed) and neither contains the other - nothing can reject here.
      double px0 = pq[0], px1 = pq[0], py0 = pq[1], py1 = pq[1];
      for (int k = 1; k < 4; k++) {
        px0 = fmin(px0, pq[2 * k]); px1 = fmax(px1, pq[2 * k]);
        py0 = fmin(py0, pq[2 * k + 1]); py1 = fmax(py1, pq[2 * k + 1]);
      }
      if (px1 < qx0 || qx1 < px0 || py1 < qy0 || qy1 < py0) continue;
    }
    double pc[2 * kPolyMax];
    int npc = clip_convex(pq, 4, img, 4, pc);
    double inter = npc >= 3 ? det_shape_area(quad, indent, pc, npc) : 0.0;
    if (MTGV_DDIV(MTGV_DSUB(vis_area, inter), card_area) < min_visible) { visible = false; break; }
    double p_area = poly_area(pq, 4);
    if (MTGV_DDIV(MTGV_DSUB(p_area, inter), p_area) < min_visible) { visible = false; break; }
    bool p_contains_vis = nv >= 3;
    for (int k = 0; k < nv && p_contains_vis; k++) p_contains_vis = poly_inside_convex(vis_pts[2 * k], vis_pts[2 * k + 1], pq, 4);
    bool vis_contains_p = true;
    for (int k = 0; k < 4 && vis_contains_p; k++)
      vis_contains_p = poly_inside_convex(pq[2 * k], pq[2 * k + 1], img, 4) && poly_inside_convex(pq[2 * k], pq[2 * k + 1], quad, 4);
    if (vis_contains_p && indent) {
      double t[2 * kPolyMax];
      int nt = clip_convex(pq, 4, indent, 4, t);
      vis_contains_p = poly_area(t, nt) == 0.0;
    }
    if ((no_contains && p_contains_vis) || vis_contains_p) visible = false;  // precedence as written (:369)
  }
  return visible;
}

// float32 corner targets of one placement attempt (od_datasets.py:336-349) when the host did not
// supply them: corner_jitter_2d -> rotate_2d -> translate_2d with the platform's libm.
MTGV_HDN void det_attempt_dst(const mtgv_det_attempt* a, int ch, int cw, float* dst) {
  const double src[4][2] = {{0, 0}, {(double)cw, 0}, {(double)cw, (double)ch}, {0, (double)ch}};
  const double scale = a->area / ((double)ch * (double)cw);
  double cx = 0, cy = 0;
  for (int k = 0; k < 4; k++) { cx += src[k][0]; cy += src[k][1]; }
  cx /= 4.0; cy /= 4.0;
  double M[6];
  const double ang = a->deg * (3.141592653589793238462643383279502884 / 180.0);
  rotation_from_ab((double)(float)(cw / 2.0), (double)(float)(ch / 2.0), cos(ang) * scale, sin(ang) * scale, M);
  const double tx = (double)a->cx - (cw / 2.0) * scale, ty = (double)a->cy - (ch / 2.0) * scale;
  for (int k = 0; k < 4; k++) {
    const double dx = src[k][0] - cx, dy = src[k][1] - cy;
    const double d = sqrt(dx * dx + dy * dy) * a->jitter[k], th = atan2(dy, dx);
    const double jx = cx + d * cos(th), jy = cy + d * sin(th);
    const double rx = MTGV_DADD(MTGV_DFMA(jy, M[1], MTGV_DMUL(jx, M[0])), M[2]);
    const double ry = MTGV_DADD(MTGV_DFMA(jy, M[4], MTGV_DMUL(jx, M[3])), M[5]);
    dst[2 * k] = (float)(rx + tx);
    dst[2 * k + 1] = (float)(ry + ty);
  }
}

MTGV_HD void det_photo_clear(DetPhotoX* o) {
  o->code = MTGV_PH_NONE;
  for (int k = 0; k < 5; k++) o->i[k] = 0;
  for (int k = 0; k < 4; k++) o->f[k] = 0.f;
  for (int k = 0; k < kBlurHalfMax; k++) o->k[k] = 0.f;
  o->slot = 0; o->_pad = 0;
  o->field = MTGV_FIELD_PHILOX;
}

// tape op -> kernel-ready op; returns false for ops that expand to nothing
MTGV_HDN bool det_expand_photo(const mtgv_photo_op* t, int slot, DetPhotoX* o) {
  det_photo_clear(o);
  o->slot = slot;
  o->field = t->field;
  switch (t->code) {
    case MTGV_PH_RBC:
      o->code = MTGV_PH_RBC; o->f[0] = (float)t->d[0]; o->f[1] = (float)t->d[1];
      return true;
    case MTGV_PH_HSV:
      if (t->d[0] == 0.0 && t->d[1] == 0.0 && t->d[2] == 0.0) return false;
      o->code = MTGV_PH_HSV;
      o->f[0] = (float)t->d[0]; o->f[1] = (float)MTGV_DDIV(t->d[1], 255.0); o->f[2] = (float)MTGV_DDIV(t->d[2], 255.0);
      return true;
    case MTGV_PH_GAUSS_NOISE:
      o->code = MTGV_PH_GAUSS_NOISE; o->f[0] = (float)t->d[0];
      return true;
    case MTGV_PH_GAUSS_BLUR: {
      const double sigma = t->d[0];
      if (!(sigma > 0.0)) return false;
      int ksize = (int)MTGV_DADD(MTGV_DMUL(sigma, 6.0), 1.0) | 1;
      if (ksize < 3) ksize = 3;
      if (ksize > 2 * kBlurHalfMax - 1) ksize = 2 * kBlurHalfMax - 1;
      const int r = ksize / 2;
      double w[kBlurHalfMax], sum = 0.0;
      // numpy: k = exp(-(x*x) / (2 sigma^2)); k / k.sum() with x ascending from -r (pairwise-free: <= 19 terms)
      const double den = MTGV_DMUL(MTGV_DMUL(2.0, sigma), sigma);
      for (int j = 0; j <= r; j++) w[j] = exp(MTGV_DDIV(-(double)(j * j), den));
      for (int x = -r; x <= r; x++) sum = MTGV_DADD(sum, w[x < 0 ? -x : x]);
      o->code = MTGV_PH_GAUSS_BLUR;
      o->i[0] = ksize;
      for (int j = 0; j <= r; j++) o->k[j] = (float)MTGV_DDIV(w[j], sum);
      return true;
    }
    case MTGV_PH_ERASE:
      if (t->i[2] <= 0 || t->i[3] <= 0) return false;
      o->code = MTGV_PH_ERASE;
      for (int k = 0; k < 5; k++) o->i[k] = t->i[k];
      for (int k = 0; k < 3; k++) o->f[k] = t->i[4] == 2 ? 1.f : (t->i[4] == 3 ? 0.f : (float)t->d[k]);
      return true;
    case MTGV_PH_ISO_NOISE:
      o->code = MTGV_PH_ISO_NOISE;
      o->f[0] = (float)MTGV_DMUL(MTGV_DMUL(t->d[0], 360.0), t->d[1]);  // np.float32(color_shift * 360.0 * intensity)
      o->f[1] = (float)MTGV_DMUL(t->d[1], 255.0);                      // Poisson rate = std(L) * intensity * 255
      return true;
    case MTGV_PH_SHOT_NOISE:
      if (!(t->d[0] > 0.0)) return false;
      o->code = MTGV_PH_SHOT_NOISE;
      o->f[0] = (float)t->d[0];
      return true;
    case MTGV_PH_MEDIAN_BLUR:
      if (t->i[0] != 3 && t->i[0] != 5 && t->i[0] != 7) return false;
      o->code = MTGV_PH_MEDIAN_BLUR;
      o->i[0] = t->i[0];
      return true;
    case MTGV_PH_MOTION_BLUR: {
      const int ks = t->i[0];
      if (ks < 3 || ks > 11 || !(ks & 1)) return false;
      int cnt = 0;
      for (int b = 0; b < ks * ks; b++) cnt += (t->i[1 + (b >> 5)] >> (b & 31)) & 1;
      if (cnt == 0) return false;
      o->code = MTGV_PH_MOTION_BLUR;
      for (int k = 0; k < 5; k++) o->i[k] = t->i[k];
      o->f[0] = 1.f / (float)cnt;  // kernel.astype(float32) / float32(sum): every set cell holds this weight
      return true;
    }
    default:
      return false;
  }
}

struct DetPlaceState {  // accepted cards so far (shared memory per warp on the device, stack on the host)
  double collide[MTGV_DET_MAX_CARDS * 8];   // apply_transform_2d(bbox, M) of every accepted card (:585)
  double placedM[MTGV_DET_MAX_CARDS][9];
  int placed_card[MTGV_DET_MAX_CARDS], placed_src[MTGV_DET_MAX_CARDS];
  int n_placed;
};

MTGV_HD double det_min_edge(const mtgv_det_config* cfg) {
  double e = cfg->min_visible_edges < 0.0 ? cfg->min_visible : cfg->min_visible_edges;  // None -> min_visible (:312-314)
  return e < cfg->min_visible ? cfg->min_visible : e;
}

// make_background: get_rotate_over_output_transform(mode="cover") (:85-118), inverted like cv::warpPerspective does
MTGV_HDN void det_bg_transform(const mtgv_det_tape* t, const mtgv_det_config* cfg, const int32_t* bg_hw, double* Minv) {
  const int S_h = cfg->size_h, S_w = cfg->size_w;
  const int h = bg_hw[2 * t->bg], w = bg_hw[2 * t->bg + 1];
  const double oh = (double)S_h, ow = (double)S_w, mx = oh > ow ? oh : ow;
  const double a = MTGV_DDIV(oh, mx), b = MTGV_DDIV(ow, mx);
  const double hyp = sqrt(MTGV_DADD(MTGV_DMUL(a, a), MTGV_DMUL(b, b)));  // math.hypot
  const double scale = MTGV_DDIV(MTGV_DMUL(hyp, mx), (double)(h < w ? h : w));
  double alpha, beta;
  if (t->bg_ab_given) { alpha = t->bg_ab[0]; beta = t->bg_ab[1]; }
  else {
    const double ang = MTGV_DMUL((double)t->bg_deg, 3.141592653589793238462643383279502884 / 180.0);
    alpha = MTGV_DMUL(cos(ang), scale); beta = MTGV_DMUL(sin(ang), scale);
  }
  double R[6], M[9];
  rotation_from_ab((double)(w / 2), (double)(h / 2), alpha, beta, R);
  const int tx = (S_w - w) >= 0 ? (S_w - w) / 2 : -((w - S_w + 1) / 2);  // python floor division
  const int ty = (S_h - h) >= 0 ? (S_h - h) / 2 : -((h - S_h + 1) / 2);
  M[0] = R[0]; M[1] = R[1]; M[2] = MTGV_DADD(R[2], (double)tx);
  M[3] = R[3]; M[4] = R[4]; M[5] = MTGV_DADD(R[5], (double)ty);
  M[6] = 0.0; M[7] = 0.0; M[8] = 1.0;
  invert3x3(M, Minv);
}

// one placement attempt (:317-372): homography from the corner targets, warped test polygon, visibility tests
MTGV_HDN bool det_try_attempt(const mtgv_det_attempt* a, const mtgv_det_config* cfg, const DetKeypoints* kp, int card_h, int card_w,
                              const DetPlaceState* st, double min_edge, double* M) {
  const float src[8] = {0.f, 0.f, (float)card_w, 0.f, (float)card_w, (float)card_h, 0.f, (float)card_h};
  float dst[8];
  if (a->dst_given) for (int k = 0; k < 8; k++) dst[k] = a->dst[k];
  else det_attempt_dst(a, card_h, card_w, dst);
  if (!get_perspective_transform(src, dst, M)) return false;
  double kp0[2 * MTGV_DET_MAX_KP];
  for (int k = 0; k < kp->n_pts; k++) det_apply(M, kp->pts[0][k][0], kp->pts[0][k][1], &kp0[2 * k], &kp0[2 * k + 1]);
  return det_visible(kp0, cfg->kind, cfg->size_h, cfg->size_w, st->collide, st->n_placed, cfg->min_visible, min_edge,
                     cfg->no_contains != 0);
}

MTGV_HD void det_commit(DetPlaceState* st, const DetKeypoints* kp, const double* M, int card, int ci) {
  const int n = st->n_placed;
  for (int k = 0; k < 9; k++) st->placedM[n][k] = M[k];
  for (int k = 0; k < 4; k++) det_apply(M, kp->bbox[k][0], kp->bbox[k][1], &st->collide[8 * n + 2 * k], &st->collide[8 * n + 2 * k + 1]);
  st->placed_card[n] = card;
  st->placed_src[n] = ci;
  st->n_placed = n + 1;
}

// r-th card of the output / composite order = placement order reversed (:594-601): labels + composite record
MTGV_HDN void det_emit_card(const mtgv_det_tape* t, const mtgv_det_config* cfg, const DetKeypoints* kp, const DetPlaceState* st, int r,
                            DetParams* P, double* keypoints, int32_t* labels) {
  const int pi = st->n_placed - 1 - r;
  const double* M = st->placedM[pi];
  for (int q = 0; q < kp->n_poly; q++) {
    const int out = r * kp->n_poly + q;
    double* dstp = keypoints + ((size_t)out * MTGV_DET_MAX_KP) * 2;
    for (int k = 0; k < kp->n_pts; k++) det_apply(M, kp->pts[q][k][0], kp->pts[q][k][1], &dstp[2 * k], &dstp[2 * k + 1]);
    for (int k = kp->n_pts; k < MTGV_DET_MAX_KP; k++) { dstp[2 * k] = 0.0; dstp[2 * k + 1] = 0.0; }
    labels[out] = q;
  }
  DetCardX* cx = &P->cards[r];
  invert3x3(M, cx->Minv);
  double minx = 1e300, maxx = -1e300, miny = 1e300, maxy = -1e300;
  for (int k = 0; k < 4; k++) {
    const double x = st->collide[8 * pi + 2 * k], y = st->collide[8 * pi + 2 * k + 1];
    minx = fmin(minx, x); maxx = fmax(maxx, x); miny = fmin(miny, y); maxy = fmax(maxy, y);
  }
  // bilinear taps reach one source pixel beyond the card edge: 2 px of slack in destination space
  minx = fmax(minx - 2.0, 0.0); miny = fmax(miny - 2.0, 0.0);
  maxx = fmin(maxx + 3.0, (double)cfg->size_w); maxy = fmin(maxy + 3.0, (double)cfg->size_h);
  cx->x0 = (int)floor(minx); cx->y0 = (int)floor(miny);
  cx->x1 = maxx > minx ? (int)ceil(maxx) : cx->x0; cx->y1 = maxy > miny ? (int)ceil(maxy) : cx->y0;
  cx->card = st->placed_card[pi];
  cx->n_ops = 0;
  const mtgv_det_card* c = &t->cards[st->placed_src[pi]];
  for (int k = 0; k < c->n_photo && k < MTGV_DET_MAX_CARD_OPS; k++)
    if (cfg->photometrics && det_expand_photo(&c->photo[k], 32 + 4 * st->placed_src[pi] + k, &cx->ops[cx->n_ops])) cx->n_ops++;
}

// one tape op -> 0, 1 or (GlassBlur) up to 4 program entries at prog[n...]; returns the new length or -1 when the program is full
MTGV_HDN int det_append_photo(const mtgv_photo_op* t, int slot, DetPhotoX* prog, int n) {
  if (t->code == MTGV_PH_GLASS_BLUR) {
    const int md = t->i[0], rounds = t->i[1];
    if (!(t->d[0] > 0.0) || md < 1 || md > 8 || rounds < 0 || rounds > 2) return n;
    if (n + 2 + rounds > kDetProgMax) return -1;
    mtgv_photo_op b = *t;
    b.code = MTGV_PH_GAUSS_BLUR;
    b.field = MTGV_FIELD_PHILOX;
    if (!det_expand_photo(&b, slot, &prog[n])) return n;
    n++;
    for (int r = 0; r < rounds; r++) {
      det_photo_clear(&prog[n]);
      prog[n].code = kPhGlassSwap;
      prog[n].i[0] = md; prog[n].i[1] = r; prog[n].i[2] = rounds;
      prog[n].slot = slot;
      prog[n].field = t->field;
      n++;
    }
    prog[n] = prog[n - 1 - rounds];  // the closing blur: same kernel
    return n + 1;
  }
  if (n >= kDetProgMax) return -1;
  return det_expand_photo(t, slot, &prog[n]) ? n + 1 : n;
}

// scene program: pre ops, cards, post ops
MTGV_HDN int det_emit_program(const mtgv_det_tape* t, const mtgv_det_config* cfg, DetParams* P) {
  int n = 0;
  if (cfg->photometrics)
    for (int k = 0; k < t->n_pre && k < MTGV_DET_MAX_PRE && n >= 0; k++) n = det_append_photo(&t->pre[k], k, P->prog, n);
  if (n < 0 || n >= kDetProgMax) return P->status = MTGV_ERR_LIMIT;
  det_photo_clear(&P->prog[n]);
  P->prog[n].code = kPhCards;
  n++;
  if (cfg->photometrics)
    for (int k = 0; k < t->n_post && k < MTGV_DET_MAX_POST && n >= 0; k++) n = det_append_photo(&t->post[k], 8 + k, P->prog, n);
  if (n < 0) return P->status = MTGV_ERR_LIMIT;
  P->n_prog = n;
  int nb = 0;
  for (int k = 0; k < n; k++) nb += det_is_boundary(P->prog[k].code);
  P->n_blur = nb;
  if (nb > kDetMaxBlur) return P->status = MTGV_ERR_LIMIT;
  return 0;
}

MTGV_HD bool det_scene_header(const mtgv_det_tape* t, const mtgv_det_config* cfg, int n_bgs, DetParams* P) {
  P->status = 0; P->bg = t->bg; P->n_placed = 0; P->n_prog = 0; P->size_h = cfg->size_h; P->size_w = cfg->size_w; P->n_blur = 0;
  P->_pad = 0; P->seed = t->seed;
  if (t->bg < 0 || t->bg >= n_bgs || t->n_cards < 0 || t->n_cards > MTGV_DET_MAX_CARDS) { P->status = MTGV_ERR_INVALID; return false; }
  return true;
}

// One scene, sequentially (the specification; the device kernel k_det_place evaluates the attempts of a
// card in parallel with the same routines and picks the first accepted one).
MTGV_HD int det_place_scene(const mtgv_det_tape* t, const mtgv_det_config* cfg, const DetKeypoints* kp, int card_h, int card_w,
                            int n_cards_pool, int n_bgs, const int32_t* bg_hw, DetParams* P, int32_t* accepted,
                            double* keypoints, int32_t* labels, int32_t* count) {
  for (int k = 0; k < MTGV_DET_MAX_CARDS; k++) accepted[k] = -1;
  for (int k = 0; k < MTGV_DET_MAX_CARDS * MTGV_DET_MAX_KPOLY; k++) labels[k] = -1;
  *count = 0;
  if (!det_scene_header(t, cfg, n_bgs, P)) return P->status;
  det_bg_transform(t, cfg, bg_hw, P->bg_Minv);
  const double min_edge = det_min_edge(cfg);
  DetPlaceState st;
  st.n_placed = 0;
  if (!t->bg_only) {
    for (int ci = 0; ci < t->n_cards; ci++) {
      const mtgv_det_card* c = &t->cards[ci];
      if (c->card < 0 || c->card >= n_cards_pool) return P->status = MTGV_ERR_INVALID;
      const int na = c->n_attempts < cfg->max_attempts ? c->n_attempts : cfg->max_attempts;
      for (int ai = 0; ai < na && ai < MTGV_DET_MAX_ATTEMPTS; ai++) {
        double M[9];
        if (!det_try_attempt(&c->att[ai], cfg, kp, card_h, card_w, &st, min_edge, M)) continue;
        accepted[ci] = ai;
        det_commit(&st, kp, M, c->card, ci);
        break;
      }
    }
  }
  P->n_placed = st.n_placed;
  for (int r = 0; r < st.n_placed; r++) det_emit_card(t, cfg, kp, &st, r, P, keypoints, labels);
  *count = st.n_placed * kp->n_poly;
  return det_emit_program(t, cfg, P);
}

}  // namespace mtgv
