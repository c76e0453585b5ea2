// mtgv_geom.cuh - fp64 parameter-expansion arithmetic shared by device kernels and the
// host-side unit-test harness (tests/host_harness.cpp).
//
// Everything here decides integer outputs (pixel coordinates on the 1/32 grid, canvas
// sizes, accept/reject) and therefore has to reproduce the reference's arithmetic bit
// for bit: OpenCV's getPerspectiveTransform / invert / getRotationMatrix2D / warp*
// coordinate generation and the numpy expressions around
// them.  Rules: every fp64 multiply/add is a separate rounding (translation units are
// compiled with -fmad=false and the critical chains use explicit _rn intrinsics); float32
// products that the reference forms in float32 use __fmul_rn.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MTGV_HD __host__ __device__ __forceinline__
#define MTGV_HDN __host__ __device__ __noinline__  /* big, cold routines: keep one copy */
#else
#define MTGV_HD inline
#define MTGV_HDN inline
#endif

#if defined(__CUDA_ARCH__)
#define MTGV_DMUL(a, b) __dmul_rn((a), (b))
#define MTGV_DADD(a, b) __dadd_rn((a), (b))
#define MTGV_DSUB(a, b) __dsub_rn((a), (b))
#define MTGV_DDIV(a, b) __ddiv_rn((a), (b))
#define MTGV_FMUL(a, b) __fmul_rn((a), (b))
#define MTGV_FADD(a, b) __fadd_rn((a), (b))
#define MTGV_FSUB(a, b) __fsub_rn((a), (b))
#define MTGV_DFMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define MTGV_DMUL(a, b) ((a) * (b))
#define MTGV_DADD(a, b) ((a) + (b))
#define MTGV_DSUB(a, b) ((a) - (b))
#define MTGV_DDIV(a, b) ((a) / (b))
#define MTGV_FMUL(a, b) ((float)(a) * (float)(b))
#define MTGV_FADD(a, b) ((float)(a) + (float)(b))
#define MTGV_FSUB(a, b) ((float)(a) - (float)(b))
#define MTGV_DFMA(a, b, c) fma((a), (b), (c))
#endif

namespace mtgv {

constexpr int kInterBits = 5;
constexpr int kInterTab = 32;
constexpr int kAbBits = 10;
constexpr int kAbScale = 1024;
constexpr int kAreaMaxTaps = 8;  // register-resident tap lists (area_taps, templated reducers): INTER_AREA scale factors up to 6
constexpr int kAreaCompactMaxTaps = 250;  // compact entries (start, count, wl, wm, wr) carry any count that fits their 8-bit field
constexpr int kBgMaxAreaScale = 24;       // largest INTER_AREA reduction of the background chain (k_background's generic reducer)

// ---------------------------------------------------------------------------------------
// cv::hal::LU64f + back substitution, n = 8, one right-hand side
// (cv::getPerspectiveTransform -> cv::solve(DECOMP_LU))
// ---------------------------------------------------------------------------------------
MTGV_HD bool lu_solve8(double A[8][8], double b[8]) {
  const double eps = 2.220446049250313e-16 * 100;
  for (int i = 0; i < 8; i++) {
    int k = i;
    for (int j = i + 1; j < 8; j++)
      if (fabs(A[j][i]) > fabs(A[k][i])) k = j;
    if (fabs(A[k][i]) < eps) return false;
    if (k != i) {
      for (int j = i; j < 8; j++) {
        double t = A[i][j];
        A[i][j] = A[k][j];
        A[k][j] = t;
      }
      double t = b[i];
      b[i] = b[k];
      b[k] = t;
    }
    double d = MTGV_DDIV(-1.0, A[i][i]);
    for (int j = i + 1; j < 8; j++) {
      double alpha = MTGV_DMUL(A[j][i], d);
      for (int c = i + 1; c < 8; c++) A[j][c] = MTGV_DADD(A[j][c], MTGV_DMUL(alpha, A[i][c]));
      b[j] = MTGV_DADD(b[j], MTGV_DMUL(alpha, b[i]));
    }
  }
  for (int i = 7; i >= 0; i--) {
    double s = b[i];
    for (int k = i + 1; k < 8; k++) s = MTGV_DSUB(s, MTGV_DMUL(A[i][k], b[k]));
    b[i] = MTGV_DDIV(s, A[i][i]);
  }
  return true;
}

// cv::getPerspectiveTransform(src f32[4][2], dst f32[4][2]) -> M[9] (row major).
MTGV_HD bool get_perspective_transform(const float* src, const float* dst, double* M) {
  double A[8][8];
  double b[8];
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 8; j++) A[i][j] = 0.0;
  for (int i = 0; i < 4; i++) {
    float x = src[2 * i], y = src[2 * i + 1];
    float u = dst[2 * i], v = dst[2 * i + 1];
    A[i][0] = A[i + 4][3] = x;
    A[i][1] = A[i + 4][4] = y;
    A[i][2] = A[i + 4][5] = 1.0;
    A[i][6] = (double)MTGV_FMUL(-x, u);  // products formed in float32 (Point2f arithmetic)
    A[i][7] = (double)MTGV_FMUL(-y, u);
    A[i + 4][6] = (double)MTGV_FMUL(-x, v);
    A[i + 4][7] = (double)MTGV_FMUL(-y, v);
    b[i] = u;
    b[i + 4] = v;
  }
  if (!lu_solve8(A, b)) {
    for (int i = 0; i < 9; i++) M[i] = 0.0;
    return false;
  }
  for (int i = 0; i < 8; i++) M[i] = b[i];
  M[8] = 1.0;
  return true;
}

// cv::invert on a 3x3 CV_64F (cofactors times 1/det); SURVEY 8a-note 1.
MTGV_HD bool invert3x3(const double* m, double* t) {
  double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  double ei_fh = MTGV_DSUB(MTGV_DMUL(e, i), MTGV_DMUL(f, h));
  double di_fg = MTGV_DSUB(MTGV_DMUL(d, i), MTGV_DMUL(f, g));
  double dh_eg = MTGV_DSUB(MTGV_DMUL(d, h), MTGV_DMUL(e, g));
  double det = MTGV_DADD(MTGV_DSUB(MTGV_DMUL(a, ei_fh), MTGV_DMUL(b, di_fg)), MTGV_DMUL(c, dh_eg));
  if (det == 0.0) {
    for (int k = 0; k < 9; k++) t[k] = 0.0;
    return false;
  }
  double s = MTGV_DDIV(1.0, det);
  t[0] = MTGV_DMUL(ei_fh, s);
  t[1] = MTGV_DMUL(MTGV_DSUB(MTGV_DMUL(c, h), MTGV_DMUL(b, i)), s);
  t[2] = MTGV_DMUL(MTGV_DSUB(MTGV_DMUL(b, f), MTGV_DMUL(c, e)), s);
  t[3] = MTGV_DMUL(MTGV_DSUB(MTGV_DMUL(f, g), MTGV_DMUL(d, i)), s);
  t[4] = MTGV_DMUL(MTGV_DSUB(MTGV_DMUL(a, i), MTGV_DMUL(c, g)), s);
  t[5] = MTGV_DMUL(MTGV_DSUB(MTGV_DMUL(c, d), MTGV_DMUL(a, f)), s);
  t[6] = MTGV_DMUL(dh_eg, s);
  t[7] = MTGV_DMUL(MTGV_DSUB(MTGV_DMUL(b, g), MTGV_DMUL(a, h)), s);
  t[8] = MTGV_DMUL(MTGV_DSUB(MTGV_DMUL(a, e), MTGV_DMUL(b, d)), s);
  return true;
}

// The affine inversion inlined in cv::warpAffine; SURVEY 8a-note 2.
MTGV_HD void invert_affine(const double* A, double* M) {
  for (int k = 0; k < 6; k++) M[k] = A[k];
  double D = MTGV_DSUB(MTGV_DMUL(M[0], M[4]), MTGV_DMUL(M[1], M[3]));
  D = D != 0.0 ? MTGV_DDIV(1.0, D) : 0.0;
  double A11 = MTGV_DMUL(M[4], D), A22 = MTGV_DMUL(M[0], D);
  M[0] = A11;
  M[1] = MTGV_DMUL(M[1], -D);
  M[3] = MTGV_DMUL(M[3], -D);
  M[4] = A22;
  double b1 = MTGV_DSUB(MTGV_DMUL(-M[0], M[2]), MTGV_DMUL(M[1], M[5]));
  double b2 = MTGV_DSUB(MTGV_DMUL(-M[3], M[2]), MTGV_DMUL(M[4], M[5]));
  M[2] = b1;
  M[5] = b2;
}

// cv::getRotationMatrix2D after cos/sin: centre already rounded to float32; 8a-note 4.
MTGV_HD void rotation_from_ab(double cx, double cy, double alpha, double beta, double* M) {
  M[0] = alpha;
  M[1] = beta;
  M[2] = MTGV_DSUB(MTGV_DMUL(MTGV_DSUB(1.0, alpha), cx), MTGV_DMUL(beta, cy));
  M[3] = -beta;
  M[4] = alpha;
  M[5] = MTGV_DADD(MTGV_DMUL(beta, cx), MTGV_DMUL(MTGV_DSUB(1.0, alpha), cy));
}

// Column-block width of WarpPerspectiveInvoker (BLOCK_SZ = 32): X0/Y0/W0 are evaluated
// at the block origin and advanced per column inside the block.
MTGV_HD int persp_block_w(int dst_h, int dst_w) {
  int bh0 = dst_h < 16 ? dst_h : 16;
  int bw0 = 1024 / bh0;
  return bw0 < dst_w ? bw0 : dst_w;
}

// One destination pixel of cv::warpPerspective's coordinate generation -> (X, Y) in 1/32 px.
MTGV_HD void persp_coord(const double* M, int x, int y, int bw0, int* X, int* Y) {
  int bxi = (x / bw0) * bw0;
  double bx = (double)bxi, x1 = (double)(x - bxi), yy = (double)y;
  double X0 = MTGV_DADD(MTGV_DADD(MTGV_DMUL(M[0], bx), MTGV_DMUL(M[1], yy)), M[2]);
  double Y0 = MTGV_DADD(MTGV_DADD(MTGV_DMUL(M[3], bx), MTGV_DMUL(M[4], yy)), M[5]);
  double W0 = MTGV_DADD(MTGV_DADD(MTGV_DMUL(M[6], bx), MTGV_DMUL(M[7], yy)), M[8]);
  double W = MTGV_DADD(W0, MTGV_DMUL(M[6], x1));
  W = W != 0.0 ? MTGV_DDIV(32.0, W) : 0.0;
  double fX = MTGV_DMUL(MTGV_DADD(X0, MTGV_DMUL(M[0], x1)), W);
  double fY = MTGV_DMUL(MTGV_DADD(Y0, MTGV_DMUL(M[3], x1)), W);
  fX = fmax(-2147483648.0, fmin(2147483647.0, fX));
  fY = fmax(-2147483648.0, fmin(2147483647.0, fY));
  *X = (int)rint(fX);
  *Y = (int)rint(fY);
}

MTGV_HD int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

// cv::warpAffine fixed-point tables: per-column deltas and per-row origins.
MTGV_HD int affine_col_delta(double m, int x) { return (int)rint(MTGV_DMUL(MTGV_DMUL(m, (double)x), 1024.0)); }
MTGV_HD int affine_row_origin(double m1, double m2, int y) {
  return (int)rint(MTGV_DMUL(MTGV_DADD(MTGV_DMUL(m1, (double)y), m2), 1024.0)) + 16;
}

// One destination index of computeResizeAreaTab (cv::resize INTER_AREA, scale >= 1).
// Returns the number of taps; taps are consecutive source indices from *start.
MTGV_HD int area_taps(int ssize, int dsize, int d, int* start, float* w) {
  double scale = MTGV_DDIV(1.0, MTGV_DDIV((double)dsize, (double)ssize));  // cv::resize: 1./inv_scale
  double fsx1 = MTGV_DMUL((double)d, scale);
  double fsx2 = MTGV_DADD(fsx1, scale);
  double cell = fmin(scale, MTGV_DSUB((double)ssize, fsx1));
  int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
  sx2 = sx2 < ssize - 1 ? sx2 : ssize - 1;
  sx1 = sx1 < sx2 ? sx1 : sx2;
  int n = 0;
  *start = sx1;
  if (MTGV_DSUB((double)sx1, fsx1) > 1e-3) {
    *start = sx1 - 1;
    w[n++] = (float)MTGV_DDIV(MTGV_DSUB((double)sx1, fsx1), cell);
  }
  for (int sx = sx1; sx < sx2 && n < kAreaMaxTaps; sx++) w[n++] = (float)MTGV_DDIV(1.0, cell);
  if (MTGV_DSUB(fsx2, (double)sx2) > 1e-3 && n < kAreaMaxTaps) {
    double t = fmin(fmin(MTGV_DSUB(fsx2, (double)sx2), 1.0), cell);
    w[n++] = (float)MTGV_DDIV(t, cell);
  }
  return n;
}

// The same table entry in compact form: taps are [left partial?] full* [right partial?], every
// full tap weighs 1/cell.  *n = taps | left << 8 | right << 9; weights as area_taps emits them.
MTGV_HD void area_compact(int ssize, int dsize, int d, int* start, int* n, float* wl, float* wm, float* wr) {
  double scale = MTGV_DDIV(1.0, MTGV_DDIV((double)dsize, (double)ssize));
  double fsx1 = MTGV_DMUL((double)d, scale);
  double fsx2 = MTGV_DADD(fsx1, scale);
  double cell = fmin(scale, MTGV_DSUB((double)ssize, fsx1));
  int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
  sx2 = sx2 < ssize - 1 ? sx2 : ssize - 1;
  sx1 = sx1 < sx2 ? sx1 : sx2;
  int cnt = 0, flags = 0;
  *start = sx1;
  *wl = *wr = 0.f;
  if (MTGV_DSUB((double)sx1, fsx1) > 1e-3) {
    *start = sx1 - 1;
    *wl = (float)MTGV_DDIV(MTGV_DSUB((double)sx1, fsx1), cell);
    flags |= 256;
    cnt++;
  }
  *wm = (float)MTGV_DDIV(1.0, cell);
  int mid = sx2 - sx1;
  if (mid > kAreaCompactMaxTaps - cnt) mid = kAreaCompactMaxTaps - cnt;
  if (mid > 0) cnt += mid;
  if (MTGV_DSUB(fsx2, (double)sx2) > 1e-3 && cnt < kAreaCompactMaxTaps) {
    double t = fmin(fmin(MTGV_DSUB(fsx2, (double)sx2), 1.0), cell);
    *wr = (float)MTGV_DDIV(t, cell);
    flags |= 512;
    cnt++;
  }
  *n = cnt | flags;
}

// cv::resize(INTER_AREA) when either axis is enlarged: BOTH axes fall back to the 2-tap linear kernel with the
// "area" coefficient rule (resize.cpp: area_mode): sx = floor(dx*scale), fx = (dx+1) - (sx+1)*inv_scale, clamped to
// [0,1) by "fx <= 0 ? 0 : fx - floor(fx)"; at the last source sample the single tap S[ssize-1] is used.
// Same compact form as area_compact: taps S[start] * wl (+ S[start+1] * wr).
MTGV_HD void area_linear_compact(int ssize, int dsize, int d, int* start, int* n, float* wl, float* wm, float* wr) {
  double inv_scale = MTGV_DDIV((double)dsize, (double)ssize);
  double scale = MTGV_DDIV(1.0, inv_scale);
  int sx = (int)floor(MTGV_DMUL((double)d, scale));
  float fx = (float)MTGV_DSUB((double)(d + 1), MTGV_DMUL((double)(sx + 1), inv_scale));
  fx = fx <= 0.f ? 0.f : fx - floorf(fx);
  if (sx < 0) { sx = 0; fx = 0.f; }
  *wm = 0.f;
  if (sx >= ssize - 1) {
    *start = ssize - 1;
    *wl = 1.f; *wr = 0.f;
    *n = 1 | 256;
    return;
  }
  *start = sx;
  *wl = 1.f - fx;
  *wr = fx;
  *n = 2 | 256 | 512;
}

// crop_to_size integer geometry (mtgvision/util/image.py:359-376).
MTGV_HD void crop_geometry(int ih, int iw, int sh, int sw, bool pad, int* rh, int* rw, int* y0, int* x0) {
  double fh = MTGV_DDIV((double)ih, (double)sh), fw = MTGV_DDIV((double)iw, (double)sw);
  double r = pad ? fmax(fh, fw) : fmin(fh, fw);
  *rh = (int)MTGV_DDIV((double)ih, r);
  *rw = (int)MTGV_DDIV((double)iw, r);
  if (pad) {
    *y0 = (sh - *rh) / 2;  // python // on non-negative operands
    *x0 = (sw - *rw) / 2;
  } else {
    *y0 = (*rh - sh) / 2;
    *x0 = (*rw - sw) / 2;
  }
}

// Mutate.warp control points (mtgvision/encoder_datasets.py:94-107); h = H-1, w = W-1.
MTGV_HD void mutate_warp_points(int H, int W, const double* u, double ratio, double ratio_min, float* src, float* dst) {
  float h = (float)(H - 1), w = (float)(W - 1);
  const float sgn[8] = {h, w, h, -w, -h, w, -h, -w};
  src[0] = 0.f; src[1] = 0.f; src[2] = 0.f; src[3] = w; src[4] = h; src[5] = 0.f; src[6] = h; src[7] = w;
  double span = MTGV_DMUL(fabs(MTGV_DSUB(ratio, ratio_min)), 0.5);
  for (int k = 0; k < 8; k++) {
    double ran = MTGV_DADD(ratio_min, MTGV_DMUL(u[k], span));
    dst[k] = (float)MTGV_DADD(MTGV_DMUL(ran, (double)sgn[k]), (double)src[k]);
  }
}

// Mutate.perspective_transform control points (encoder_datasets.py:380-401).
MTGV_HD void mutate_perspective_points(int rows, int cols, const double* u, float* src, float* dst) {
  double c = (double)cols, r = (double)rows;
  src[0] = 0.f; src[1] = 0.f; src[2] = (float)cols; src[3] = 0.f;
  src[4] = 0.f; src[5] = (float)rows; src[6] = (float)cols; src[7] = (float)rows;
  dst[0] = (float)MTGV_DMUL(u[0], c);
  dst[1] = (float)MTGV_DMUL(u[1], r);
  dst[2] = (float)MTGV_DADD(c, MTGV_DMUL(u[2], c));
  dst[3] = (float)MTGV_DMUL(u[3], r);
  dst[4] = (float)MTGV_DMUL(u[4], c);
  dst[5] = (float)MTGV_DADD(r, MTGV_DMUL(u[5], r));
  dst[6] = (float)MTGV_DADD(c, MTGV_DMUL(u[6], c));
  dst[7] = (float)MTGV_DADD(r, MTGV_DMUL(u[7], r));
}

// Mutate.affine_transform matrix (encoder_datasets.py:366-374): Shear @ [R; 0 0 1] then
// translation.  numpy's 3x3 float64 matmul goes through BLAS dgemm whose k-loop is an FMA
// chain on every x86 host with FMA: row 0 = fma(shear, R1j, R0j).
MTGV_HD void mutate_affine_matrix(int rows, int cols, double alpha, double beta, double tx, double ty, double shear,
                                  double* M) {
  double R[6];
  rotation_from_ab((double)(float)(cols / 2.0), (double)(float)(rows / 2.0), alpha, beta, R);
  M[0] = MTGV_DFMA(shear, R[3], R[0]);
  M[1] = MTGV_DFMA(shear, R[4], R[1]);
  M[2] = MTGV_DADD(MTGV_DFMA(shear, R[5], R[2]), tx);
  M[3] = R[3];
  M[4] = R[4];
  M[5] = MTGV_DADD(R[5], ty);
}

// uimg.rotate_bounded (mtgvision/util/image.py:380-398): canvas size and shifted matrix.
MTGV_HD void rotate_bounded_matrix(int h, int w, double alpha, double beta, double* M, int* nh, int* nw) {
  int cy = h / 2, cx = w / 2;
  rotation_from_ab((double)cx, (double)cy, alpha, beta, M);
  double c = fabs(M[0]), s = fabs(M[1]);
  *nw = (int)MTGV_DADD(MTGV_DMUL((double)h, s), MTGV_DMUL((double)w, c));
  *nh = (int)MTGV_DADD(MTGV_DMUL((double)h, c), MTGV_DMUL((double)w, s));
  M[2] = MTGV_DADD(M[2], MTGV_DSUB(MTGV_DDIV((double)*nw, 2.0), (double)cx));
  M[5] = MTGV_DADD(M[5], MTGV_DSUB(MTGV_DDIV((double)*nh, 2.0), (double)cy));
}

}  // namespace mtgv
