// mtgv_persp.cuh - cv::warpPerspective coordinate generation on the device: exact restatement
//  plus a guarded fast path that provably returns the same integers.
#pragma once

#include "mtgv_geom.cuh"

namespace mtgv {

// cvRound(fX), cvRound(fY) of WarpPerspectiveInvoker for x1 columns past a block origin (X0,Y0,W0).
// Exact restatement.
static __device__ __noinline__ int2 persp_exact(double X0, double Y0, double W0, double m0, double m3, double m6, double x1) {
  double W = __dadd_rn(W0, __dmul_rn(m6, x1));
  W = W != 0.0 ? __ddiv_rn(32.0, W) : 0.0;
  int2 r;
  r.x = __double2int_rn(__dmul_rn(__dadd_rn(X0, __dmul_rn(m0, x1)), W));  // saturating, like cv2's clamp
  r.y = __double2int_rn(__dmul_rn(__dadd_rn(Y0, __dmul_rn(m3, x1)), W));
  return r;
}

// Guarded fast path: reciprocal by rcp.approx + two Newton steps (relative error ~2^-52), results scaled by
// 2^16 so the distance to the nearest rounding tie is visible in the low bits.  Whenever either coordinate
// is within 4/65536 of a tie, saturates, or the reciprocal is not finite, the exact routine decides.  The
// approximate value differs from cv2's by < 1e-6 of those 1/65536 units, so the outputs are identical.
static __device__ __forceinline__ int2 persp_xy(double X0, double Y0, double W0, double m0, double m3, double m6, double x1) {
  const double W = __fma_rn(m6, x1, W0);
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(W));
  double e = __fma_rn(-W, r, 1.0);
  r = __fma_rn(r, e, r);
  e = __fma_rn(-W, r, 1.0);
  r = __fma_rn(r, e, r);
  const double s = __dmul_rn(r, 2097152.0);  // 32 * 2^16
  const int xq = __double2int_rn(__dmul_rn(__fma_rn(m0, x1, X0), s));
  const int yq = __double2int_rn(__dmul_rn(__fma_rn(m3, x1, Y0), s));
  const unsigned tx = ((unsigned)xq + 32772u) & 0xFFFFu, ty = ((unsigned)yq + 32772u) & 0xFFFFu;
  const bool finite = (__double2hiint(r) & 0x7FF00000) != 0x7FF00000;
  const bool in_range = (unsigned)xq + 0x7FF00000u < 0xFFE00000u && (unsigned)yq + 0x7FF00000u < 0xFFE00000u;
  if (finite && in_range && tx > 8u && ty > 8u) return make_int2((int)((unsigned)xq + 32768u) >> 16, (int)((unsigned)yq + 32768u) >> 16);
  return persp_exact(X0, Y0, W0, m0, m3, m6, x1);
}

// X0, Y0, W0 of WarpPerspectiveInvoker at column-block origin bx of destination row y (inverse matrix M)
static __device__ __forceinline__ void persp_origin(const double* M, double bx, double yy, double* o) {
  o[0] = __dadd_rn(__dadd_rn(__dmul_rn(M[0], bx), __dmul_rn(M[1], yy)), M[2]);
  o[1] = __dadd_rn(__dadd_rn(__dmul_rn(M[3], bx), __dmul_rn(M[4], yy)), M[5]);
  o[2] = __dadd_rn(__dadd_rn(__dmul_rn(M[6], bx), __dmul_rn(M[7], yy)), M[8]);
}

}  // namespace mtgv

namespace mtgv {

// ---- segment-relative coordinates ------------------------------------------------------------
// Along a destination row, fX(x) = 32 * NX(x) / W(x) with NX, W linear in x.  Around a reference
// column c:  fX(c + u) - fX(c) = u * D / (1 + e * u),  D = (32 * m0 - m6 * fX(c)) / W(c),  e = m6 / W(c)
// (exact algebra).  A segment stores cv2's own fX(c), fY(c) (exact fp64 evaluation, split into
// integer and float32 fraction) and D, e in float32; the other columns of the segment (|u| <= 8) are
// evaluated in float32.  With |u * D| <= 300 the float32 error of the sum is below 1.4e-4 (2.5 ulp on the
// product, 1 ulp on the quotient, 1/2 ulp of 512 on the sum), so whenever the result is further than kPerspGuard from a
// rounding tie it rounds to the same integer as cv2's double; otherwise (or for degenerate
// segments, which carry NaN) the caller evaluates the exact routine.
constexpr int kPerspSegShift = 4;           // 16 columns per segment, reference at +8
constexpr float kPerspGuard = 4.0e-4f;

struct PerspSeg {
  int xc, yc;
  float fx, fy, dx, dy, e, _pad;
};

static __device__ __forceinline__ void persp_seg_build(const double* M, int seg, int y, int bw0, PerspSeg* s) {
  const int xc = (seg << kPerspSegShift) + (1 << (kPerspSegShift - 1));
  const int bxi = (bw0 & (bw0 - 1)) == 0 ? (xc & ~(bw0 - 1)) : (xc / bw0) * bw0;
  double o[3];
  persp_origin(M, (double)bxi, (double)y, o);
  const double x1 = (double)(xc - bxi);
  const double W = __dadd_rn(o[2], __dmul_rn(M[6], x1));
  const double Wi = W != 0.0 ? __ddiv_rn(32.0, W) : 0.0;
  const double fX = __dmul_rn(__dadd_rn(o[0], __dmul_rn(M[0], x1)), Wi);
  const double fY = __dmul_rn(__dadd_rn(o[1], __dmul_rn(M[3], x1)), Wi);
  const int Xc = __double2int_rn(fX), Yc = __double2int_rn(fY);
  s->xc = Xc; s->yc = Yc;
  s->fx = (float)(fX - (double)Xc);
  s->fy = (float)(fY - (double)Yc);
  const double iw = Wi * 0.03125;  // 1 / W up to one rounding; D and e only need float32 accuracy
  float dx = (float)((32.0 * M[0] - M[6] * fX) * iw), dy = (float)((32.0 * M[3] - M[6] * fY) * iw);
  const float e = (float)(M[6] * iw);
  const float half = (float)(1 << (kPerspSegShift - 1));
  const bool ok = W != 0.0 && fabs(fX) < 1.0e9 && fabs(fY) < 1.0e9 && fabsf(dx) * half <= 300.f && fabsf(dy) * half <= 300.f &&
                  fabsf(e) * half <= 0.25f;
  if (!ok) dx = dy = __int_as_float(0x7fc00000);  // NaN: every evaluation falls back to the exact routine
  s->dx = dx; s->dy = dy; s->e = e; s->_pad = 0.f;
}

// u = column - reference column of the segment.  Returns false when the exact routine must decide.
// Rounding by the 1.5 * 2^23 trick: (s + K) - K is rint(s) for |s| < 2^22, and the low mantissa bits of
// s + K are that integer.
static __device__ __forceinline__ bool persp_seg_eval(const PerspSeg& s, float u, int* X, int* Y) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fmaf_rn(s.e, u, 1.f)));  // |e * u| <= 1/4: no denormals, 1 ulp
  const float q = u * r;
  const float sx = __fmaf_rn(q, s.dx, s.fx), sy = __fmaf_rn(q, s.dy, s.fy);
  const float K = 12582912.f;
  const float kx = sx + K, ky = sy + K;
  *X = s.xc + (__float_as_int(kx) - 0x4B400000);
  *Y = s.yc + (__float_as_int(ky) - 0x4B400000);
  const float ex = fabsf(sx - (kx - K)), ey = fabsf(sy - (ky - K));
  return ex < 0.5f - kPerspGuard && ey < 0.5f - kPerspGuard;  // false for NaN (degenerate segments carry NaN slopes)
}

}  // namespace mtgv
