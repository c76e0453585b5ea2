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
