// mtgv_poly.cuh - convex polygon clipping in fp64 (placeholder until the detection path
// lands; see mtgv_det.cu).  Host/device for the same reason as mtgv_geom.cuh.
#pragma once
#include "mtgv_geom.cuh"

namespace mtgv {

constexpr int kPolyMax = 16;

// Shoelace area, absolute value; vertices p[2k], p[2k+1].
MTGV_HD double poly_area(const double* p, int n) {
  if (n < 3) return 0.0;
  double s = 0.0;
  for (int i = 0; i < n; i++) {
    int j = i + 1 == n ? 0 : i + 1;
    s = MTGV_DADD(s, MTGV_DSUB(MTGV_DMUL(p[2 * i], p[2 * j + 1]), MTGV_DMUL(p[2 * j], p[2 * i + 1])));
  }
  return fabs(MTGV_DMUL(s, 0.5));
}

// signed twice-area orientation of polygon
MTGV_HD double poly_signed2(const double* p, int n) {
  double s = 0.0;
  for (int i = 0; i < n; i++) {
    int j = i + 1 == n ? 0 : i + 1;
    s = MTGV_DADD(s, MTGV_DSUB(MTGV_DMUL(p[2 * i], p[2 * j + 1]), MTGV_DMUL(p[2 * j], p[2 * i + 1])));
  }
  return s;
}

// Sutherland-Hodgman: subject (any simple polygon) clipped by a CONVEX polygon.
// out must hold 2*kPolyMax doubles; returns the vertex count (<= kPolyMax).
MTGV_HDN int clip_convex(const double* subj, int ns, const double* clip, int nc, double* out) {
  double a[2 * kPolyMax], b[2 * kPolyMax];
  int na = ns;
  for (int k = 0; k < 2 * ns; k++) a[k] = subj[k];
  double orient = poly_signed2(clip, nc) >= 0.0 ? 1.0 : -1.0;
  for (int e = 0; e < nc && na > 0; e++) {
    int e2 = e + 1 == nc ? 0 : e + 1;
    double ex = clip[2 * e], ey = clip[2 * e + 1];
    double dx = MTGV_DSUB(clip[2 * e2], ex), dy = MTGV_DSUB(clip[2 * e2 + 1], ey);
    int nb = 0;
    for (int i = 0; i < na; i++) {
      int j = i + 1 == na ? 0 : i + 1;
      double px = a[2 * i], py = a[2 * i + 1], qx = a[2 * j], qy = a[2 * j + 1];
      double sp = MTGV_DMUL(orient, MTGV_DSUB(MTGV_DMUL(dx, MTGV_DSUB(py, ey)), MTGV_DMUL(dy, MTGV_DSUB(px, ex))));
      double sq = MTGV_DMUL(orient, MTGV_DSUB(MTGV_DMUL(dx, MTGV_DSUB(qy, ey)), MTGV_DMUL(dy, MTGV_DSUB(qx, ex))));
      bool pin = sp >= 0.0, qin = sq >= 0.0;
      if (pin && nb < kPolyMax) {
        b[2 * nb] = px;
        b[2 * nb + 1] = py;
        nb++;
      }
      if (pin != qin && nb < kPolyMax) {
        double t = MTGV_DDIV(sp, MTGV_DSUB(sp, sq));
        b[2 * nb] = MTGV_DADD(px, MTGV_DMUL(t, MTGV_DSUB(qx, px)));
        b[2 * nb + 1] = MTGV_DADD(py, MTGV_DMUL(t, MTGV_DSUB(qy, py)));
        nb++;
      }
    }
    na = nb;
    for (int k = 0; k < 2 * nb; k++) a[k] = b[k];
  }
  for (int k = 0; k < 2 * na; k++) out[k] = a[k];
  return na;
}

}  // namespace mtgv
