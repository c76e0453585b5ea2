// mtgv_mask.h - host-side rounded-rectangle mask (runs once per card pool).
#pragma once
#include <stddef.h>

namespace mtgv {

// cv::Circle (drawing.cpp) restated for the filled quarter disc of round_rect_mask
// (mtgvision/util/image.py:406-425); host side, runs once per pool.
inline void host_round_rect_mask(int h, int w, int radius, float* out) {
  for (size_t i = 0; i < (size_t)h * w; i++) out[i] = 1.f;
  const int r = radius;
  float* corner = new float[(size_t)r * r]();
  int err = 0, dx = r, dy = 0, plus = 1, minus = (r << 1) - 1;
  while (dx >= dy) {
    const int rows[2] = {dy, dx}, half[2] = {dx, dy};
    for (int k = 0; k < 2; k++)
      if (rows[k] >= 0 && rows[k] < r)
        for (int x = 0; x <= (half[k] < r - 1 ? half[k] : r - 1); x++) corner[rows[k] * r + x] = 1.f;
    dy++;
    err += plus;
    plus += 2;
    int mask = (err <= 0) - 1;
    err -= minus & mask;
    dx += mask;
    minus -= mask & 2;
  }
  // corner(y,x) is the top-left-origin quarter disc (centre at (0,0)) = bottom-right piece
  for (int y = 0; y < r; y++)
    for (int x = 0; x < r; x++) {
      float v = corner[y * r + x];
      // np.rot90(c,1)[i][j] = c[j][r-1-i]; rot90(c,2)[i][j] = c[r-1-i][r-1-j]; rot90(c,3)[i][j] = c[r-1-j][i]
      out[(size_t)(h - r + y) * w + (w - r + x)] = v;                             // br: rot90(corner, 0)
      out[(size_t)y * w + (w - r + x)] = corner[x * r + (r - 1 - y)];             // tr: rot90(corner, 1)
      out[(size_t)y * w + x] = corner[(r - 1 - y) * r + (r - 1 - x)];             // tl: rot90(corner, 2)
      out[(size_t)(h - r + y) * w + x] = corner[(r - 1 - x) * r + y];             // bl: rot90(corner, 3)
    }
  delete[] corner;
}


}  // namespace mtgv
