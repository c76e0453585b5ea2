// mtgv_bg.cu - k_background: the background chain of make_virtual for sm_100a.
//
// make_bg (mtgvision/encoder_datasets.py:774-784) = flip -> rotate_bounded (warpAffine onto
// an (nh,nw) canvas, util/image.py:380-398) -> warp_inv (warpPerspective, same canvas,
// encoder_datasets.py:113-116) -> crop_to_size (INTER_AREA to (rh,rw), centre crop,
// util/image.py:349-377), with tint / fade / brightness_contrast (:165-193) scheduled before
// or after the geometric group.  Only the out_h x out_w crop survives, so the chain is
// evaluated backwards per output tile: the tile's INTER_AREA window of the warp_inv image and,
// under the inverse homography, the bounding box of that window in the rotate canvas.  Both
// are staged in shared memory as float32; every pixel of either stage is computed once per
// tile.
//
// What is exact and what is toleranced.  Source COORDINATES decide which texels are read and
// are reproduced bit for bit: warpAffine's 10-bit fixed point and
// warpPerspective's fp64 X/W on the 1/32 grid (note 1; a guarded fast path below recomputes
// with cv2's exact operation order whenever its own result is within 2^-14 of a rounding
// tie).  Pixel VALUES are float32 with fused multiply-adds and the u8 -> [0,1] scaling folded
// into the interpolation; they differ from cv2's separately rounded sums by ~1e-7, far inside
// the +-1 uint8 LSB bar of BASELINE.json (tests/test_gpu_encoder.py states the bar).
//
// Layouts: background pool = RGBX uint8 words (one 4-byte load per bilinear tap fetches the
// three channels), rotate-canvas tile = float4 per pixel (one 16-byte shared load per tap),
// warp_inv tile = three float planes.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "mtgv_internal.cuh"
#include "mtgv_persp.cuh"

namespace mtgv {

#ifndef MTGV_BG_SKIP
#define MTGV_BG_SKIP 0  // timing experiments only: 1 skip stage R, 2 skip stage W, 4 skip stage A, 8 skip tile set-up
#endif
#ifndef MTGV_BG_THREADS
#define MTGV_BG_THREADS 256
#endif
constexpr int kBgThreads = MTGV_BG_THREADS;
#ifndef MTGV_BG_BIG
#define MTGV_BG_BIG 0  // 1: one 512-thread CTA per SM working on 16x32 tiles (experiment)
#endif
constexpr int kBgCtas = MTGV_BG_BIG ? 1 : 2;
constexpr int kBgTR = MTGV_BG_BIG ? 16 : 8, kBgTC = 32;   // output pixels per tile
constexpr int kBgWCap = MTGV_BG_BIG ? 6400 : 3200;          // warp_inv pixels staged per tile
constexpr int kBgRCap = MTGV_BG_BIG ? 5632 : 2816;          // rotate-canvas pixels staged per tile
constexpr int kBgWRows = MTGV_BG_BIG ? 64 : 32, kBgWSegs = 8;  // window rows / 16-column coordinate segments per tile
constexpr int kBgCanvas = 704;         // rotate canvases up to this extent use per-item fixed-point tables
constexpr int kBgBandRows = 32;        // output rows per work item
constexpr int kBgMaxOW = 256;

struct AreaEnt {  // one destination index of computeResizeAreaTab, compact: taps = [left?] mid* [right?]
  int start, n;   // n = taps | left partial << 8 | right partial << 9
  float wl, wm, wr;
};

struct BgSmem {
  float4 rtile[kBgRCap];
  float wtile[3][kBgWCap];
  PerspSeg seg[kBgWRows * kBgWSegs];   // reference coordinates of the warp_inv stage per (window row, 16-column segment)
  int colA[kBgCanvas], colB[kBgCanvas], rowX[kBgCanvas], rowY[kBgCanvas];
  AreaEnt ax[kBgMaxOW], ay[kBgBandRows];
  // the sample
  int status, kind, bg, out_h, out_w, bg_h, bg_w, flip_h, flip_v, nh, nw, bg_rh, bg_rw, bg_y0, bg_x0, n_fg, n_pre, n_post;
  double rot_inv[6], winv[9];
  // elementwise chains: value = step1(step0(byte)); step = fma(v, a, b) [+ saturate]
  float pre_a[2][3], pre_b[2][3];
  int pre_clip[2], pre_steps, pre_linear;
  float lin_a[3], lin_b[3];
  float post_a[2][3], post_b[2][3];
  int post_clip[2], post_steps, post_linear;
  float plin_a[3], plin_b[3];
  int max_nx;
  int tile[8];  // rx0, ry0, rtw, rth, division magics of rtw, WW, tile width
  int item;
};

__device__ __forceinline__ float sat01(float v) { return __saturatef(v); }

__device__ __noinline__ int2 persp_coord_nl(const double* M, int x, int y, int bw0) {
  int2 r;
  persp_coord(M, x, y, bw0, &r.x, &r.y);
  return r;
}

__device__ __forceinline__ float byte_f(uint32_t w, int c) {  // exact u8 -> float without I2F
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u + c)) - 8388608.f;
}

__device__ __forceinline__ void bilinear_w(int ax, int ay, float* w) {  // initInterTab2D(INTER_LINEAR) products, exact
  const float fx = (float)ax * 0.03125f, fy = (float)ay * 0.03125f, gx = 1.f - fx, gy = 1.f - fy;
  w[0] = gy * gx; w[1] = gy * fx; w[2] = fy * gx; w[3] = fy * fx;
}

struct BgSrc {
  const uint32_t* px;  // RGBX words
  int h, w, pitchw, fh, fv;
};

// elementwise ops scheduled before the geometric group, applied to one source byte of channel c
__device__ __forceinline__ float pre_chain(const BgSmem& S, float b, int c) {
  float v = __fmaf_rn(b, S.pre_a[0][c], S.pre_b[0][c]);
  if (S.pre_clip[0]) v = sat01(v);
  if (S.pre_steps > 1) {
    v = __fmaf_rn(v, S.pre_a[1][c], S.pre_b[1][c]);
    if (S.pre_clip[1]) v = sat01(v);
  }
  return v;
}

// General rotate_bounded output pixel from its fixed-point source coordinates: any tap may fall
// outside the source (BORDER_CONSTANT 0), cv2.flip folded into the index.
__device__ __noinline__ float4 rot_px_general(const BgSmem& S, const BgSrc& b, int X, int Y) {
  const int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
  float w[4];
  bilinear_w(X & 31, Y & 31, w);
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int t = 0; t < 4; t++) {
    const int y = sy + (t >> 1), x = sx + (t & 1);
    if ((unsigned)y < (unsigned)b.h && (unsigned)x < (unsigned)b.w) {
      const int yy = b.fv ? b.h - 1 - y : y, xx = b.fh ? b.w - 1 - x : x;
      const uint32_t word = __ldg(b.px + (size_t)yy * b.pitchw + xx);
#pragma unroll
      for (int c = 0; c < 3; c++) acc[c] = __fmaf_rn(pre_chain(S, byte_f(word, c), c), w[t], acc[c]);
    }
  }
  return make_float4(acc[0], acc[1], acc[2], 0.f);
}

__device__ __forceinline__ void rot_coords_general(const BgSmem& S, int ry, int rx, int* X, int* Y) {
  int cA, cB, oX, oY;
  if (rx < kBgCanvas) { cA = S.colA[rx]; cB = S.colB[rx]; }
  else { cA = affine_col_delta(S.rot_inv[0], rx); cB = affine_col_delta(S.rot_inv[3], rx); }
  if (ry < kBgCanvas) { oX = S.rowX[ry]; oY = S.rowY[ry]; }
  else { oX = affine_row_origin(S.rot_inv[1], S.rot_inv[2], ry); oY = affine_row_origin(S.rot_inv[4], S.rot_inv[5], ry); }
  *X = (oX + cA) >> 5;
  *Y = (oY + cB) >> 5;
}

// rotate canvas pixel (ry, rx) for the warp_inv stage: zero outside the canvas, the staged tile when inside
// it, else recomputed (keeps the tiling a pure optimisation)
__device__ __noinline__ float4 rot_at(const BgSmem& S, const BgSrc& b, int ry, int rx) {
  if ((unsigned)ry >= (unsigned)S.nh || (unsigned)rx >= (unsigned)S.nw) return make_float4(0.f, 0.f, 0.f, 0.f);
  const int ty = ry - S.tile[1], tx = rx - S.tile[0];
  if ((unsigned)ty < (unsigned)S.tile[3] && (unsigned)tx < (unsigned)S.tile[2]) return S.rtile[ty * S.tile[2] + tx];
  int X, Y;
  rot_coords_general(S, ry, rx, &X, &Y);
  return rot_px_general(S, b, X, Y);
}

__device__ __forceinline__ float area_w(const AreaEnt& e, int k) {
  if (k == 0 && (e.n & 256)) return e.wl;
  if (k == (e.n & 255) - 1 && (e.n & 512)) return e.wr;
  return e.wm;
}

// k / d == (k * magic) >> 24 for k * d < 2^24 (tile-local indices and extents: k < 2^15, d < 2^9)
__device__ __forceinline__ unsigned div_magic(int d) { return d > 0 ? ((1u << 24) + (unsigned)d - 1u) / (unsigned)d : 0u; }
__device__ __forceinline__ int div_by(int k, unsigned magic) { return (int)(((unsigned)k * magic) >> 24); }

// ---- stage R: rotate_bounded's warpAffine output over the staged footprint ----
// LINEAR: the elementwise chain in front of the geometric group is affine on the byte range, so the four
// taps are interpolated as integers (weights (32-ax)(32-ay).. sum to 1024, exact) and the chain is applied once.
struct RTap {
  int X, Y;
  uint32_t t0, t1, t2, t3;
  bool in;
};

// MODE 1 (LINEAR): integer taps, affine chain applied once.  MODE 2: the frequent non-affine chain - first step clipped, second
// step (identity when absent) not clipped - with its constants in registers (la/lb = step 0, qa/qb = step 1): per tap and
// channel PRMT, FADD, FFMA.SAT, FFMA, FFMA.  MODE 0: any chain, flags read from shared memory.
template <int MODE, bool TABLES>
__device__ __forceinline__ void stage_rotate(BgSmem& S, const BgSrc& b, int rx0, int ry0, int rtw, int rth, unsigned magic,
                                             const float* la, const float* lb, const float* qa, const float* qb, int tid) {
  constexpr bool LINEAR = MODE == 1;
  const int npx = rtw * rth;
  const int rs = b.fv ? -b.pitchw : b.pitchw, cs = b.fh ? -1 : 1;
  const unsigned in_w = (unsigned)(b.w - 1), in_h = (unsigned)(b.h - 1);
  const int ybase = b.fv ? b.h - 1 : 0, ysign = b.fv ? -1 : 1, xbase = b.fh ? b.w - 1 : 0, xsign = b.fh ? -1 : 1;
  auto fetch = [&](int k) {
    RTap t;
    const int ty = div_by(k, magic), tx = k - ty * rtw;
    if (TABLES) {
      t.X = (S.rowX[ry0 + ty] + S.colA[rx0 + tx]) >> 5;
      t.Y = (S.rowY[ry0 + ty] + S.colB[rx0 + tx]) >> 5;
    } else {
      rot_coords_general(S, ry0 + ty, rx0 + tx, &t.X, &t.Y);
    }
    const int sx = t.X >> 5, sy = t.Y >> 5;  // canvases are far below the int16 saturation of cv2's remap
    t.in = (unsigned)sx < in_w && (unsigned)sy < in_h;
    t.t0 = t.t1 = t.t2 = t.t3 = 0u;
    if (t.in) {
      // all four taps inside the source: cv2.flip folded into the base index and strides
      const int i0 = (ybase + ysign * sy) * b.pitchw + (xbase + xsign * sx);
      t.t0 = __ldg(b.px + i0); t.t1 = __ldg(b.px + (i0 + cs)); t.t2 = __ldg(b.px + (i0 + rs)); t.t3 = __ldg(b.px + (i0 + rs + cs));
    }
    return t;
  };
  auto finish = [&](int k, const RTap& t) {
    float4 v;
    if (t.in) {
      const int ax = t.X & 31, ay = t.Y & 31;
      if (LINEAR) {
        const unsigned pxw = (unsigned)(32 - ax) + ((unsigned)ax << 16);
        const unsigned wtop = pxw * (unsigned)(32 - ay), wbot = pxw * (unsigned)ay;  // (w00 | w01 << 16), (w10 | w11 << 16)
        const unsigned rg_t = __byte_perm(t.t0, t.t1, 0x5140), rg_b = __byte_perm(t.t2, t.t3, 0x5140);  // [R0 R1 G0 G1]
        const unsigned b_t = __byte_perm(t.t0, t.t1, 0x0062), b_b = __byte_perm(t.t2, t.t3, 0x0062);    // [B0 B1 . .]
        const unsigned rw = __byte_perm(rg_t, rg_b, 0x5410), gw = __byte_perm(rg_t, rg_b, 0x7632), bw = __byte_perm(b_t, b_b, 0x5410);
        const unsigned sr = __dp2a_hi(wbot, rw, __dp2a_lo(wtop, rw, 0u));
        const unsigned sg = __dp2a_hi(wbot, gw, __dp2a_lo(wtop, gw, 0u));
        const unsigned sb = __dp2a_hi(wbot, bw, __dp2a_lo(wtop, bw, 0u));
        v = make_float4(__fmaf_rn((float)sr, la[0], lb[0]), __fmaf_rn((float)sg, la[1], lb[1]), __fmaf_rn((float)sb, la[2], lb[2]), 0.f);
      } else {
        float w[4];
        bilinear_w(ax, ay, w);
        float o[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
          float a;
          if (MODE == 2) {
            auto ch = [&](uint32_t word) { return __fmaf_rn(sat01(__fmaf_rn(byte_f(word, c), la[c], lb[c])), qa[c], qb[c]); };
            a = ch(t.t0) * w[0];
            a = __fmaf_rn(ch(t.t1), w[1], a);
            a = __fmaf_rn(ch(t.t2), w[2], a);
            a = __fmaf_rn(ch(t.t3), w[3], a);
          } else {
            a = pre_chain(S, byte_f(t.t0, c), c) * w[0];
            a = __fmaf_rn(pre_chain(S, byte_f(t.t1, c), c), w[1], a);
            a = __fmaf_rn(pre_chain(S, byte_f(t.t2, c), c), w[2], a);
            a = __fmaf_rn(pre_chain(S, byte_f(t.t3, c), c), w[3], a);
          }
          o[c] = a;
        }
        v = make_float4(o[0], o[1], o[2], 0.f);
      }
    } else {
      // BORDER_CONSTANT: all four taps outside the source (the empty corners of the rotate canvas) -> 0
      const int sx = t.X >> 5, sy = t.Y >> 5;
      if (sx < -1 || sy < -1 || sx >= b.w || sy >= b.h) v = make_float4(0.f, 0.f, 0.f, 0.f);
      else v = rot_px_general(S, b, t.X, t.Y);
    }
    S.rtile[k] = v;
  };
  // software-pipelined: the four tap words of the next two pixels are in flight while the current two are blended
  if (tid >= npx) return;
  RTap na = fetch(tid), nc = fetch(tid + kBgThreads < npx ? tid + kBgThreads : tid);
  for (int k = tid; k < npx; k += 2 * kBgThreads) {
    const int k2 = k + kBgThreads, k3 = k + 2 * kBgThreads, k4 = k + 3 * kBgThreads;
    const RTap a = na, c = nc;
    if (k3 < npx) {
      na = fetch(k3);
      nc = fetch(k4 < npx ? k4 : k3);
    }
    finish(k, a);
    if (k2 < npx) finish(k2, c);
  }
}

// ---- stage A: INTER_AREA reduction (cv::ResizeArea_: horizontal taps, then weighted rows) + img_clip ----
template <int NX>
__device__ __forceinline__ void stage_area(const BgSmem& S, int ty0, int ty1, int tx0, int tx1, int by0, int wy0, int wx0, int WW,
                                           unsigned magic, float* __restrict__ outp, int OH, int OW, int tid) {
  const int tw = tx1 - tx0, npx = (ty1 - ty0) * tw;
  for (int k = tid; k < npx; k += kBgThreads) {
    const int r = div_by(k, magic), cidx = k - r * tw;
    const AreaEnt ey = S.ay[ty0 - by0 + r], ex = S.ax[tx0 + cidx];
    const int ny = ey.n & 255, nx = ex.n & 255;
    float wx[NX];
#pragma unroll
    for (int i = 0; i < NX; i++) wx[i] = i < nx ? area_w(ex, i) : 0.f;
    float sum[3] = {0.f, 0.f, 0.f};
    int q = (ey.start - wy0) * WW + (ex.start - wx0);
    for (int j = 0; j < ny; j++, q += WW) {
      float h[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < NX; i++) {
        if (i < nx) {
          h[0] = __fmaf_rn(S.wtile[0][q + i], wx[i], h[0]);
          h[1] = __fmaf_rn(S.wtile[1][q + i], wx[i], h[1]);
          h[2] = __fmaf_rn(S.wtile[2][q + i], wx[i], h[2]);
        }
      }
      const float wyj = area_w(ey, j);
#pragma unroll
      for (int c = 0; c < 3; c++) sum[c] = __fmaf_rn(wyj, h[c], sum[c]);
    }
    const int o = (ty0 + r) * OW + tx0 + cidx;
#pragma unroll
    for (int c = 0; c < 3; c++) outp[(size_t)c * OH * OW + o] = sat01(sum[c]);  // img_clip (util/image.py:334)
  }
}

// any tap count (reductions beyond x6: large photographs): weights selected per tap instead of held in registers
__device__ __noinline__ void stage_area_any(const BgSmem& S, int ty0, int ty1, int tx0, int tx1, int by0, int wy0, int wx0, int WW,
                                            float* __restrict__ outp, int OH, int OW, int tid) {
  const int tw = tx1 - tx0, npx = (ty1 - ty0) * tw;
  for (int k = tid; k < npx; k += kBgThreads) {
    const int r = k / tw, cidx = k - r * tw;
    const AreaEnt ey = S.ay[ty0 - by0 + r], ex = S.ax[tx0 + cidx];
    const int ny = ey.n & 255, nx = ex.n & 255;
    float sum[3] = {0.f, 0.f, 0.f};
    int q = (ey.start - wy0) * WW + (ex.start - wx0);
    for (int j = 0; j < ny; j++, q += WW) {
      float h[3] = {0.f, 0.f, 0.f};
      for (int i = 0; i < nx; i++) {
        const float wxi = area_w(ex, i);
        h[0] = __fmaf_rn(S.wtile[0][q + i], wxi, h[0]);
        h[1] = __fmaf_rn(S.wtile[1][q + i], wxi, h[1]);
        h[2] = __fmaf_rn(S.wtile[2][q + i], wxi, h[2]);
      }
      const float wyj = area_w(ey, j);
      for (int c = 0; c < 3; c++) sum[c] = __fmaf_rn(wyj, h[c], sum[c]);
    }
    const int o = (ty0 + r) * OW + tx0 + cidx;
    for (int c = 0; c < 3; c++) outp[(size_t)c * OH * OW + o] = sat01(sum[c]);
  }
}

// window larger than the staging buffers (very large backgrounds): every tap evaluated directly
__device__ __noinline__ void stage_area_direct(const BgSmem& S, const BgSrc& b, int ty0, int ty1, int tx0, int tx1, int by0, int bw0,
                                               float* __restrict__ outp, int OH, int OW, int tid) {
  const int tw = tx1 - tx0, npx = (ty1 - ty0) * tw;
  for (int k = tid; k < npx; k += kBgThreads) {
    const int r = k / tw, cidx = k - r * tw;
    const AreaEnt ey = S.ay[ty0 - by0 + r], ex = S.ax[tx0 + cidx];
    const int ny = ey.n & 255, nx = ex.n & 255;
    float sum[3] = {0.f, 0.f, 0.f};
    for (int j = 0; j < ny; j++) {
      float h[3] = {0.f, 0.f, 0.f};
      for (int i = 0; i < nx; i++) {
        int X, Y;
        persp_coord(S.winv, ex.start + i, ey.start + j, bw0, &X, &Y);
        const int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
        const float4 t0 = rot_at(S, b, sy, sx), t1 = rot_at(S, b, sy, sx + 1), t2 = rot_at(S, b, sy + 1, sx), t3 = rot_at(S, b, sy + 1, sx + 1);
        float w[4], v[3];
        bilinear_w(X & 31, Y & 31, w);
        v[0] = __fmaf_rn(t3.x, w[3], __fmaf_rn(t2.x, w[2], __fmaf_rn(t1.x, w[1], t0.x * w[0])));
        v[1] = __fmaf_rn(t3.y, w[3], __fmaf_rn(t2.y, w[2], __fmaf_rn(t1.y, w[1], t0.y * w[0])));
        v[2] = __fmaf_rn(t3.z, w[3], __fmaf_rn(t2.z, w[2], __fmaf_rn(t1.z, w[1], t0.z * w[0])));
        for (int q = 0; q < S.post_steps; q++)
          for (int c = 0; c < 3; c++) {
            v[c] = __fmaf_rn(v[c], S.post_a[q][c], S.post_b[q][c]);
            if (S.post_clip[q]) v[c] = sat01(v[c]);
          }
        const float wxi = area_w(ex, i);
        for (int c = 0; c < 3; c++) h[c] = __fmaf_rn(v[c], wxi, h[c]);
      }
      const float wyj = area_w(ey, j);
      for (int c = 0; c < 3; c++) sum[c] = __fmaf_rn(wyj, h[c], sum[c]);
    }
    const int o = (ty0 + r) * OW + tx0 + cidx;
    for (int c = 0; c < 3; c++) outp[(size_t)c * OH * OW + o] = sat01(sum[c]);
  }
}

__global__ void __launch_bounds__(kBgThreads, kBgCtas) k_background(const mtgv_enc_params* __restrict__ params, int n, int n_bands,
                                                              const uint8_t* __restrict__ bg_pool, const int64_t* __restrict__ bg_off,
                                                              float* __restrict__ bg_out, int* __restrict__ work_counter) {
  extern __shared__ __align__(16) unsigned char bg_smem_raw[];
  BgSmem& S = *reinterpret_cast<BgSmem*>(bg_smem_raw);
  const int tid = threadIdx.x, nt = kBgThreads, lane = tid & 31;
  for (;;) {
    __syncthreads();
    if (tid == 0) { S.item = atomicAdd(work_counter, 1); S.max_nx = 0; }
    __syncthreads();
    const int item = S.item;
    if (item >= n * n_bands) break;
    const int s = item / n_bands, band = item - s * n_bands;
    const mtgv_enc_params* gp = params + s;
    if (tid == 0) {
      S.status = gp->status; S.kind = gp->kind; S.bg = gp->bg; S.out_h = gp->out_h; S.out_w = gp->out_w;
      S.bg_h = gp->bg_h; S.bg_w = gp->bg_w; S.flip_h = gp->flip_h; S.flip_v = gp->flip_v;
      S.nh = gp->rot_nh; S.nw = gp->rot_nw; S.bg_rh = gp->bg_rh; S.bg_rw = gp->bg_rw; S.bg_y0 = gp->bg_y0; S.bg_x0 = gp->bg_x0;
      S.n_fg = gp->n_fg; S.n_pre = gp->n_pre; S.n_post = gp->n_post;
    } else if (tid >= 32 && tid < 38) {
      S.rot_inv[tid - 32] = gp->rot_inv[tid - 32];
    } else if (tid >= 64 && tid < 73) {
      S.winv[tid - 64] = gp->winv[tid - 64];
    }
    __syncthreads();
    if (S.status != 0 || S.kind == MTGV_KIND_CROPPED) continue;
    const int OH = S.out_h, OW = S.out_w, nh = S.nh, nw = S.nw;
    const int by0 = band * kBgBandRows, by1 = min(OH, by0 + kBgBandRows);
    if (by0 >= by1) continue;
    // ---- per-item tables ----
    if (tid < 3) {
      // elementwise chains (expand_elementwise, mtgv_expand.cuh): pre ops act on source bytes, post ops on warp_inv output
      const int c = tid;
      const mtgv_x_op* ops = gp->ops + S.n_fg;
      const int n_pre = S.n_pre, n_post = S.n_post;
      float a0 = 1.f / 255.f, b0 = 0.f, a1 = 1.f, b1 = 0.f;
      int c0 = 0, c1 = 0;
      if (n_pre >= 1 && ((ops[0].i[0] >> c) & 1)) { a0 = ops[0].f[c] * (1.f / 255.f); b0 = ops[0].f[4 + c]; c0 = ops[0].i[1]; }
      if (n_pre >= 2 && ((ops[1].i[0] >> c) & 1)) { a1 = ops[1].f[c]; b1 = ops[1].f[4 + c]; c1 = ops[1].i[1]; }
      S.pre_a[0][c] = a0; S.pre_b[0][c] = b0; S.pre_a[1][c] = a1; S.pre_b[1][c] = b1;
      // a chain is affine on its input range when no clipped step can leave [0,1] at either end of the range
      const float lo0 = b0, hi0 = __fmaf_rn(255.f, a0, b0);
      bool lin = !c0 || (fminf(lo0, hi0) >= 0.f && fmaxf(lo0, hi0) <= 1.f);
      const float lo1 = __fmaf_rn(lo0, a1, b1), hi1 = __fmaf_rn(hi0, a1, b1);
      lin = lin && (!c1 || (fminf(lo1, hi1) >= 0.f && fmaxf(lo1, hi1) <= 1.f));
      S.lin_a[c] = a0 * a1 * (1.f / 1024.f);  // applied to the integer tap sum (weights sum to 1024)
      S.lin_b[c] = __fmaf_rn(b0, a1, b1);
      const unsigned all = __ballot_sync(0x7u, lin);
      // post chain on values in [0,1]
      float pa0 = 1.f, pb0 = 0.f, pa1 = 1.f, pb1 = 0.f;
      int pc0 = 0, pc1 = 0;
      if (n_post >= 1 && ((ops[n_pre].i[0] >> c) & 1)) { pa0 = ops[n_pre].f[c]; pb0 = ops[n_pre].f[4 + c]; pc0 = ops[n_pre].i[1]; }
      if (n_post >= 2 && ((ops[n_pre + 1].i[0] >> c) & 1)) { pa1 = ops[n_pre + 1].f[c]; pb1 = ops[n_pre + 1].f[4 + c]; pc1 = ops[n_pre + 1].i[1]; }
      S.post_a[0][c] = pa0; S.post_b[0][c] = pb0; S.post_a[1][c] = pa1; S.post_b[1][c] = pb1;
      const float ql0 = pb0, qh0 = pa0 + pb0;
      bool plin = !pc0 || (fminf(ql0, qh0) >= 0.f && fmaxf(ql0, qh0) <= 1.f);
      const float ql1 = __fmaf_rn(ql0, pa1, pb1), qh1 = __fmaf_rn(qh0, pa1, pb1);
      plin = plin && (!pc1 || (fminf(ql1, qh1) >= 0.f && fmaxf(ql1, qh1) <= 1.f));
      S.plin_a[c] = pa0 * pa1;
      S.plin_b[c] = __fmaf_rn(pb0, pa1, pb1);
      const unsigned pall = __ballot_sync(0x7u, plin);
      if (c == 0) {
        S.pre_clip[0] = (n_pre >= 1) ? ops[0].i[1] : 0;
        S.pre_clip[1] = (n_pre >= 2) ? ops[1].i[1] : 0;
        S.pre_steps = n_pre >= 2 ? 2 : 1;
        S.pre_linear = all == 0x7u;
        S.post_steps = n_post;
        S.post_clip[0] = n_post >= 1 ? ops[n_pre].i[1] : 0;
        S.post_clip[1] = n_post >= 2 ? ops[n_pre + 1].i[1] : 0;
        S.post_linear = pall == 0x7u;
      }
    }
    for (int k = tid; k < kBgCanvas; k += nt) {  // warpAffine fixed-point tables (cv::warpAffine adelta/bdelta, X0/Y0)
      if (k < nw) { S.colA[k] = affine_col_delta(S.rot_inv[0], k); S.colB[k] = affine_col_delta(S.rot_inv[3], k); }
      if (k < nh) { S.rowX[k] = affine_row_origin(S.rot_inv[1], S.rot_inv[2], k); S.rowY[k] = affine_row_origin(S.rot_inv[4], S.rot_inv[5], k); }
    }
    // crop_to_size: INTER_AREA (nh,nw)->(bg_rh,bg_rw) tables for the band's rows and all columns
    const bool enlarge = S.bg_rh > nh || S.bg_rw > nw;  // cv::resize: INTER_AREA enlarging = 2-tap area-linear on both axes
    for (int k = tid; k < OW + (by1 - by0); k += nt) {
      AreaEnt e;
      if (k < OW) {
        if (enlarge) area_linear_compact(nw, S.bg_rw, S.bg_x0 + k, &e.start, &e.n, &e.wl, &e.wm, &e.wr);
        else area_compact(nw, S.bg_rw, S.bg_x0 + k, &e.start, &e.n, &e.wl, &e.wm, &e.wr);
        S.ax[k] = e;
        atomicMax(&S.max_nx, e.n & 255);
      } else {
        if (enlarge) area_linear_compact(nh, S.bg_rh, S.bg_y0 + by0 + (k - OW), &e.start, &e.n, &e.wl, &e.wm, &e.wr);
        else area_compact(nh, S.bg_rh, S.bg_y0 + by0 + (k - OW), &e.start, &e.n, &e.wl, &e.wm, &e.wr);
        S.ay[k - OW] = e;
      }
    }
    BgSrc b;
    b.h = S.bg_h; b.w = S.bg_w; b.pitchw = (S.bg_w + 3) & ~3; b.fh = S.flip_h; b.fv = S.flip_v;
    b.px = reinterpret_cast<const uint32_t*>(bg_pool + bg_off[S.bg]);
    const int bw0 = persp_block_w(nh, nw);
    const bool tables = nh <= kBgCanvas && nw <= kBgCanvas;
    // tile shape: shrink until a tile's warp_inv window fits the staging buffer
    int TR = kBgTR, TC = kBgTC;
    {
      const double sy = (double)nh / S.bg_rh, sx = (double)nw / S.bg_rw;
      while (TR * TC > 1) {
        const int wh = (int)(TR * sy) + 3, ww = (int)(TC * sx) + 3;
        if (wh * ww <= kBgWCap && wh <= kBgWRows && (ww >> kPerspSegShift) + 2 <= kBgWSegs) break;
        if (TC * sx >= TR * sy && TC > 1) TC >>= 1; else if (TR > 1) TR >>= 1; else TC >>= 1;
      }
    }
    float* outp = bg_out + (size_t)s * 3 * OH * OW;
    __syncthreads();
    const bool pre_linear = S.pre_linear != 0, post_linear = S.post_linear != 0;
    // the frequent non-affine chain: step 0 clipped, step 1 (identity when absent) not clipped
    const bool pre_sat_nosat = !pre_linear && S.pre_clip[0] != 0 && (S.pre_steps < 2 || S.pre_clip[1] == 0);
    const int max_nx = S.max_nx;
    float la[3], lb[3], pa[3], pb[3], qa[3], qb[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      pa[c] = S.plin_a[c]; pb[c] = S.plin_b[c];
      if (pre_sat_nosat) {
        la[c] = S.pre_a[0][c]; lb[c] = S.pre_b[0][c];
        qa[c] = S.pre_steps > 1 ? S.pre_a[1][c] : 1.f; qb[c] = S.pre_steps > 1 ? S.pre_b[1][c] : 0.f;
      } else {
        la[c] = S.lin_a[c]; lb[c] = S.lin_b[c]; qa[c] = 1.f; qb[c] = 0.f;
      }
    }

    // Tiles of the band are software-pipelined: the set-up of tile t+1 (bounding box, coordinate segments) shares
    // a barrier interval with the INTER_AREA reduction of tile t, which reads none of what the set-up writes.
    struct TileGeom { int ty0, ty1, tx0, tx1, wy0, wy1, wx0, wx1, WH, WW, seg0, nseg; bool staged; };
    const int tiles_x = (OW + TC - 1) / TC, n_tiles = tiles_x * ((by1 - by0 + TR - 1) / TR);
    auto geom = [&](int t) {
      TileGeom g;
      const int tyi = t / tiles_x;
      g.ty0 = by0 + tyi * TR; g.tx0 = (t - tyi * tiles_x) * TC;
      g.ty1 = min(g.ty0 + TR, by1); g.tx1 = min(g.tx0 + TC, OW);
      const AreaEnt eY0 = S.ay[g.ty0 - by0], eY1 = S.ay[g.ty1 - 1 - by0], eX0 = S.ax[g.tx0], eX1 = S.ax[g.tx1 - 1];
      g.wy0 = eY0.start; g.wy1 = eY1.start + (eY1.n & 255); g.wx0 = eX0.start; g.wx1 = eX1.start + (eX1.n & 255);
      g.WH = g.wy1 - g.wy0; g.WW = g.wx1 - g.wx0;
      g.seg0 = g.wx0 >> kPerspSegShift; g.nseg = ((g.wx1 - 1) >> kPerspSegShift) - g.seg0 + 1;
      g.staged = g.WH * g.WW <= kBgWCap && g.WH <= kBgWRows && g.nseg <= kBgWSegs;
      return g;
    };
    auto setup = [&](const TileGeom& g) {
      if (tid < 32) {
        // bounding box, in rotate-canvas pixels, of the window's image under the inverse homography
        // (a projective map with W > 0 sends the window rectangle into the convex hull of its corners)
        int X, Y;
        persp_coord(S.winv, (lane & 1) ? g.wx1 - 1 : g.wx0, (lane & 2) ? g.wy1 - 1 : g.wy0, bw0, &X, &Y);
        int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
        int minx = sx, maxx = sx, miny = sy, maxy = sy;
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
          minx = min(minx, __shfl_xor_sync(0xffffffffu, minx, o)); maxx = max(maxx, __shfl_xor_sync(0xffffffffu, maxx, o));
          miny = min(miny, __shfl_xor_sync(0xffffffffu, miny, o)); maxy = max(maxy, __shfl_xor_sync(0xffffffffu, maxy, o));
        }
        if (lane == 0) {
          const int x0 = max(minx - 1, 0), x1 = min(maxx + 3, nw), y0 = max(miny - 1, 0), y1 = min(maxy + 3, nh);
          int rtw = max(x1 - x0, 0), rth = max(y1 - y0, 0);
          if (rtw * rth > kBgRCap || !g.staged) rtw = rth = 0;
          S.tile[0] = x0; S.tile[1] = y0; S.tile[2] = rtw; S.tile[3] = rth;
          S.tile[4] = (int)div_magic(rtw);
        } else if (lane == 1) {
          S.tile[5] = (int)div_magic(g.WW);
        } else if (lane == 2) {
          S.tile[6] = (int)div_magic(g.tx1 - g.tx0);
        }
      } else if (g.staged) {
        // exact cv2 coordinates at the reference column of every (window row, segment)
        for (int k = tid - 32; k < g.WH * g.nseg; k += nt - 32) {
          const int wr = k / g.nseg;
          persp_seg_build(S.winv, g.seg0 + (k - wr * g.nseg), g.wy0 + wr, bw0, &S.seg[k]);
        }
      }
    };
    TileGeom g = geom(0);
    setup(g);
    __syncthreads();
    for (int t = 0; t < n_tiles; t++) {
      {
        const int ty0 = g.ty0, ty1 = g.ty1, tx0 = g.tx0, tx1 = g.tx1, wy0 = g.wy0, wx0 = g.wx0, WH = g.WH, WW = g.WW;
        const int seg0 = g.seg0, nseg = g.nseg;
        const bool staged = g.staged;
        const int rx0 = S.tile[0], ry0 = S.tile[1], rtw = S.tile[2], rth = S.tile[3];
        const unsigned mg_r = (unsigned)S.tile[4], mg_w = (unsigned)S.tile[5], mg_a = (unsigned)S.tile[6];
        if (MTGV_BG_SKIP & 1) {}
        else if (!tables) stage_rotate<0, false>(S, b, rx0, ry0, rtw, rth, mg_r, la, lb, qa, qb, tid);
        else if (pre_linear) stage_rotate<1, true>(S, b, rx0, ry0, rtw, rth, mg_r, la, lb, qa, qb, tid);
        else if (pre_sat_nosat) stage_rotate<2, true>(S, b, rx0, ry0, rtw, rth, mg_r, la, lb, qa, qb, tid);
        else stage_rotate<0, true>(S, b, rx0, ry0, rtw, rth, mg_r, la, lb, qa, qb, tid);
        __syncthreads();
        if (staged && !(MTGV_BG_SKIP & 2)) {
          // ---- stage W: warp_inv output over the window + elementwise ops scheduled after the geometric group ----
          const int npx = WH * WW;
          const unsigned fast_w = rtw > 0 ? rtw - 1 : 0, fast_h = rth > 0 ? rth - 1 : 0;
          auto coords = [&](int k, int2* XY) {  // returns false when the exact routine has to decide
            const int r = div_by(k, mg_w), cx = k - r * WW;
            const int wx = wx0 + cx;
            const PerspSeg sg = S.seg[r * nseg + ((wx >> kPerspSegShift) - seg0)];
            return persp_seg_eval(sg, (float)((wx & ((1 << kPerspSegShift) - 1)) - (1 << (kPerspSegShift - 1))), &XY->x, &XY->y);
          };
          auto coords_exact = [&](int k) {
            const int r = div_by(k, mg_w);
            return persp_coord_nl(S.winv, wx0 + (k - r * WW), wy0 + r, bw0);
          };
          auto finish = [&](int k, int2 XY) {
            const int sx = XY.x >> 5, sy = XY.y >> 5;
            const int tx = sx - rx0, ty = sy - ry0;
            float4 t0, t1, t2, t3;
            if ((unsigned)tx < fast_w && (unsigned)ty < fast_h) {
              const float4* q = S.rtile + ty * rtw + tx;  // 2x2 footprint inside the staged tile (hence inside the canvas)
              t0 = q[0]; t1 = q[1]; t2 = q[rtw]; t3 = q[rtw + 1];
            } else {
              const int ssx = sat_short(sx), ssy = sat_short(sy);
              t0 = rot_at(S, b, ssy, ssx); t1 = rot_at(S, b, ssy, ssx + 1);
              t2 = rot_at(S, b, ssy + 1, ssx); t3 = rot_at(S, b, ssy + 1, ssx + 1);
            }
            float w[4];
            bilinear_w(XY.x & 31, XY.y & 31, w);
            float v[3];
            v[0] = __fmaf_rn(t3.x, w[3], __fmaf_rn(t2.x, w[2], __fmaf_rn(t1.x, w[1], t0.x * w[0])));
            v[1] = __fmaf_rn(t3.y, w[3], __fmaf_rn(t2.y, w[2], __fmaf_rn(t1.y, w[1], t0.y * w[0])));
            v[2] = __fmaf_rn(t3.z, w[3], __fmaf_rn(t2.z, w[2], __fmaf_rn(t1.z, w[1], t0.z * w[0])));
            if (post_linear) {
#pragma unroll
              for (int c = 0; c < 3; c++) v[c] = __fmaf_rn(v[c], pa[c], pb[c]);
            } else {
#pragma unroll
              for (int c = 0; c < 3; c++) {
                v[c] = __fmaf_rn(v[c], S.post_a[0][c], S.post_b[0][c]);
                if (S.post_clip[0]) v[c] = sat01(v[c]);
                v[c] = __fmaf_rn(v[c], S.post_a[1][c], S.post_b[1][c]);
                if (S.post_clip[1]) v[c] = sat01(v[c]);
              }
            }
            S.wtile[0][k] = v[0]; S.wtile[1][k] = v[1]; S.wtile[2][k] = v[2];
          };
          // the trip count is uniform over the warp (inactive lanes recompute pixel 0 and store nothing), so the rare
          // exact-coordinate fallback sits behind one warp vote instead of per-lane divergence bookkeeping
          for (int kb = 0; kb < npx; kb += 2 * nt) {
            const int k = kb + tid, k2 = k + nt;
            const bool one = k < npx, two = k2 < npx;
            int2 xa, xb;
            const bool oka = coords(one ? k : 0, &xa), okb = coords(two ? k2 : 0, &xb);
            if (__any_sync(0xffffffffu, !(oka && okb))) {
              if (!oka) xa = coords_exact(one ? k : 0);
              if (!okb) xb = coords_exact(two ? k2 : 0);
            }
            if (one) finish(k, xa);
            if (two) finish(k2, xb);
          }
        }
        __syncthreads();
        if (MTGV_BG_SKIP & 4) {}
        else if (!staged) stage_area_direct(S, b, ty0, ty1, tx0, tx1, by0, bw0, outp, OH, OW, tid);
        else if (max_nx <= 4) stage_area<4>(S, ty0, ty1, tx0, tx1, by0, wy0, wx0, WW, mg_a, outp, OH, OW, tid);
        else if (max_nx <= 6) stage_area<6>(S, ty0, ty1, tx0, tx1, by0, wy0, wx0, WW, mg_a, outp, OH, OW, tid);
        else if (max_nx <= kAreaMaxTaps) stage_area<kAreaMaxTaps>(S, ty0, ty1, tx0, tx1, by0, wy0, wx0, WW, mg_a, outp, OH, OW, tid);
        else stage_area_any(S, ty0, ty1, tx0, tx1, by0, wy0, wx0, WW, outp, OH, OW, tid);
      }
      if (t + 1 < n_tiles) {
        g = geom(t + 1);
        if (!(MTGV_BG_SKIP & 8)) setup(g);
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------ //
// pool ingest: HWC uint8 RGB -> RGBX words, rows padded to 16 B                          //
// ------------------------------------------------------------------------------------ //

__global__ void k_interleave(const uint8_t* __restrict__ hwc, uint32_t* __restrict__ words, int h, int w, int pitchw) {
  const size_t img = blockIdx.y;
  const uint8_t* src = hwc + img * (size_t)h * w * 3;
  uint32_t* dst = words + img * (size_t)h * pitchw;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)h * pitchw; i += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / pitchw), x = (int)(i % pitchw);
    uint32_t v = 0;
    if (x < w) {
      const uint8_t* p = src + ((size_t)y * w + x) * 3;
      v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
    }
    dst[i] = v;
  }
}

int pool_interleave(mtgv_ctx* ctx, const uint8_t* hwc, uint8_t* words, int n, int h, int w, cudaStream_t st) {
  if (n <= 0) return MTGV_OK;
  const int pitchw = (w + 3) & ~3;
  for (int i0 = 0; i0 < n; i0 += 32768) {
    const int cnt = n - i0 < 32768 ? n - i0 : 32768;
    dim3 grid(64, cnt);
    k_interleave<<<grid, 256, 0, st>>>(hwc + (size_t)i0 * h * w * 3, (uint32_t*)words + (size_t)i0 * h * pitchw, h, w, pitchw);
    ctx->launches++;
  }
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  return MTGV_OK;
}

size_t bg_image_bytes(int h, int w) { return (size_t)h * ((w + 3) & ~3) * 4; }

int bg_launch(mtgv_ctx* ctx, const mtgv_enc_params* params, int m, int OH, int OW, float* bg_out, cudaStream_t st) {
  if (OW > kBgMaxOW) return fail(ctx, MTGV_ERR_LIMIT, "x_size_hw width > 256");
  if (!(ctx->attrs_set & 1u)) {
    MTGV_CUDA_OK(ctx, cudaFuncSetAttribute(k_background, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BgSmem)));
    MTGV_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->bg_blocks_per_sm, k_background, kBgThreads, sizeof(BgSmem)));
    ctx->attrs_set |= 1u;
  }
  if (!ctx->bg_counter) MTGV_CUDA_OK(ctx, cudaMalloc(&ctx->bg_counter, 4));
  MTGV_CUDA_OK(ctx, cudaMemsetAsync(ctx->bg_counter, 0, 4, st));
  const int n_bands = (OH + kBgBandRows - 1) / kBgBandRows;
  int grid = ctx->sm_count * (ctx->bg_blocks_per_sm > 0 ? ctx->bg_blocks_per_sm : 1);
  if (grid > m * n_bands) grid = m * n_bands;
  k_background<<<grid, kBgThreads, sizeof(BgSmem), st>>>(params, m, n_bands, ctx->bg_planes, ctx->bg_off, bg_out, ctx->bg_counter);
  ctx->launches++;
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  return MTGV_OK;
}

}  // namespace mtgv
