// mtgv_det.cu - detection-scene generator for sm_100a (mtgvision/od_datasets.py).
//
//   k_det_sample  Philox tape of Gen.random / generate_synthetic_image draws        (production)
//   k_det_place   subsystem (4): per scene, sequential rejection-sampled placement with the
//                 shapely tests restated as fp64 convex clipping, homographies, warped keypoint
//                 labels, composite list and the per-scene pixel program                (mtgv_det.cuh)
//   k_det_pixels  subsystems (2)(3)(5): one thread per output pixel walks the scene program in
//                 registers - background cover-warp, pre-augments, every placed card whose
//                 footprint covers the pixel (warp of image + mask, alpha composite in reverse
//                 placement order), post-augments, cast - and writes the pixel ONCE.  The
//                 reference writes S*S pixels per card (od_datasets.py:594-599).  Only Gaussian
//                 blurs need neighbours: the program is cut at each blur into passes that
//                 ping-pong through a float32 scratch image.
#include <cuda_fp16.h>

#include <type_traits>

#include "mtgv_det.cuh"
#include "mtgv_internal.cuh"
#include "mtgv_persp.cuh"

namespace mtgv {

struct DetState {
  mtgv_det_config cfg{};
  bool set = false;
  mtgv_det_config* cfg_dev = nullptr;
  DetKeypoints kp{};
  DetKeypoints* kp_dev = nullptr;
  float* scratch[2] = {nullptr, nullptr};
  size_t scratch_cap = 0;  // floats per buffer
  int32_t* lists = nullptr;  // pass_count[kDetMaxBlur+1] followed by pass_list[kDetMaxBlur+1][chunk]
  size_t list_cap = 0;
  float* stats = nullptr;    // [chunk][kDetMaxBlur+1] image statistics for ISONoise boundaries
  // detection's own card copy: RGBA words [n][h][w], A = round_rect_mask(card_hw, 0.046) * 255 - image and mask
  // of make_card_with_mask (od_datasets.py:218-235) are warped by the same matrix, one 4-byte load per tap serves both
  uint32_t* card_rgba = nullptr;
  size_t rgba_cap = 0;
  uint64_t rgba_epoch = 0;
};

static DetState* det_state(mtgv_ctx* ctx) {
  if (!ctx->det) ctx->det = new DetState();
  return (DetState*)ctx->det;
}

int det_destroy(mtgv_ctx* ctx) {
  if (!ctx->det) return MTGV_OK;
  DetState* d = (DetState*)ctx->det;
  cudaFree(d->cfg_dev); cudaFree(d->kp_dev); cudaFree(d->scratch[0]); cudaFree(d->scratch[1]); cudaFree(d->lists); cudaFree(d->card_rgba);
  cudaFree(d->stats);
  delete d;
  ctx->det = nullptr;
  return MTGV_OK;
}

// ------------------------------------------------------------------------------------ //
// placement kernel                                                                      //
// ------------------------------------------------------------------------------------ //

// One warp per scene.  Cards are placed sequentially (card k is tested against cards 0..k-1,
// od_datasets.py:558-587) but the <= 10 attempts of a card are independent given the accepted
// set: lanes evaluate them in parallel and the lowest accepted attempt wins, which is exactly the
// reference's "first attempt that passes".  Labels / composite records are emitted one card per lane.
constexpr int kPlaceWarps = 4;

__global__ void __launch_bounds__(kPlaceWarps * 32) k_det_place(const mtgv_det_tape* __restrict__ tape, int n,
                                                                const mtgv_det_config* __restrict__ cfg,
                                                                const DetKeypoints* __restrict__ kp, int card_h, int card_w,
                                                                int n_cards, int n_bgs, const int32_t* __restrict__ bg_hw,
                                                                DetParams* params, int32_t* accepted, double* keypoints,
                                                                int32_t* labels, int32_t* counts) {
  __shared__ DetPlaceState s_state[kPlaceWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * kPlaceWarps + warp;
  if (s >= n) return;
  const mtgv_det_tape* t = &tape[s];
  DetParams* P = &params[s];
  DetPlaceState* st = &s_state[warp];
  const size_t nk = (size_t)MTGV_DET_MAX_CARDS * MTGV_DET_MAX_KPOLY;
  int32_t* acc = accepted + (size_t)s * MTGV_DET_MAX_CARDS;
  int32_t* lab = labels + (size_t)s * nk;
  double* kps = keypoints + (size_t)s * nk * MTGV_DET_MAX_KP * 2;
  acc[lane] = -1;
  for (int k = lane; k < (int)nk; k += 32) lab[k] = -1;
  bool ok = true;
  if (lane == 0) {
    counts[s] = 0;
    st->n_placed = 0;
    ok = det_scene_header(t, cfg, n_bgs, P);
    if (ok) det_bg_transform(t, cfg, bg_hw, P->bg_Minv);
  }
  ok = __shfl_sync(0xffffffffu, ok, 0);
  if (!ok) return;
  __syncwarp();
  const double min_edge = det_min_edge(cfg);
  const int n_scene_cards = t->bg_only ? 0 : t->n_cards;
  for (int ci = 0; ci < n_scene_cards; ci++) {
    const mtgv_det_card* c = &t->cards[ci];
    if (c->card < 0 || c->card >= n_cards) {
      if (lane == 0) P->status = MTGV_ERR_INVALID;
      return;
    }
    int na = c->n_attempts < cfg->max_attempts ? c->n_attempts : cfg->max_attempts;
    na = na < MTGV_DET_MAX_ATTEMPTS ? na : MTGV_DET_MAX_ATTEMPTS;
    double M[9];
    const bool pass = lane < na && det_try_attempt(&c->att[lane], cfg, kp, card_h, card_w, st, min_edge, M);
    const unsigned votes = __ballot_sync(0xffffffffu, pass);
    if (votes) {
      const int first = __ffs(votes) - 1;
      if (lane == first) {
        acc[ci] = first;
        det_commit(st, kp, M, c->card, ci);
      }
    }
    __syncwarp();
  }
  const int n_placed = st->n_placed;
  if (lane < n_placed) det_emit_card(t, cfg, kp, st, lane, P, kps, lab);
  if (lane == 0) {
    P->n_placed = n_placed;
    counts[s] = n_placed * kp->n_poly;
    det_emit_program(t, cfg, P);
  }
}

// ------------------------------------------------------------------------------------ //
// pixel helpers                                                                         //
// ------------------------------------------------------------------------------------ //

__device__ __forceinline__ float dclip01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

__device__ __forceinline__ float d_u8_over_255(uint32_t b) {  // np.divide(u8, 255.0, dtype=float32), exact
  const float f = __uint_as_float(0x4B000000u | b) - 8388608.f;
  const float r = 0.003921568859368563f;
  const float q = __fmul_rn(f, r);
  return __fmaf_rn(__fmaf_rn(-255.f, q, f), r, q);
}

__device__ __forceinline__ float d_bilinear(float v0, float v1, float v2, float v3, int ax, int ay) {
  const float fx = (float)ax * 0.03125f, fy = (float)ay * 0.03125f, gx = 1.f - fx, gy = 1.f - fy;
  float s = __fmul_rn(v0, gy * gx);
  s = __fadd_rn(s, __fmul_rn(v1, gy * fx));
  s = __fadd_rn(s, __fmul_rn(v2, fy * gx));
  return __fadd_rn(s, __fmul_rn(v3, fy * fx));
}

// cv2.warpPerspective coordinates of destination (x, y): block origin + guarded fast path (mtgv_persp.cuh)
__device__ __forceinline__ int2 d_persp(const double* M, int x, int y, int bw0, int bw_shift) {
  const int bxi = bw_shift >= 0 ? (x >> bw_shift) << bw_shift : (x / bw0) * bw0;
  double o[3];
  persp_origin(M, (double)bxi, (double)y, o);
  return persp_xy(o[0], o[1], o[2], M[0], M[3], M[6], (double)(x - bxi));
}

// bilinear blend of four uint8 taps [b00 b01 b10 b11] with the 1/32-px weights, exact in integers
// ((32-ax)(32-ay) ... sum to 1024), scaled to [0,1] once
__device__ __forceinline__ float d_bilinear_u8(uint32_t taps, int ax, int ay) {
  const unsigned pxw = (unsigned)(32 - ax) + ((unsigned)ax << 16);
  const unsigned s = __dp2a_hi(pxw * (unsigned)ay, taps, __dp2a_lo(pxw * (unsigned)(32 - ay), taps, 0u));
  return (float)s * (1.0f / (1024.0f * 255.0f));
}

// cv2.cvtColor(float32 RGB -> HSV): H in degrees [0,360), S, V in [0,1] (RGB2HSV_f, color_hsv.simd.hpp)
__device__ __forceinline__ void d_rgb2hsv(float r, float g, float b, float* h, float* s, float* v) {
  float vmax = fmaxf(r, fmaxf(g, b)), vmin = fminf(r, fminf(g, b));
  float diff = vmax - vmin;
  *v = vmax;
  *s = diff / (fabsf(vmax) + 1.1920929e-07f);
  diff = 60.f / (diff + 1.1920929e-07f);
  float hh;
  if (vmax == r) hh = (g - b) * diff;
  else if (vmax == g) hh = (b - r) * diff + 120.f;
  else hh = (r - g) * diff + 240.f;
  if (hh < 0.f) hh += 360.f;
  *h = hh;
}

// cv2.cvtColor(float32 HSV -> RGB) (HSV2RGB_native)
__device__ __forceinline__ void d_hsv2rgb(float h, float s, float v, float* r, float* g, float* b) {
  if (s == 0.f) { *r = *g = *b = v; return; }
  h *= (6.f / 360.f);
  if (h < 0.f) do h += 6.f; while (h < 0.f);
  else if (h >= 6.f) do h -= 6.f; while (h >= 6.f);
  int sector = (int)floorf(h);
  h -= (float)sector;
  if ((unsigned)sector >= 6u) { sector = 0; h = 0.f; }
  const float t0 = v, t1 = v * (1.f - s), t2 = v * (1.f - s * h), t3 = v * (1.f - s * (1.f - h));
  switch (sector) {  // sector_data {b,g,r} = {1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0}
    case 0: *b = t1; *g = t3; *r = t0; break;
    case 1: *b = t1; *g = t0; *r = t2; break;
    case 2: *b = t3; *g = t0; *r = t1; break;
    case 3: *b = t0; *g = t2; *r = t1; break;
    case 4: *b = t0; *g = t1; *r = t3; break;
    default: *b = t2; *g = t1; *r = t0; break;
  }
}

// cv2.cvtColor(float32 RGB -> HLS): H in degrees [0,360), L, S in [0,1] (RGB2HLS_f, color_hsv.simd.hpp)
__device__ __forceinline__ void d_rgb2hls(float r, float g, float b, float* h, float* l, float* s) {
  const float vmax = fmaxf(r, fmaxf(g, b)), vmin = fminf(r, fminf(g, b));
  float diff = vmax - vmin;
  *l = (vmax + vmin) * 0.5f;
  if (diff > 1.1920929e-07f) {
    *s = *l < 0.5f ? diff / (vmax + vmin) : diff / (2.f - vmax - vmin);
    diff = 60.f / diff;
    float hh;
    if (vmax == r) hh = (g - b) * diff;
    else if (vmax == g) hh = (b - r) * diff + 120.f;
    else hh = (r - g) * diff + 240.f;
    if (hh < 0.f) hh += 360.f;
    *h = hh;
  } else {
    *h = 0.f; *s = 0.f;
  }
}

// cv2.cvtColor(float32 HLS -> RGB) (HLS2RGB_f)
__device__ __forceinline__ void d_hls2rgb(float h, float l, float s, float* r, float* g, float* b) {
  if (s == 0.f) { *r = *g = *b = l; return; }
  const float p2 = l <= 0.5f ? l * (1.f + s) : l + s - l * s;
  const float p1 = 2.f * l - p2;
  h *= (6.f / 360.f);
  if (h < 0.f) do h += 6.f; while (h < 0.f);
  else if (h >= 6.f) do h -= 6.f; while (h >= 6.f);
  int sector = (int)floorf(h);
  h -= (float)sector;
  if ((unsigned)sector >= 6u) { sector = 0; h = 0.f; }
  const float t0 = p2, t1 = p1, t2 = p1 + (p2 - p1) * (1.f - h), t3 = p1 + (p2 - p1) * h;
  switch (sector) {  // sector_data {b,g,r} = {1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0}
    case 0: *b = t1; *g = t3; *r = t0; break;
    case 1: *b = t1; *g = t0; *r = t2; break;
    case 2: *b = t3; *g = t0; *r = t1; break;
    case 3: *b = t0; *g = t2; *r = t1; break;
    case 4: *b = t0; *g = t1; *r = t3; break;
    default: *b = t2; *g = t1; *r = t0; break;
  }
}

// Poisson(lam) by inversion of the cumulative distribution (production fields; lam <= ~64 on this path)
__device__ __forceinline__ float d_poisson(float lam, float u) {
  if (!(lam > 0.f)) return 0.f;
  float p = __expf(-lam), F = p;
  int k = 0;
  while (u > F && k < 512) {
    k++;
    p *= lam / (float)k;
    F += p;
  }
  return (float)k;
}

__device__ __forceinline__ void d_philox(uint64_t seed, int slot, uint32_t idx, uint32_t sub, uint32_t* r) {
  Philox ph;
  ph.key[0] = (uint32_t)seed ^ (0x85EBCA6Bu * (uint32_t)(slot + 1));
  ph.key[1] = (uint32_t)(seed >> 32) ^ 0x64657421u;
  ph(idx, sub, 0x6d746776u, 0, r);
}
__device__ __forceinline__ float d_unit(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

// pointwise photometric op on one RGB pixel at (y, x) of an image of width W (od_datasets.py:420-512)
__device__ __forceinline__ void d_photo_point(const DetPhotoX& op, float* rgb, int y, int x, int W, uint64_t seed,
                                              const uint32_t* __restrict__ fields) {
  switch (op.code) {
    case MTGV_PH_RBC:  // RandomBrightnessContrast on float32: clip(img*alpha + beta)
#pragma unroll
      for (int c = 0; c < 3; c++) rgb[c] = dclip01(__fadd_rn(__fmul_rn(rgb[c], op.f[0]), op.f[1]));
      break;
    case MTGV_PH_HSV: {  // HueSaturationValue through cv2's RGB<->HSV float conversion
      float h, s, v;
      d_rgb2hsv(rgb[0], rgb[1], rgb[2], &h, &s, &v);
      h = h + op.f[0];
      h = h - 360.f * floorf(h / 360.f);  // np.mod(h, 360)
      s = dclip01(s + op.f[1]);
      v = dclip01(v + op.f[2]);
      d_hsv2rgb(h, s, v, &rgb[0], &rgb[1], &rgb[2]);
      break;
    }
    case MTGV_PH_GAUSS_NOISE: {  // GaussNoise: clip(img + sigma * N(0,1))
      const uint32_t p = (uint32_t)(y * W + x);
      float g[3];
      if (op.field != MTGV_FIELD_PHILOX) {
        const float* f = (const float*)(fields + op.field) + (size_t)p * 3;
        g[0] = f[0]; g[1] = f[1]; g[2] = f[2];
      } else {
        uint32_t r[4];
        d_philox(seed, op.slot, p, 0, r);  // one call: two Box-Muller pairs, three of the four normals used
        const float rad0 = sqrtf(-2.f * __logf(((float)(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f)));
        const float rad1 = sqrtf(-2.f * __logf(((float)(r[2] >> 8) + 0.5f) * (1.0f / 16777216.0f)));
        float sn, cs;
        sincospif(2.f * d_unit(r[1]), &sn, &cs);
        g[0] = rad0 * cs; g[1] = rad0 * sn; g[2] = rad1 * cospif(2.f * d_unit(r[3]));
      }
#pragma unroll
      for (int c = 0; c < 3; c++) rgb[c] = dclip01(__fadd_rn(rgb[c], __fmul_rn(g[c], op.f[0])));
      break;
    }
    case MTGV_PH_SHOT_NOISE: {  // ShotNoise: photon noise in linear light (gamma 2.2), Poisson counts * scale, back to gamma
      const uint32_t p = (uint32_t)(y * W + x);
      float cnt[3];
      if (op.field != MTGV_FIELD_PHILOX) {
        const float* f = (const float*)(fields + op.field) + (size_t)p * 3;
        cnt[0] = f[0]; cnt[1] = f[1]; cnt[2] = f[2];
      } else {
        uint32_t r[4];
        d_philox(seed, op.slot, p, 5, r);
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const float lin = powf(dclip01(rgb[c]), 2.2f);
          cnt[c] = d_poisson(__fdiv_rn(__fadd_rn(lin, __fmul_rn(op.f[0], 1e-6f)), op.f[0]), d_unit(r[c]));
        }
      }
#pragma unroll
      for (int c = 0; c < 3; c++) rgb[c] = powf(dclip01(__fmul_rn(cnt[c], op.f[0])), 1.0f / 2.2f);
      break;
    }
    case MTGV_PH_ERASE: {  // Erasing: rectangle fill
      const int ty = y - op.i[0], tx = x - op.i[1];
      if ((unsigned)ty < (unsigned)op.i[2] && (unsigned)tx < (unsigned)op.i[3]) {
        if (op.i[4] == 0) {
          const uint32_t p = (uint32_t)(ty * op.i[3] + tx);
          if (op.field != MTGV_FIELD_PHILOX) {
            const float* f = (const float*)(fields + op.field) + (size_t)p * 3;
            rgb[0] = f[0]; rgb[1] = f[1]; rgb[2] = f[2];
          } else {
            uint32_t r[4];
            d_philox(seed, op.slot, p, 2, r);
            rgb[0] = d_unit(r[0]); rgb[1] = d_unit(r[1]); rgb[2] = d_unit(r[2]);
          }
        } else {
          rgb[0] = op.f[0]; rgb[1] = op.f[1]; rgb[2] = op.f[2];
        }
      }
      break;
    }
    default:
      break;
  }
}

// ISONoise on one pixel; std_l = cv2.meanStdDev(HLS image)[1] of the image the op receives
__device__ __forceinline__ void d_iso_point(const DetPhotoX& op, float* rgb, int y, int x, int W, uint64_t seed,
                                            const uint32_t* __restrict__ fields, float std_l) {
  float h, l, s;
  d_rgb2hls(rgb[0], rgb[1], rgb[2], &h, &l, &s);
  const uint32_t p = (uint32_t)(y * W + x);
  float lum, col;
  if (op.field != MTGV_FIELD_PHILOX) {
    const float* f = (const float*)(fields + op.field) + (size_t)p * 2;
    lum = f[0]; col = f[1];
  } else {
    uint32_t r[4];
    d_philox(seed, op.slot, p, 4, r);
    const float rad = sqrtf(-2.f * __logf(((float)(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f)));
    col = rad * cospif(2.f * d_unit(r[1]));
    lum = d_poisson(__fmul_rn(std_l, op.f[1]), d_unit(r[2]));
  }
  float hue = fmodf(__fadd_rn(h, __fmul_rn(col, op.f[0])), 360.f);  // np.mod(x, 360): sign of the divisor
  if (hue < 0.f) hue += 360.f;
  l = __fadd_rn(l, __fmul_rn(__fdiv_rn(lum, 255.f), __fsub_rn(1.f, l)));
  d_hls2rgb(hue, l, s, &rgb[0], &rgb[1], &rgb[2]);
#pragma unroll
  for (int c = 0; c < 3; c++) rgb[c] = dclip01(rgb[c]);
}

__global__ void k_det_rgba(const uint8_t* __restrict__ planes, int pitch, const float* __restrict__ mask, uint32_t* __restrict__ rgba,
                           int h, int w) {
  const size_t card = blockIdx.y;
  const uint8_t* src = planes + card * 3 * (size_t)h * pitch;
  uint32_t* dst = rgba + card * (size_t)h * w;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < h * w; i += gridDim.x * blockDim.x) {
    const int y = i / w, x = i - y * w;
    const uint8_t* p = src + (size_t)y * pitch + x;
    const uint32_t a = mask[i] != 0.f ? 255u : 0u;  // the mask is {0,1}-valued (cv::circle fill)
    dst[i] = (uint32_t)p[0] | ((uint32_t)p[(size_t)h * pitch] << 8) | ((uint32_t)p[2 * (size_t)h * pitch] << 16) | (a << 24);
  }
}

// ------------------------------------------------------------------------------------ //
// pixel kernel                                                                          //
// ------------------------------------------------------------------------------------ //

constexpr int kDetTW = 32, kDetTH = 32;   // pixels per tile
constexpr int kDetBW = 32, kDetBH = 8;    // threads per CTA; each thread owns kDetTH / kDetBH rows of the tile
constexpr int kDetRows = kDetTH / kDetBH;
constexpr int kDetHalo = kBlurHalfMax - 1;

struct DetLaunch {
  const DetParams* params;
  int n, pass;
  const int32_t* pass_count;  // [kDetMaxBlur + 1] scenes that have a pass p (pass 0: all scenes, no list)
  const int32_t* pass_list;   // [kDetMaxBlur + 1][n]
  const uint32_t* card_rgba;  // [n][h][w] RGBA words, A = mask * 255
  int card_h, card_w;
  const uint8_t* bg_planes;
  const int64_t* bg_off;
  const int32_t* bg_hw;
  const float* src;   // scratch written by the previous pass  [n,S,S,3] float32
  float* dst;         // scratch for the next pass
  float* stats;       // [n][kDetMaxBlur + 1]: std of the HLS L channel of `src`, for scenes whose boundary op is ISONoise
  void* out;
  int out_dtype;
  const uint32_t* fields;
};

struct DetTileSmem {
  int32_t status, bg, n_prog, n_blur, n_cull, first, last, blur_idx, is_final, has_cards, _p0, _p1;
  double bg_Minv[9];
  uint64_t seed;
  int32_t cull[MTGV_DET_MAX_CARDS];
  DetPhotoX prog[kDetProgMax];
  float halo[(kDetTH + 2 * kDetHalo) * (kDetTW + 2 * kDetHalo) * 3];
  float hrow[(kDetTH + 2 * kDetHalo) * kDetTW * 3];
};

// scenes that still have work in pass p (one Gaussian blur per pass boundary)
__global__ void k_det_lists(const DetParams* __restrict__ params, int n, int32_t* pass_count, int32_t* pass_list) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int nb = params[s].status == 0 ? params[s].n_blur : 0;
  for (int p = 1; p <= nb && p <= kDetMaxBlur; p++) pass_list[(size_t)p * n + atomicAdd(&pass_count[p], 1)] = s;
}

__device__ __forceinline__ void det_store(const DetLaunch& L, int s, int S_h, int S_w, int y, int x, const float* rgb, bool final_) {
  const size_t o = (size_t)y * S_w + x;
  if (!final_) {
    float* d = L.dst + ((size_t)s * S_h * S_w + o) * 3;
    d[0] = rgb[0]; d[1] = rgb[1]; d[2] = rgb[2];
    return;
  }
  const size_t plane = (size_t)S_h * S_w;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const size_t idx = ((size_t)s * 3 + c) * plane + o;
    if (L.out_dtype == MTGV_OUT_F16) ((__half*)L.out)[idx] = __float2half_rn(rgb[c]);
    else if (L.out_dtype == MTGV_OUT_U8) ((uint8_t*)L.out)[idx] = (uint8_t)(dclip01(rgb[c]) * 255.f);  // imwrite: (img*255).astype(uint8)
    else ((float*)L.out)[idx] = rgb[c];
  }
}

// cv2.medianBlur of the tile from the packed uint8 halo in T.halo (window origin = pixel): results to T.hrow[(ty*32+tx)*3+c]
__device__ __noinline__ void det_median_tile(DetTileSmem& T, int ks, int hw) {
          const uint2* H2 = reinterpret_cast<const uint2*>(T.halo);
#pragma unroll
          for (int i = 0; i < kDetRows; i++) {
            const uint2* c0 = H2 + (threadIdx.y + i * kDetBH) * hw + threadIdx.x;
            uint32_t med[3];
            if (ks == 3) {
              // 9 values per channel: the 19-exchange median network (Paeth), integer min / max
#pragma unroll
              for (int c = 0; c < 3; c++) {
                int v[9];
#pragma unroll
                for (int q = 0; q < 9; q++) {
                  const uint2 e = c0[(q / 3) * hw + (q % 3)];
                  v[q] = (int)(c == 0 ? (e.x & 0xFFFFu) : (c == 1 ? (e.x >> 16) : e.y));
                }
#define MTGV_CE(a, b) { const int lo_ = min(v[a], v[b]), hi_ = max(v[a], v[b]); v[a] = lo_; v[b] = hi_; }
                MTGV_CE(1, 2) MTGV_CE(4, 5) MTGV_CE(7, 8) MTGV_CE(0, 1) MTGV_CE(3, 4) MTGV_CE(6, 7) MTGV_CE(1, 2) MTGV_CE(4, 5) MTGV_CE(7, 8)
                MTGV_CE(0, 3) MTGV_CE(5, 8) MTGV_CE(4, 7) MTGV_CE(3, 6) MTGV_CE(1, 4) MTGV_CE(2, 5) MTGV_CE(4, 7) MTGV_CE(4, 2) MTGV_CE(6, 4) MTGV_CE(4, 2)
#undef MTGV_CE
                med[c] = (uint32_t)v[4];
              }
            } else {
              // smallest v with #(values <= v) > n / 2 by bisection on the byte, three channels at once: with bit 8 of a lane set,
              // (mid | 0x100) - x keeps that bit exactly when mid >= x, and the masked differences add up to the counts
              const int need = (ks * ks) / 2 + 1;
              int lo[3] = {0, 0, 0}, hi[3] = {255, 255, 255};
#pragma unroll 1
              for (int round = 0; round < 8; round++) {
                const int m0 = (lo[0] + hi[0]) >> 1, m1 = (lo[1] + hi[1]) >> 1, m2 = (lo[2] + hi[2]) >> 1;
                const uint32_t mrg = (uint32_t)m0 | ((uint32_t)m1 << 16) | 0x01000100u, mb = (uint32_t)m2 | 0x100u;
                uint32_t arg = 0u, ab = 0u;
                auto count = [&](auto KS) {  // window size as a compile-time constant: the element loop unrolls into plain loads
                  constexpr int K = decltype(KS)::value;
#pragma unroll
                  for (int dy = 0; dy < K; dy++) {
                    const uint2* row = c0 + dy * hw;
#pragma unroll
                    for (int dx = 0; dx < K; dx++) {
                      const uint2 e = row[dx];
                      arg += (mrg - e.x) & 0x01000100u;
                      ab += (mb - e.y) & 0x100u;
                    }
                  }
                };
                if (ks == 5) count(std::integral_constant<int, 5>{}); else count(std::integral_constant<int, 7>{});
                const int c0n = (int)((arg >> 8) & 0xFFu), c1n = (int)(arg >> 24), c2n = (int)(ab >> 8);
                if (c0n >= need) hi[0] = m0; else lo[0] = m0 + 1;
                if (c1n >= need) hi[1] = m1; else lo[1] = m1 + 1;
                if (c2n >= need) hi[2] = m2; else lo[2] = m2 + 1;
              }
              med[0] = (uint32_t)lo[0]; med[1] = (uint32_t)lo[1]; med[2] = (uint32_t)lo[2];
            }
            float* o = T.hrow + ((threadIdx.y + i * kDetBH) * kDetTW + threadIdx.x) * 3;
#pragma unroll
            for (int c = 0; c < 3; c++) o[c] = d_u8_over_255(med[c]);  // np.divide(u8, 255.0, dtype=float32)
          }
}

// one swap round of GlassBlur for the tile (see the call site): results to T.hrow[(ty*32+tx)*3+c]
__device__ __noinline__ void det_glass_tile(DetTileSmem& T, const DetLaunch& L, const DetPhotoX& bl, const float* __restrict__ img, int S_h,
                                            int S_w, int tx0, int ty0, int tid, int nt) {
  const int md = bl.i[0], rnd = bl.i[1], rounds = bl.i[2];
  const int nH = S_h - 2 * md;
  const int32_t* dxy = bl.field != MTGV_FIELD_PHILOX ? (const int32_t*)(L.fields + bl.field) : nullptr;
        // the draws of the tile and its halo, once: (dy + md) << 8 | (dx + md) per interior pixel, 0xFFFF outside the interior
        const int thw = kDetTW + 2 * md, thh = kDetTH + 2 * md;
        uint16_t* tab = reinterpret_cast<uint16_t*>(T.halo);
        for (int k = tid; k < thw * thh; k += nt) {
          const int h = ty0 - md + k / thw, w = tx0 - md + k % thw;
          uint16_t e = 0xFFFFu;
          if (h > md && h <= S_h - md && w > md && w <= S_w - md) {
            const uint32_t n = (uint32_t)((S_w - md - w) * nH + (S_h - md - h));  // albumentations' pixel order: columns, then rows, descending
            int dy, dx;
            if (dxy) {
              dy = dxy[((size_t)n * rounds + rnd) * 2]; dx = dxy[((size_t)n * rounds + rnd) * 2 + 1];
            } else {
              uint32_t r[4];
              d_philox(T.seed, bl.slot, n, 6 + rnd, r);
              dy = (int)(((uint64_t)r[0] * (uint32_t)(2 * md)) >> 32) - md;
              dx = (int)(((uint64_t)r[1] * (uint32_t)(2 * md)) >> 32) - md;
            }
            e = (uint16_t)(((dy + md) << 8) | (dx + md));
          }
          tab[k] = e;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kDetRows; i++) {
          const int x = tx0 + threadIdx.x, y = ty0 + threadIdx.y + i * kDetBH;
          float* o = T.hrow + ((threadIdx.y + i * kDetBH) * kDetTW + threadIdx.x) * 3;
          o[0] = o[1] = o[2] = 0.f;
          if (x >= S_w || y >= S_h) continue;
          const int ly = y - ty0 + md, lx = x - tx0 + md;  // this pixel in the table
          int sy = y, sx = x;
          bool found = false;
          // sources (h, w) = (y + a, x + b), a, b in (-md, md], whose draw is (-a, -b): smallest w first, then smallest h = the
          // last one in index order
          for (int b = -md + 1; b <= md && !found; b++)
            for (int a = -md + 1; a <= md; a++) {
              if (tab[(ly + a) * thw + lx + b] == (uint16_t)(((md - a) << 8) | (md - b))) { sy = y + a; sx = x + b; found = true; break; }
            }
          if (!found) {
            const uint16_t e = tab[ly * thw + lx];
            if (e != 0xFFFFu) { sy = y + (int)(e >> 8) - md; sx = x + (int)(e & 0xFFu) - md; }
          }
          const float* p = img + ((size_t)sy * S_w + sx) * 3;
          o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
        }
}

// cv2.meanStdDev(HLS image)[1] for the scenes of pass `L.pass` whose boundary op is ISONoise: population standard deviation of
// the L channel ((max + min) / 2 of RGB) over the previous pass's image, accumulated in double like cv2.  One CTA per scene.
__global__ void __launch_bounds__(1024) k_det_stats(DetLaunch L, int S_h, int S_w) {
  __shared__ double red[2][32];
  const int li = blockIdx.x;
  if (li >= L.pass_count[L.pass]) return;
  const int s = L.pass_list[(size_t)L.pass * L.n + li];
  const DetParams& P = L.params[s];
  int seen = 0, code = MTGV_PH_NONE;
  for (int k = 0; k < P.n_prog; k++)
    if (det_is_boundary(P.prog[k].code) && ++seen == L.pass) { code = P.prog[k].code; break; }
  if (code != MTGV_PH_ISO_NOISE) return;
  const float* img = L.src + (size_t)s * S_h * S_w * 3;
  const int n = S_h * S_w;
  double sum = 0.0, sq = 0.0;
  auto add = [&](float r, float g, float b) {
    const double l = (double)((fmaxf(r, fmaxf(g, b)) + fminf(r, fminf(g, b))) * 0.5f);
    sum += l; sq += l * l;
  };
  // four pixels = three 16-byte loads per iteration (the scratch image is 16-byte aligned per scene when S_h*S_w*3 is a
  // multiple of 4; otherwise the scalar tail loop takes everything)
  const bool vec = (((size_t)S_h * S_w * 3) & 3) == 0;
  const int n4 = vec ? n >> 2 : 0;
  const float4* img4 = reinterpret_cast<const float4*>(img);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const float4 a = img4[3 * (size_t)i], b = img4[3 * (size_t)i + 1], c = img4[3 * (size_t)i + 2];
    add(a.x, a.y, a.z); add(a.w, b.x, b.y); add(b.z, b.w, c.x); add(c.y, c.z, c.w);
  }
  for (int i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) add(img[3 * (size_t)i], img[3 * (size_t)i + 1], img[3 * (size_t)i + 2]);
  for (int o = 16; o; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sum; red[1][threadIdx.x >> 5] = sq; }
  __syncthreads();
  if (threadIdx.x < 32) {
    sum = threadIdx.x < (blockDim.x >> 5) ? red[0][threadIdx.x] : 0.0;
    sq = threadIdx.x < (blockDim.x >> 5) ? red[1][threadIdx.x] : 0.0;
    for (int o = 16; o; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); }
    if (threadIdx.x == 0) {
      const double mean = sum / n, var = sq / n - mean * mean;
      L.stats[(size_t)s * (kDetMaxBlur + 1) + L.pass] = (float)sqrt(var > 0.0 ? var : 0.0);
    }
  }
}

#ifndef MTGV_DET_BLOCKS
#define MTGV_DET_BLOCKS 4  // 64 registers per thread: the pixel walk is latency-bound, occupancy pays (measured 2 -> 4: +18 %)
#endif
__global__ void __launch_bounds__(kDetBW * kDetBH, MTGV_DET_BLOCKS) k_det_pixels(DetLaunch L, int S_h, int S_w) {
  extern __shared__ __align__(16) unsigned char det_smem_raw[];
  DetTileSmem& T = *reinterpret_cast<DetTileSmem*>(det_smem_raw);
  const int tid = threadIdx.y * kDetBW + threadIdx.x, nt = kDetBW * kDetBH;
  const int tiles_x = (S_w + kDetTW - 1) / kDetTW, tiles_y = (S_h + kDetTH - 1) / kDetTH, tiles = tiles_x * tiles_y;
  const int n_scenes = L.pass == 0 ? L.n : L.pass_count[L.pass];
  const int bw0 = persp_block_w(S_h, S_w);
  const int bw_shift = (bw0 & (bw0 - 1)) == 0 ? 31 - __clz(bw0) : -1;
  for (long long work = blockIdx.x; work < (long long)n_scenes * tiles; work += gridDim.x) {
    const int li = (int)(work / tiles), tile = (int)(work - (long long)li * tiles);
    const int s = L.pass == 0 ? li : L.pass_list[(size_t)L.pass * L.n + li];
    const DetParams& P = L.params[s];
    const int tx0 = (tile % tiles_x) * kDetTW, ty0 = (tile / tiles_x) * kDetTH;
    __syncthreads();
    if (tid == 0) {
      T.status = P.status; T.bg = P.bg; T.n_prog = P.n_prog; T.n_blur = P.n_blur; T.seed = P.seed;
      // program segment of this pass: after the pass-th blur (or from the start) up to the next blur
      int first = 0, seen = 0, blur_idx = -1;
      if (L.pass > 0) {
        first = P.n_prog;
        for (int k = 0; k < P.n_prog; k++)
          if (det_is_boundary(P.prog[k].code) && ++seen == L.pass) { first = k + 1; blur_idx = k; break; }
      }
      int last = P.n_prog, has_cards = 0;
      for (int k = first; k < P.n_prog; k++) {
        if (det_is_boundary(P.prog[k].code)) { last = k; break; }
        has_cards |= P.prog[k].code == kPhCards;
      }
      T.first = first; T.last = last; T.blur_idx = blur_idx; T.is_final = last == P.n_prog; T.has_cards = has_cards;
    } else if (tid >= 32 && tid < 41) {
      T.bg_Minv[tid - 32] = P.bg_Minv[tid - 32];
    }
    __syncthreads();
    const bool live_scene = T.status == 0 && L.pass <= T.n_blur;
    if (T.status != 0 && L.pass == 0) {  // failed scenes produce zeros (the host raises from the status word)
      const float z[3] = {0.f, 0.f, 0.f};
      for (int i = 0; i < kDetRows; i++) {
        const int x = tx0 + threadIdx.x, y = ty0 + threadIdx.y + i * kDetBH;
        if (x < S_w && y < S_h) det_store(L, s, S_h, S_w, y, x, z, true);
      }
    }
    if (!live_scene) continue;
    for (int k = tid; k < (int)(sizeof(DetPhotoX) / 4) * T.n_prog; k += nt) ((uint32_t*)T.prog)[k] = ((const uint32_t*)P.prog)[k];
    if (tid < 32) {  // cull the placed cards against the tile (ballot compaction keeps composite order)
      const bool hit = T.has_cards && tid < P.n_placed && P.cards[tid].x0 < tx0 + kDetTW && P.cards[tid].x1 > tx0 &&
                       P.cards[tid].y0 < ty0 + kDetTH && P.cards[tid].y1 > ty0;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) T.cull[__popc(m & ((1u << tid) - 1))] = tid;
      if (tid == 0) T.n_cull = __popc(m);
    }
    __syncthreads();

    float rgb[kDetRows][3];
    if (L.pass == 0) {
      // make_background: cv2.warpPerspective(bg, M, (S,S)) with the cover transform (od_datasets.py:195-203)
      const int bh = L.bg_hw[2 * T.bg], bw = L.bg_hw[2 * T.bg + 1], pitchw = (bw + 3) & ~3;
      const uint32_t* src = reinterpret_cast<const uint32_t*>(L.bg_planes + L.bg_off[T.bg]);  // RGBX words
#pragma unroll
      for (int i = 0; i < kDetRows; i++) {
        const int x = tx0 + threadIdx.x, y = ty0 + threadIdx.y + i * kDetBH;
        rgb[i][0] = rgb[i][1] = rgb[i][2] = 0.f;
        if (x >= S_w || y >= S_h) continue;
        const int2 XY = d_persp(T.bg_Minv, x, y, bw0, bw_shift);
        const int X = XY.x, Y = XY.y;
        const int sx = X >> 5, sy = Y >> 5;
        if ((unsigned)sx < (unsigned)(bw - 1) && (unsigned)sy < (unsigned)(bh - 1)) {
          // all four taps inside: RGBX words, channels blended exactly in integers
          const uint32_t* p = src + (size_t)sy * pitchw + sx;
          const uint32_t t0 = __ldg(p), t1 = __ldg(p + 1), t2 = __ldg(p + pitchw), t3 = __ldg(p + pitchw + 1);
          const unsigned rg_t = __byte_perm(t0, t1, 0x5140), rg_b = __byte_perm(t2, t3, 0x5140);
          const unsigned b_t = __byte_perm(t0, t1, 0x0062), b_b = __byte_perm(t2, t3, 0x0062);
          rgb[i][0] = d_bilinear_u8(__byte_perm(rg_t, rg_b, 0x5410), X & 31, Y & 31);
          rgb[i][1] = d_bilinear_u8(__byte_perm(rg_t, rg_b, 0x7632), X & 31, Y & 31);
          rgb[i][2] = d_bilinear_u8(__byte_perm(b_t, b_b, 0x5410), X & 31, Y & 31);
          continue;
        }
        float v[4][3];
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const int yy = sy + (t >> 1), xx = sx + (t & 1);
          const bool in = (unsigned)yy < (unsigned)bh && (unsigned)xx < (unsigned)bw;
          const uint32_t word = in ? __ldg(src + (size_t)yy * pitchw + xx) : 0u;
#pragma unroll
          for (int c = 0; c < 3; c++) v[t][c] = in ? d_u8_over_255((word >> (8 * c)) & 255u) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < 3; c++) rgb[i][c] = d_bilinear(v[0][c], v[1][c], v[2][c], v[3][c], X & 31, Y & 31);
      }
    } else {
      // the boundary op that ended the previous segment, evaluated over this tile from the previous pass's image
      const DetPhotoX& bl = T.prog[T.blur_idx];
      const float* img = L.src + (size_t)s * S_h * S_w * 3;
      if (bl.code == kPhGlassSwap) {
        // one swap round of GlassBlur as a gather: a pixel that some interior pixel targets takes that pixel's value (the
        // LAST such source in albumentations' order - columns descending, rows descending inside a column - wins); an
        // untargeted interior pixel takes the value at its own (dy, dx); border pixels that nobody targets stay
        det_glass_tile(T, L, bl, img, S_h, S_w, tx0, ty0, tid, nt);  // out of line like the median; results through T.hrow
#pragma unroll
        for (int i = 0; i < kDetRows; i++) {
          const float* o = T.hrow + ((threadIdx.y + i * kDetBH) * kDetTW + threadIdx.x) * 3;
          rgb[i][0] = o[0]; rgb[i][1] = o[1]; rgb[i][2] = o[2];
        }
      } else if (bl.code == MTGV_PH_ISO_NOISE) {
        const float std_l = L.stats[(size_t)s * (kDetMaxBlur + 1) + L.pass];
#pragma unroll
        for (int i = 0; i < kDetRows; i++) {
          const int x = tx0 + threadIdx.x, y = ty0 + threadIdx.y + i * kDetBH;
          rgb[i][0] = rgb[i][1] = rgb[i][2] = 0.f;
          if (x >= S_w || y >= S_h) continue;
          const float* p = img + ((size_t)y * S_w + x) * 3;
          rgb[i][0] = p[0]; rgb[i][1] = p[1]; rgb[i][2] = p[2];
          d_iso_point(bl, rgb[i], y, x, S_w, T.seed, L.fields, std_l);
        }
      } else {
        const int r = bl.i[0] / 2;
        const int hw = kDetTW + 2 * r, hh = kDetTH + 2 * r;
        const bool median = bl.code == MTGV_PH_MEDIAN_BLUR;
        for (int k = tid; k < hw * hh; k += nt) {
          int yy = ty0 - r + k / hw, xx = tx0 - r + k % hw;
          if (median) {  // cv2.medianBlur: BORDER_REPLICATE; the image as uint8 = rint(clip(x) * 255)
            yy = yy < 0 ? 0 : (yy >= S_h ? S_h - 1 : yy);
            xx = xx < 0 ? 0 : (xx >= S_w ? S_w - 1 : xx);
          } else {
            while (yy < 0 || yy >= S_h) yy = yy < 0 ? -yy : 2 * S_h - 2 - yy;  // reflect101
            while (xx < 0 || xx >= S_w) xx = xx < 0 ? -xx : 2 * S_w - 2 - xx;
          }
          const float* p = img + ((size_t)yy * S_w + xx) * 3;
          if (median) {
            // bytes in 16-bit lanes: word 0 = R | G << 16, word 1 = B (one 8-byte load per window element, carry-free lane arithmetic)
            const uint32_t r8 = (uint32_t)rintf(dclip01(p[0]) * 255.f), g8 = (uint32_t)rintf(dclip01(p[1]) * 255.f), b8 = (uint32_t)rintf(dclip01(p[2]) * 255.f);
            reinterpret_cast<uint2*>(T.halo)[k] = make_uint2(r8 | (g8 << 16), b8);
          } else {
            T.halo[3 * k] = p[0]; T.halo[3 * k + 1] = p[1]; T.halo[3 * k + 2] = p[2];
          }
        }
        __syncthreads();
        if (bl.code == MTGV_PH_GAUSS_BLUR) {
          // GaussianBlur: separable, BORDER_REFLECT_101, symmetric pairing
          for (int k = tid; k < hh * kDetTW; k += nt) {
            const int yy = k / kDetTW, xx = k % kDetTW;
            const float* c0 = T.halo + (yy * hw + xx + r) * 3;
#pragma unroll
            for (int c = 0; c < 3; c++) {
              float acc = c0[c] * bl.k[0];
              for (int j = 1; j <= r; j++) acc += (c0[c - 3 * j] + c0[c + 3 * j]) * bl.k[j];
              T.hrow[3 * k + c] = acc;
            }
          }
          __syncthreads();
#pragma unroll
          for (int i = 0; i < kDetRows; i++) {
            const float* c0 = T.hrow + ((threadIdx.y + i * kDetBH + r) * kDetTW + threadIdx.x) * 3;
#pragma unroll
            for (int c = 0; c < 3; c++) {
              float acc = c0[c] * bl.k[0];
              for (int j = 1; j <= r; j++) acc += (c0[c - 3 * j * kDetTW] + c0[c + 3 * j * kDetTW]) * bl.k[j];
              rgb[i][c] = acc;
            }
          }
        } else if (median) {
          det_median_tile(T, bl.i[0], hw);  // rare and register-hungry: kept out of line, results come back through T.hrow
#pragma unroll
          for (int i = 0; i < kDetRows; i++) {
            const float* o = T.hrow + ((threadIdx.y + i * kDetBH) * kDetTW + threadIdx.x) * 3;
            rgb[i][0] = o[0]; rgb[i][1] = o[1]; rgb[i][2] = o[2];
          }
        } else {
          // MotionBlur: cv2.filter2D (correlation, anchor at the centre) with the normalised line kernel
          const int ks = bl.i[0];
          const float wgt = bl.f[0];
#pragma unroll
          for (int i = 0; i < kDetRows; i++) {
            const float* c0 = T.halo + ((threadIdx.y + i * kDetBH) * hw + threadIdx.x) * 3;
            float acc[3] = {0.f, 0.f, 0.f};
            for (int wi = 0; wi < 4; wi++) {  // the set cells in row-major order (cv2.filter2D's coefficient order)
              uint32_t bits = (uint32_t)bl.i[1 + wi];
              while (bits) {
                const int b = 32 * wi + __ffs((int)bits) - 1;
                bits &= bits - 1;
                const int dy = b / ks, dx = b - dy * ks;
                const float* q = c0 + (dy * hw + dx) * 3;
#pragma unroll
                for (int c = 0; c < 3; c++) acc[c] = __fadd_rn(acc[c], __fmul_rn(q[c], wgt));
              }
            }
            rgb[i][0] = acc[0]; rgb[i][1] = acc[1]; rgb[i][2] = acc[2];
          }
        }
      }
    }

    const int ch = L.card_h, cw = L.card_w;
#pragma unroll
    for (int i = 0; i < kDetRows; i++) {
      const int x = tx0 + threadIdx.x, y = ty0 + threadIdx.y + i * kDetBH;
      if (x >= S_w || y >= S_h) continue;
      float* px_rgb = rgb[i];
      for (int k = T.first; k < T.last; k++) {
        const DetPhotoX& op = T.prog[k];
        if (op.code != kPhCards) {
          d_photo_point(op, px_rgb, y, x, S_w, T.seed, L.fields);
          continue;
        }
        // composite every card covering this pixel, reverse placement order (od_datasets.py:594-599)
        for (int q = 0; q < T.n_cull; q++) {
          const DetCardX& c = P.cards[T.cull[q]];
          if (x < c.x0 || x >= c.x1 || y < c.y0 || y >= c.y1) continue;
          const int2 XY = d_persp(c.Minv, x, y, bw0, bw_shift);
          const int X = XY.x, Y = XY.y;
          const int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
          if (sx < -1 || sx >= cw || sy < -1 || sy >= ch) continue;  // all four taps outside: mask = 0
          const bool x0 = sx >= 0, x1 = sx + 1 < cw, y0 = sy >= 0, y1 = sy + 1 < ch;
          const uint32_t* tp = L.card_rgba + (size_t)c.card * ch * cw + (long long)sy * cw + sx;
          const uint32_t t0 = (y0 && x0) ? __ldg(tp) : 0u, t1 = (y0 && x1) ? __ldg(tp + 1) : 0u;
          const uint32_t t2 = (y1 && x0) ? __ldg(tp + cw) : 0u, t3 = (y1 && x1) ? __ldg(tp + cw + 1) : 0u;
          const int ax = X & 31, ay = Y & 31;
          // mask: {0,255} bytes blended exactly; taps outside the card are 0 like the warp's constant border
          const unsigned mtaps = (__byte_perm(__byte_perm(t0, t1, 0x0073), __byte_perm(t2, t3, 0x0073), 0x5410) >> 7) & 0x01010101u;
          const unsigned pxw = (unsigned)(32 - ax) + ((unsigned)ax << 16);
          const float m = (float)__dp2a_hi(pxw * (unsigned)ay, mtaps, __dp2a_lo(pxw * (unsigned)(32 - ay), mtaps, 0u)) * (1.0f / 1024.0f);  // exact
          if (m == 0.f) continue;  // mask*img + (1-mask)*bg == bg exactly
          const float im = __fsub_rn(1.f, m);
          if (c.n_ops == 0) {
            // no per-texel card augmentation: integer bilinear per channel (outside taps are 0, as in the float path)
            const unsigned rg_t = __byte_perm(t0, t1, 0x5140), rg_b = __byte_perm(t2, t3, 0x5140);
            const unsigned b_t = __byte_perm(t0, t1, 0x0062), b_b = __byte_perm(t2, t3, 0x0062);
            const float wr = d_bilinear_u8(__byte_perm(rg_t, rg_b, 0x5410), ax, ay);
            const float wg = d_bilinear_u8(__byte_perm(rg_t, rg_b, 0x7632), ax, ay);
            const float wb = d_bilinear_u8(__byte_perm(b_t, b_b, 0x5410), ax, ay);
            px_rgb[0] = __fadd_rn(__fmul_rn(m, wr), __fmul_rn(im, px_rgb[0]));
            px_rgb[1] = __fadd_rn(__fmul_rn(m, wg), __fmul_rn(im, px_rgb[1]));
            px_rgb[2] = __fadd_rn(__fmul_rn(m, wb), __fmul_rn(im, px_rgb[2]));
            continue;
          }
          float v[4][3];
          const uint32_t tw[4] = {t0, t1, t2, t3};
#pragma unroll
          for (int t = 0; t < 4; t++) {
            const int yy = sy + (t >> 1), xx = sx + (t & 1);
            const bool in = (t >> 1 ? y1 : y0) && (t & 1 ? x1 : x0);
            if (in) {
              float px[3];
#pragma unroll
              for (int cc = 0; cc < 3; cc++) px[cc] = d_u8_over_255((tw[t] >> (8 * cc)) & 255u);
              for (int o = 0; o < c.n_ops; o++) d_photo_point(c.ops[o], px, yy, xx, cw, T.seed, L.fields);  // pre_transform_card (:581)
              v[t][0] = px[0]; v[t][1] = px[1]; v[t][2] = px[2];
            } else {
              v[t][0] = v[t][1] = v[t][2] = 0.f;
            }
          }
#pragma unroll
          for (int cc = 0; cc < 3; cc++) {
            const float w = d_bilinear(v[0][cc], v[1][cc], v[2][cc], v[3][cc], ax, ay);
            px_rgb[cc] = __fadd_rn(__fmul_rn(m, w), __fmul_rn(im, px_rgb[cc]));
          }
        }
      }
      det_store(L, s, S_h, S_w, y, x, px_rgb, T.is_final != 0);
    }
  }
}

// ------------------------------------------------------------------------------------ //
// production sampler                                                                    //
// ------------------------------------------------------------------------------------ //

struct DRng {
  Philox ph;
  uint32_t c0, c1, ctr, buf[4];
  int have;
  __device__ DRng(uint64_t seed, uint64_t index) {
    ph.key[0] = (uint32_t)seed ^ 0x64657400u; ph.key[1] = (uint32_t)(seed >> 32);
    c0 = (uint32_t)index; c1 = (uint32_t)(index >> 32); ctr = 0; have = 0;
  }
  __device__ uint32_t u32() { if (!have) { ph(c0, c1, 7u, ctr++, buf); have = 4; } return buf[--have]; }
  __device__ double uniform() {
    uint32_t a = u32() >> 5, b = u32() >> 6;
    return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
  }
  __device__ double uniform(double lo, double hi) { return lo + (hi - lo) * uniform(); }
  __device__ int below(int n) { return (int)(((uint64_t)u32() * (uint32_t)n) >> 32); }
};

__device__ void ph_init(mtgv_photo_op* o, int code) {
  o->code = code;
  for (int k = 0; k < 5; k++) o->i[k] = 0;
  for (int k = 0; k < 3; k++) o->d[k] = 0.0;
  o->field = MTGV_FIELD_PHILOX;
}
__device__ int ph_rbc(DRng& r, mtgv_photo_op* o, double p, double b, double c0, double c1) {
  if (!(r.uniform() < p)) return 0;
  ph_init(o, MTGV_PH_RBC);
  o->d[0] = 1.0 + r.uniform(c0, c1);
  o->d[1] = r.uniform(-b, b);
  return 1;
}
__device__ int ph_hsv(DRng& r, mtgv_photo_op* o, double p, double val) {
  if (!(r.uniform() < p)) return 0;
  ph_init(o, MTGV_PH_HSV);
  o->d[0] = r.uniform(-30, 30); o->d[1] = r.uniform(-40, 40); o->d[2] = r.uniform(-val, val);
  return 1;
}
__device__ int ph_noise(DRng& r, mtgv_photo_op* o, double p, double smax) {
  if (!(r.uniform() < p)) return 0;
  ph_init(o, MTGV_PH_GAUSS_NOISE);
  o->d[0] = r.uniform(0.0, smax);
  return 1;
}
__device__ int ph_blur(DRng& r, mtgv_photo_op* o, double p, double smax) {
  if (!(r.uniform() < p)) return 0;
  ph_init(o, MTGV_PH_GAUSS_BLUR);
  o->d[0] = r.uniform(0.0, smax);
  return 1;
}
__device__ int ph_erase(DRng& r, mtgv_photo_op* o, double p, double s0, double s1, int fill, int h, int w) {
  if (!(r.uniform() < p)) return 0;
  const double area = r.uniform(s0, s1) * h * w, aspect = exp(r.uniform(log(0.3), log(3.3)));
  const int eh = (int)rint(sqrt(area * aspect)), ew = (int)rint(sqrt(area / aspect));
  if (eh < 1 || ew < 1 || eh >= h || ew >= w) return 0;
  ph_init(o, MTGV_PH_ERASE);
  o->i[0] = r.below(h - eh + 1); o->i[1] = r.below(w - ew + 1); o->i[2] = eh; o->i[3] = ew; o->i[4] = fill;
  if (fill == 1) for (int k = 0; k < 3; k++) o->d[k] = r.uniform();
  return 1;
}
__device__ int ph_fill(DRng& r, double p0, double p1, double p2) {  // np.random.choice(4, p)
  const double u = r.uniform();
  return u < p0 ? 0 : (u < p0 + p1 ? 1 : (u < p0 + p1 + p2 ? 2 : 3));
}
__device__ void ph_perm(DRng& r, int* idx, int n) {
  for (int i = 0; i < n; i++) idx[i] = i;
  for (int i = n - 1; i >= 1; i--) { int j = r.below(i + 1); int t = idx[i]; idx[i] = idx[j]; idx[j] = t; }
}
// one_of(noise family) / one_of(blur family) of get_bg_transform (od_datasets.py:443-457)
__device__ int ph_noise_family(DRng& r, mtgv_photo_op* o, double p) {
  const int c = r.below(3);
  if (c == 0) return ph_noise(r, o, p, 0.2);
  if (!(r.uniform() < p)) return 0;
  if (c == 1) {  // ISONoise(color_shift=(0.01, 0.4)), default intensity=(0.1, 0.5)
    ph_init(o, MTGV_PH_ISO_NOISE);
    o->d[0] = r.uniform(0.01, 0.4); o->d[1] = r.uniform(0.1, 0.5);
  } else {       // ShotNoise(scale_range=(0.1, 0.3))
    ph_init(o, MTGV_PH_SHOT_NOISE);
    o->d[0] = r.uniform(0.1, 0.3);
  }
  return 1;
}
__device__ int ph_blur_family(DRng& r, mtgv_photo_op* o, double p) {
  const int c = r.below(5);
  if (c == 0) return ph_blur(r, o, p, 3.0);
  if (c == 4) {  // GlassBlur(sigma=0.5, max_delta=4, iterations=2, p = p * 2 / 3)
    if (!(r.uniform() < p / 3 * 2)) return 0;
    ph_init(o, MTGV_PH_GLASS_BLUR);
    o->d[0] = 0.5; o->i[0] = 4; o->i[1] = 2;
    return 1;
  }
  if (!(r.uniform() < p)) return 0;
  if (c == 1) {  // MedianBlur(blur_limit=(3, 7))
    ph_init(o, MTGV_PH_MEDIAN_BLUR);
    o->i[0] = 3 + 2 * r.below(3);
    return 1;
  }
  // MotionBlur(blur_limit=(3, 11)), twice in the family
  ph_init(o, MTGV_PH_MOTION_BLUR);
  const int ks = 3 + 2 * r.below(5);
  const int x1 = r.below(ks), x2 = r.below(ks);
  int y1, y2;
  if (x1 == x2) { y1 = r.below(ks); y2 = r.below(ks - 1); y2 += y2 >= y1; }
  else { y1 = r.below(ks); y2 = r.below(ks); }
  o->i[0] = ks;
  det_line_mask(ks, x1, y1, x2, y2, &o->i[1]);
  return 1;
}

// Gen._get_bg_ds + ran_path (od_datasets.py:656-672): the dataset with probability p, then an image uniformly inside it
__device__ __forceinline__ int det_draw_bg(DRng& r, const mtgv_det_config* cfg, int n_bgs) {
  const int nf = cfg->n_bgs_first;
  if (nf <= 0 || nf >= n_bgs) return r.below(n_bgs);
  if (r.uniform() < cfg->bg_first_prob) return r.below(nf);
  return nf + r.below(n_bgs - nf);
}

__global__ void k_det_sample(uint64_t seed, int64_t first, int n, const mtgv_det_config* __restrict__ cfg, int n_cards_pool,
                             int n_bgs, int card_h, int card_w, mtgv_det_tape* tape) {
  // one warp per scene, lane 0 working: scenes take divergent paths through the augmentation graphs
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n || (threadIdx.x & 31) != 0) return;
  mtgv_det_tape* t = &tape[i];
  DRng r(seed, (uint64_t)(first + i));
  const int S_h = cfg->size_h, S_w = cfg->size_w;
  t->seed = seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(first + i + 1));
  t->bg_ab_given = 0; t->_pad = 0; t->bg_ab[0] = t->bg_ab[1] = 0.0;
  t->n_pre = t->n_post = 0; t->n_cards = 0;
  t->bg_only = cfg->ratio_bg > 0.0 && r.uniform() < cfg->ratio_bg;  // Gen.random (:686)
  const bool ph = cfg->photometrics != 0;
  int idx[8];
  if (t->bg_only) {  // random_bg / make_aug_background (:206-210)
    t->bg = det_draw_bg(r, cfg, n_bgs);
    t->bg_deg = r.below(360);
    const int fl = ph_fill(r, 0.1, 0.5, 0.2);
    if (ph) {
      ph_perm(r, idx, 4);
      for (int q = 0; q < 3; q++) {
        mtgv_photo_op* o = &t->pre[t->n_pre];
        switch (idx[q]) {
          case 0: t->n_pre += ph_rbc(r, o, 0.5, 0.4, -0.4, 0.4); break;
          case 1: t->n_pre += ph_blur(r, o, 0.2, 2.0); break;
          case 2: t->n_pre += ph_noise(r, o, 0.2, 0.1); break;
          default: t->n_pre += ph_erase(r, o, 0.4, 0.02, 0.2, fl, S_h, S_w); break;
        }
      }
    }
    const int fe = ph_fill(r, 0.1, 0.1, 0.4);
    if (ph) {
      ph_perm(r, idx, 7);
      for (int q = 0; q < 4; q++) {
        mtgv_photo_op* o = &t->post[t->n_post];
        switch (idx[q]) {
          case 0: t->n_post += ph_rbc(r, o, 0.5, 0.7, -0.5, 0.5); break;
          case 1: t->n_post += ph_hsv(r, o, 0.5, 30.0); break;
          case 2: t->n_post += ph_noise_family(r, o, 0.5); break;
          case 3: t->n_post += ph_blur_family(r, o, 0.5); break;
          case 4: t->n_post += ph_noise_family(r, o, 0.1); break;
          case 5: t->n_post += ph_blur_family(r, o, 0.1); break;
          default: t->n_post += ph_erase(r, o, 0.4, 0.02, 0.4, fe, S_h, S_w); break;
        }
      }
    }
    return;
  }
  // generate_synthetic_image (:520-611)
  const int fill_light = ph_fill(r, 0.1, 0.5, 0.2), fill_card = ph_fill(r, 0.1, 0.5, 0.2);
  t->bg = det_draw_bg(r, cfg, n_bgs);
  t->bg_deg = r.below(360);
  if (ph) {
    ph_perm(r, idx, 4);
    for (int q = 0; q < 3; q++) {
      mtgv_photo_op* o = &t->pre[t->n_pre];
      switch (idx[q]) {
        case 0: t->n_pre += ph_rbc(r, o, 0.5, 0.4, -0.4, 0.4); break;
        case 1: t->n_pre += ph_blur(r, o, 0.2, 2.0); break;
        case 2: t->n_pre += ph_noise(r, o, 0.2, 0.1); break;
        default: t->n_pre += ph_erase(r, o, 0.4, 0.02, 0.2, fill_light, S_h, S_w); break;
      }
    }
  }
  const int span = cfg->num_cards_max - cfg->num_cards_min;
  t->n_cards = cfg->num_cards_min + (span > 0 ? r.below(span) : 0);
  if (t->n_cards > MTGV_DET_MAX_CARDS) t->n_cards = MTGV_DET_MAX_CARDS;
  double edge = cfg->min_visible_edges < 0.0 ? cfg->min_visible : cfg->min_visible_edges;
  if (edge < cfg->min_visible) edge = cfg->min_visible;
  const int diag = (int)sqrt((double)card_h * card_h + (double)card_w * card_w);
  const int pad = diag / 2, ovr = (int)(diag * (1.0 - edge));
  const int lox = pad - ovr, hix = S_w - pad + ovr, loy = pad - ovr, hiy = S_h - pad + ovr;
  const double amin = (double)S_h * S_w * cfg->min_area_ratio, amax = (double)S_h * S_w * cfg->max_area_ratio;
  const double la = log(amin), lb = log(amax);
  for (int ci = 0; ci < t->n_cards; ci++) {
    mtgv_det_card* c = &t->cards[ci];
    c->card = r.below(n_cards_pool);
    c->n_attempts = cfg->max_attempts < MTGV_DET_MAX_ATTEMPTS ? cfg->max_attempts : MTGV_DET_MAX_ATTEMPTS;
    c->n_photo = 0; c->_pad = 0;
    for (int ai = 0; ai < c->n_attempts; ai++) {
      mtgv_det_attempt* a = &c->att[ai];
      a->cx = lox + r.below(hix - lox + 1);  // random.randint is inclusive
      a->cy = loy + r.below(hiy - loy + 1);
      a->dst_given = 0; a->_pad = 0;
      a->deg = r.uniform(0.0, 360.0);
      a->area = cfg->size_sample_mode == 1 ? r.uniform(amin, amax) : exp(r.uniform(la, lb));  // place_card_on_background_get_transform (:327-332)
      for (int k = 0; k < 4; k++) a->jitter[k] = r.uniform(1.0 - cfg->jitter_ratio, 1.0 + cfg->jitter_ratio);
      for (int k = 0; k < 8; k++) a->dst[k] = 0.f;
    }
    if (ph) {  // get_card_transform (:490-512): 2 of 3
      ph_perm(r, idx, 3);
      for (int q = 0; q < 2; q++) {
        mtgv_photo_op* o = &c->photo[c->n_photo];
        switch (idx[q]) {
          case 0: c->n_photo += ph_rbc(r, o, 0.8, 0.2, -0.4, -0.4); break;
          case 1: c->n_photo += ph_hsv(r, o, 0.8, 0.0); break;
          default: c->n_photo += ph_erase(r, o, 0.3, 0.02, 0.2, fill_card, card_h, card_w); break;
        }
      }
    }
  }
  if (ph) {  // get_bg_transform (:441-487): 4 of 6
    ph_perm(r, idx, 6);
    for (int q = 0; q < 4; q++) {
      mtgv_photo_op* o = &t->post[t->n_post];
      switch (idx[q]) {
        case 0: t->n_post += ph_rbc(r, o, 0.5, 0.4, -0.5, 0.5); break;
        case 1: t->n_post += ph_hsv(r, o, 0.5, 0.0); break;
        case 2: t->n_post += ph_noise_family(r, o, 0.5); break;
        case 3: t->n_post += ph_blur_family(r, o, 0.5); break;
        case 4: t->n_post += ph_noise_family(r, o, 0.1); break;
        default: t->n_post += ph_blur_family(r, o, 0.1); break;
      }
    }
  }
}

// ------------------------------------------------------------------------------------ //
// host entry points                                                                     //
// ------------------------------------------------------------------------------------ //

static int det_ready(mtgv_ctx* ctx, DetState** out) {
  if (!ctx) return MTGV_ERR_INVALID;
  DetState* d = det_state(ctx);
  if (!d->set) return fail(ctx, MTGV_ERR_STATE, "detection config not set (mtgv_set_det_config)");
  if (!ctx->n_cards) return fail(ctx, MTGV_ERR_STATE, "card pool not set (mtgv_set_card_pool)");
  if (!ctx->n_bgs) return fail(ctx, MTGV_ERR_STATE, "background pool not set (mtgv_set_bg_pool)");
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e != cudaSuccess) return fail(ctx, MTGV_ERR_CUDA, cudaGetErrorString(e));
  *out = d;
  return MTGV_OK;
}

static int det_refresh_keypoints(mtgv_ctx* ctx, DetState* d) {
  if (!d->set || !ctx->n_cards) return MTGV_OK;
  det_keypoints(ctx->card_h, ctx->card_w, d->cfg.kind, &d->kp);
  if (!d->kp_dev) MTGV_CUDA_OK(ctx, cudaMalloc(&d->kp_dev, sizeof(DetKeypoints)));
  MTGV_CUDA_OK(ctx, cudaMemcpy(d->kp_dev, &d->kp, sizeof(DetKeypoints), cudaMemcpyHostToDevice));
  return MTGV_OK;
}

}  // namespace mtgv

using namespace mtgv;

extern "C" {

int mtgv_det_params_size(void) { return (int)sizeof(DetParams); }

int mtgv_set_det_config(mtgv_ctx* ctx, const mtgv_det_config* cfg) {
  if (!ctx || !cfg) return MTGV_ERR_INVALID;
  if (cfg->size_h < 8 || cfg->size_w < 8 || cfg->size_h > 8192 || cfg->size_w > 8192)
    return fail(ctx, MTGV_ERR_INVALID, "mtgv_set_det_config: bg_size_hw out of range");
  if (cfg->num_cards_min < 0 || cfg->num_cards_max - 1 > MTGV_DET_MAX_CARDS || cfg->num_cards_max < cfg->num_cards_min)
    return fail(ctx, MTGV_ERR_LIMIT, "mtgv_set_det_config: at most 32 cards per scene");
  if (cfg->max_attempts < 1 || cfg->max_attempts > MTGV_DET_MAX_ATTEMPTS)
    return fail(ctx, MTGV_ERR_LIMIT, "mtgv_set_det_config: card_max_place_attempts must be 1..10");
  if (cfg->kind != 0 && cfg->kind != 1) return fail(ctx, MTGV_ERR_INVALID, "mtgv_set_det_config: kind must be obb (0) or seg (1)");
  MTGV_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  DetState* d = det_state(ctx);
  if (ctx->n_cards) {
    // random.randint(pad - ovr, S - pad + ovr) raises "empty range" in the reference when the card
    // diagonal does not fit (od_datasets.py:320-323)
    double edge = cfg->min_visible_edges < 0.0 ? cfg->min_visible : cfg->min_visible_edges;
    if (edge < cfg->min_visible) edge = cfg->min_visible;
    const int diag = (int)sqrt((double)ctx->card_h * ctx->card_h + (double)ctx->card_w * ctx->card_w);
    const int pad = diag / 2, ovr = (int)(diag * (1.0 - edge));
    if (cfg->size_w - pad + ovr < pad - ovr || cfg->size_h - pad + ovr < pad - ovr)
      return fail(ctx, MTGV_ERR_INVALID, "empty range in randrange: card diagonal does not fit bg_size_hw with this card_min_visible_ratio_edges");
  }
  d->cfg = *cfg;
  d->set = true;
  if (!d->cfg_dev) MTGV_CUDA_OK(ctx, cudaMalloc(&d->cfg_dev, sizeof(mtgv_det_config)));
  MTGV_CUDA_OK(ctx, cudaMemcpy(d->cfg_dev, cfg, sizeof(*cfg), cudaMemcpyHostToDevice));
  return det_refresh_keypoints(ctx, d);
}

int mtgv_sample_det_tape(mtgv_ctx* ctx, uint64_t seed, int64_t first_index, int n, mtgv_det_tape* tape, void* stream) {
  DetState* d;
  int rc = det_ready(ctx, &d);
  if (rc) return rc;
  if (!tape || n < 0) return fail(ctx, MTGV_ERR_INVALID, "mtgv_sample_det_tape: bad arguments");
  if (n == 0) return MTGV_OK;
  k_det_sample<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>(seed, first_index, n, d->cfg_dev, ctx->n_cards, ctx->n_bgs,
                                                               ctx->card_h, ctx->card_w, tape);
  ctx->launches++;
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  return MTGV_OK;
}

int mtgv_det_place(mtgv_ctx* ctx, const mtgv_det_tape* tape, int n, void* params, int32_t* accepted, double* keypoints,
                   int32_t* labels, int32_t* counts, void* stream) {
  DetState* d;
  int rc = det_ready(ctx, &d);
  if (rc) return rc;
  if (!tape || !params || !accepted || !keypoints || !labels || !counts || n < 0)
    return fail(ctx, MTGV_ERR_INVALID, "mtgv_det_place: bad arguments");
  if (n == 0) return MTGV_OK;
  rc = det_refresh_keypoints(ctx, d);
  if (rc) return rc;
  k_det_place<<<(n + kPlaceWarps - 1) / kPlaceWarps, kPlaceWarps * 32, 0, (cudaStream_t)stream>>>(tape, n, d->cfg_dev, d->kp_dev, ctx->card_h, ctx->card_w, ctx->n_cards,
                                                              ctx->n_bgs, ctx->bg_hw, (DetParams*)params, accepted, keypoints,
                                                              labels, counts);
  ctx->launches++;
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  return MTGV_OK;
}

int mtgv_det_batch(mtgv_ctx* ctx, const void* params, int n, void* images, int out_dtype, const void* fields, void* stream) {
  DetState* d;
  int rc = det_ready(ctx, &d);
  if (rc) return rc;
  if (!params || !images || n < 0 || out_dtype < 0 || out_dtype > 2) return fail(ctx, MTGV_ERR_INVALID, "mtgv_det_batch: bad arguments");
  if (n == 0) return MTGV_OK;
  const int S_h = d->cfg.size_h, S_w = d->cfg.size_w;
  const size_t per = (size_t)S_h * S_w * 3;
  cudaStream_t st = (cudaStream_t)stream;
  static int blocks_per_sm = 1;
  if (!(ctx->attrs_set & 2u)) {
    MTGV_CUDA_OK(ctx, cudaFuncSetAttribute(k_det_pixels, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DetTileSmem)));
    MTGV_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_det_pixels, kDetBW * kDetBH, sizeof(DetTileSmem)));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    ctx->attrs_set |= 2u;
  }
  // scenes are processed in chunks so the blur scratch (2 float32 images per scene) stays bounded
  size_t chunk = ((size_t)3 << 30) / (per * 4 * 2);
  chunk = chunk < 1 ? 1 : (chunk > (size_t)n ? (size_t)n : chunk);
  const int passes = d->cfg.photometrics ? 1 + kDetMaxBlur : 1;
  if (d->cfg.photometrics && chunk * per > d->scratch_cap) {
    cudaFree(d->scratch[0]); cudaFree(d->scratch[1]);
    d->scratch[0] = d->scratch[1] = nullptr; d->scratch_cap = 0;
    MTGV_CUDA_OK(ctx, cudaMalloc(&d->scratch[0], chunk * per * 4));
    MTGV_CUDA_OK(ctx, cudaMalloc(&d->scratch[1], chunk * per * 4));
    d->scratch_cap = chunk * per;
  }
  if (chunk > d->list_cap) {
    cudaFree(d->lists); cudaFree(d->stats);
    d->lists = nullptr; d->stats = nullptr; d->list_cap = 0;
    MTGV_CUDA_OK(ctx, cudaMalloc(&d->lists, (size_t)(kDetMaxBlur + 1) * (chunk + 1) * 4));
    MTGV_CUDA_OK(ctx, cudaMalloc(&d->stats, (size_t)(kDetMaxBlur + 1) * chunk * 4));
    d->list_cap = chunk;
  }
  // (re)build detection's RGBA card copy when the pool changed since the last batch
  {
    const size_t words = (size_t)ctx->n_cards * ctx->card_h * ctx->card_w;
    if (words > d->rgba_cap) {
      cudaFree(d->card_rgba);
      d->card_rgba = nullptr; d->rgba_cap = 0;
      MTGV_CUDA_OK(ctx, cudaMalloc(&d->card_rgba, words * 4));
      d->rgba_cap = words;
      d->rgba_epoch = 0;
    }
    if (d->rgba_epoch != ctx->card_epoch) {
      for (int c0 = 0; c0 < ctx->n_cards; c0 += 32768) {
        const int cnt = ctx->n_cards - c0 < 32768 ? ctx->n_cards - c0 : 32768;
        k_det_rgba<<<dim3(64, cnt), 256, 0, st>>>(ctx->card_planes + (size_t)c0 * 3 * ctx->card_h * ctx->card_pitch, ctx->card_pitch,
                                                   ctx->mask_det, d->card_rgba + (size_t)c0 * ctx->card_h * ctx->card_w, ctx->card_h,
                                                   ctx->card_w);
        ctx->launches++;
      }
      d->rgba_epoch = ctx->card_epoch;
    }
  }
  const size_t elem = out_dtype == MTGV_OUT_F16 ? 2 : (out_dtype == MTGV_OUT_U8 ? 1 : 4);
  const int tiles = ((S_w + kDetTW - 1) / kDetTW) * ((S_h + kDetTH - 1) / kDetTH);
  for (size_t base = 0; base < (size_t)n; base += chunk) {
    const int m = (int)((size_t)n - base < chunk ? (size_t)n - base : chunk);
    int32_t* pass_count = d->lists;
    int32_t* pass_list = d->lists + (kDetMaxBlur + 1);
    if (passes > 1) {
      // one pass per Gaussian blur in a scene program plus one; k_det_place rejects programs with more
      // than kDetMaxBlur blurs (the reference graphs hold at most 3: one in bg_light, two blur families)
      MTGV_CUDA_OK(ctx, cudaMemsetAsync(pass_count, 0, (kDetMaxBlur + 1) * 4, st));
      k_det_lists<<<(m + 127) / 128, 128, 0, st>>>((const DetParams*)params + base, m, pass_count, pass_list);
      ctx->launches++;
    }
    for (int pass = 0; pass < passes; pass++) {
      DetLaunch L;
      L.params = (const DetParams*)params + base; L.n = m; L.pass = pass; L.pass_count = pass_count; L.pass_list = pass_list;
      L.card_rgba = d->card_rgba; L.card_h = ctx->card_h; L.card_w = ctx->card_w;
      L.bg_planes = ctx->bg_planes; L.bg_off = ctx->bg_off; L.bg_hw = ctx->bg_hw;
      L.src = pass > 0 ? d->scratch[(pass - 1) & 1] : nullptr;
      L.dst = d->scratch[pass & 1];
      L.out = (char*)images + base * per * elem; L.out_dtype = out_dtype; L.fields = (const uint32_t*)fields;
      L.stats = d->stats;
      if (pass > 0) {
        // image statistics for ISONoise boundaries: one CTA per scene of the chunk; it leaves at once when it is beyond the
        // pass's list (whose length lives on the device) or the scene's boundary op is something else
        k_det_stats<<<m, 1024, 0, st>>>(L, S_h, S_w);
        ctx->launches++;
      }
      long long want = (long long)m * tiles;
      long long cap = (long long)ctx->sm_count * blocks_per_sm * 4;
      int grid = (int)(want < cap ? want : cap);
      k_det_pixels<<<grid, dim3(kDetBW, kDetBH), sizeof(DetTileSmem), st>>>(L, S_h, S_w);
      ctx->launches++;
    }
  }
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  return MTGV_OK;
}

}  // extern "C"
