// mtgv_enc.cu - k_encoder (the plane interpreter of the encoder-pair generator), the tape sampler,
// the parameter expansion launch and the batch driver for sm_100a.
//
// One encoder batch = k_sample_tape -> k_expand -> [k_background (mtgv_bg.cu) | k_foreground
// (mtgv_fg.cu)] -> k_encoder.  The first two pixel kernels write float32 planes to scratch; k_encoder
// runs one persistent CTA per SM over a work queue of (sample, plane) items.  A plane is one channel
// of one augmented sample held entirely in shared memory as float32 (2 ping-pong planes of
// out_h*out_w*4 B = 2*96 KiB for 192x128): the foreground ops (downscale/upscale, warp, photometrics),
// the composite over the background plane and the shuffled post-augments all run on-chip, and the
// plane leaves as NCHW fp16/u8.  The three colour planes and the alpha plane of a sample are
// independent except at the composite, where colour CTAs consume the alpha plane published by the alpha
// CTA (L2 scratch + release/acquire flag; the queue order guarantees the producer is running).
//
// Reference semantics (mtgvision/encoder_datasets.py, mtgvision/util/image.py) are cited at each
// stage; the cv2 arithmetic follows SURVEY.md section 8a-notes and is restated in
// oracle/cv2_restate.py.  Compiled with -fmad=false; fused multiply-adds appear only where written
// explicitly (value paths, see DESIGN.md "What is exact and what is toleranced").
#include <cuda_fp16.h>
#include <stdlib.h>

#include "mtgv_internal.cuh"
#include "mtgv_persp.cuh"

namespace mtgv {

#ifndef MTGV_ENC_THREADS
#define MTGV_ENC_THREADS 512
#endif
constexpr int kThreads = MTGV_ENC_THREADS;

// ------------------------------------------------------------------------------------ //
// small device helpers                                                                  //
// ------------------------------------------------------------------------------------ //

__device__ __forceinline__ float clip01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

__device__ __forceinline__ float bilinear_weights_sum(float v0, float v1, float v2, float v3, int ax, int ay) {
  // initInterTab2D(INTER_LINEAR): w = (1-fy|fy)*(1-fx|fx), fx = ax/32 -> all products exact
  float fx = (float)ax * 0.03125f, fy = (float)ay * 0.03125f;
  float gx = 1.f - fx, gy = 1.f - fy;
  float s = __fmul_rn(v0, gy * gx);
  s = __fadd_rn(s, __fmul_rn(v1, gy * fx));
  s = __fadd_rn(s, __fmul_rn(v2, fy * gx));
  s = __fadd_rn(s, __fmul_rn(v3, fy * fx));
  return s;
}

// remapBilinear<float>, BORDER_CONSTANT 0, source = float plane (shared or global)
__device__ __forceinline__ float bilinear_plane(const float* __restrict__ S, int H, int W, int X, int Y) {
  int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
  bool x0 = (unsigned)sx < (unsigned)W, x1 = (unsigned)(sx + 1) < (unsigned)W;
  bool y0 = (unsigned)sy < (unsigned)H, y1 = (unsigned)(sy + 1) < (unsigned)H;
  const float* p = S + sy * W + sx;
  float v0 = (y0 && x0) ? p[0] : 0.f;
  float v1 = (y0 && x1) ? p[1] : 0.f;
  float v2 = (y1 && x0) ? p[W] : 0.f;
  float v3 = (y1 && x1) ? p[W + 1] : 0.f;
  return bilinear_weights_sum(v0, v1, v2, v3, X & 31, Y & 31);
}

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ float u32_to_unit(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

// per-pixel Philox draw for device-generated random fields: (seed, op slot) key, (pixel, channel) counter
__device__ __forceinline__ void field_draw(uint64_t seed, int slot, uint32_t pixel, uint32_t chan, uint32_t sub, uint32_t* r) {
  Philox ph;
  ph.key[0] = (uint32_t)seed ^ (0x9E3779B9u * (uint32_t)(slot + 1));
  ph.key[1] = (uint32_t)(seed >> 32);
  ph(pixel, chan, sub, 0x6d746776u, r);
}

// four standard normals from one Philox call (two Box-Muller pairs)
__device__ __forceinline__ void normals4(const uint32_t* r, float* g) {
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const float u1 = ((float)(r[2 * q] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float rad = sqrtf(-2.f * __logf(u1));
    float sn, cs;
    sincospif(2.f * u32_to_unit(r[2 * q + 1]), &sn, &cs);
    g[2 * q] = rad * cs;
    g[2 * q + 1] = rad * sn;
  }
}

// ------------------------------------------------------------------------------------ //
// plane interpreter                                                                     //
// ------------------------------------------------------------------------------------ //

struct Vm {
  float* P[2];
  int cur;
  int H, W;
  int chan;  // 0..2 colour, 3 alpha
  const uint32_t* fields;
  uint64_t seed;
  double* aux;  // vm_aux_bytes(H, W) of scratch: per (row, column block) perspective origins
  __device__ float* cur_p() const { return P[cur]; }
  __device__ float* oth_p() const { return P[cur ^ 1]; }
};

__host__ __device__ inline size_t vm_aux_bytes(int H, int W) {
  const int bw0 = persp_block_w(H, W);
  return (size_t)H * ((W + bw0 - 1) / bw0) * 4 * sizeof(double);
}

// row-major walk of an h x w plane by the whole block without per-pixel divisions
#define MTGV_FOR_PIXELS(i, x, y, h, w)                                                                       \
  for (int i = threadIdx.x, x = threadIdx.x % (w), y = threadIdx.x / (w), dx__ = blockDim.x % (w),          \
           dy__ = blockDim.x / (w);                                                                        \
       i < (h) * (w); i += blockDim.x, x += dx__, y += dy__ + (x >= (w) ? 1 : 0), x -= (x >= (w) ? (w) : 0))

// cv2.resize restated per destination pixel (oracle/cv2_restate.py resize_nearest/linear/cubic)
__device__ __forceinline__ void cubic_coeffs(float x, float* c) {
  const float A = -0.75f;
  c[0] = ((A * (x + 1.f) - 5.f * A) * (x + 1.f) + 8.f * A) * (x + 1.f) - 4.f * A;
  c[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  c[2] = ((A + 2.f) * (1.f - x) - (A + 3.f)) * (1.f - x) * (1.f - x) + 1.f;
  c[3] = 1.f - c[0] - c[1] - c[2];
}

__device__ __forceinline__ void linear_ofs(int d, int ssize, double scale, int* s, float* f) {
  float fx = (float)(((double)d + 0.5) * scale - 0.5);
  int sx = (int)floorf(fx);
  fx -= (float)sx;
  if (sx < 0) { sx = 0; fx = 0.f; }
  if (sx >= ssize - 1) { sx = ssize - 1; fx = 0.f; }
  *s = sx;
  *f = fx;
}

__device__ float resize_px(const float* __restrict__ S, int sh, int sw, int dh, int dw, int interp, int y, int x) {
  if (interp == 0) {  // INTER_NEAREST
    double ifx = 1.0 / ((double)dw / (double)sw), ify = 1.0 / ((double)dh / (double)sh);
    int sx = min((int)floor((double)x * ifx), sw - 1), sy = min((int)floor((double)y * ify), sh - 1);
    return S[sy * sw + sx];
  }
  double scx = 1.0 / ((double)dw / (double)sw), scy = 1.0 / ((double)dh / (double)sh);  // cv::resize: 1./inv_scale
  if (interp == 1) {  // INTER_LINEAR
    int sx, sy;
    float fx, fy;
    linear_ofs(x, sw, scx, &sx, &fx);
    linear_ofs(y, sh, scy, &sy, &fy);
    int sx1 = min(sx + 1, sw - 1), sy1 = min(sy + 1, sh - 1);
    float gx = 1.f - fx, gy = 1.f - fy;
    float r0 = S[sy * sw + sx] * gx + S[sy * sw + sx1] * fx;
    float r1 = S[sy1 * sw + sx] * gx + S[sy1 * sw + sx1] * fx;
    return r0 * gy + r1 * fy;
  }
  // INTER_CUBIC
  float fx = (float)(((double)x + 0.5) * scx - 0.5), fy = (float)(((double)y + 0.5) * scy - 0.5);
  int sx = (int)floorf(fx), sy = (int)floorf(fy);
  fx -= (float)sx;
  fy -= (float)sy;
  float cx[4], cy[4];
  cubic_coeffs(fx, cx);
  cubic_coeffs(fy, cy);
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    int yy = min(max(sy - 1 + j, 0), sh - 1);
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int xx = min(max(sx - 1 + i, 0), sw - 1);
      r = r + S[yy * sw + xx] * cx[i];
    }
    acc = acc + r * cy[j];
  }
  return acc;
}

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

// Executes one expanded op on the current plane.  Block-cooperative; returns with all
// writes visible (trailing __syncthreads).
__device__ void vm_run_op(Vm& vm, const mtgv_x_op& op, int slot) {
  const int H = vm.H, W = vm.W, HW = H * W, c = vm.chan;
  float* cur = vm.cur_p();
  float* oth = vm.oth_p();
  const int tid = threadIdx.x, nt = blockDim.x;
  switch (op.code) {
    case MTGV_X_ELEM: {
      if (!((op.i[0] >> c) & 1)) break;
      const float a = op.f[c], b = op.f[4 + c];
      const bool clip = op.i[1] != 0;
      for (int i = tid; i < HW; i += nt) {
        float v = __fadd_rn(__fmul_rn(a, cur[i]), b);
        cur[i] = clip ? clip01(v) : v;
      }
      break;
    }
    case MTGV_X_DOWNUP: {  // Mutate.downscale_upscale: two cv2.resize calls
      const int n = op.i[0], h2 = H >> n, w2 = W >> n;
      MTGV_FOR_PIXELS(i, x, y, h2, w2) oth[i] = resize_px(cur, H, W, h2, w2, op.i[1], y, x);
      __syncthreads();
      MTGV_FOR_PIXELS(i, x, y, H, W) cur[i] = resize_px(oth, h2, w2, H, W, op.i[2], y, x);
      break;
    }
    case MTGV_X_WARP_PERSP: {  // cv2.warpPerspective, same size
      const int bw0 = persp_block_w(H, W), nblk = (W + bw0 - 1) / bw0;
      for (int k = tid; k < H * nblk; k += nt) persp_origin(op.d, (double)((k % nblk) * bw0), (double)(k / nblk), vm.aux + 4 * k);
      __syncthreads();
      const double m0 = op.d[0], m3 = op.d[3], m6 = op.d[6];
      const int bw_shift = (bw0 & (bw0 - 1)) == 0 ? 31 - __clz(bw0) : -1;
      MTGV_FOR_PIXELS(i, x, y, H, W) {
        const int bi = bw_shift >= 0 ? x >> bw_shift : x / bw0;
        const double* o = vm.aux + 4 * (y * nblk + bi);
        const int2 XY = persp_xy(o[0], o[1], o[2], m0, m3, m6, (double)(x - bi * bw0));
        oth[i] = bilinear_plane(cur, H, W, XY.x, XY.y);
      }
      vm.cur ^= 1;
      break;
    }
    case MTGV_X_WARP_AFFINE: {  // cv2.warpAffine, same size
      MTGV_FOR_PIXELS(i, x, y, H, W) {
        int X = (affine_row_origin(op.d[1], op.d[2], y) + affine_col_delta(op.d[0], x)) >> 5;
        int Y = (affine_row_origin(op.d[4], op.d[5], y) + affine_col_delta(op.d[3], x)) >> 5;
        oth[i] = bilinear_plane(cur, H, W, X, Y);
      }
      vm.cur ^= 1;
      break;
    }
    case MTGV_X_BLUR3: {  // cv2.GaussianBlur((3,3),0): [1/4,1/2,1/4] separable, REFLECT_101
      MTGV_FOR_PIXELS(i, x, y, H, W) {
        int xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);
        float r[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const float* row = cur + reflect101(y - 1 + j, H) * W;
          r[j] = row[xm] * 0.25f + row[x] * 0.5f + row[xp] * 0.25f;
        }
        oth[i] = r[0] * 0.25f + r[1] * 0.5f + r[2] * 0.25f;
      }
      vm.cur ^= 1;
      break;
    }
    case MTGV_X_SHARPEN: {  // filter2D [[0,-1,0],[-1,5,-1],[0,-1,0]] + clip (Mutate.sharpen)
      MTGV_FOR_PIXELS(i, x, y, H, W) {
        float v = 5.f * cur[i] - cur[reflect101(y - 1, H) * W + x] - cur[y * W + reflect101(x - 1, W)] -
                  cur[y * W + reflect101(x + 1, W)] - cur[reflect101(y + 1, H) * W + x];
        oth[i] = clip01(v);
      }
      vm.cur ^= 1;
      break;
    }
    case MTGV_X_NOISE: {  // Mutate.noise: noisy variant blended by ratio into channels :3
      if (c > 2) break;
      const int kind = op.i[0];
      const float ra = op.f[0], rb = op.f[1];
      const bool inj = op.field != MTGV_FIELD_PHILOX;
      if (kind == 2) {  // noise_salt_pepper(strength .1, svp .5) -> clip -> blend
        for (int i = tid; i < HW; i += nt) oth[i] = cur[i];
        __syncthreads();
        for (int pass = 0; pass < 2; pass++) {
          const int npts = pass ? op.n_field2 : op.n_field;
          const int32_t* pts = inj ? (const int32_t*)(vm.fields + (pass ? op.field2 : op.field)) : nullptr;
          for (int k = tid; k < npts; k += nt) {
            int py, px, pc;
            if (inj) {
              py = pts[3 * k], px = pts[3 * k + 1], pc = pts[3 * k + 2];
            } else {  // np.random.randint(0, i-1) per axis of (H, W, 3)
              uint32_t r[4];
              field_draw(vm.seed, slot, (uint32_t)k, 0, 10 + pass, r);
              py = (int)(((uint64_t)r[0] * (uint32_t)(H - 1)) >> 32);
              px = (int)(((uint64_t)r[1] * (uint32_t)(W - 1)) >> 32);
              pc = (int)(r[2] & 1u);
            }
            if (pc == c) oth[py * W + px] = pass ? 0.f : 1.f;
          }
          __syncthreads();
        }
        for (int i = tid; i < HW; i += nt) cur[i] = __fadd_rn(__fmul_rn(ra, clip01(oth[i])), __fmul_rn(rb, cur[i]));
        break;
      }
      const float* fld = inj ? (const float*)(vm.fields + op.field) : nullptr;
      // device-generated fields: one Philox call per group of four consecutive pixels
      for (int i0 = 4 * tid; i0 < HW; i0 += 4 * nt) {
        float g4[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t r[4] = {0u, 0u, 0u, 0u};
        if (!inj) {
          field_draw(vm.seed, slot, (uint32_t)(i0 >> 2), (uint32_t)c, 0, r);
          if (kind != 3) normals4(r, g4);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = i0 + u;
          if (i >= HW) break;
          float x = cur[i], noisy;
          float g;
          if (inj) {
            g = fld[(size_t)i * 3 + c];
          } else if (kind == 3) {  // Poisson(lam = clip(x)*0.8) by inversion
            float lam = clip01(x) * 0.8f, p = __expf(-lam), F = p, uu = u32_to_unit(r[u]);
            int k = 0;
            while (uu > F && k < 32) {
              k++;
              p *= lam / (float)k;
              F += p;
            }
            g = (float)k;
          } else {
            g = kind == 1 ? g4[u] * 0.22360679774997896f : g4[u];  // sqrt(var=0.05)
          }
          if (kind == 0)  // noise_speckle(strength 0.3): x * (1 + g*0.3)
            noisy = clip01(__fmul_rn(x, __fadd_rn(1.f, __fmul_rn(g, 0.3f))));
          else if (kind == 1)  // noise_gaussian(var 0.05)
            noisy = clip01(__fadd_rn(x, g));
          else {  // noise_poisson(peak .8, amount .5): clip(.5*clip(x) + .5*(counts/.8))
            float xs = clip01(x);
            noisy = clip01(__fadd_rn(__fmul_rn(0.5f, xs), __fmul_rn(0.5f, (float)((double)g / 0.8))));
          }
          cur[i] = __fadd_rn(__fmul_rn(ra, noisy), __fmul_rn(rb, x));
        }
      }
      break;
    }
    case MTGV_X_GAUSS_NOISE: {  // Mutate.gaussian_noise(sigma .25): clip(img + noise)
      const bool inj = op.field != MTGV_FIELD_PHILOX;
      const float* fld = inj ? (const float*)(vm.fields + op.field) : nullptr;
      const int nc = 3;
      if (c > 2) break;
      for (int i0 = 4 * tid; i0 < HW; i0 += 4 * nt) {
        float g4[4] = {0.f, 0.f, 0.f, 0.f};
        if (!inj) {
          uint32_t r[4];
          field_draw(vm.seed, slot, (uint32_t)(i0 >> 2), (uint32_t)c, 0, r);
          normals4(r, g4);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = i0 + u;
          if (i >= HW) break;
          const float g = inj ? fld[(size_t)i * nc + c] : g4[u] * 0.25f;
          cur[i] = clip01(__fadd_rn(cur[i], g));
        }
      }
      break;
    }
    case MTGV_X_SALT_PEPPER: {  // Mutate.salt_pepper_noise: all channels, salt then pepper
      if (c > 2) break;
      const bool inj = op.field != MTGV_FIELD_PHILOX;
      for (int pass = 0; pass < 2; pass++) {
        const int npts = pass ? op.n_field2 : op.n_field;
        const int32_t* pts = inj ? (const int32_t*)(vm.fields + (pass ? op.field2 : op.field)) : nullptr;
        for (int k = tid; k < npts; k += nt) {
          int py, px;
          if (inj) {
            py = pts[2 * k], px = pts[2 * k + 1];
          } else {
            uint32_t r[4];
            field_draw(vm.seed, slot, (uint32_t)k, 0, 10 + pass, r);
            py = (int)(((uint64_t)r[0] * (uint32_t)(H - 1)) >> 32);
            px = (int)(((uint64_t)r[1] * (uint32_t)(W - 1)) >> 32);
          }
          cur[py * W + px] = pass ? 0.f : 1.f;
        }
        __syncthreads();
      }
      break;
    }
    case MTGV_X_ERASE: {  // Mutate.random_erasing
      if (c > 2) break;
      const int y0 = op.i[0], y1 = op.i[1], x0 = op.i[2], x1 = op.i[3], mode = op.i[4];
      const int bw = x1 - x0, n = (y1 - y0) * bw;
      float fill = op.f[c];
      if (mode == 4) {  // block mean (float32)
        float part = 0.f;
        for (int k = tid; k < n; k += nt) part += cur[(y0 + k / bw) * W + x0 + k % bw];
        for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        __syncthreads();
        if ((tid & 31) == 0) oth[tid >> 5] = part;
        __syncthreads();
        float tot = 0.f;
        for (int k = 0; k < (nt >> 5); k++) tot += oth[k];
        fill = tot / (float)n;
        __syncthreads();
      }
      const bool inj = op.field != MTGV_FIELD_PHILOX;
      const float* fld = (mode == 0 && inj) ? (const float*)(vm.fields + op.field) : nullptr;
      for (int k = tid; k < n; k += nt) {
        float v = fill;
        if (mode == 0) {
          if (inj) {
            v = fld[(size_t)k * 3 + c];
          } else {
            uint32_t r[4];
            field_draw(vm.seed, slot, (uint32_t)k, (uint32_t)c, 3, r);
            v = u32_to_unit(r[0]);
          }
        }
        cur[(y0 + k / bw) * W + x0 + k % bw] = v;
      }
      break;
    }
    case MTGV_X_CUTOUT: {  // Mutate.cutout: 8 holes [y-4,y+4)x[x-4,x+4) -> 0
      if (c > 2) break;
      for (int k = tid; k < 8 * 64; k += nt) {
        int hole = k >> 6, dy = (k >> 3) & 7, dx = k & 7;
        int y = op.i[2 * hole] - 4 + dy, x = op.i[2 * hole + 1] - 4 + dx;
        if ((unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W) cur[y * W + x] = 0.f;
      }
      break;
    }
    default:
      break;
  }
  __syncthreads();
}

struct EncLaunch {
  const mtgv_enc_params* params;
  int n;
  const float* alpha0;
  const float* bg_scratch;  // [n,3,H,W] float32 from k_background
  const float* fg_scratch;  // [n,3,H,W] float32 from k_foreground
  float* alpha_scratch;
  int* sync_words;
  void* out;
  int out_dtype;
  const uint32_t* fields;
};

// rgba_over_rgb (util/image.py:246-290): cur = clip(bg*(1-a) + fg*a); bg from k_background's output
__device__ void stage_composite(float* cur, int HW, const float* __restrict__ bg, const float* __restrict__ alpha) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if ((HW & 3) == 0) {
    // 16-byte loads, four vectors (of each operand) in flight per thread before the first use
    const float4* bg4 = reinterpret_cast<const float4*>(bg);
    const float4* al4 = reinterpret_cast<const float4*>(alpha);
    float4* cur4 = reinterpret_cast<float4*>(cur);
    const int n4 = HW >> 2;
    for (int j0 = tid; j0 < n4; j0 += 4 * nt) {
      float4 b[4], a[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int j = j0 + u * nt;
        if (j < n4) {
          b[u] = __ldcg(bg4 + j);
          a[u] = alpha ? __ldcg(al4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int j = j0 + u * nt;
        if (j < n4) {
          float4 c = cur4[j];
          c.x = clip01(__fadd_rn(__fmul_rn(b[u].x, __fsub_rn(1.f, a[u].x)), __fmul_rn(c.x, a[u].x)));
          c.y = clip01(__fadd_rn(__fmul_rn(b[u].y, __fsub_rn(1.f, a[u].y)), __fmul_rn(c.y, a[u].y)));
          c.z = clip01(__fadd_rn(__fmul_rn(b[u].z, __fsub_rn(1.f, a[u].z)), __fmul_rn(c.z, a[u].z)));
          c.w = clip01(__fadd_rn(__fmul_rn(b[u].w, __fsub_rn(1.f, a[u].w)), __fmul_rn(c.w, a[u].w)));
          cur4[j] = c;
        }
      }
    }
    return;
  }
  for (int o = tid; o < HW; o += nt) {
    const float a = alpha ? __ldcg(alpha + o) : 0.f;
    const float bgv = __ldcg(bg + o);
    cur[o] = clip01(__fadd_rn(__fmul_rn(bgv, __fsub_rn(1.f, a)), __fmul_rn(cur[o], a)));
  }
}

// foreground plane from k_foreground into shared memory, zero outside the pasted rectangle
// (crop_to_size(pad=True), util/image.py:366-371)
__device__ void stage_load_fg(float* P, int OH, int OW, const float* __restrict__ fg, bool none, int y0, int y1, int x0, int x1) {
  const int tid = threadIdx.x, nt = blockDim.x, HW = OH * OW;
  if ((OW & 3) == 0) {
    const int W4 = OW >> 2, n4 = HW >> 2;
    const unsigned magic = ((1u << 24) + (unsigned)W4 - 1u) / (unsigned)W4;  // j / W4 for j * W4 < 2^24
    const float4* fg4 = reinterpret_cast<const float4*>(fg);
    float4* P4 = reinterpret_cast<float4*>(P);
    for (int j0 = tid; j0 < n4; j0 += 4 * nt) {
      float4 v[4];
      int yy[4], xx[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int j = j0 + u * nt;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        yy[u] = (int)(((unsigned)j * magic) >> 24);
        xx[u] = (j - yy[u] * W4) << 2;
        if (j < n4 && !none && yy[u] >= y0 && yy[u] < y1 && xx[u] + 3 >= x0 && xx[u] < x1) v[u] = __ldcg(fg4 + j);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int j = j0 + u * nt;
        if (j < n4) {
          const bool row = !none && yy[u] >= y0 && yy[u] < y1;
          float4 c = v[u];
          c.x = (row && xx[u] >= x0 && xx[u] < x1) ? c.x : 0.f;
          c.y = (row && xx[u] + 1 >= x0 && xx[u] + 1 < x1) ? c.y : 0.f;
          c.z = (row && xx[u] + 2 >= x0 && xx[u] + 2 < x1) ? c.z : 0.f;
          c.w = (row && xx[u] + 3 >= x0 && xx[u] + 3 < x1) ? c.w : 0.f;
          P4[j] = c;
        }
      }
    }
    return;
  }
  for (int i = tid; i < HW; i += nt) {
    const int y = i / OW, x = i - y * OW;
    P[i] = (!none && y >= y0 && y < y1 && x >= x0 && x < x1) ? __ldcg(fg + i) : 0.f;
  }
}

// ------------------------------------------------------------------------------------ //
// plane store (subsystem 5): NCHW fp16 / uint8 / fp32, 16-byte vectors                  //
// ------------------------------------------------------------------------------------ //

__device__ void store_plane(const float* __restrict__ P, int HW, void* out, size_t plane_index, int dtype) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (dtype == MTGV_OUT_F16) {
    __half* o = (__half*)out + plane_index * HW;
    if ((HW & 7) == 0) {
      for (int i = tid * 8; i < HW; i += nt * 8) {
        __half2 h0 = __floats2half2_rn(P[i], P[i + 1]), h1 = __floats2half2_rn(P[i + 2], P[i + 3]);
        __half2 h2 = __floats2half2_rn(P[i + 4], P[i + 5]), h3 = __floats2half2_rn(P[i + 6], P[i + 7]);
        uint4 v;
        v.x = *(uint32_t*)&h0; v.y = *(uint32_t*)&h1; v.z = *(uint32_t*)&h2; v.w = *(uint32_t*)&h3;
        *(uint4*)(o + i) = v;
      }
    } else {
      for (int i = tid; i < HW; i += nt) o[i] = __float2half_rn(P[i]);
    }
  } else if (dtype == MTGV_OUT_U8) {
    uint8_t* o = (uint8_t*)out + plane_index * HW;
    if ((HW & 15) == 0) {
      for (int i = tid * 16; i < HW; i += nt * 16) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          uint32_t acc = 0;
#pragma unroll
          for (int r = 0; r < 4; r++) acc |= (uint32_t)__float2int_rn(clip01(P[i + 4 * q + r]) * 255.f) << (8 * r);
          w[q] = acc;
        }
        *(uint4*)(o + i) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    } else {
      for (int i = tid; i < HW; i += nt) o[i] = (uint8_t)__float2int_rn(clip01(P[i]) * 255.f);
    }
  } else {
    float* o = (float*)out + plane_index * HW;
    for (int i = tid; i < HW; i += nt) o[i] = P[i];
  }
}

// ------------------------------------------------------------------------------------ //
// the persistent encoder kernel                                                         //
// ------------------------------------------------------------------------------------ //

__device__ __forceinline__ bool op_touches_alpha(const mtgv_x_op& op) {
  switch (op.code) {
    case MTGV_X_ELEM: return (op.i[0] >> 3) & 1;
    case MTGV_X_DOWNUP:
    case MTGV_X_WARP_PERSP:
    case MTGV_X_WARP_AFFINE: return true;
    default: return false;
  }
}

struct SmemLayout {
  float* P0; float* P1; double* aux; mtgv_enc_params* sp; int* item;
};

__host__ __device__ inline size_t enc_smem_bytes(int OH, int OW) {
  size_t HW = (size_t)OH * OW;
  size_t b = 2 * HW * 4;
  b = (b + 15) & ~(size_t)15;
  b += vm_aux_bytes(OH, OW);
  b = (b + 15) & ~(size_t)15;
  b += sizeof(mtgv_enc_params);
  b = (b + 15) & ~(size_t)15;
  b += 64;
  return b;
}

__device__ inline SmemLayout carve(unsigned char* raw, int OH, int OW) {
  SmemLayout s;
  size_t HW = (size_t)OH * OW;
  s.P0 = (float*)raw;
  s.P1 = s.P0 + HW;
  size_t off = (2 * HW * 4 + 15) & ~(size_t)15;
  s.aux = (double*)(raw + off);
  off = (off + vm_aux_bytes(OH, OW) + 15) & ~(size_t)15;
  s.sp = (mtgv_enc_params*)(raw + off);
  off = (off + sizeof(mtgv_enc_params) + 15) & ~(size_t)15;
  s.item = (int*)(raw + off + 32);
  return s;
}

__global__ void __launch_bounds__(kThreads, 1) k_encoder(EncLaunch L, int OH, int OW) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemLayout S = carve(smem_raw, OH, OW);
  const int tid = threadIdx.x, nt = blockDim.x, HW = OH * OW;
  int* counter = L.sync_words;
  int* flags = L.sync_words + 1;

  for (;;) {
    if (tid == 0) *S.item = atomicAdd(counter, 1);
    __syncthreads();
    const int item = *S.item;
    if (item >= 4 * L.n) break;
    const int s = item >> 2, plane = (item & 3) - 1;  // -1: alpha plane, 0..2: R,G,B
    {  // stage the sample's parameters
      const uint32_t* src = (const uint32_t*)(L.params + s);
      uint32_t* dst = (uint32_t*)S.sp;
      for (int k = tid; k < (int)(sizeof(mtgv_enc_params) / 4); k += nt) dst[k] = src[k];
    }
    __syncthreads();
    const mtgv_enc_params& sp = *S.sp;
    const bool bad = sp.status != 0 || sp.out_h != OH || sp.out_w != OW;
    bool alpha_static = true;
    for (int k = 0; k < sp.n_fg; k++) alpha_static = alpha_static && !op_touches_alpha(sp.ops[k]);

    if (plane < 0) {
      // ---- alpha plane: static rounded-rect alpha through the geometric foreground ops ----
      if (!bad && sp.kind == MTGV_KIND_VIRTUAL && !alpha_static) {
        for (int i = tid; i < HW; i += nt) S.P0[i] = L.alpha0[i];
        __syncthreads();
        Vm vm{{S.P0, S.P1}, 0, OH, OW, 3, L.fields, sp.seed, S.aux};
        for (int k = 0; k < sp.n_fg; k++) vm_run_op(vm, sp.ops[k], k);
        float* dst = L.alpha_scratch + (size_t)s * HW;
        const float* srcp = vm.cur_p();
        for (int i = tid; i < HW; i += nt) __stcg(dst + i, srcp[i]);
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(flags + s, 1);
      }
      __syncthreads();
      continue;
    }

    // ---- colour plane ----
    if (bad) {
      for (int i = tid; i < HW; i += nt) S.P0[i] = 0.f;
      __syncthreads();
      store_plane(S.P0, HW, L.out, (size_t)s * 3 + plane, L.out_dtype);
      __syncthreads();
      continue;
    }
    const bool bg_only = sp.kind == MTGV_KIND_BG_ONLY;
    if (sp.kind != MTGV_KIND_CROPPED) {
      // the composite reads this sample's background plane (and the static alpha) some microseconds from now:
      // pull the lines into L2 while the foreground ops run
      const char* bgp = (const char*)(L.bg_scratch + ((size_t)s * 3 + plane) * HW);
      for (int o = tid * 128; o < HW * 4; o += nt * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(bgp + o));
    }
    stage_load_fg(S.P0, OH, OW, L.fg_scratch + ((size_t)fg_plane_owner(L.params, L.n, s) * 3 + plane) * HW, bg_only, sp.fg_y0, sp.fg_y0 + sp.fg_rh, sp.fg_x0,
                  sp.fg_x0 + sp.fg_rw);
    __syncthreads();
    Vm vm{{S.P0, S.P1}, 0, OH, OW, plane, L.fields, sp.seed, S.aux};
    if (sp.kind != MTGV_KIND_CROPPED) {
      for (int k = 0; k < sp.n_fg; k++) vm_run_op(vm, sp.ops[k], k);
      const float* alpha = bg_only ? nullptr : L.alpha0;
      if (!alpha_static && !bg_only) {
        alpha = L.alpha_scratch + (size_t)s * HW;
        if (tid == 0)
          while (ld_acquire(flags + s) == 0) __nanosleep(200);
        __syncthreads();
      }
      stage_composite(vm.cur_p(), HW, L.bg_scratch + ((size_t)s * 3 + plane) * HW, alpha);
      __syncthreads();
      const int base = sp.n_fg + sp.n_pre + sp.n_post;
      for (int k = 0; k < sp.n_vrtl; k++) vm_run_op(vm, sp.ops[base + k], base + k);
    }
    store_plane(vm.cur_p(), HW, L.out, (size_t)s * 3 + plane, L.out_dtype);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------ //
// static alpha: pad(INTER_AREA(round_rect_mask)) for (card_hw -> out_hw)                //
// ------------------------------------------------------------------------------------ //

__global__ void k_static_alpha(const float* __restrict__ mask, int ch, int cw, int OH, int OW, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= OH * OW) return;
  int rh, rw, y0, x0;
  if (ch == OH && cw == OW) { rh = OH; rw = OW; y0 = x0 = 0; }
  else crop_geometry(ch, cw, OH, OW, true, &rh, &rw, &y0, &x0);
  int ry = i / OW - y0, rx = i % OW - x0;
  float v = 0.f;
  if ((unsigned)ry < (unsigned)rh && (unsigned)rx < (unsigned)rw) {
    int sy0, sx0;
    float wy[kAreaMaxTaps], wx[kAreaMaxTaps];
    int ny = area_taps(ch, rh, ry, &sy0, wy), nx = area_taps(cw, rw, rx, &sx0, wx);
    float sum = 0.f;
    for (int j = 0; j < ny; j++) {
      float h = 0.f;
      for (int k = 0; k < nx; k++) h = __fadd_rn(h, __fmul_rn(mask[(size_t)(sy0 + j) * cw + sx0 + k], wx[k]));
      sum = j == 0 ? __fmul_rn(wy[0], h) : __fadd_rn(sum, __fmul_rn(wy[j], h));
    }
    v = clip01(sum);
  }
  out[i] = v;
}

int enc_build_static_alpha(mtgv_ctx* ctx, cudaStream_t st) {
  if (!ctx->cfg_set || !ctx->mask_enc) return MTGV_OK;
  const int OH = ctx->cfg.out_h, OW = ctx->cfg.out_w;
  if (ctx->alpha0) cudaFree(ctx->alpha0);
  MTGV_CUDA_OK(ctx, cudaMalloc(&ctx->alpha0, (size_t)OH * OW * 4));
  k_static_alpha<<<(OH * OW + 255) / 256, 256, 0, st>>>(ctx->mask_enc, ctx->card_h, ctx->card_w, OH, OW, ctx->alpha0);
  ctx->launches++;
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  return MTGV_OK;
}

// ------------------------------------------------------------------------------------ //
// pool ingest: HWC uint8 -> planar rows padded to 16 B                                  //
// ------------------------------------------------------------------------------------ //

__global__ void k_planarize(const uint8_t* __restrict__ hwc, uint8_t* __restrict__ planes, int h, int w, int pitch) {
  const size_t img = blockIdx.y;
  const uint8_t* src = hwc + img * (size_t)h * w * 3;
  uint8_t* dst = planes + img * (size_t)3 * h * pitch;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)h * pitch; i += (size_t)gridDim.x * blockDim.x) {
    int y = (int)(i / pitch), x = (int)(i % pitch);
#pragma unroll
    for (int c = 0; c < 3; c++) dst[((size_t)c * h + y) * pitch + x] = x < w ? src[((size_t)y * w + x) * 3 + c] : 0;
  }
}

int pool_planarize(mtgv_ctx* ctx, const uint8_t* hwc, uint8_t* planes, int n, int h, int w, int pitch, cudaStream_t st) {
  if (n <= 0) return MTGV_OK;
  for (int i0 = 0; i0 < n; i0 += 32768) {
    int cnt = n - i0 < 32768 ? n - i0 : 32768;
    dim3 grid(64, cnt);
    k_planarize<<<grid, 256, 0, st>>>(hwc + (size_t)i0 * h * w * 3, planes + (size_t)i0 * 3 * h * pitch, h, w, pitch);
    ctx->launches++;
  }
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  return MTGV_OK;
}

// ------------------------------------------------------------------------------------ //
// subsystem (1): tape -> params                                                         //
// ------------------------------------------------------------------------------------ //

__global__ void k_expand(const mtgv_enc_tape* tape, int n, const mtgv_enc_config* cfg, PoolMeta pm, mtgv_enc_params* params,
                         int64_t* labels) {
  // one warp per sample, lane 0 working: every sample takes its own path through the op expansions, so 32
  // samples in one warp serialise (measured 4.6 active lanes per instruction); one sample per warp spreads
  // the same serial chains over all SMs instead
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s >= n || (threadIdx.x & 31) != 0) return;
  expand_encoder_sample(&tape[s], cfg, pm, &params[s]);
  if (labels) {
    int card = params[s].card;
    bool ok = params[s].status == 0 && card >= 0 && card < pm.n_cards;
    for (int k = 0; k < 3; k++) labels[(size_t)s * 3 + k] = ok ? (int64_t)pm.labels3[card * 3 + k] : -1;
  }
}

static PoolMeta pool_meta(const mtgv_ctx* ctx) {
  PoolMeta pm;
  pm.card_h = ctx->card_h; pm.card_w = ctx->card_w; pm.n_cards = ctx->n_cards; pm.n_bgs = ctx->n_bgs;
  pm.labels3 = ctx->labels3; pm.grp_off = ctx->grp_off; pm.grp_mem = ctx->grp_mem; pm.bg_hw = ctx->bg_hw;
  return pm;
}

int enc_expand(mtgv_ctx* ctx, const mtgv_enc_tape* tape, int n, mtgv_enc_params* params, int64_t* labels, cudaStream_t st) {
  if (n <= 0) return MTGV_OK;
  k_expand<<<(n + 3) / 4, 128, 0, st>>>(tape, n, ctx->cfg_dev, pool_meta(ctx), params, labels);
  ctx->launches++;
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  return MTGV_OK;
}

// ------------------------------------------------------------------------------------ //
// production sampler (Philox): what _random_image_batch/_make_image_batch/make_virtual   //
// draw from `random` / `np.random`, one thread per pair           //
// ------------------------------------------------------------------------------------ //

struct Rng {
  Philox ph;
  uint32_t c0, c1, c2;
  uint32_t ctr;
  uint32_t buf[4];
  int have;
  __device__ Rng(uint64_t seed, uint64_t index, uint32_t stream) {
    ph.key[0] = (uint32_t)seed; ph.key[1] = (uint32_t)(seed >> 32);
    c0 = (uint32_t)index; c1 = (uint32_t)(index >> 32); c2 = stream; ctr = 0; have = 0;
  }
  __device__ uint32_t u32() {
    if (!have) { ph(c0, c1, c2, ctr++, buf); have = 4; }
    return buf[--have];
  }
  __device__ double uniform() {  // 53-bit like random.random()
    uint32_t a = u32() >> 5, b = u32() >> 6;
    return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
  }
  __device__ double uniform(double lo, double hi) { return lo + (hi - lo) * uniform(); }
  __device__ int below(int n) { return (int)(((uint64_t)u32() * (uint32_t)n) >> 32); }
};

__device__ void tape_op_init(mtgv_tape_op* o, int code) {
  o->code = code; o->n_field = o->n_field2 = 0; o->_pad = 0;
  for (int k = 0; k < 16; k++) o->i[k] = 0;
  for (int k = 0; k < 8; k++) o->d[k] = 0.0;
  o->field = o->field2 = MTGV_FIELD_PHILOX;
}

__device__ int sample_fade(Rng& r, mtgv_tape_op* o) {  // ApplyChoice(fade_black, fade_white, brightness_contrast, None)
  int c = r.below(4);
  if (c == 3) return 0;
  if (c == 0) { tape_op_init(o, MTGV_OP_FADE_BLACK); o->d[0] = r.uniform(); }
  else if (c == 1) { tape_op_init(o, MTGV_OP_FADE_WHITE); o->d[0] = r.uniform(); }
  else { tape_op_init(o, MTGV_OP_BC); o->d[0] = r.uniform(-0.2, 0.2); o->d[1] = r.uniform(-0.2, 0.2); }
  return 1;
}
__device__ int sample_tint(Rng& r, mtgv_tape_op* o) {  // ApplyChoice(tint, None)
  if (r.below(2) != 0) return 0;
  tape_op_init(o, MTGV_OP_TINT);
  for (int k = 0; k < 3; k++) o->d[k] = r.uniform();
  return 1;
}
__device__ int sample_downup(Rng& r, mtgv_tape_op* o) {  // ApplyChoice(downscale_upscale, None, None, None)
  if (r.below(4) != 0) return 0;
  tape_op_init(o, MTGV_OP_DOWNUP);
  o->i[0] = r.below(3); o->i[1] = r.below(3); o->i[2] = r.below(3);
  return 1;
}
__device__ int sample_noise6(Rng& r, mtgv_tape_op* o, int H, int W) {
  int c = r.below(6);
  switch (c) {
    case 0: tape_op_init(o, MTGV_OP_NOISE); o->i[0] = r.below(4); o->d[0] = r.uniform(); return 1;
    case 1: tape_op_init(o, MTGV_OP_GAUSS_NOISE); return 1;
    case 2: tape_op_init(o, MTGV_OP_SALT_PEPPER); return 1;
    case 3: {
      tape_op_init(o, MTGV_OP_ERASE);
      o->d[0] = r.uniform(0.2, 0.4); o->d[1] = r.uniform(1.0, 3.0); o->d[2] = r.uniform();
      double area = o->d[0] * (double)(H * W), aspect = o->d[1];
      if (o->d[2] < 0.5) aspect = 1.0 / aspect;
      int bw = (int)sqrt(area / aspect), bh = (int)sqrt(area * aspect);
      o->i[4] = 1; o->i[5] = bw; o->i[6] = bh;
      int mx = -(bw / 2), Mx = W + bw / 2, my = -(bh / 2), My = H + bh / 2;
      if (Mx <= mx || My <= my) { o->i[0] = 0; return 1; }
      int cx = mx + r.below(Mx - mx), cy = my + r.below(My - my);
      o->i[1] = cx; o->i[2] = cy;
      int x0 = max(0, cx - bw / 2), y0 = max(0, cy - bh / 2), x1 = min(W, cx + bw / 2), y1 = min(H, cy + bh / 2);
      if (y1 <= y0 || x1 <= x0) { o->i[0] = 1; return 1; }
      o->i[0] = 2;
      o->i[3] = r.below(5);
      if (o->i[3] == 1) for (int k = 0; k < 3; k++) o->d[3 + k] = r.uniform();
      return 1;
    }
    case 4:
      tape_op_init(o, MTGV_OP_CUTOUT);
      for (int k = 0; k < 8; k++) { o->i[2 * k] = r.below(H); o->i[2 * k + 1] = r.below(W); }
      return 1;
    default: return 0;
  }
}

__device__ void shuffle_n(Rng& r, int* idx, int n) {  // random.shuffle (Fisher-Yates from the end)
  for (int i = n - 1; i >= 1; i--) { int j = r.below(i + 1); int t = idx[i]; idx[i] = idx[j]; idx[j] = t; }
}

__device__ void sample_virtual(Rng& r, mtgv_enc_tape* t, const mtgv_enc_config* cfg) {
  const int H = cfg->out_h, W = cfg->out_w;
  t->upsidedown = cfg->half_upsidedown ? (r.below(2) == 0) : 0;
  int n = 0;
  // _RAN_FG
  n += sample_downup(r, &t->ops[n]);
  {
    int c = r.below(4);
    if (c == 0) { tape_op_init(&t->ops[n], MTGV_OP_WARP); for (int k = 0; k < 8; k++) t->ops[n].d[k] = r.uniform(); n++; }
    else if (c == 1) {
      mtgv_tape_op* o = &t->ops[n++];
      tape_op_init(o, MTGV_OP_AFFINE);
      o->d[0] = r.uniform(-5.0, 5.0); o->d[1] = r.uniform(-10.0, 10.0); o->d[2] = r.uniform(-10.0, 10.0);
      double s = fmin(1.0 + 0.1, 1.0 / (1.0 + 0.1));
      o->d[3] = r.uniform(s, 1.0 / s); o->d[4] = r.uniform(-0.3, 0.3);
    } else if (c == 2) {
      tape_op_init(&t->ops[n], MTGV_OP_PERSPECTIVE);
      for (int k = 0; k < 8; k++) t->ops[n].d[k] = r.uniform(-0.1, 0.1);
      n++;
    }
  }
  n += sample_tint(r, &t->ops[n]);
  n += sample_fade(r, &t->ops[n]);
  t->n_fg = n;
  // _RAN_BG
  int g3[3] = {0, 1, 2};
  shuffle_n(r, g3, 3);
  for (int q = 0; q < 3; q++) {
    if (g3[q] == 0) {
      tape_op_init(&t->ops[n], MTGV_OP_FLIP);
      t->ops[n].i[0] = r.uniform() >= 0.5; t->ops[n].i[1] = r.uniform() >= 0.5; n++;
      tape_op_init(&t->ops[n], MTGV_OP_ROTATE); t->ops[n].d[0] = r.uniform(); n++;
      tape_op_init(&t->ops[n], MTGV_OP_WARP_INV); for (int k = 0; k < 8; k++) t->ops[n].d[k] = r.uniform(); n++;
    } else if (g3[q] == 1) n += sample_tint(r, &t->ops[n]);
    else n += sample_fade(r, &t->ops[n]);
  }
  t->n_bg = n - t->n_fg;
  // _RAN_VRTL
  int g7[7] = {0, 1, 2, 3, 4, 5, 6};
  shuffle_n(r, g7, 7);
  const int base = n;
  for (int q = 0; q < 7; q++) {
    switch (g7[q]) {
      case 0: n += sample_downup(r, &t->ops[n]); break;
      case 1: if (r.below(3) == 0) { tape_op_init(&t->ops[n], MTGV_OP_BLUR); t->ops[n].i[0] = r.below(2) * 2 + 1; n++; } break;
      case 2: if (r.below(3) == 0) { tape_op_init(&t->ops[n], MTGV_OP_SHARPEN); n++; } break;
      case 3: n += sample_noise6(r, &t->ops[n], H, W); break;
      case 4: if (r.below(2) == 0) n += sample_noise6(r, &t->ops[n], H, W); break;
      case 5: n += sample_tint(r, &t->ops[n]); break;
      default: n += sample_fade(r, &t->ops[n]); break;
    }
  }
  t->n_vrtl = n - base;
}

__global__ void k_sample_tape(uint64_t seed, int64_t first, int n_pairs, const mtgv_enc_config* cfg, PoolMeta pm,
                              const int32_t* cards, const int32_t* bgs, double p_tii, double p_neg, mtgv_enc_tape* tape) {
  // one warp per x-sample (lane 0 working), see k_expand; `which` selects x or x2 of pair i
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_xs = cfg->paired ? 2 : 1;
  const int i = w / n_xs, only = w - i * n_xs;
  if (i >= n_pairs || (threadIdx.x & 31) != 0) return;
  const uint64_t g = (uint64_t)(first + i);
  Rng sel(seed, g, 1);  // card / background selection stream (re-derivable by other samples)
  int card = sel.below(pm.n_cards);
  int bg = sel.below(pm.n_bgs);
  if (cards) card = cards[i];
  if (bgs) bg = bgs[i];
  if (p_tii < 0.0) p_tii = cfg->target_is_input_prob;
  if (p_neg < 0.0) p_neg = cfg->similar_neg_prob;
  for (int which = only; which <= only; which++) {
    mtgv_enc_tape* t = &tape[which * n_pairs + i];
    Rng r(seed, g, 2 + which);
    t->card = card; t->bg = bg; t->swap_choice = -1; t->upsidedown = 0;
    t->n_fg = t->n_bg = t->n_vrtl = 0;
    t->seed = seed ^ (0x9E3779B97F4A7C15ull * (2 * g + which + 1));
    if (which == 1) {
      // hard negative (encoder_train.py:217-221) and bg1 = random.choice(bg_imgs) (:224)
      if (r.uniform() < p_neg) {
        int m = pm.grp_off[card + 1] - pm.grp_off[card] - 1;
        if (m > 0) t->swap_choice = r.below(m);
      }
      int slot = r.below(n_pairs);
      if (bgs) {
        t->bg = bgs[slot];
      } else {
        Rng other(seed, (uint64_t)(first + slot), 1);
        other.below(pm.n_cards);
        t->bg = other.below(pm.n_bgs);
      }
    }
    if (r.uniform() < p_tii) {
      t->kind = MTGV_KIND_CROPPED;
    } else {
      t->kind = MTGV_KIND_VIRTUAL;
      sample_virtual(r, t, cfg);
    }
  }
}

int enc_sample_tape(mtgv_ctx* ctx, uint64_t seed, int64_t first, int n_pairs, const int32_t* cards, const int32_t* bgs,
                    double p_tii, double p_neg, mtgv_enc_tape* tape, cudaStream_t st) {
  if (n_pairs <= 0) return MTGV_OK;
  const int n_warps = n_pairs * (ctx->cfg.paired ? 2 : 1);
  k_sample_tape<<<(n_warps + 3) / 4, 128, 0, st>>>(seed, first, n_pairs, ctx->cfg_dev, pool_meta(ctx), cards, bgs, p_tii,
                                                    p_neg, tape);
  ctx->launches++;
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  return MTGV_OK;
}

// ------------------------------------------------------------------------------------ //
// batch launch                                                                          //
// ------------------------------------------------------------------------------------ //

static int ensure_scratch(mtgv_ctx* ctx, int n, size_t HW) {
  if ((size_t)n * HW > ctx->alpha_cap) {  // alpha_cap counts floats
    if (ctx->alpha_scratch) cudaFree(ctx->alpha_scratch);
    ctx->alpha_scratch = nullptr;
    ctx->alpha_cap = 0;
    MTGV_CUDA_OK(ctx, cudaMalloc(&ctx->alpha_scratch, (size_t)n * HW * 4));
    ctx->alpha_cap = (size_t)n * HW;
  }
  if ((size_t)n + 1 > ctx->sync_cap) {
    if (ctx->sync_words) cudaFree(ctx->sync_words);
    ctx->sync_words = nullptr;
    ctx->sync_cap = 0;
    MTGV_CUDA_OK(ctx, cudaMalloc(&ctx->sync_words, ((size_t)n + 1) * 4));
    ctx->sync_cap = (size_t)n + 1;
  }
  return MTGV_OK;
}

static int enc_batch_sized(mtgv_ctx* ctx, const mtgv_enc_params* params, int n, void* out, int out_dtype, const void* fields,
                           int OH, int OW, cudaStream_t st);

int enc_batch(mtgv_ctx* ctx, const mtgv_enc_params* params, int n, void* out, int out_dtype, const void* fields,
              cudaStream_t st) {
  return enc_batch_sized(ctx, params, n, out, out_dtype, fields, ctx->cfg.out_h, ctx->cfg.out_w, st);
}

static int enc_batch_sized(mtgv_ctx* ctx, const mtgv_enc_params* params, int n, void* out, int out_dtype, const void* fields,
                           int OH, int OW, cudaStream_t st) {
  if (n <= 0) return MTGV_OK;
  const size_t HW = (size_t)OH * OW;
  const size_t smem = enc_smem_bytes(OH, OW);
  if ((int)smem > ctx->max_smem_optin)
    return fail(ctx, MTGV_ERR_LIMIT, "x_size_hw too large: two float32 planes must fit in 227 KB of shared memory");
  // Chunk = as many samples as 1.5 GB of float32 scratch (background + foreground planes) holds: the
  // persistent kernels lose more to per-launch tails than the pipeline gains from L2 residency of the
  // scratch (measured: 6 chunks of 190 samples 79k x-samples/s, one chunk of 1024 92k).
  int chunk = (int)((size_t)1536 * 1024 * 1024 / (6 * HW * 4));
  if (const char* e = getenv("MTGV_CHUNK")) chunk = atoi(e);
  chunk = chunk < 1 ? 1 : (chunk > n ? n : chunk);
  int rc = ensure_scratch(ctx, chunk, HW);
  if (rc) return rc;
  if ((size_t)chunk * 6 * HW > ctx->bg_cap) {  // background planes followed by foreground planes
    if (ctx->bg_scratch) cudaFree(ctx->bg_scratch);
    ctx->bg_scratch = nullptr; ctx->bg_cap = 0;
    MTGV_CUDA_OK(ctx, cudaMalloc(&ctx->bg_scratch, (size_t)chunk * 6 * HW * 4));
    ctx->bg_cap = (size_t)chunk * 6 * HW;
  }
  float* fg_scratch = ctx->bg_scratch + (size_t)chunk * 3 * HW;
  if (!(ctx->attrs_set & 4u)) {
    MTGV_CUDA_OK(ctx, cudaFuncSetAttribute(k_encoder, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin));
    ctx->attrs_set |= 4u;
  }
  const size_t elem = out_dtype == MTGV_OUT_F16 ? 2 : (out_dtype == MTGV_OUT_U8 ? 1 : 4);
  for (int base = 0; base < n; base += chunk) {
    const int m = n - base < chunk ? n - base : chunk;
    MTGV_CUDA_OK(ctx, cudaMemsetAsync(ctx->sync_words, 0, ((size_t)m + 1) * 4, st));
    if (ctx->n_bgs > 0) {
      int rc2 = bg_launch(ctx, params + base, m, OH, OW, ctx->bg_scratch, st);
      if (rc2) return rc2;
    }
    {
      int rc3 = fg_launch(ctx, params + base, m, OH, OW, fg_scratch, st);
      if (rc3) return rc3;
    }
    EncLaunch L;
    L.params = params + base; L.n = m; L.fg_scratch = fg_scratch;
    L.alpha0 = ctx->alpha0; L.bg_scratch = ctx->bg_scratch; L.alpha_scratch = ctx->alpha_scratch;
    L.sync_words = ctx->sync_words; L.out = (char*)out + (size_t)base * 3 * HW * elem; L.out_dtype = out_dtype;
    L.fields = (const uint32_t*)fields;
    int grid = 4 * m < ctx->sm_count ? 4 * m : ctx->sm_count;
    k_encoder<<<grid, kThreads, smem, st>>>(L, OH, OW);
    ctx->launches++;
  }
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  return MTGV_OK;
}

__global__ void k_target_params(const int32_t* cards, int n, const mtgv_enc_config* cfg, PoolMeta pm, mtgv_enc_params* params) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  mtgv_enc_tape t;
  t.kind = MTGV_KIND_CROPPED; t.card = cards[s]; t.swap_choice = -1; t.bg = 0; t.upsidedown = 0;
  t.n_fg = t.n_bg = t.n_vrtl = 0; t.seed = 0;
  mtgv_enc_config c = *cfg;
  c.out_h = cfg->y_h; c.out_w = cfg->y_w;
  expand_encoder_sample(&t, &c, pm, &params[s]);
}

int enc_targets(mtgv_ctx* ctx, const int32_t* cards, int n, void* out, int out_dtype, cudaStream_t st) {
  if (n <= 0) return MTGV_OK;
  if ((size_t)n > ctx->tmp_params_cap) {
    if (ctx->tmp_params) cudaFree(ctx->tmp_params);
    ctx->tmp_params = nullptr; ctx->tmp_params_cap = 0;
    MTGV_CUDA_OK(ctx, cudaMalloc(&ctx->tmp_params, (size_t)n * sizeof(mtgv_enc_params)));
    ctx->tmp_params_cap = n;
  }
  k_target_params<<<(n + 63) / 64, 64, 0, st>>>(cards, n, ctx->cfg_dev, pool_meta(ctx), ctx->tmp_params);
  ctx->launches++;
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  return enc_batch_sized(ctx, ctx->tmp_params, n, out, out_dtype, nullptr, ctx->cfg.y_h, ctx->cfg.y_w, st);
}

// ------------------------------------------------------------------------------------ //
// parity / debug entries                                                                //
// ------------------------------------------------------------------------------------ //

__global__ void k_warp_perspective(const float* __restrict__ src, int sh, int sw, int c, const double* __restrict__ M,
                                   float* __restrict__ dst, int dh, int dw) {
  __shared__ double Mi[9];
  const int img = blockIdx.y;
  if (threadIdx.x == 0) invert3x3(M + (size_t)img * 9, Mi);  // cv::invert inside cv::warpPerspective
  __syncthreads();
  const int bw0 = persp_block_w(dh, dw);
  const float* S = src + (size_t)img * sh * sw * c;
  float* D = dst + (size_t)img * dh * dw * c;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < dh * dw; i += gridDim.x * blockDim.x) {
    // the segment-relative fast path of the production kernels (mtgv_persp.cuh), with the exact routine as its
    // fallback: this entry is the one compared bit for bit against cv2.warpPerspective
    int X, Y;
    {
      const int x = i % dw, y = i / dw;
      PerspSeg sg;
      persp_seg_build(Mi, x >> kPerspSegShift, y, bw0, &sg);
      if (!persp_seg_eval(sg, (float)((x & ((1 << kPerspSegShift) - 1)) - (1 << (kPerspSegShift - 1))), &X, &Y))
        persp_coord(Mi, x, y, bw0, &X, &Y);
    }
    int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
    bool x0 = (unsigned)sx < (unsigned)sw, x1 = (unsigned)(sx + 1) < (unsigned)sw;
    bool y0 = (unsigned)sy < (unsigned)sh, y1 = (unsigned)(sy + 1) < (unsigned)sh;
    const float* p = S + ((size_t)sy * sw + sx) * c;
    for (int k = 0; k < c; k++) {
      float v0 = (y0 && x0) ? p[k] : 0.f, v1 = (y0 && x1) ? p[c + k] : 0.f;
      float v2 = (y1 && x0) ? p[(size_t)sw * c + k] : 0.f, v3 = (y1 && x1) ? p[(size_t)sw * c + c + k] : 0.f;
      D[(size_t)i * c + k] = bilinear_weights_sum(v0, v1, v2, v3, X & 31, Y & 31);
    }
  }
}

int enc_warp_perspective(mtgv_ctx* ctx, const float* src, int n, int sh, int sw, int c, const double* M, float* dst, int dh,
                         int dw, cudaStream_t st) {
  if (n <= 0) return MTGV_OK;
  dim3 grid((dh * dw + 255) / 256 < 592 ? (dh * dw + 255) / 256 : 592, n);
  k_warp_perspective<<<grid, 256, 0, st>>>(src, sh, sw, c, M, dst, dh, dw);
  ctx->launches++;
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  return MTGV_OK;
}

__global__ void __launch_bounds__(kThreads, 1) k_run_plane_ops(float* img, int h, int w, int c, const mtgv_x_op* ops, int n_ops,
                                                              const uint32_t* fields, uint64_t seed) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int HW = h * w;
  float* P0 = (float*)smem_raw;
  float* P1 = P0 + HW;
  double* aux = (double*)(smem_raw + (((size_t)2 * HW * 4 + 15) & ~(size_t)15));
  mtgv_x_op* sop = (mtgv_x_op*)((unsigned char*)aux + ((vm_aux_bytes(h, w) + 15) & ~(size_t)15));
  const int im = blockIdx.x / c, ch = blockIdx.x % c;
  float* base = img + (size_t)im * HW * c;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) P0[i] = base[(size_t)i * c + ch];
  Vm vm{{P0, P1}, 0, h, w, ch, fields, seed + (uint64_t)im, aux};
  __syncthreads();
  for (int k = 0; k < n_ops; k++) {
    for (int q = threadIdx.x; q < (int)(sizeof(mtgv_x_op) / 4); q += blockDim.x) ((uint32_t*)sop)[q] = ((const uint32_t*)(ops + k))[q];
    __syncthreads();
    vm_run_op(vm, *sop, k);
  }
  const float* r = vm.cur_p();
  for (int i = threadIdx.x; i < HW; i += blockDim.x) base[(size_t)i * c + ch] = r[i];
}

int enc_run_plane_ops(mtgv_ctx* ctx, float* img, int n, int h, int w, int c, const mtgv_x_op* ops, int n_ops,
                      const void* fields, uint64_t seed, cudaStream_t st) {
  if (n <= 0) return MTGV_OK;
  size_t smem = (((size_t)2 * h * w * 4 + 15) & ~(size_t)15) + ((vm_aux_bytes(h, w) + 15) & ~(size_t)15) + sizeof(mtgv_x_op) + 16;
  if ((int)smem > ctx->max_smem_optin) return fail(ctx, MTGV_ERR_LIMIT, "image too large for the plane interpreter");
  if (c < 1 || c > 4) return fail(ctx, MTGV_ERR_INVALID, "channels must be 1..4");
  if (!(ctx->attrs_set & 8u)) {
    MTGV_CUDA_OK(ctx, cudaFuncSetAttribute(k_run_plane_ops, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin));
    ctx->attrs_set |= 8u;
  }
  k_run_plane_ops<<<n * c, kThreads, smem, st>>>(img, h, w, c, ops, n_ops, (const uint32_t*)fields, seed);
  ctx->launches++;
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  return MTGV_OK;
}

}  // namespace mtgv
