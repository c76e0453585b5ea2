// tex_probe.cu - GPU-side probes that decide kernel design (run through gpurun, results in profiles/):
//  1. precision of hardware bilinear filtering on RGBA8 textures (normalized-float read) at the
//     1/32-pixel weights cv2.warpPerspective/warpAffine use, against exact arithmetic;
//  2. zero border (cudaAddressModeBorder) behaviour at the image edge;
//  3. throughput of TEX (linear filter) vs 4x LDG.32 gathers under a rotated access pattern;
//  4. fp64 vs fp32 FMA issue rate on this part.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void k_filter(cudaTextureObject_t tex, int w, int h, int n, const int* __restrict__ X, const int* __restrict__ Y, float4* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = (float)(X[i] >> 5) + (float)(X[i] & 31) * 0.03125f + 0.5f;
  float y = (float)(Y[i] >> 5) + (float)(Y[i] & 31) * 0.03125f + 0.5f;
  out[i] = tex2D<float4>(tex, x, y);
}

__global__ void k_point(cudaTextureObject_t tex, int n, const int* __restrict__ X, const int* __restrict__ Y, float4* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = tex2D<float4>(tex, (float)(X[i] >> 5) + 0.5f, (float)(Y[i] >> 5) + 0.5f);
}

// rotated sweep: every thread samples a W x H canvas rotated by ~37 degrees, like rotate_bounded
__global__ void k_tex_rot(cudaTextureObject_t tex, int W, int H, float ca, float sa, float4* out, int reps) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int r = 0; r < reps; r++) {
    float fx = ca * x + sa * (y + r) + 0.5f, fy = -sa * x + ca * (y + r) + 100.5f;
    float4 v = tex2D<float4>(tex, fx, fy);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  if (x < W) out[(size_t)y * W + x] = acc;
}

__global__ void k_ldg_rot(const uint32_t* __restrict__ img, int iw, int ih, int pitch_words, int W, int H, float ca, float sa, float4* out, int reps) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int r = 0; r < reps; r++) {
    float fx = ca * x + sa * (y + r), fy = -sa * x + ca * (y + r) + 100.f;
    int X = __float2int_rn(fx * 32.f), Y = __float2int_rn(fy * 32.f);
    int sx = X >> 5, sy = Y >> 5;
    float ax = (X & 31) * 0.03125f, ay = (Y & 31) * 0.03125f;
    float w00 = (1 - ay) * (1 - ax), w01 = (1 - ay) * ax, w10 = ay * (1 - ax), w11 = ay * ax;
    uint32_t t00 = 0, t01 = 0, t10 = 0, t11 = 0;
    bool x0 = (unsigned)sx < (unsigned)iw, x1 = (unsigned)(sx + 1) < (unsigned)iw, y0 = (unsigned)sy < (unsigned)ih, y1 = (unsigned)(sy + 1) < (unsigned)ih;
    const uint32_t* p = img + (size_t)sy * pitch_words + sx;
    if (y0 && x0) t00 = __ldg(p);
    if (y0 && x1) t01 = __ldg(p + 1);
    if (y1 && x0) t10 = __ldg(p + pitch_words);
    if (y1 && x1) t11 = __ldg(p + pitch_words + 1);
    const float k = 1.f / 255.f;
    acc.x += ((t00 & 255) * w00 + (t01 & 255) * w01 + (t10 & 255) * w10 + (t11 & 255) * w11) * k;
    acc.y += (((t00 >> 8) & 255) * w00 + ((t01 >> 8) & 255) * w01 + ((t10 >> 8) & 255) * w10 + ((t11 >> 8) & 255) * w11) * k;
    acc.z += (((t00 >> 16) & 255) * w00 + ((t01 >> 16) & 255) * w01 + ((t10 >> 16) & 255) * w10 + ((t11 >> 16) & 255) * w11) * k;
  }
  if (x < W) out[(size_t)y * W + x] = acc;
}

template <typename T>
__global__ void k_fma_rate(T* out, int iters) {
  T a = (T)threadIdx.x * (T)1e-3, b = (T)1.0000001, c = (T)1e-7;
  T a2 = a + 1, a3 = a + 2, a4 = a + 3;
  for (int i = 0; i < iters; i++) {
    a = a * b + c; a2 = a2 * b + c; a3 = a3 * b + c; a4 = a4 * b + c;
    a = a * b + c; a2 = a2 * b + c; a3 = a3 * b + c; a4 = a4 * b + c;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + a2 + a3 + a4;
}

static cudaTextureObject_t make_tex(void* dev, int w, int h, size_t pitch, cudaTextureFilterMode fm, cudaArray_t arr) {
  cudaResourceDesc rd{};
  if (arr) { rd.resType = cudaResourceTypeArray; rd.res.array.array = arr; }
  else {
    rd.resType = cudaResourceTypePitch2D; rd.res.pitch2D.devPtr = dev; rd.res.pitch2D.width = w; rd.res.pitch2D.height = h;
    rd.res.pitch2D.pitchInBytes = pitch; rd.res.pitch2D.desc = cudaCreateChannelDesc<uchar4>();
  }
  cudaTextureDesc td{};
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
  td.filterMode = fm; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 0;
  cudaTextureObject_t t; CK(cudaCreateTextureObject(&t, &rd, &td, nullptr));
  return t;
}

int main() {
  const int w = 500, h = 375;
  size_t pitch = ((size_t)w * 4 + 511) & ~(size_t)511;
  std::vector<uint8_t> img(pitch * h);
  srand(7);
  for (auto& b : img) b = rand() & 255;
  uint8_t* d_img; CK(cudaMalloc(&d_img, pitch * h)); CK(cudaMemcpy(d_img, img.data(), pitch * h, cudaMemcpyHostToDevice));
  cudaArray_t arr; cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
  CK(cudaMallocArray(&arr, &cd, w, h)); CK(cudaMemcpy2DToArray(arr, 0, 0, img.data(), pitch, (size_t)w * 4, h, cudaMemcpyHostToDevice));
  // sample points: all 1024 sub-pixel phases at random interior positions + edge positions
  std::vector<int> X, Y;
  for (int rep = 0; rep < 64; rep++)
    for (int ay = 0; ay < 32; ay++) for (int ax = 0; ax < 32; ax++) {
      int sx = rand() % (w + 4) - 2, sy = rand() % (h + 4) - 2;  // includes -2..-1 and w..w+1 (border)
      X.push_back(sx * 32 + ax); Y.push_back(sy * 32 + ay);
    }
  int n = (int)X.size();
  int *dX, *dY; float4* dout;
  CK(cudaMalloc(&dX, n * 4)); CK(cudaMalloc(&dY, n * 4)); CK(cudaMalloc(&dout, n * sizeof(float4)));
  CK(cudaMemcpy(dX, X.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dY, Y.data(), n * 4, cudaMemcpyHostToDevice));
  std::vector<float4> out(n);
  for (int mode = 0; mode < 2; mode++) {
    cudaTextureObject_t tl = make_tex(d_img, w, h, pitch, cudaFilterModeLinear, mode ? arr : nullptr);
    k_filter<<<(n + 255) / 256, 256>>>(tl, w, h, n, dX, dY, dout); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), dout, n * sizeof(float4), cudaMemcpyDeviceToHost));
    double max_err = 0, max_err_f32 = 0; int n_bits_exact = 0; long bad_border = 0;
    for (int i = 0; i < n; i++) {
      int sx = X[i] >> 5, sy = Y[i] >> 5, ax = X[i] & 31, ay = Y[i] & 31;
      for (int c = 0; c < 4; c++) {
        auto px = [&](int yy, int xx) -> int { return (xx < 0 || yy < 0 || xx >= w || yy >= h) ? 0 : img[yy * pitch + xx * 4 + c]; };
        int b00 = px(sy, sx), b01 = px(sy, sx + 1), b10 = px(sy + 1, sx), b11 = px(sy + 1, sx + 1);
        double exact = ((32 - ay) * ((32 - ax) * b00 + ax * b01) + ay * ((32 - ax) * b10 + ax * b11)) / (1024.0 * 255.0);
        // cv2-order float32 reference: sum of tap*weight with separate roundings
        float fx = ax * 0.03125f, fy = ay * 0.03125f, gx = 1.f - fx, gy = 1.f - fy;
        volatile float s = (float)(b00 / 255.0) * (gy * gx); s = s + (float)(b01 / 255.0) * (gy * fx);
        s = s + (float)(b10 / 255.0) * (fy * gx); s = s + (float)(b11 / 255.0) * (fy * fx);
        float got = ((float*)&out[i])[c];
        max_err = fmax(max_err, fabs(got - exact)); max_err_f32 = fmax(max_err_f32, fabs((double)got - (double)s));
        if (got == s) n_bits_exact++;
        if ((sx < -1 || sy < -1 || sx >= w || sy >= h) && got != 0.f) bad_border++;
      }
    }
    printf("{\"probe\": \"tex_linear_rgba8\", \"resource\": \"%s\", \"samples\": %d, \"max_abs_err_vs_exact\": %.3e, \"max_abs_err_vs_cv2_f32_order\": %.3e, \"bit_equal_frac\": %.4f, \"nonzero_outside_border\": %ld}\n",
           mode ? "array" : "pitch2D", n * 4, max_err, max_err_f32, n_bits_exact / (4.0 * n), bad_border);
    cudaTextureObject_t tp = make_tex(d_img, w, h, pitch, cudaFilterModePoint, mode ? arr : nullptr);
    k_point<<<(n + 255) / 256, 256>>>(tp, n, dX, dY, dout); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), dout, n * sizeof(float4), cudaMemcpyDeviceToHost));
    int neq = 0; double perr = 0;
    for (int i = 0; i < n; i++) {
      int sx = X[i] >> 5, sy = Y[i] >> 5;
      for (int c = 0; c < 4; c++) {
        int b = (sx < 0 || sy < 0 || sx >= w || sy >= h) ? 0 : img[sy * pitch + sx * 4 + c];
        float ref = (float)(b / 255.0); float got = ((float*)&out[i])[c];
        if (got == ref) neq++; perr = fmax(perr, fabs((double)got - b / 255.0));
      }
    }
    printf("{\"probe\": \"tex_point_rgba8\", \"resource\": \"%s\", \"equals_fl(b/255)_frac\": %.4f, \"max_abs_err\": %.3e}\n", mode ? "array" : "pitch2D", neq / (4.0 * n), perr);
    cudaDestroyTextureObject(tl); cudaDestroyTextureObject(tp);
  }
  // throughput
  {
    const int W = 640, H = 4096, reps = 16;
    float4* big; CK(cudaMalloc(&big, (size_t)W * H * sizeof(float4)));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ca = cosf(0.65f) * 0.7f, sa = sinf(0.65f) * 0.7f;
    for (int mode = 0; mode < 3; mode++) {
      cudaTextureObject_t tl = mode < 2 ? make_tex(d_img, w, h, pitch, cudaFilterModeLinear, mode ? arr : nullptr) : 0;
      float best = 1e9;
      for (int it = 0; it < 5; it++) {
        cudaEventRecord(e0);
        if (mode < 2) k_tex_rot<<<dim3(W / 128, H), 128>>>(tl, W, H, ca, sa, big, reps);
        else k_ldg_rot<<<dim3(W / 128, H), 128>>>((const uint32_t*)d_img, w, h, (int)(pitch / 4), W, H, ca, sa, big, reps);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms);
      }
      double samples = (double)W * H * reps;
      printf("{\"probe\": \"rotated_bilinear_rgba8_throughput\", \"path\": \"%s\", \"ms\": %.3f, \"Gsamples_per_s\": %.2f}\n",
             mode == 0 ? "tex_linear_pitch2D" : mode == 1 ? "tex_linear_array" : "4xLDG32+fp32", best, samples / best * 1e-6);
    }
  }
  // fma issue rates
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    void* o; CK(cudaMalloc(&o, (size_t)blocks * threads * 8));
    for (int mode = 0; mode < 2; mode++) {
      float best = 1e9;
      for (int it = 0; it < 4; it++) {
        cudaEventRecord(e0);
        if (mode) k_fma_rate<double><<<blocks, threads>>>((double*)o, iters); else k_fma_rate<float><<<blocks, threads>>>((float*)o, iters);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms);
      }
      double fma = (double)blocks * threads * iters * 8;
      printf("{\"probe\": \"fma_rate\", \"type\": \"%s\", \"ms\": %.3f, \"TFMA_per_s\": %.3f}\n", mode ? "f64" : "f32", best, fma / best * 1e-9);
    }
  }
  return 0;
}
