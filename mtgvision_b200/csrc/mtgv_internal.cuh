// mtgv_internal.cuh - context object and helpers shared by the translation units of libmtgv.so
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/mtgv.h"
#include "mtgv_expand.cuh"
#include "mtgv_mask.h"

struct mtgv_ctx {
  int device = 0;
  int sm_count = 148;
  int max_smem_optin = 0;
  std::string err;
  int64_t launches = 0;
  unsigned attrs_set = 0;
  uint64_t card_epoch = 0;  // bumped whenever card pixels change (set / update): derived copies rebuild lazily  // per-context (= per-device) bits: which kernels had their function attributes set

  // card pool (planar copy: [n][3][h][pitch] uint8, pitch multiple of 16)
  uint8_t* card_planes = nullptr;
  int n_cards = 0, card_h = 0, card_w = 0, card_pitch = 0;
  int32_t* labels3 = nullptr;
  int32_t* grp_off = nullptr;
  int32_t* grp_mem = nullptr;
  float* mask_enc = nullptr;  // round_rect_mask(card_hw, 0.05) float32, encoder make_masked
  float* mask_det = nullptr;  // round_rect_mask(card_hw, 0.046) float32, detection make_card_with_mask

  // background pool: per image [h][pitchw] RGBX uint8 words, pitchw = round_up(w,4) (rows are whole 16-byte vectors)
  uint8_t* bg_planes = nullptr;
  int64_t* bg_off = nullptr;  // [n] byte offset of image j in bg_planes
  int32_t* bg_hw = nullptr;   // [n][2] (h, w)
  int n_bgs = 0;
  std::vector<int64_t> bg_off_host;
  std::vector<int32_t> bg_hw_host;

  // encoder
  mtgv_enc_config cfg{};
  bool cfg_set = false;
  mtgv_enc_config* cfg_dev = nullptr;
  float* alpha0 = nullptr;  // static foreground alpha [out_h,out_w]: pad(area_resize(mask_enc))
  float* alpha_scratch = nullptr;
  float* bg_scratch = nullptr;  // k_background output for one chunk: [chunk,3,H,W] float32
  size_t bg_cap = 0;            // floats
  int bg_blocks_per_sm = 0;
  int* bg_counter = nullptr;    // work-queue word of k_background
  int* fg_counter = nullptr;    // work-queue word of k_foreground
  size_t alpha_cap = 0;  // samples
  int32_t* sync_words = nullptr;  // [0] work counter, [1..] per-sample alpha-ready flags
  size_t sync_cap = 0;
  mtgv_enc_params* tmp_params = nullptr;  // scratch for mtgv_encoder_targets
  size_t tmp_params_cap = 0;

  // detection state lives in mtgv_det.cu
  void* det = nullptr;
  // JPEG decode scratch lives in mtgv_jpeg.cu
  void* jpeg = nullptr;
  // JPEG encode scratch lives in mtgv_jpegenc.cu
  void* jpegenc = nullptr;
};

namespace mtgv {

inline int fail(mtgv_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

#define MTGV_CUDA_OK(ctx, expr)                                                                   \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      return mtgv::fail((ctx), MTGV_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based so every sample / pixel is addressable ----
struct Philox {
  uint32_t key[2];
  __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
    uint64_t p = (uint64_t)a * b;
    *hi = (uint32_t)(p >> 32);
    *lo = (uint32_t)p;
  }
  __host__ __device__ inline void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t* out) const {
    uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
    for (int r = 0; r < 10; r++) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(0xD2511F53u, c0, &hi0, &lo0);
      mulhilo(0xCD9E8D57u, c2, &hi1, &lo1);
      uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};

}  // namespace mtgv

#if defined(__CUDACC__)
namespace mtgv {
// x and x2 of a positive pair usually resize the SAME card with the same geometry (make_virtual of the same
// card, no hard-negative swap): the area-resized foreground planes are identical, so sample s of the second half
// of a launch reuses the planes of sample s - n/2 when every input of the resize matches.  Evaluated
// identically by k_foreground (skip) and k_encoder (read the partner's planes).
__device__ __forceinline__ int fg_plane_owner(const mtgv_enc_params* __restrict__ p, int n, int s) {
  if ((n & 1) || s < (n >> 1)) return s;
  const mtgv_enc_params& a = p[s];
  const mtgv_enc_params& b = p[s - (n >> 1)];
  const bool same = a.status == 0 && b.status == 0 && a.kind == b.kind && a.kind != MTGV_KIND_BG_ONLY && a.card == b.card &&
                    a.upsidedown == b.upsidedown && a.out_h == b.out_h && a.out_w == b.out_w && a.src_y0 == b.src_y0 &&
                    a.src_x0 == b.src_x0 && a.src_h == b.src_h && a.src_w == b.src_w && a.fg_rh == b.fg_rh && a.fg_rw == b.fg_rw &&
                    a.fg_y0 == b.fg_y0 && a.fg_x0 == b.fg_x0;
  return same ? s - (n >> 1) : s;
}
}  // namespace mtgv
#endif

// entry points implemented in mtgv_enc.cu, called from mtgv_api.cu
namespace mtgv {
int enc_build_static_alpha(mtgv_ctx* ctx, cudaStream_t st);
int enc_sample_tape(mtgv_ctx* ctx, uint64_t seed, int64_t first, int n_pairs, const int32_t* cards, const int32_t* bgs, double p_tii,
                    double p_neg, mtgv_enc_tape* tape, cudaStream_t st);
int enc_expand(mtgv_ctx* ctx, const mtgv_enc_tape* tape, int n, mtgv_enc_params* params, int64_t* labels, cudaStream_t st);
int enc_batch(mtgv_ctx* ctx, const mtgv_enc_params* params, int n, void* out, int out_dtype, const void* fields,
              cudaStream_t st);
int enc_targets(mtgv_ctx* ctx, const int32_t* cards, int n, void* out, int out_dtype, cudaStream_t st);
int enc_warp_perspective(mtgv_ctx* ctx, const float* src, int n, int sh, int sw, int c, const double* M, float* dst, int dh,
                         int dw, cudaStream_t st);
int enc_run_plane_ops(mtgv_ctx* ctx, float* img, int n, int h, int w, int c, const mtgv_x_op* ops, int n_ops,
                      const void* fields, uint64_t seed, cudaStream_t st);
int pool_planarize(mtgv_ctx* ctx, const uint8_t* hwc, uint8_t* planes, int n, int h, int w, int pitch, cudaStream_t st);
// mtgv_bg.cu
int pool_interleave(mtgv_ctx* ctx, const uint8_t* hwc, uint8_t* words, int n, int h, int w, cudaStream_t st);
size_t bg_image_bytes(int h, int w);
// mtgv_dewarp.cu
int dewarp_u8(mtgv_ctx* ctx, const uint8_t* frame, int fh, int fw, int fc, const float* quads, int n, const float* dst_rect,
              uint8_t* out, int oh, int ow, cudaStream_t st);
// mtgv_jpeg.cu
int jpeg_destroy(mtgv_ctx* ctx);
int jpeg_last_kernel_ms(mtgv_ctx* ctx, float* ms);
int jpeg_info(mtgv_ctx* ctx, const uint8_t* file, int64_t len, int32_t* hw);
struct JpegDst {  // where one decoded file goes: byte offset from the call's `out`, expected frame size, layout (JpegImg::out_layout), pitch
  int64_t off;
  int32_t h, w, layout, pitch;
};
int jpeg_decode_batch_ex(mtgv_ctx* ctx, const uint8_t* files, const int64_t* file_off, int n, uint8_t* out, const JpegDst* dst,
                         cudaStream_t stream);
int jpeg_info_batch(mtgv_ctx* ctx, const uint8_t* files, const int64_t* file_off, int n, int32_t* hw);
int jpeg_decode_batch(mtgv_ctx* ctx, const uint8_t* files, const int64_t* file_off, int n, uint8_t* out, const int64_t* out_off,
                      const int32_t* hw, cudaStream_t st);
// mtgv_jpegenc.cu
int jpegenc_destroy(mtgv_ctx* ctx);
int jpegenc_last_kernel_ms(mtgv_ctx* ctx, float* ms);
int jpegenc_batch(mtgv_ctx* ctx, const uint8_t* images, int n, int h, int w, int layout, int quality, uint8_t* out, int64_t cap,
                  int32_t* out_len, cudaStream_t st);
int jpegenc_compact(mtgv_ctx* ctx, const uint8_t* slots, int64_t cap, const int32_t* out_len, int n, uint8_t* compact, int64_t* offsets,
                    cudaStream_t st);
// mtgv_fg.cu
int fg_launch(mtgv_ctx* ctx, const mtgv_enc_params* params, int m, int OH, int OW, float* fg_out, cudaStream_t st);
int bg_launch(mtgv_ctx* ctx, const mtgv_enc_params* params, int m, int OH, int OW, float* bg_out, cudaStream_t st);
}  // namespace mtgv
