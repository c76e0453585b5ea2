// mtgv_api.cu - C ABI of libmtgv.so (include/mtgv.h): context, pools, entry points.
#include <math.h>

#include <vector>
#include <algorithm>
#include <thread>
#include <cstring>

#include "mtgv_internal.cuh"

using namespace mtgv;

namespace mtgv {
int det_destroy(mtgv_ctx* ctx);
}

extern "C" {

int mtgv_abi_version(void) { return MTGV_ABI_VERSION; }

int mtgv_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(mtgv_tape_op);
    case 1: return (int)sizeof(mtgv_enc_tape);
    case 2: return (int)sizeof(mtgv_x_op);
    case 3: return (int)sizeof(mtgv_enc_params);
    case 4: return (int)sizeof(mtgv_enc_config);
    case 5: return (int)sizeof(mtgv_photo_op);
    case 6: return (int)sizeof(mtgv_det_attempt);
    case 7: return (int)sizeof(mtgv_det_card);
    case 8: return (int)sizeof(mtgv_det_tape);
    case 9: return (int)sizeof(mtgv_det_config);
    default: return -1;
  }
}

mtgv_ctx* mtgv_create(int device) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return nullptr;
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  mtgv_ctx* ctx = new mtgv_ctx();
  ctx->device = device;
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  if (cudaMalloc(&ctx->cfg_dev, sizeof(mtgv_enc_config)) != cudaSuccess) {
    delete ctx;
    return nullptr;
  }
  return ctx;
}

static void free_cards(mtgv_ctx* ctx) {
  cudaFree(ctx->card_planes); cudaFree(ctx->labels3); cudaFree(ctx->grp_off); cudaFree(ctx->grp_mem);
  cudaFree(ctx->mask_enc); cudaFree(ctx->mask_det);
  ctx->card_planes = nullptr; ctx->labels3 = ctx->grp_off = ctx->grp_mem = nullptr;
  ctx->mask_enc = ctx->mask_det = nullptr;
  ctx->n_cards = 0;
}
static void free_bgs(mtgv_ctx* ctx) {
  cudaFree(ctx->bg_planes); cudaFree(ctx->bg_off); cudaFree(ctx->bg_hw);
  ctx->bg_planes = nullptr; ctx->bg_off = nullptr; ctx->bg_hw = nullptr; ctx->n_bgs = 0;
}

void mtgv_destroy(mtgv_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  det_destroy(ctx);
  jpeg_destroy(ctx);
  jpegenc_destroy(ctx);
  free_cards(ctx);
  free_bgs(ctx);
  cudaFree(ctx->cfg_dev); cudaFree(ctx->alpha0); cudaFree(ctx->alpha_scratch); cudaFree(ctx->sync_words);
  cudaFree(ctx->tmp_params); cudaFree(ctx->bg_scratch); cudaFree(ctx->bg_counter); cudaFree(ctx->fg_counter);
  delete ctx;
}

const char* mtgv_last_error(const mtgv_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int64_t mtgv_launch_count(const mtgv_ctx* ctx) { return ctx ? ctx->launches : 0; }

int mtgv_set_card_pool(mtgv_ctx* ctx, const uint8_t* cards, int n, int h, int w, const int32_t* labels3, const int32_t* grp_off,
                       const int32_t* grp_mem, int n_mem) {
  if (!ctx) return MTGV_ERR_INVALID;
  if (!cards || n <= 0 || h < 8 || w < 8 || !labels3 || !grp_off || !grp_mem || n_mem < n)
    return fail(ctx, MTGV_ERR_INVALID, "mtgv_set_card_pool: bad arguments");
  MTGV_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  free_cards(ctx);
  ctx->card_h = h; ctx->card_w = w; ctx->card_pitch = round_up(w, 16);
  MTGV_CUDA_OK(ctx, cudaMalloc(&ctx->card_planes, (size_t)n * 3 * h * ctx->card_pitch));
  MTGV_CUDA_OK(ctx, cudaMalloc(&ctx->labels3, (size_t)n * 3 * 4));
  MTGV_CUDA_OK(ctx, cudaMalloc(&ctx->grp_off, ((size_t)n + 1) * 4));
  MTGV_CUDA_OK(ctx, cudaMalloc(&ctx->grp_mem, (size_t)n_mem * 4));
  MTGV_CUDA_OK(ctx, cudaMemcpy(ctx->labels3, labels3, (size_t)n * 3 * 4, cudaMemcpyDefault));
  MTGV_CUDA_OK(ctx, cudaMemcpy(ctx->grp_off, grp_off, ((size_t)n + 1) * 4, cudaMemcpyDefault));
  MTGV_CUDA_OK(ctx, cudaMemcpy(ctx->grp_mem, grp_mem, (size_t)n_mem * 4, cudaMemcpyDefault));
  int rc = pool_planarize(ctx, cards, ctx->card_planes, n, h, w, ctx->card_pitch, 0);
  if (rc) return rc;
  // static masks: make_masked uses radius_ratio 0.05 (encoder_datasets.py:763), make_card_with_mask 0.046
  // (od_datasets.py:223,234); radius = ceil(max(h,w)*ratio) (util/image.py:413-414)
  std::vector<float> m((size_t)h * w);
  const int hw_max = h > w ? h : w;
  const double ratios[2] = {0.05, 0.046};
  float** dst[2] = {&ctx->mask_enc, &ctx->mask_det};
  for (int k = 0; k < 2; k++) {
    int radius = (int)ceil((double)hw_max * ratios[k]);
    if (2 * radius > h || 2 * radius > w) return fail(ctx, MTGV_ERR_LIMIT, "card too small for the rounded-corner mask");
    host_round_rect_mask(h, w, radius, m.data());
    MTGV_CUDA_OK(ctx, cudaMalloc(dst[k], (size_t)h * w * 4));
    MTGV_CUDA_OK(ctx, cudaMemcpy(*dst[k], m.data(), (size_t)h * w * 4, cudaMemcpyHostToDevice));
  }
  ctx->n_cards = n;
  ctx->card_epoch++;
  rc = enc_build_static_alpha(ctx, 0);
  if (rc) return rc;
  MTGV_CUDA_OK(ctx, cudaDeviceSynchronize());
  return MTGV_OK;
}

int mtgv_set_bg_pool(mtgv_ctx* ctx, const uint8_t* bgs, const int64_t* offsets_host, const int32_t* hw_host, int n) {
  if (!ctx) return MTGV_ERR_INVALID;
  if (!bgs || !offsets_host || !hw_host || n <= 0) return fail(ctx, MTGV_ERR_INVALID, "mtgv_set_bg_pool: bad arguments");
  MTGV_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  free_bgs(ctx);
  std::vector<int64_t> off(n);
  size_t total = 0;
  for (int j = 0; j < n; j++) {
    int h = hw_host[2 * j], w = hw_host[2 * j + 1];
    if (h < 2 || w < 2) return fail(ctx, MTGV_ERR_INVALID, "mtgv_set_bg_pool: image smaller than 2x2");
    off[j] = (int64_t)total;
    total += bg_image_bytes(h, w);
  }
  MTGV_CUDA_OK(ctx, cudaMalloc(&ctx->bg_planes, total));
  MTGV_CUDA_OK(ctx, cudaMalloc(&ctx->bg_off, (size_t)n * 8));
  MTGV_CUDA_OK(ctx, cudaMalloc(&ctx->bg_hw, (size_t)n * 2 * 4));
  ctx->bg_off_host = off;
  ctx->bg_hw_host.assign(hw_host, hw_host + 2 * (size_t)n);
  MTGV_CUDA_OK(ctx, cudaMemcpy(ctx->bg_off, off.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
  MTGV_CUDA_OK(ctx, cudaMemcpy(ctx->bg_hw, hw_host, (size_t)n * 2 * 4, cudaMemcpyHostToDevice));
  // runs of equally-sized, contiguous images are ingested with one launch each
  int j = 0;
  while (j < n) {
    int h = hw_host[2 * j], w = hw_host[2 * j + 1], k = j + 1;
    while (k < n && hw_host[2 * k] == h && hw_host[2 * k + 1] == w &&
           offsets_host[k] == offsets_host[k - 1] + (int64_t)h * w * 3)
      k++;
    int rc = pool_interleave(ctx, bgs + offsets_host[j], ctx->bg_planes + off[j], k - j, h, w, 0);
    if (rc) return rc;
    j = k;
  }
  ctx->n_bgs = n;
  MTGV_CUDA_OK(ctx, cudaDeviceSynchronize());
  return MTGV_OK;
}

int mtgv_set_encoder_config(mtgv_ctx* ctx, const mtgv_enc_config* cfg) {
  if (!ctx || !cfg) return MTGV_ERR_INVALID;
  if (cfg->out_h < 8 || cfg->out_w < 8 || cfg->y_h < 8 || cfg->y_w < 8)
    return fail(ctx, MTGV_ERR_INVALID, "mtgv_set_encoder_config: sizes must be >= 8");
  MTGV_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  ctx->cfg = *cfg;
  ctx->cfg_set = true;
  MTGV_CUDA_OK(ctx, cudaMemcpy(ctx->cfg_dev, cfg, sizeof(*cfg), cudaMemcpyHostToDevice));
  int rc = enc_build_static_alpha(ctx, 0);
  if (rc) return rc;
  MTGV_CUDA_OK(ctx, cudaDeviceSynchronize());
  return MTGV_OK;
}

static int need_encoder(mtgv_ctx* ctx, bool need_bg) {
  if (!ctx) return MTGV_ERR_INVALID;
  if (!ctx->cfg_set) return fail(ctx, MTGV_ERR_STATE, "encoder config not set (mtgv_set_encoder_config)");
  if (!ctx->n_cards) return fail(ctx, MTGV_ERR_STATE, "card pool not set (mtgv_set_card_pool)");
  if (need_bg && !ctx->n_bgs) return fail(ctx, MTGV_ERR_STATE, "background pool not set (mtgv_set_bg_pool)");
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e != cudaSuccess) return fail(ctx, MTGV_ERR_CUDA, cudaGetErrorString(e));
  return MTGV_OK;
}

int mtgv_sample_encoder_tape(mtgv_ctx* ctx, uint64_t seed, int64_t first_index, int n_pairs, mtgv_enc_tape* tape, void* stream) {
  int rc = need_encoder(ctx, true);
  if (rc) return rc;
  if (n_pairs == 0) return MTGV_OK;  // empty batch: nothing to write (the caller's buffer may be a null pointer)
  if (!tape || n_pairs < 0) return fail(ctx, MTGV_ERR_INVALID, "mtgv_sample_encoder_tape: bad arguments");
  return enc_sample_tape(ctx, seed, first_index, n_pairs, nullptr, nullptr, -1.0, -1.0, tape, (cudaStream_t)stream);
}

int mtgv_update_card_images(mtgv_ctx* ctx, const uint8_t* cards, int first, int n, void* stream) {
  if (!ctx) return MTGV_ERR_INVALID;
  if (!ctx->n_cards) return fail(ctx, MTGV_ERR_STATE, "card pool not set (mtgv_set_card_pool)");
  if (!cards || first < 0 || n < 0 || first + n > ctx->n_cards)
    return fail(ctx, MTGV_ERR_INVALID, "mtgv_update_card_images: range outside the pool");
  MTGV_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  ctx->card_epoch++;
  return pool_planarize(ctx, cards, ctx->card_planes + (size_t)first * 3 * ctx->card_h * ctx->card_pitch, n, ctx->card_h,
                        ctx->card_w, ctx->card_pitch, (cudaStream_t)stream);
}

int mtgv_update_bg_images(mtgv_ctx* ctx, const uint8_t* bgs, int first, int n, void* stream) {
  if (!ctx) return MTGV_ERR_INVALID;
  if (!ctx->n_bgs) return fail(ctx, MTGV_ERR_STATE, "background pool not set (mtgv_set_bg_pool)");
  if (!bgs || first < 0 || n < 0 || first + n > ctx->n_bgs)
    return fail(ctx, MTGV_ERR_INVALID, "mtgv_update_bg_images: range outside the pool");
  if (n == 0) return MTGV_OK;
  const int h = ctx->bg_hw_host[2 * first], w = ctx->bg_hw_host[2 * first + 1];
  for (int j = first; j < first + n; j++)
    if (ctx->bg_hw_host[2 * j] != h || ctx->bg_hw_host[2 * j + 1] != w)
      return fail(ctx, MTGV_ERR_INVALID, "mtgv_update_bg_images: entries in the range differ in size");
  MTGV_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  return pool_interleave(ctx, bgs, ctx->bg_planes + ctx->bg_off_host[first], n, h, w, (cudaStream_t)stream);
}

int mtgv_sample_encoder_tape_ex(mtgv_ctx* ctx, uint64_t seed, int64_t first_index, int n_pairs, const int32_t* cards,
                                const int32_t* bgs, double target_is_input_prob, double similar_neg_prob,
                                mtgv_enc_tape* tape, void* stream) {
  int rc = need_encoder(ctx, true);
  if (rc) return rc;
  if (n_pairs == 0) return MTGV_OK;
  if (!tape || n_pairs < 0) return fail(ctx, MTGV_ERR_INVALID, "mtgv_sample_encoder_tape_ex: bad arguments");
  return enc_sample_tape(ctx, seed, first_index, n_pairs, cards, bgs, target_is_input_prob, similar_neg_prob, tape,
                         (cudaStream_t)stream);
}

int mtgv_get_mask(mtgv_ctx* ctx, int which, float* out, void* stream) {
  if (!ctx) return MTGV_ERR_INVALID;
  if (!ctx->n_cards) return fail(ctx, MTGV_ERR_STATE, "card pool not set (mtgv_set_card_pool)");
  if (!out || which < 0 || which > 1) return fail(ctx, MTGV_ERR_INVALID, "mtgv_get_mask: bad arguments");
  MTGV_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  MTGV_CUDA_OK(ctx, cudaMemcpyAsync(out, which ? ctx->mask_det : ctx->mask_enc, (size_t)ctx->card_h * ctx->card_w * 4,
                                    cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return MTGV_OK;
}

int mtgv_expand_params(mtgv_ctx* ctx, const mtgv_enc_tape* tape, int n, mtgv_enc_params* params, int64_t* labels, void* stream) {
  int rc = need_encoder(ctx, false);
  if (rc) return rc;
  if (n == 0) return MTGV_OK;
  if (!tape || !params || n < 0) return fail(ctx, MTGV_ERR_INVALID, "mtgv_expand_params: bad arguments");
  return enc_expand(ctx, tape, n, params, labels, (cudaStream_t)stream);
}

int mtgv_encoder_batch(mtgv_ctx* ctx, const mtgv_enc_params* params, int n, void* out, int out_dtype, const void* fields,
                       void* stream) {
  int rc = need_encoder(ctx, false);
  if (rc) return rc;
  if (n == 0) return MTGV_OK;
  if (!params || !out || n < 0 || out_dtype < 0 || out_dtype > 2)
    return fail(ctx, MTGV_ERR_INVALID, "mtgv_encoder_batch: bad arguments");
  return enc_batch(ctx, params, n, out, out_dtype, fields, (cudaStream_t)stream);
}

int mtgv_encoder_targets(mtgv_ctx* ctx, const int32_t* cards, int n, void* out, int out_dtype, void* stream) {
  int rc = need_encoder(ctx, false);
  if (rc) return rc;
  if (n == 0) return MTGV_OK;
  if (!cards || !out || n < 0 || out_dtype < 0 || out_dtype > 2)
    return fail(ctx, MTGV_ERR_INVALID, "mtgv_encoder_targets: bad arguments");
  return enc_targets(ctx, cards, n, out, out_dtype, (cudaStream_t)stream);
}

int mtgv_warp_perspective(mtgv_ctx* ctx, const float* src, int n, int sh, int sw, int c, const double* M, float* dst, int dh,
                          int dw, void* stream) {
  if (!ctx) return MTGV_ERR_INVALID;
  if (!src || !M || !dst || n < 0 || sh < 1 || sw < 1 || dh < 1 || dw < 1 || c < 1 || c > 4)
    return fail(ctx, MTGV_ERR_INVALID, "mtgv_warp_perspective: bad arguments");
  MTGV_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  return enc_warp_perspective(ctx, src, n, sh, sw, c, M, dst, dh, dw, (cudaStream_t)stream);
}

int mtgv_extract_dewarped(mtgv_ctx* ctx, const uint8_t* frame, int frame_h, int frame_w, int channels, const float* quads, int n,
                          const float* dst_rect, uint8_t* out, int out_h, int out_w, void* stream) {
  if (!ctx) return MTGV_ERR_INVALID;
  if (n == 0) return MTGV_OK;
  if (!frame || !quads || !dst_rect || !out || n < 0 || frame_h < 1 || frame_w < 1 || out_h < 1 || out_w < 1 || channels < 1 ||
      channels > 4)
    return fail(ctx, MTGV_ERR_INVALID, "mtgv_extract_dewarped: bad arguments");
  MTGV_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  return dewarp_u8(ctx, frame, frame_h, frame_w, channels, quads, n, dst_rect, out, out_h, out_w, (cudaStream_t)stream);
}

int mtgv_jpeg_info(mtgv_ctx* ctx, const uint8_t* file, int64_t len, int32_t* hw) {
  if (!ctx) return MTGV_ERR_INVALID;
  if (!file || !hw || len < 0) return fail(ctx, MTGV_ERR_INVALID, "mtgv_jpeg_info: bad arguments");
  return jpeg_info(ctx, file, len, hw);
}

int mtgv_jpeg_info_batch(mtgv_ctx* ctx, const uint8_t* files, const int64_t* file_off, int n, int32_t* hw) {
  if (!ctx) return MTGV_ERR_INVALID;
  if (n == 0) return MTGV_OK;
  if (!files || !file_off || !hw || n < 0) return fail(ctx, MTGV_ERR_INVALID, "mtgv_jpeg_info_batch: bad arguments");
  return jpeg_info_batch(ctx, files, file_off, n, hw);
}

int mtgv_gather_files(mtgv_ctx* ctx, const uint8_t* const* srcs, const int64_t* lens, int n, uint8_t* dst, int64_t dst_cap, int64_t* file_off) {
  // host-only: works without a context (ctx only receives the error message)
  auto bad = [&](const std::string& m) { return ctx ? fail(ctx, MTGV_ERR_INVALID, "mtgv_gather_files: " + m) : (int)MTGV_ERR_INVALID; };
  if (!file_off || n < 0 || (n > 0 && (!srcs || !lens || !dst))) return bad("bad arguments");
  file_off[0] = 0;
  for (int i = 0; i < n; i++) {
    if (lens[i] < 0 || !srcs[i]) return bad("bad buffer " + std::to_string(i));
    file_off[i + 1] = file_off[i] + lens[i];
  }
  if (file_off[n] > dst_cap) return bad("destination too small");
  // equal byte shares per thread (files differ a lot in size), split inside files where needed
  const int64_t total = file_off[n];
  unsigned hc = std::thread::hardware_concurrency();
  const int nt = total < (4 << 20) ? 1 : (int)(hc < 2 ? 1 : (hc > 8 ? 8 : hc));
  auto run = [&](int t) {
    const int64_t b0 = total * t / nt, b1 = total * (t + 1) / nt;
    int i = (int)(std::upper_bound(file_off, file_off + n + 1, b0) - file_off) - 1;  // the file byte b0 lies in
    int64_t b = b0;
    while (b < b1 && i < n) {
      const int64_t e = file_off[i + 1] < b1 ? file_off[i + 1] : b1;
      if (e > b) {
        memcpy(dst + b, srcs[i] + (b - file_off[i]), (size_t)(e - b));
        b = e;
      }
      if (b >= file_off[i + 1]) i++;
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; t++) th.emplace_back(run, t);
  run(0);
  for (auto& x : th) x.join();
  return MTGV_OK;
}

int mtgv_jpeg_last_kernel_ms(mtgv_ctx* ctx, float* ms3) {
  if (!ctx || !ms3) return MTGV_ERR_INVALID;
  MTGV_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  return jpeg_last_kernel_ms(ctx, ms3);
}

int mtgv_decode_jpeg_batch(mtgv_ctx* ctx, const uint8_t* files, const int64_t* file_off, int n, uint8_t* out, const int64_t* out_off,
                           const int32_t* hw, void* stream) {
  if (!ctx) return MTGV_ERR_INVALID;
  if (n == 0) return MTGV_OK;
  if (!files || !file_off || !out || !out_off || !hw || n < 0) return fail(ctx, MTGV_ERR_INVALID, "mtgv_decode_jpeg_batch: bad arguments");
  MTGV_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  return jpeg_decode_batch(ctx, files, file_off, n, out, out_off, hw, (cudaStream_t)stream);
}

int mtgv_decode_jpeg_to_pools(mtgv_ctx* ctx, const uint8_t* files, const int64_t* file_off, int n_cards, int first_card, int n_bgs,
                              int first_bg, void* stream) {
  if (!ctx) return MTGV_ERR_INVALID;
  const int n = n_cards + n_bgs;
  if (n_cards < 0 || n_bgs < 0) return fail(ctx, MTGV_ERR_INVALID, "mtgv_decode_jpeg_to_pools: bad arguments");
  if (n == 0) return MTGV_OK;
  if (!files || !file_off) return fail(ctx, MTGV_ERR_INVALID, "mtgv_decode_jpeg_to_pools: bad arguments");
  if (n > 65535) return fail(ctx, MTGV_ERR_LIMIT, "mtgv_decode_jpeg_to_pools: more than 65535 files per call (one grid row per file); split the batch");
  if (n_cards && (!ctx->n_cards || first_card < 0 || first_card + n_cards > ctx->n_cards))
    return fail(ctx, MTGV_ERR_INVALID, "mtgv_decode_jpeg_to_pools: card range outside the pool (mtgv_set_card_pool)");
  if (n_bgs && (!ctx->n_bgs || first_bg < 0 || first_bg + n_bgs > ctx->n_bgs))
    return fail(ctx, MTGV_ERR_INVALID, "mtgv_decode_jpeg_to_pools: background range outside the pool (mtgv_set_bg_pool)");
  MTGV_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  // destinations as absolute device addresses (two pools, two allocations): out = null + offset
  std::vector<JpegDst> dst((size_t)n);
  for (int i = 0; i < n_cards; i++) {
    const uint8_t* p = ctx->card_planes + (size_t)(first_card + i) * 3 * ctx->card_h * ctx->card_pitch;
    dst[i] = JpegDst{(int64_t)(uintptr_t)p, ctx->card_h, ctx->card_w, 1, ctx->card_pitch};
  }
  for (int j = 0; j < n_bgs; j++) {
    const int slot = first_bg + j, h = ctx->bg_hw_host[2 * slot], w = ctx->bg_hw_host[2 * slot + 1];
    const uint8_t* p = ctx->bg_planes + ctx->bg_off_host[slot];
    dst[n_cards + j] = JpegDst{(int64_t)(uintptr_t)p, h, w, 2, (w + 3) & ~3};
  }
  if (n_cards) ctx->card_epoch++;
  return jpeg_decode_batch_ex(ctx, files, file_off, n, (uint8_t*)nullptr, dst.data(), (cudaStream_t)stream);
}

int mtgv_encode_jpeg_batch(mtgv_ctx* ctx, const uint8_t* images, int n, int h, int w, int layout, int quality, uint8_t* out,
                           int64_t cap, int32_t* out_len, void* stream) {
  if (!ctx) return MTGV_ERR_INVALID;
  if (n == 0) return MTGV_OK;
  if (!images || !out || !out_len || n < 0 || h < 1 || w < 1 || (layout != MTGV_LAYOUT_NHWC && layout != MTGV_LAYOUT_NCHW) ||
      quality < 1 || quality > 100 || cap < 1024 || (cap & 3))
    return fail(ctx, MTGV_ERR_INVALID, "mtgv_encode_jpeg_batch: bad arguments");
  if (h > 16384 || w > 16384) return fail(ctx, MTGV_ERR_LIMIT, "mtgv_encode_jpeg_batch: image larger than 16384 pixels per side");
  if (n > 65535) return fail(ctx, MTGV_ERR_LIMIT, "mtgv_encode_jpeg_batch: more than 65535 images per call (one grid row per image); split the batch");
  MTGV_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  return jpegenc_batch(ctx, images, n, h, w, layout, quality, out, cap, out_len, (cudaStream_t)stream);
}

int mtgv_compact_jpeg_files(mtgv_ctx* ctx, const uint8_t* slots, int64_t cap, const int32_t* out_len, int n, uint8_t* compact,
                            int64_t* offsets, void* stream) {
  if (!ctx) return MTGV_ERR_INVALID;
  if (!offsets || n < 0 || (n > 0 && (!slots || !out_len || !compact || cap < 4 || (cap & 3))))
    return fail(ctx, MTGV_ERR_INVALID, "mtgv_compact_jpeg_files: bad arguments");
  if (n > 65535) return fail(ctx, MTGV_ERR_LIMIT, "mtgv_compact_jpeg_files: more than 65535 files per call (one grid row per file); split the batch");
  MTGV_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  return jpegenc_compact(ctx, slots, cap, out_len, n, compact, offsets, (cudaStream_t)stream);
}

int mtgv_jpeg_encode_last_kernel_ms(mtgv_ctx* ctx, float* ms2) {
  if (!ctx || !ms2) return MTGV_ERR_INVALID;
  MTGV_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  return jpegenc_last_kernel_ms(ctx, ms2);
}

int mtgv_run_plane_ops(mtgv_ctx* ctx, float* img, int n, int h, int w, int c, const mtgv_x_op* ops, int n_ops, const void* fields,
                       uint64_t seed, void* stream) {
  if (!ctx) return MTGV_ERR_INVALID;
  if (!img || !ops || n < 0 || n_ops < 0 || h < 2 || w < 2)
    return fail(ctx, MTGV_ERR_INVALID, "mtgv_run_plane_ops: bad arguments");
  MTGV_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  return enc_run_plane_ops(ctx, img, n, h, w, c, ops, n_ops, fields, seed, (cudaStream_t)stream);
}

}  // extern "C"
