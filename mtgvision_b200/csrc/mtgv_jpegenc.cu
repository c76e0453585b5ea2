// mtgv_jpegenc.cu - batched baseline JPEG encode of generated scenes.
// Replaces the cv2.imwrite behind save_sample / imwrite (mtgvision/od_datasets.py:794-832, util/image.py:95-104)
// for uint8 images whose sides are multiples of 16; arithmetic in mtgv_jpegenc.cuh, FILE BYTES identical to cv2's
// (libjpeg-turbo: quality-scaled Annex K tables, 4:2:0, ISLOW forward DCT, standard Huffman tables, JFIF header).
// Two kernels per batch, every image of the batch in each launch:
//   k_jpegenc_dct   a CTA takes four MCUs: RGB -> YCbCr with the 2x2 chroma box filter (one pixel quad per thread),
//                   then 8 threads per 8x8 block run the row and column passes of the forward DCT through shared
//                   memory, quantise and store the block in zigzag order (coalesced 3 KiB per CTA).
//   k_jpegenc_huff  one CTA per image.  Huffman coding is a serial bit stream, made parallel in four steps:
//                   (A) every thread sizes one block's code (DC difference against its predecessor block, read
//                   straight from the coefficient array), a CTA scan turns sizes into bit offsets; (B) every thread
//                   writes its block's bits at its offset into a zeroed word buffer (atomicOr on the two words it
//                   shares with its neighbours, plain stores between); (C) the last byte is padded with one bits;
//                   (D) 0xFF bytes are counted per word, scanned, and the stream is copied behind the header with the
//                   stuffed zero bytes inserted; EOI and the file length close the image.
#include "mtgv_internal.cuh"
#include "mtgv_jpegenc.cuh"

namespace mtgv {

struct JpegEncState {
  int16_t* coef = nullptr;   size_t coef_cap = 0;
  uint32_t* bits = nullptr;  size_t bits_cap = 0;   // unstuffed streams, `cap` bytes per image
  uint32_t* boff = nullptr;  size_t boff_cap = 0;   // bit offset of every block
  JpegEncTables* tables = nullptr;
  uint8_t* header = nullptr;
  int header_len = 0, quality = -1, h = 0, w = 0;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  bool timed = false;
};

static int enc_grow(mtgv_ctx* ctx, void** p, size_t* cap, size_t need) {
  if (need <= *cap) return MTGV_OK;
  if (*p) MTGV_CUDA_OK(ctx, cudaFree(*p));
  *p = nullptr; *cap = 0;
  MTGV_CUDA_OK(ctx, cudaMalloc(p, need));
  *cap = need;
  return MTGV_OK;
}

__constant__ uint8_t c_zz_of_natural[64];  // position in zigzag order of natural index k

struct ImgLayout {
  int64_t img_stride;
  int row_stride, px_stride, ch_stride;
};

__global__ void __launch_bounds__(256) k_jpegenc_dct(const uint8_t* __restrict__ images, ImgLayout L, int mcux, int nmcu,
                                                     const JpegEncTables* __restrict__ T, int16_t* __restrict__ coef) {
  __shared__ int16_t samp[24][64];   // level-shifted samples, natural order; blocks = 4 MCUs x (Y00 Y01 Y10 Y11 Cb Cr)
  __shared__ int ws[24][8][9];
  __shared__ __align__(16) int16_t outb[24][64];  // quantised, zigzag order
  const int tid = threadIdx.x;
  const int img = blockIdx.y;
  const int m0 = blockIdx.x * 4;
  {  // colour conversion + chroma box filter: one 2x2 pixel quad per thread
    const int lm = tid >> 6, q = tid & 63, qy = q >> 3, qx = q & 7, m = m0 + lm;
    if (m < nmcu) {
      const int my = m / mcux, mx = m - my * mcux;
      const uint8_t* base = images + (int64_t)img * L.img_stride;
      int sb = 0, sr = 0;
#pragma unroll
      for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
          const int y = my * 16 + 2 * qy + dy, x = mx * 16 + 2 * qx + dx;
          const uint8_t* p = base + (int64_t)y * L.row_stride + (int64_t)x * L.px_stride;
          int Y, cb, cr;
          jpegenc_ycc(__ldg(p), __ldg(p + L.ch_stride), __ldg(p + 2 * (int64_t)L.ch_stride), &Y, &cb, &cr);
          const int yy = 2 * qy + dy, xx = 2 * qx + dx;
          samp[lm * 6 + (yy >> 3) * 2 + (xx >> 3)][(yy & 7) * 8 + (xx & 7)] = (int16_t)(Y - 128);
          sb += cb; sr += cr;
        }
      const int bias = ((mx * 8 + qx) & 1) ? 2 : 1;
      samp[lm * 6 + 4][qy * 8 + qx] = (int16_t)(((sb + bias) >> 2) - 128);
      samp[lm * 6 + 5][qy * 8 + qx] = (int16_t)(((sr + bias) >> 2) - 128);
    }
  }
  __syncthreads();
  const int b = tid >> 3, t = tid & 7;
  if (b < 24) {
    int d[8], o[8];
#pragma unroll
    for (int k = 0; k < 8; k++) d[k] = samp[b][t * 8 + k];
    jpegenc_fdct8(d, o, true);  // row t
#pragma unroll
    for (int k = 0; k < 8; k++) ws[b][t][k] = o[k];
  }
  __syncthreads();
  if (b < 24) {
    int d[8], o[8];
#pragma unroll
    for (int k = 0; k < 8; k++) d[k] = ws[b][k][t];
    jpegenc_fdct8(d, o, false);  // column t
    const uint16_t* q = T->q[(b % 6) < 4 ? 0 : 1];
#pragma unroll
    for (int k = 0; k < 8; k++) outb[b][c_zz_of_natural[k * 8 + t]] = (int16_t)jpegenc_quant(o[k], q[k * 8 + t]);
  }
  __syncthreads();
  const int live = (nmcu - m0 < 4 ? nmcu - m0 : 4) * 6 * 64 / 8;  // uint4 words of the live MCUs
  uint4* dst = (uint4*)(coef + ((int64_t)img * nmcu + m0) * 6 * 64);
  for (int i = tid; i < live; i += 256) dst[i] = ((const uint4*)outb)[i];
}

struct CountPut {
  int n = 0;
  __device__ __forceinline__ void operator()(unsigned, int size) { n += size; }
};

struct WordPut {  // writes bits at an arbitrary bit offset of a zeroed big-endian word stream
  uint32_t* words;
  uint32_t cap_words;
  uint32_t w;       // next word to write
  uint64_t acc = 0;
  int n;            // pending bits in acc (the first n0 are the neighbour's: zeros here)
  bool first = true;
  __device__ __forceinline__ WordPut(uint32_t* base, uint32_t cap, uint32_t bit_off) : words(base), cap_words(cap), w(bit_off >> 5), n((int)(bit_off & 31u)) {}
  __device__ __forceinline__ void emit(uint32_t v, bool shared_word) {
    if (w < cap_words) {
      const uint32_t be = __byte_perm(v, 0, 0x0123);  // stream order = byte order in memory
      if (shared_word) atomicOr(words + w, be);
      else words[w] = be;
    }
    w++;
  }
  __device__ __forceinline__ void operator()(unsigned code, int size) {
    acc = (acc << size) | code;
    n += size;
    if (n >= 32) {
      emit((uint32_t)(acc >> (n - 32)), first);
      first = false;
      n -= 32;
    }
  }
  __device__ __forceinline__ void flush() {
    if (n > 0) emit((uint32_t)(acc << (32 - n)), true);
  }
};

__device__ __forceinline__ int block_scan_excl(int v, int* warp_sums, int* total) {  // 1024 threads
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  __syncthreads();  // warp_sums reuse
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  int woff = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 32; w++) {
    const int s = warp_sums[w];
    if (w < warp) woff += s;
    tot += s;
  }
  *total = tot;
  return woff + incl - v;
}

__global__ void __launch_bounds__(1024) k_jpegenc_huff(const int16_t* __restrict__ coef, int nmcu, const JpegEncTables* __restrict__ T,
                                                       const uint8_t* __restrict__ header, int header_len, uint32_t* __restrict__ bits,
                                                       uint32_t* __restrict__ boff, uint8_t* __restrict__ out, int64_t cap,
                                                       int32_t* __restrict__ out_len) {
  __shared__ uint32_t dc[2][16], ac[2][256];
  __shared__ int warp_sums[32];
  const int tid = threadIdx.x, img = blockIdx.x, nblk = nmcu * 6;
  for (int i = tid; i < 32; i += 1024) ((uint32_t*)dc)[i] = ((const uint32_t*)T->dc)[i];
  for (int i = tid; i < 512; i += 1024) ((uint32_t*)ac)[i] = ((const uint32_t*)T->ac)[i];
  const int16_t* C = coef + (int64_t)img * nblk * 64;
  uint32_t* W = bits + (int64_t)img * (cap >> 2);
  uint32_t* O = boff + (int64_t)img * nblk;
  uint8_t* dst = out + (int64_t)img * cap;
  const uint32_t cap_words = (uint32_t)(cap >> 2);
  __syncthreads();
  // (A) code size of every block -> bit offsets
  unsigned run = 0;
  for (int base = 0; base < nblk; base += 1024) {
    const int b = base + tid;
    int len = 0;
    if (b < nblk) {
      const int m = b / 6, j = b - m * 6, pb = jpegenc_pred_block(m, j), t = j < 4 ? 0 : 1;
      CountPut cp;
      jpegenc_block(C + (int64_t)b * 64, pb < 0 ? 0 : (int)C[(int64_t)pb * 64], dc[t], ac[t], cp);
      len = cp.n;
    }
    int tot;
    const int ex = block_scan_excl(len, warp_sums, &tot);
    if (b < nblk) O[b] = run + (unsigned)ex;
    run += (unsigned)tot;
  }
  const unsigned total_bits = run;
  const unsigned nbytes = (total_bits + 7u) >> 3;
  const bool fits = (int64_t)nbytes + 4 <= cap;  // the word buffer holds the whole stream
  __syncthreads();
  // (B) the bits
  if (fits) {
    for (int b = tid; b < nblk; b += 1024) {
      const int m = b / 6, j = b - m * 6, pb = jpegenc_pred_block(m, j), t = j < 4 ? 0 : 1;
      WordPut wp(W, cap_words, O[b]);
      jpegenc_block(C + (int64_t)b * 64, pb < 0 ? 0 : (int)C[(int64_t)pb * 64], dc[t], ac[t], wp);
      wp.flush();
    }
    // (C) one bits up to the byte boundary (jchuff.c flush_bits)
    if (tid == 0 && (total_bits & 7u)) {
      const unsigned pad = 8u - (total_bits & 7u), sh = 32u - (total_bits & 31u) - pad;
      atomicOr(W + (total_bits >> 5), __byte_perm(((1u << pad) - 1u) << sh, 0, 0x0123));
    }
  }
  __syncthreads();
  // (D) byte stuffing behind the header
  unsigned pos = (unsigned)header_len;
  if (fits) {
    const unsigned nwords = (nbytes + 3u) >> 2;
    for (unsigned base = 0; base < nwords; base += 1024) {
      const unsigned wi = base + tid;
      uint32_t v = 0;
      int cnt = 0, nb = 0;
      if (wi < nwords) {
        v = W[wi];  // memory order = stream order: byte k of the word is (v >> 8k) & 255
        nb = nbytes - wi * 4u < 4u ? (int)(nbytes - wi * 4u) : 4;
        for (int k = 0; k < nb; k++) cnt += ((v >> (8 * k)) & 255u) == 255u;
      }
      int tot;
      const int ex = block_scan_excl(nb + cnt, warp_sums, &tot);
      unsigned p = pos + (unsigned)ex;
      for (int k = 0; k < nb; k++) {
        const uint8_t byte = (uint8_t)(v >> (8 * k));
        if ((int64_t)p < cap) dst[p] = byte;
        p++;
        if (byte == 255u) {
          if ((int64_t)p < cap) dst[p] = 0;
          p++;
        }
      }
      pos += (unsigned)tot;
    }
  }
  const bool ok = fits && (int64_t)pos + 2 <= cap;
  if (ok) {
    for (int i = tid; i < header_len; i += 1024) dst[i] = header[i];
    if (tid == 0) { dst[pos] = 0xFF; dst[pos + 1] = 0xD9; }
  }
  if (tid == 0) out_len[img] = ok ? (int32_t)(pos + 2) : -1;
}

// file lengths -> byte offsets of the compact layout (one CTA; files that did not fit count as empty)
__global__ void __launch_bounds__(1024) k_jpegenc_offsets(const int32_t* __restrict__ out_len, int n, int64_t* __restrict__ offsets) {
  __shared__ int warp_sums[32];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n && out_len[i] > 0 ? out_len[i] : 0;
    int tot;
    const int ex = block_scan_excl(v, warp_sums, &tot);
    const long long c = carry;
    if (i < n) offsets[i] = c + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n] = carry;
}

// grid (chunks, n): file i's bytes from its cap-sized slot to offsets[i] of the compact buffer
__global__ void __launch_bounds__(256) k_jpegenc_compact(const uint8_t* __restrict__ slots, int64_t cap, const int32_t* __restrict__ out_len,
                                                         const int64_t* __restrict__ offsets, uint8_t* __restrict__ compact) {
  const int i = blockIdx.y, len = out_len[i];
  if (len <= 0) return;
  const uint8_t* src = slots + (int64_t)i * cap;
  uint8_t* dst = compact + offsets[i];
  const int head = (int)((4 - ((uintptr_t)dst & 3)) & 3);  // destination word alignment; the slot itself is word aligned
  for (int k = blockIdx.x * 256 + threadIdx.x; k < head && k < len; k += gridDim.x * 256) dst[k] = src[k];
  const int nw = len > head ? (len - head) >> 2 : 0;
  for (int k = blockIdx.x * 256 + threadIdx.x; k < nw; k += gridDim.x * 256) {
    const uint8_t* p = src + head + 4 * k;
    ((uint32_t*)(dst + head))[k] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
  }
  for (int k = head + 4 * nw + blockIdx.x * 256 + threadIdx.x; k < len; k += gridDim.x * 256) dst[k] = src[k];
}

int jpegenc_compact(mtgv_ctx* ctx, const uint8_t* slots, int64_t cap, const int32_t* out_len, int n, uint8_t* compact, int64_t* offsets,
                    cudaStream_t stream) {
  k_jpegenc_offsets<<<1, 1024, 0, stream>>>(out_len, n, offsets);
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  k_jpegenc_compact<<<dim3(32, n), 256, 0, stream>>>(slots, cap, out_len, offsets, compact);
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 2;
  return MTGV_OK;
}

int jpegenc_destroy(mtgv_ctx* ctx) {
  JpegEncState* st = (JpegEncState*)ctx->jpegenc;
  if (!st) return MTGV_OK;
  cudaFree(st->coef); cudaFree(st->bits); cudaFree(st->boff); cudaFree(st->tables); cudaFree(st->header);
  for (auto& e : st->ev) if (e) cudaEventDestroy(e);
  delete st;
  ctx->jpegenc = nullptr;
  return MTGV_OK;
}

int jpegenc_last_kernel_ms(mtgv_ctx* ctx, float* ms) {
  JpegEncState* st = (JpegEncState*)ctx->jpegenc;
  if (!st || !st->timed) return fail(ctx, MTGV_ERR_INVALID, "mtgv_jpeg_encode_last_kernel_ms: no batch was encoded yet");
  MTGV_CUDA_OK(ctx, cudaEventSynchronize(st->ev[2]));
  for (int k = 0; k < 2; k++) MTGV_CUDA_OK(ctx, cudaEventElapsedTime(&ms[k], st->ev[k], st->ev[k + 1]));
  return MTGV_OK;
}

int jpegenc_batch(mtgv_ctx* ctx, const uint8_t* images, int n, int h, int w, int layout, int quality, uint8_t* out, int64_t cap,
                  int32_t* out_len, cudaStream_t stream) {
  if (!ctx->jpegenc) {
    JpegEncState* s = new JpegEncState();
    ctx->jpegenc = s;
    uint8_t inv[64];
    for (int k = 0; k < 64; k++) inv[kJpegZigzag[k]] = (uint8_t)k;
    MTGV_CUDA_OK(ctx, cudaMemcpyToSymbol(c_zz_of_natural, inv, 64));
    MTGV_CUDA_OK(ctx, cudaMalloc((void**)&s->tables, sizeof(JpegEncTables)));
    MTGV_CUDA_OK(ctx, cudaMalloc((void**)&s->header, 1024));
    for (auto& e : s->ev) MTGV_CUDA_OK(ctx, cudaEventCreate(&e));
  }
  JpegEncState* st = (JpegEncState*)ctx->jpegenc;
  if (st->quality != quality || st->h != h || st->w != w) {
    JpegEncTables T;
    jpegenc_tables(quality, &T);
    const std::vector<uint8_t> hdr = jpegenc_header(h, w, T);
    if (hdr.size() > 1024) return fail(ctx, MTGV_ERR_LIMIT, "mtgv_encode_jpeg_batch: header too long");
    MTGV_CUDA_OK(ctx, cudaStreamSynchronize(stream));
    MTGV_CUDA_OK(ctx, cudaMemcpy(st->tables, &T, sizeof(T), cudaMemcpyHostToDevice));
    MTGV_CUDA_OK(ctx, cudaMemcpy(st->header, hdr.data(), hdr.size(), cudaMemcpyHostToDevice));
    st->header_len = (int)hdr.size();
    st->quality = quality; st->h = h; st->w = w;
  }
  const int mcux = w / 16, nmcu = mcux * (h / 16), nblk = nmcu * 6;
  int rc;
  if ((rc = enc_grow(ctx, (void**)&st->coef, &st->coef_cap, (size_t)n * nblk * 64 * sizeof(int16_t)))) return rc;
  if ((rc = enc_grow(ctx, (void**)&st->bits, &st->bits_cap, (size_t)n * (size_t)cap))) return rc;
  if ((rc = enc_grow(ctx, (void**)&st->boff, &st->boff_cap, (size_t)n * nblk * sizeof(uint32_t)))) return rc;
  ImgLayout L;
  if (layout == MTGV_LAYOUT_NCHW) { L.img_stride = (int64_t)3 * h * w; L.row_stride = w; L.px_stride = 1; L.ch_stride = h * w; }
  else { L.img_stride = (int64_t)3 * h * w; L.row_stride = 3 * w; L.px_stride = 3; L.ch_stride = 1; }
  MTGV_CUDA_OK(ctx, cudaEventRecord(st->ev[0], stream));
  k_jpegenc_dct<<<dim3((nmcu + 3) / 4, n), 256, 0, stream>>>(images, L, mcux, nmcu, st->tables, st->coef);
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  MTGV_CUDA_OK(ctx, cudaEventRecord(st->ev[1], stream));
  MTGV_CUDA_OK(ctx, cudaMemsetAsync(st->bits, 0, (size_t)n * (size_t)cap, stream));
  k_jpegenc_huff<<<n, 1024, 0, stream>>>(st->coef, nmcu, st->tables, st->header, st->header_len, st->bits, st->boff, out, cap, out_len);
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  MTGV_CUDA_OK(ctx, cudaEventRecord(st->ev[2], stream));
  st->timed = true;
  ctx->launches += 2;
  return MTGV_OK;
}

}  // namespace mtgv
