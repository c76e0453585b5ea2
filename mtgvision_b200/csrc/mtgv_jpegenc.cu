// mtgv_jpegenc.cu - batched baseline JPEG encode of generated scenes.
// Replaces the cv2.imwrite behind save_sample / imwrite (mtgvision/od_datasets.py:794-832, util/image.py:95-104)
// for uint8 RGB images of any size; arithmetic in mtgv_jpegenc.cuh, FILE BYTES identical to cv2's
// (libjpeg-turbo: quality-scaled Annex K tables, 4:2:0, ISLOW forward DCT, standard Huffman tables, JFIF header).
// Five kernels per batch, every image of the batch in each launch:
//   k_jpegenc_dct   a CTA takes four MCUs: RGB -> YCbCr with the 2x2 chroma box filter (one pixel quad per thread),
//                   then 8 threads per 8x8 block run the row and column passes of the forward DCT through shared
//                   memory, quantise and store the block in zigzag order (coalesced 3 KiB per CTA).
//   k_jpegenc_size / _scan / _emit / _stuff.  Huffman coding is a serial bit stream, made parallel in four steps:
//                   (A) a warp per block sizes its code - lanes own two coefficients each, the zero run in front of a
//                   nonzero coefficient comes from the block's nonzero mask (ballot), the DC difference from the
//                   predecessor block in the coefficient array - and a CTA scan turns sizes into bit offsets; (B) the
//                   warps assemble each block's bits in shared memory (a warp prefix sum places every lane's symbols)
//                   and copy them to a zeroed word buffer (atomicOr on the two words shared with the neighbouring
//                   blocks, plain stores between); (C) the last byte is padded with one bits;
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

template <bool RAGGED>  // RAGGED: the image is not a whole number of 16x16 MCUs (edge replication, dummy blocks)
__global__ void __launch_bounds__(256) k_jpegenc_dct(const uint8_t* __restrict__ images, ImgLayout L, int H, int W, int mcux, int nmcu,
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
      const bool edge = RAGGED && ((my + 1) * 16 > H || (mx + 1) * 16 > W);  // an MCU that sticks out of the image
      int sb = 0, sr = 0;
#pragma unroll
      for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
          // edge rules (mtgv_jpegenc.cuh): inside the image both rows are the pixel's own
          int y = my * 16 + 2 * qy + dy, x = mx * 16 + 2 * qx + dx, yl = y, yc = y;
          if (edge) { x = jpegenc_col(x, W); yl = jpegenc_luma_row(y, H); yc = jpegenc_chroma_row(y, H); }
          const uint8_t* p = base + (int64_t)yl * L.row_stride + (int64_t)x * L.px_stride;
          int Y, cb, cr;
          jpegenc_ycc(__ldg(p), __ldg(p + L.ch_stride), __ldg(p + 2 * (int64_t)L.ch_stride), &Y, &cb, &cr);
          if (RAGGED && yc != yl) {
            const uint8_t* pc = base + (int64_t)yc * L.row_stride + (int64_t)x * L.px_stride;
            int Yc;
            jpegenc_ycc(__ldg(pc), __ldg(pc + L.ch_stride), __ldg(pc + 2 * (int64_t)L.ch_stride), &Yc, &cb, &cr);
          }
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
  if (RAGGED && ((((W + 7) >> 3) | ((H + 7) >> 3)) & 1)) {  // an odd number of luma block columns / rows: the last MCUs hold dummy blocks
    if (tid < 4 && m0 + tid < nmcu) {  // luma blocks wholly outside the image: zero AC, DC of the block before (jccoefct.c)
      const int m = m0 + tid, my = m / mcux, mx = m - my * mcux;
      jpegenc_dummy_blocks(&outb[tid * 6][0], mx == mcux - 1 && (((W + 7) >> 3) & 1), my == nmcu / mcux - 1 && (((H + 7) >> 3) & 1));
    }
    __syncthreads();
  }
  const int live = (nmcu - m0 < 4 ? nmcu - m0 : 4) * 6 * 64 / 8;  // uint4 words of the live MCUs
  uint4* dst = (uint4*)(coef + ((int64_t)img * nmcu + m0) * 6 * 64);
  for (int i = tid; i < live; i += 256) dst[i] = ((const uint4*)outb)[i];
}

// encode_one_block (jchuff.c) by one warp: lane l owns zigzag coefficients l and l + 32.  The run in front of a nonzero
// coefficient is the distance to the previous set bit of the block's nonzero mask, so every lane knows its own symbols
// (ZRLs, run/size code, value bits) and their length without walking the block; warp prefix sums give the bit positions.
struct WarpBlock {
  uint32_t zrl, code0, code1;   // ZRL code; this lane's run/size (or DC) codes, length << 16 | code; 0 = nothing to emit
  int nz0, nz1;                 // ZRLs in front of the two symbols
  uint32_t val0, val1;          // value bits
  int nb0, nb1;                 // number of value bits
  int len0, len1;               // bits this lane emits for coefficient l / l + 32
  int eob;                      // lane 0: length << 16 | code of the EOB symbol, 0 if the block ends on coefficient 63
};

__device__ __forceinline__ WarpBlock warp_block(const int16_t* __restrict__ zz, int last_dc, const uint32_t* dc_tab, const uint32_t* ac_tab,
                                                int lane) {
  WarpBlock B;
  int v0 = zz[lane], v1 = zz[lane + 32];
  if (lane == 0) v0 -= last_dc;
  // bit l of m_lo / m_hi: AC coefficient l / l + 32 (zigzag order) is nonzero
  const unsigned m_lo = __ballot_sync(0xffffffffu, v0 != 0 && lane != 0), m_hi = __ballot_sync(0xffffffffu, v1 != 0);
  B.zrl = ac_tab[0xF0];
  const int zl = (int)(B.zrl >> 16);
  const unsigned lt = (1u << lane) - 1u;
  const int last_lo = m_lo ? 31 - __clz(m_lo) : 0;
  auto one = [&](bool dc, int v, int prev, int k, uint32_t& code, int& nz, uint32_t& val, int& nb, int& len) {
    code = 0; nz = 0; val = 0; nb = 0; len = 0;
    if (dc) {  // DC difference
      nb = 32 - __clz(v < 0 ? -v : v);
      code = dc_tab[nb];
    } else if (v != 0) {
      const int run = k - prev - 1;  // prev: the previous nonzero coefficient (0 = the DC term)
      nz = run >> 4;
      nb = 32 - __clz(v < 0 ? -v : v);
      code = ac_tab[((run & 15) << 4) + nb];
    } else {
      return;
    }
    val = (unsigned)(v < 0 ? v - 1 : v) & ((1u << nb) - 1u);
    len = nz * zl + (int)(code >> 16) + nb;
  };
  const unsigned b0 = m_lo & lt, b1 = m_hi & lt;
  one(lane == 0, v0, b0 ? 31 - __clz(b0) : 0, lane, B.code0, B.nz0, B.val0, B.nb0, B.len0);
  one(false, v1, b1 ? 63 - __clz(b1) : last_lo, lane + 32, B.code1, B.nz1, B.val1, B.nb1, B.len1);
  const int last = m_hi ? 63 - __clz(m_hi) : last_lo;
  B.eob = last < 63 ? (int)ac_tab[0] : 0;
  return B;
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

__device__ __forceinline__ int warp_incl(int v, int lane) {
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  return incl;
}

// ORs the low nbits (1..32) of val into an MSB-first bit stream held as 32-bit words (word 0 bit 31 = first bit)
__device__ __forceinline__ void smem_put(uint32_t* buf, unsigned bitpos, uint32_t val, int nbits) {
  const unsigned w = bitpos >> 5, sh = bitpos & 31u;
  const uint64_t v64 = (uint64_t)val << (64 - nbits - (int)sh);
  atomicOr(&buf[w], (uint32_t)(v64 >> 32));
  if ((uint32_t)v64) atomicOr(&buf[w + 1], (uint32_t)v64);
}

constexpr int kHuffThreads = 512, kHuffWarps = kHuffThreads / 32;
constexpr int kEncBufWords = 56;  // a block's code: at most 20 + 63 * 26 bits, plus the 31 bits in front of it in its first word

__device__ __forceinline__ int block_scan_excl(int v, int* warp_sums, int* total) {  // kHuffThreads threads
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
  for (int w = 0; w < kHuffWarps; w++) {
    const int s = warp_sums[w];
    if (w < warp) woff += s;
    tot += s;
  }
  *total = tot;
  return woff + incl - v;
}

// (A) code size of every block of every image: a warp per block
__global__ void __launch_bounds__(256) k_jpegenc_size(const int16_t* __restrict__ coef, int nblk, const JpegEncTables* __restrict__ T,
                                                      uint32_t* __restrict__ boff) {
  __shared__ uint32_t dc[2][16], ac[2][256];
  for (int i = threadIdx.x; i < 32; i += 256) ((uint32_t*)dc)[i] = ((const uint32_t*)T->dc)[i];
  for (int i = threadIdx.x; i < 512; i += 256) ((uint32_t*)ac)[i] = ((const uint32_t*)T->ac)[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int img = blockIdx.y;
  const int16_t* C = coef + (int64_t)img * nblk * 64;
  for (int b = blockIdx.x * 8 + (threadIdx.x >> 5); b < nblk; b += gridDim.x * 8) {
    const int m = b / 6, j = b - m * 6, pb = jpegenc_pred_block(m, j), t = j < 4 ? 0 : 1;
    const WarpBlock B = warp_block(C + (int64_t)b * 64, pb < 0 ? 0 : (int)C[(int64_t)pb * 64], dc[t], ac[t], lane);
    const int len = warp_sum(B.len0 + B.len1) + (B.eob >> 16);
    if (lane == 0) boff[(int64_t)img * nblk + b] = (unsigned)len;
  }
}

// sizes -> bit offsets, one CTA per image; total_bits[img] = length of the image's code
__global__ void __launch_bounds__(kHuffThreads) k_jpegenc_scan(int nblk, uint32_t* __restrict__ boff, uint32_t* __restrict__ total_bits) {
  __shared__ int warp_sums[kHuffWarps];
  uint32_t* O = boff + (int64_t)blockIdx.x * nblk;
  unsigned run = 0;
  for (int base = 0; base < nblk; base += kHuffThreads) {
    const int b = base + threadIdx.x;
    const int len = b < nblk ? (int)O[b] : 0;
    int tot;
    const int ex = block_scan_excl(len, warp_sums, &tot);
    if (b < nblk) O[b] = run + (unsigned)ex;
    run += (unsigned)tot;
  }
  if (threadIdx.x == 0) total_bits[blockIdx.x] = run;
}

// (B) the bits: every lane ORs its symbols into the warp's shared-memory image of the block, which is then copied to the
// image's zeroed word buffer (the first and last word are shared with the neighbouring blocks: atomicOr; stores between)
__global__ void __launch_bounds__(256) k_jpegenc_emit(const int16_t* __restrict__ coef, int nblk,
                                                      const JpegEncTables* __restrict__ T, const uint32_t* __restrict__ boff,
                                                      const uint32_t* __restrict__ total_bits, uint32_t* __restrict__ bits, int64_t cap) {
  __shared__ uint32_t dc[2][16], ac[2][256];
  __shared__ uint32_t ebuf[8][kEncBufWords];
  for (int i = threadIdx.x; i < 32; i += 256) ((uint32_t*)dc)[i] = ((const uint32_t*)T->dc)[i];
  for (int i = threadIdx.x; i < 512; i += 256) ((uint32_t*)ac)[i] = ((const uint32_t*)T->ac)[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  uint32_t* buf = ebuf[threadIdx.x >> 5];
  const uint32_t cap_words = (uint32_t)(cap >> 2);
  const int img = blockIdx.y;
  if ((int64_t)((total_bits[img] + 7u) >> 3) + 4 > cap) return;  // the image does not fit: k_jpegenc_stuff reports it
  const int16_t* C = coef + (int64_t)img * nblk * 64;
  uint32_t* W = bits + (int64_t)img * (cap >> 2);
  for (int b = blockIdx.x * 8 + (threadIdx.x >> 5); b < nblk; b += gridDim.x * 8) {
    const int m = b / 6, j = b - m * 6, pb = jpegenc_pred_block(m, j), t = j < 4 ? 0 : 1;
    const WarpBlock B = warp_block(C + (int64_t)b * 64, pb < 0 ? 0 : (int)C[(int64_t)pb * 64], dc[t], ac[t], lane);
    const unsigned off = boff[(int64_t)img * nblk + b], s0 = off & 31u;
    const int in0 = warp_incl(B.len0, lane), in1 = warp_incl(B.len1, lane);
    const int tot0 = __shfl_sync(0xffffffffu, in0, 31);
    const int len = tot0 + __shfl_sync(0xffffffffu, in1, 31) + (B.eob >> 16);
    const int nw = (int)((s0 + (unsigned)len + 31u) >> 5);
    for (int i = lane; i < nw; i += 32) buf[i] = 0;
    __syncwarp();
    unsigned p0 = s0 + (unsigned)(in0 - B.len0), p1 = s0 + (unsigned)(tot0 + in1 - B.len1);
    if (B.len0) {  // run/size code and value bits go in as one field (at most 16 + 11 bits)
      for (int z = 0; z < B.nz0; z++, p0 += B.zrl >> 16) smem_put(buf, p0, B.zrl & 0xffffu, (int)(B.zrl >> 16));
      smem_put(buf, p0, ((B.code0 & 0xffffu) << B.nb0) | B.val0, (int)(B.code0 >> 16) + B.nb0);
    }
    if (B.len1) {
      for (int z = 0; z < B.nz1; z++, p1 += B.zrl >> 16) smem_put(buf, p1, B.zrl & 0xffffu, (int)(B.zrl >> 16));
      smem_put(buf, p1, ((B.code1 & 0xffffu) << B.nb1) | B.val1, (int)(B.code1 >> 16) + B.nb1);
    }
    if (lane == 0 && B.eob) smem_put(buf, s0 + (unsigned)len - ((unsigned)B.eob >> 16), (unsigned)B.eob & 0xffffu, B.eob >> 16);
    __syncwarp();
    const uint32_t w0 = off >> 5;
    for (int i = lane; i < nw; i += 32) {
      if (w0 + i < cap_words) {
        const uint32_t be = __byte_perm(buf[i], 0, 0x0123);  // stream order = byte order in memory
        if (i == 0 || i == nw - 1) atomicOr(W + w0 + i, be);
        else W[w0 + i] = be;
      }
    }
    __syncwarp();
  }
}

// (C) one bits up to the byte boundary, (D) byte stuffing behind the header, EOI, file length: one CTA per image
__global__ void __launch_bounds__(kHuffThreads) k_jpegenc_stuff(const uint32_t* __restrict__ total_bits_all, const uint8_t* __restrict__ header,
                                                                int header_len, uint32_t* __restrict__ bits, uint8_t* __restrict__ out, int64_t cap,
                                                                int32_t* __restrict__ out_len) {
  __shared__ int warp_sums[kHuffWarps];
  const int tid = threadIdx.x, img = blockIdx.x;
  uint32_t* W = bits + (int64_t)img * (cap >> 2);
  uint8_t* dst = out + (int64_t)img * cap;
  const unsigned total_bits = total_bits_all[img];
  const unsigned nbytes = (total_bits + 7u) >> 3;
  const bool fits = (int64_t)nbytes + 4 <= cap;  // the word buffer holds the whole stream
  if (fits && tid == 0 && (total_bits & 7u)) {  // jchuff.c flush_bits
    const unsigned pad = 8u - (total_bits & 7u), sh = 32u - (total_bits & 31u) - pad;
    W[total_bits >> 5] |= __byte_perm(((1u << pad) - 1u) << sh, 0, 0x0123);
  }
  __syncthreads();
  unsigned pos = (unsigned)header_len;
  if (fits) {
    const unsigned nwords = (nbytes + 3u) >> 2;
    for (unsigned base = 0; base < nwords; base += kHuffThreads) {
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
    for (int i = tid; i < header_len; i += kHuffThreads) dst[i] = header[i];
    if (tid == 0) { dst[pos] = 0xFF; dst[pos + 1] = 0xD9; }
  }
  if (tid == 0) out_len[img] = ok ? (int32_t)(pos + 2) : -1;
}

// file lengths -> byte offsets of the compact layout (one CTA; files that did not fit count as empty)
__global__ void __launch_bounds__(kHuffThreads) k_jpegenc_offsets(const int32_t* __restrict__ out_len, int n, int64_t* __restrict__ offsets) {
  __shared__ int warp_sums[kHuffWarps];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += kHuffThreads) {
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
  k_jpegenc_offsets<<<1, kHuffThreads, 0, stream>>>(out_len, n, offsets);  // n == 0: writes offsets[0] = 0 and nothing else
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
  if (n == 0) return MTGV_OK;
  k_jpegenc_compact<<<dim3(32, n), 256, 0, stream>>>(slots, cap, out_len, offsets, compact);
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
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
    MTGV_CUDA_OK(ctx, cudaDeviceSynchronize());  // once: the table is in place before kernels on the caller's stream read it
    MTGV_CUDA_OK(ctx, cudaMalloc((void**)&s->tables, sizeof(JpegEncTables)));
    MTGV_CUDA_OK(ctx, cudaMalloc((void**)&s->header, 1024));
    for (auto& e : s->ev) MTGV_CUDA_OK(ctx, cudaEventCreate(&e));
  }
  JpegEncState* st = (JpegEncState*)ctx->jpegenc;
  if (st->timed) MTGV_CUDA_OK(ctx, cudaEventSynchronize(st->ev[2]));  // the previous batch (on whatever stream) still owns tables and scratch
  if (st->quality != quality || st->h != h || st->w != w) {
    JpegEncTables T;
    jpegenc_tables(quality, &T);
    const std::vector<uint8_t> hdr = jpegenc_header(h, w, T);
    if (hdr.size() > 1024) return fail(ctx, MTGV_ERR_LIMIT, "mtgv_encode_jpeg_batch: header too long");
    // on the caller's stream, so that the kernels below are ordered behind the upload
    MTGV_CUDA_OK(ctx, cudaMemcpyAsync(st->tables, &T, sizeof(T), cudaMemcpyHostToDevice, stream));
    MTGV_CUDA_OK(ctx, cudaMemcpyAsync(st->header, hdr.data(), hdr.size(), cudaMemcpyHostToDevice, stream));
    MTGV_CUDA_OK(ctx, cudaStreamSynchronize(stream));  // T and hdr are locals
    st->header_len = (int)hdr.size();
    st->quality = quality; st->h = h; st->w = w;
  }
  const int mcux = (w + 15) / 16, nmcu = mcux * ((h + 15) / 16), nblk = nmcu * 6;
  int rc;
  if ((rc = enc_grow(ctx, (void**)&st->coef, &st->coef_cap, (size_t)n * nblk * 64 * sizeof(int16_t)))) return rc;
  if ((rc = enc_grow(ctx, (void**)&st->bits, &st->bits_cap, (size_t)n * (size_t)cap))) return rc;
  if ((rc = enc_grow(ctx, (void**)&st->boff, &st->boff_cap, ((size_t)n * nblk + (size_t)n) * sizeof(uint32_t)))) return rc;
  ImgLayout L;
  if (layout == MTGV_LAYOUT_NCHW) { L.img_stride = (int64_t)3 * h * w; L.row_stride = w; L.px_stride = 1; L.ch_stride = h * w; }
  else { L.img_stride = (int64_t)3 * h * w; L.row_stride = 3 * w; L.px_stride = 3; L.ch_stride = 1; }
  MTGV_CUDA_OK(ctx, cudaEventRecord(st->ev[0], stream));
  if ((h | w) & 15) k_jpegenc_dct<true><<<dim3((nmcu + 3) / 4, n), 256, 0, stream>>>(images, L, h, w, mcux, nmcu, st->tables, st->coef);
  else k_jpegenc_dct<false><<<dim3((nmcu + 3) / 4, n), 256, 0, stream>>>(images, L, h, w, mcux, nmcu, st->tables, st->coef);
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  MTGV_CUDA_OK(ctx, cudaEventRecord(st->ev[1], stream));
  MTGV_CUDA_OK(ctx, cudaMemsetAsync(st->bits, 0, (size_t)n * (size_t)cap, stream));
  {
    // per image as many 8-warp CTAs as its blocks need, capped so that the grid stays around 32 CTAs per SM
    int gx = (nblk + 7) / 8;
    const int cap_gx = (ctx->sm_count * 32 + n - 1) / n;
    if (gx > cap_gx) gx = cap_gx < 1 ? 1 : cap_gx;
    const dim3 grid((unsigned)gx, (unsigned)n);
    uint32_t* tbits = st->boff + (size_t)n * nblk;  // [n] behind the block offsets
    k_jpegenc_size<<<grid, 256, 0, stream>>>(st->coef, nblk, st->tables, st->boff);
    MTGV_CUDA_OK(ctx, cudaGetLastError());
    k_jpegenc_scan<<<n, kHuffThreads, 0, stream>>>(nblk, st->boff, tbits);
    MTGV_CUDA_OK(ctx, cudaGetLastError());
    k_jpegenc_emit<<<grid, 256, 0, stream>>>(st->coef, nblk, st->tables, st->boff, tbits, st->bits, cap);
    MTGV_CUDA_OK(ctx, cudaGetLastError());
    k_jpegenc_stuff<<<n, kHuffThreads, 0, stream>>>(tbits, st->header, st->header_len, st->bits, out, cap, out_len);
    MTGV_CUDA_OK(ctx, cudaGetLastError());
  }
  MTGV_CUDA_OK(ctx, cudaEventRecord(st->ev[2], stream));
  st->timed = true;
  ctx->launches += 5;
  return MTGV_OK;
}

}  // namespace mtgv
