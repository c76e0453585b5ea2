// mtgv_jpeg.cu - batched baseline JPEG decode into the uint8 pool format.
// Replaces the cv2.imread behind imread_float (mtgvision/util/image.py:107-114) for the images that feed
// mtgv_set_bg_pool / mtgv_set_card_pool; arithmetic in mtgv_jpeg.cuh, bit-exact with cv2 (libjpeg-turbo ISLOW +
// fancy upsampling).  Four kernels per batch, all images of the batch in each launch:
//   k_jpeg_entropy_par  one CTA per file.  Huffman decode is a serial bit walk; the CTA squeezes the stuffed zeros (and
//                   restart markers) out of the scan and then decodes it in parallel: files with restart intervals a
//                   thread per interval (independent pieces with known start states), files without - the usual case -
//                   by the self-synchronising scheme described at the kernel.  k_jpeg_dc integrates the DC differences.
//   k_jpeg_idct     8 threads per 8x8 block: dequantise, column pass, row pass through shared memory, one 8-byte
//                   store per thread into the component sample plane.
//   k_jpeg_color    one thread per output pixel: triangle-filter chroma upsampling + fixed-point YCbCr->RGB,
//                   written as HWC uint8 at the caller's offsets (= the layout mtgv_set_bg_pool ingests).
// The container (markers, tables) is parsed on the host; file bytes, tables and descriptors go up in one
// staging copy each.
#include "mtgv_internal.cuh"
#include <mutex>

#include "mtgv_jpeg.cuh"
#include "mtgv_jpeg_prog.h"

#include <thread>

namespace mtgv {

struct JpegState {
  uint8_t* files = nullptr;   size_t files_cap = 0;
  int16_t* coef = nullptr;    size_t coef_cap = 0;    // coefficient blocks of the batch (bytes)
  uint8_t* planes = nullptr;  size_t planes_cap = 0;
  uint8_t* clean = nullptr;   size_t clean_cap = 0;   // byte-unstuffed scans of the single-interval files
  uint64_t* sync = nullptr;   size_t sync_cap = 0;    // subsequence checkpoints (bytes)
  uint32_t* rstpos = nullptr; size_t rstpos_cap = 0;  // restart-interval starts in the clean scans
  int16_t* dcs = nullptr;     size_t dcs_cap = 0;     // DC term of every block of the single-interval files
  int32_t* endblk = nullptr;  size_t endblk_cap = 0;  // per file: first block the entropy decoder never reached
  uint8_t* desc = nullptr;    size_t desc_cap = 0;    // JpegImg[n] | JpegTables[n]
  int16_t* prog_host = nullptr; size_t prog_host_cap = 0;  // pinned: coefficients of the batch's progressive files (host entropy stage)
  uint8_t* desc_host = nullptr; size_t desc_host_cap = 0;  // pinned
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // around the three kernels of the last batch
  bool timed = false;
};

static int grow(mtgv_ctx* ctx, void** p, size_t* cap, size_t need) {
  if (need <= *cap) return MTGV_OK;
  if (*p) MTGV_CUDA_OK(ctx, cudaFree(*p));
  *p = nullptr; *cap = 0;
  need += need / 4;
  MTGV_CUDA_OK(ctx, cudaMalloc(p, need));
  *cap = need;
  return MTGV_OK;
}

__constant__ uint8_t c_zigzag[64];

// ---- Huffman symbol look-up shared by the entropy kernels ----
template <class Tables>
__device__ __forceinline__ int ent_symbol(const Tables& S, int ti, uint32_t win, int& used) {
  const unsigned e = S.fast[ti][win >> (32 - kJpegFastBits)];
  if (e) {
    used = (int)(e >> 8);
    return (int)(e & 255u);
  }
  const int top16 = (int)(win >> 16);
  int l = kJpegFastBits + 1;
  while (l <= 16 && (top16 >> (16 - l)) > S.maxcode[ti][l]) l++;
  if (l > 16) {  // corrupt data: libjpeg warns and returns a zero symbol
    used = 16;
    return 0;
  }
  used = l;
  return S.vals[ti][((top16 >> (16 - l)) + S.valoff[ti][l]) & 255];
}

// the s (1..15) bits after the code, sign-extended like HUFF_EXTEND; used + s <= 31
__device__ __forceinline__ int ent_extend(uint32_t win, int used, int s) {
  const int v = (int)((win << used) >> (32 - s));
  return v < (1 << (s - 1)) ? v - ((1 << s) - 1) : v;
}

// ---- entropy decode: one CTA per file, every thread a stretch of the bit stream ----
// Huffman streams resynchronise by themselves: a decoder started at an arbitrary bit in an arbitrary state falls into
// step with the true symbol boundaries after a few dozen symbols (Klein & Wiseman 2003; Weissenberger & Schmidt 2018 for
// GPUs).  The CTA first squeezes the stuffed zeros out of the scan into a clean big-endian copy, then cuts it into
// runs, one per thread, with a checkpoint every 512 bits.
//   round 0   every thread decodes its run from a cold state (bit = start of the run, block start), keeping the state
//             (bit position, zigzag index, block-in-MCU) and the number of completed blocks at every subsequence end;
//   round r   a thread whose predecessor's final state differs from the state it started from decodes again from there,
//             overwriting its checkpoints until one repeats (from there on its earlier decode was already right);
//             repeated until no thread starts over - then every checkpoint follows from the true start of the scan by
//             induction, whatever the data (worst case: as many rounds as threads).  (Measured and dropped: queueing the
//             runs that start over and decoding them one checkpoint per step with packed warps - 10 % slower, the extra
//             barriers and window reloads cost more than the idle lanes; 12-bit look-up tables built by the CTA
//             - 6 % slower: long codes are not the problem, the 34 KB of tables cost occupancy; a third stream word
//             prefetched one window ahead - neutral: the load the profile shows the warps waiting on lands in time either
//             way, what costs is the issue slots of three passes, the second of them at 8 of 32 lanes);
//   write     block counts are prefix-summed, every thread decodes its run once more and stores the coefficients.
// DC terms are stored as differences and integrated by k_jpeg_dc.
constexpr int kSubBits = 512;  // measured: 128 -> 25 % slower (per-checkpoint overhead in the cold pass), 1024 -> later break-off in the rounds
#ifndef MTGV_JPEG_FASTCOUNT
#define MTGV_JPEG_FASTCOUNT 1
#endif
#ifndef MTGV_JPEG_PAR_THREADS
#define MTGV_JPEG_PAR_THREADS 256
#endif
constexpr int kParThreads = MTGV_JPEG_PAR_THREADS;  // runs per file
constexpr uint64_t kStateMask = (1ull << 48) - 1;

struct ParSmem {
  uint16_t fast[4][1 << kJpegFastBits];
#if MTGV_JPEG_FASTCOUNT
  uint16_t cnt[4][1 << kJpegFastBits];  // for the passes that only follow the stream: bits of code + value | zigzag advance << 8
#endif
  int32_t maxcode[4][18];
  int32_t valoff[4][18];
  uint8_t vals[4][256];
  uint8_t blk_c[16], blk_y[16], blk_x[16], blk_td[16], blk_ta[16];
  int64_t j_off[16];           // coefficient index of block j of MCU (0,0)
  int32_t j_rs[16], j_cs[16];  // ... plus my * j_rs + mx * j_cs for MCU (mx, my)
  unsigned wsum[kParThreads / 32];
  int marker, changed;
  uint8_t zz[64];
};

struct ParState {
  unsigned p;  // bits consumed
  int k, j;    // next zigzag index (0: a DC term follows), block within the MCU
};

__device__ __forceinline__ uint64_t par_pack(const ParState& s, int nb) {
  return (uint64_t)s.p | ((uint64_t)s.k << 32) | ((uint64_t)s.j << 40) | ((uint64_t)(nb & 0xffff) << 48);
}
__device__ __forceinline__ ParState par_unpack(uint64_t v) {
  ParState s;
  s.p = (unsigned)v; s.k = (int)(v >> 32) & 63; s.j = (int)(v >> 40) & 15;
  return s;
}

struct ParGeom {
  int nb_mcu, mcux, nblk_scan;
};

__device__ __forceinline__ int16_t* par_blk(const ParSmem& S, int16_t* coef, int j, int mx, int my) {
  return coef + S.j_off[j] + (int64_t)my * S.j_rs[j] + mx * S.j_cs[j];
}

// decodes from st up to the bit `boundary`; returns the number of blocks completed.  WRITE: b = index (scan order) of
// the block st lies in; coefficients of blocks >= nblk_scan (garbage after the last MCU) are dropped.
// ABSDC (restart intervals, decoded whole by one call): DC predictions start at zero here and the DC TERMS are written;
// otherwise the DC DIFFERENCES are written (k_jpeg_dc integrates them).  max_blocks: stop after that many blocks.
// BOUNDED: stream words past the end read as zero (only the pass that finishes a truncated scan runs past the data; the
// others stay within one word of the end, which the CTA zeroes).
template <bool WRITE, bool ABSDC = false, bool BOUNDED = false>
__device__ __forceinline__ int par_decode(const ParSmem& S, const ParGeom& G, const uint32_t* __restrict__ cl, unsigned Lw, ParState& st,
                                          unsigned boundary, int16_t* coef, int16_t* dcs, int& b, int max_blocks = 0x7fffffff) {
  unsigned p = st.p;
  int k = st.k, j = st.j, nb = 0;
  int pred0 = 0, pred1 = 0, pred2 = 0;
  int16_t* blk = nullptr;
  int mx = 0, my = 0;
  if (WRITE) {  // MCU coordinates follow the block counter from here on without divisions
    const int mcu = b / G.nb_mcu;
    my = mcu / G.mcux;
    mx = mcu - my * G.mcux;
    blk = par_blk(S, coef, j, mx, my);
  }
  int tdc = S.blk_td[j], tac = S.blk_ta[j];  // Huffman tables of the current block
  unsigned cur = p >> 5;  // two stream words stay in registers; a symbol is at most 31 bits, so the window moves by 0 or 1 words
  uint32_t w0 = !BOUNDED || cur < Lw ? cl[cur] : 0u, w1 = !BOUNDED || cur + 1u < Lw ? cl[cur + 1u] : 0u;
  while (p < boundary && nb < max_blocks) {
    if ((p >> 5) != cur) {
      cur = p >> 5;
      w0 = w1;
      w1 = !BOUNDED || cur + 1u < Lw ? cl[cur + 1u] : 0u;
    }
    const uint32_t win = __funnelshift_l(w1, w0, p & 31u);
    const bool dc = k == 0;
    const int ti = dc ? tdc : tac;
#if MTGV_JPEG_FASTCOUNT
    unsigned ce = 0;
    if (!WRITE) ce = S.cnt[ti][win >> (32 - kJpegFastBits)];
    if (!WRITE && ce) {  // short code: its length, value bits and zigzag step come from one table entry
      p += ce & 255u;
      k += (int)(ce >> 8);
    } else
#endif
    {
      int used;
      const int sym = ent_symbol(S, ti, win, used);
      const int s = sym & 15, r = dc ? 0 : sym >> 4;
      if (WRITE && ABSDC && dc) {
        const int diff = s ? ent_extend(win, used, s) : 0;
        const int c = S.blk_c[j];
        const int pr = c == 0 ? (pred0 += diff) : (c == 1 ? (pred1 += diff) : (pred2 += diff));
        if (b < G.nblk_scan) dcs[(blk - coef) >> 6] = (int16_t)pr;
        used += s;
      } else if (s) {
        if (WRITE) {
          const int pos = dc ? 0 : k + r;
          if (pos < 64 && b < G.nblk_scan) {
            const int16_t v = (int16_t)ent_extend(win, used, s);
            if (dc) dcs[(blk - coef) >> 6] = v;  // DC differences go to the compact per-block array k_jpeg_dc integrates
            else blk[S.zz[pos]] = v;
          }
        }
        used += s;
      }
      k = dc ? 1 : (s ? k + r + 1 : (r == 15 ? k + 16 : 64));
      p += (unsigned)used;
    }
    if (k >= 64) {
      k = 0;
      nb++;
      if (++j == G.nb_mcu) {
        j = 0;
        if (WRITE && ++mx == G.mcux) { mx = 0; my++; }
      }
      tdc = S.blk_td[j]; tac = S.blk_ta[j];
      if (WRITE) {
        b++;
        blk = par_blk(S, coef, j, mx, my);
      }
    }
  }
  st.p = p; st.k = k; st.j = j;
  return nb;
}

// ---- squeeze out the stuffed zeros: 4 consecutive bytes per thread, 1 KiB per pass; the first marker ends the scan ----
// RST: restart markers (files with restart intervals) are squeezed out too and the clean position behind each is recorded
// in rp: the intervals start there, byte aligned, with fresh DC predictions.  Returns the clean length in bytes.
template <bool RST>
__device__ __forceinline__ unsigned par_unstuff(ParSmem& S, const uint8_t* __restrict__ file, int scan_off, int file_len, uint8_t* __restrict__ cl8,
                                                uint32_t* __restrict__ rp, int nseg, unsigned* nrst_out) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr bool intervals = RST;
  unsigned wp = 0, nrst = 0;
  for (int base = scan_off;; base += 4 * kParThreads) {
    const int i0 = base + tid * 4;
    unsigned b[6];  // previous byte, four own bytes, next byte
#pragma unroll
    for (int q = 0; q < 6; q++) {
      const int idx = i0 - 1 + q;
      b[q] = idx < scan_off ? 0u : (idx < file_len ? (unsigned)__ldg(file + idx) : 0x1FFu);  // past the end reads as a marker
    }
    int mk = 0x7fffffff;
    unsigned keep = 0, rst = 0;
#pragma unroll
    for (int q = 1; q <= 4; q++) {
      const bool rst_ff = intervals && b[q] == 0xFFu && (b[q + 1] & 0xF8u) == 0xD0u;      // an RSTn marker ...
      const bool rst_dn = intervals && b[q - 1] == 0xFFu && (b[q] & 0xF8u) == 0xD0u;      // ... and its second byte
      const bool marker = !rst_ff && (b[q] > 0xFFu || (b[q] == 0xFFu && b[q + 1] != 0u));
      if (marker && mk == 0x7fffffff) mk = i0 + q - 1;
      if (!(b[q - 1] == 0xFFu && b[q] == 0u) && !rst_ff && !rst_dn) keep |= 1u << (q - 1);
      if (rst_ff) rst |= 1u << (q - 1);
    }
    if (tid == 0) S.marker = 0x7fffffff;
    __syncthreads();
    if (mk != 0x7fffffff) atomicMin(&S.marker, mk);
    __syncthreads();
    const int first_marker = S.marker;
#pragma unroll
    for (int q = 0; q < 4; q++)
      if (i0 + q >= first_marker) { keep &= ~(1u << q); rst &= ~(1u << q); }
    const int cnt = __popc(keep) | (__popc(rst) << 16);  // kept bytes | restart markers, scanned together
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) S.wsum[warp] = (unsigned)incl;
    __syncthreads();
    unsigned woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kParThreads / 32; w++) {
      if (w < warp) woff += S.wsum[w];
      tot += S.wsum[w];
    }
    const unsigned before = woff + (unsigned)(incl - cnt);
    unsigned pos = wp + (before & 0xffffu), ri = nrst + (before >> 16);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      if ((keep >> q) & 1u) cl8[(pos++) ^ 3u] = (uint8_t)b[q + 1];
      if ((rst >> q) & 1u) {
        if ((int)ri < nseg - 1) rp[ri] = pos;  // the next interval starts at this clean byte
        ri++;
      }
    }
    wp += tot & 0xffffu;
    nrst += tot >> 16;
    __syncthreads();
    if (first_marker != 0x7fffffff) break;
  }
  *nrst_out = nrst;
  return wp;
}

__global__ void __launch_bounds__(kParThreads, 1536 / kParThreads) k_jpeg_entropy_par(const uint8_t* __restrict__ files, const JpegImg* __restrict__ imgs,
                                                                 const JpegTables* __restrict__ tbs,
                                                                 uint8_t* __restrict__ clean, uint64_t* __restrict__ sync,
                                                                 int16_t* __restrict__ coef, int16_t* __restrict__ dcs, int32_t* __restrict__ end_blk,
                                                                 uint32_t* __restrict__ rstpos) {
  __shared__ ParSmem S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int img = blockIdx.x;
  const JpegImg& im = imgs[img];
  if (im.par == 0) return;  // progressive file: its coefficients were decoded on the host and are copied in behind this kernel
  {
    const JpegTables* T = tbs + img;
    for (int i = tid; i < (int)(sizeof(S.fast) / 4); i += kParThreads) ((uint32_t*)S.fast)[i] = ((const uint32_t*)T->fast)[i];
    for (int i = tid; i < 4 * 18; i += kParThreads) {
      ((int32_t*)S.maxcode)[i] = ((const int32_t*)T->maxcode)[i];
      ((int32_t*)S.valoff)[i] = ((const int32_t*)T->valoff)[i];
    }
    for (int i = tid; i < 4 * 256 / 4; i += kParThreads) ((uint32_t*)S.vals)[i] = ((const uint32_t*)T->vals)[i];
    if (tid < 64) S.zz[tid] = c_zigzag[tid];
    if (tid == 0) {
      int j = 0;
      for (int c = 0; c < im.ncomp; c++)
        for (int by = 0; by < im.cv[c]; by++)
          for (int bx = 0; bx < im.ch[c]; bx++, j++) {
            S.blk_c[j] = (uint8_t)c; S.blk_y[j] = (uint8_t)by; S.blk_x[j] = (uint8_t)bx;
            S.blk_td[j] = (uint8_t)im.td[c]; S.blk_ta[j] = (uint8_t)(2 + im.ta[c]);
            S.j_off[j] = (im.coef_blk + im.blk0[c] + (int64_t)by * im.bw[c] + bx) * 64;
            S.j_rs[j] = im.cv[c] * im.bw[c] * 64;
            S.j_cs[j] = im.ch[c] * 64;
          }
    }
  }
  ParGeom G;
  G.nb_mcu = 0;
  for (int c = 0; c < im.ncomp; c++) G.nb_mcu += im.ch[c] * im.cv[c];
  G.mcux = im.mcux;
  G.nblk_scan = im.mcux * im.mcuy * G.nb_mcu;
  const uint8_t* file = files + im.file_off;
  const int file_len = im.file_len, scan_off = im.scan_off;
  uint8_t* cl8 = clean + im.clean_off;
  __syncthreads();
#if MTGV_JPEG_FASTCOUNT
  for (int i = tid; i < 4 << kJpegFastBits; i += kParThreads) {
    const int ti = i >> kJpegFastBits;
    const unsigned e = S.fast[ti][i & ((1 << kJpegFastBits) - 1)];
    const unsigned s = e & 15u, r = (e >> 4) & 15u;
    const unsigned adv = ti < 2 ? 1u : (s ? r + 1u : (r == 15u ? 16u : 64u));  // tables 0,1: DC (the block moves on to its AC terms)
    S.cnt[ti][i & ((1 << kJpegFastBits) - 1)] = e ? (uint16_t)(((e >> 8) + s) | (adv << 8)) : (uint16_t)0;
  }
  // (read after the barriers of the unstuffing pass)
#endif

  const bool intervals = im.par == 2;
  uint32_t* rp = rstpos + im.rst_off;
  unsigned nrst = 0;
  const unsigned wp = intervals ? par_unstuff<true>(S, file, scan_off, file_len, cl8, rp, im.nseg, &nrst)
                                : par_unstuff<false>(S, file, scan_off, file_len, cl8, rp, im.nseg, &nrst);
  if (tid < 16) cl8[(wp + tid) ^ 3u] = 0;  // the last word and the three behind it read as zero bits
  __syncthreads();
  const unsigned Lw = (wp + 3u) >> 2;
  const uint32_t* cl = (const uint32_t*)cl8;

  if (intervals) {
    // ---- restart intervals: independent, byte-aligned pieces with known start states - a thread per interval ----
    const int nint = (int)nrst + 1 < im.nseg ? (int)nrst + 1 : im.nseg;  // intervals whose start was found
    const int total_mcu = im.mcux * im.mcuy;
    for (int iv = tid; iv < nint; iv += kParThreads) {
      const unsigned b0 = iv == 0 ? 0u : rp[iv - 1], b1 = iv + 1 < nint ? rp[iv] : wp;
      ParState st;
      st.p = b0 * 8u; st.k = 0; st.j = 0;
      int b = iv * im.dri * G.nb_mcu;
      const int mcus = total_mcu - iv * im.dri < im.dri ? total_mcu - iv * im.dri : im.dri;
      par_decode<true, true>(S, G, cl, Lw, st, b1 * 8u, coef, dcs, b, mcus * G.nb_mcu);
    }
    if (tid == 0) end_blk[img] = G.nblk_scan;
    return;
  }

  // ---- runs: thread t owns bits [t * run_bits, (t + 1) * run_bits) of the clean scan, a checkpoint every kSubBits ----
  const unsigned total_bits = wp * 8u;
  unsigned run_bits = ((total_bits + kParThreads - 1) / kParThreads + 31u) & ~31u;
  if (run_bits < (unsigned)kSubBits) run_bits = kSubBits;
  const int cps = (int)((run_bits + kSubBits - 1) / kSubBits);
  const unsigned run0 = (unsigned)tid * run_bits;
  const bool active = run0 < total_bits || tid == 0;
  const unsigned run1 = run0 + run_bits < total_bits ? run0 + run_bits : total_bits;
  const int ncp = active ? (int)((run1 - run0 + kSubBits - 1) / kSubBits) : 0;
  uint64_t* out = sync + im.sync_off + (size_t)tid * cps;
  int dummy_b = 0;
  ParState cold;
  cold.p = run0; cold.k = 0; cold.j = 0;
  uint64_t my_in = par_pack(cold, 0);
  {
    ParState st = cold;
    for (int c = 0; c < ncp; c++) {
      const unsigned bnd = run0 + (unsigned)(c + 1) * kSubBits;
      const int nb = par_decode<false>(S, G, cl, Lw, st, bnd < run1 ? bnd : run1, nullptr, nullptr, dummy_b);
      out[c] = par_pack(st, nb);
    }
  }
  __syncthreads();
  for (;;) {
    uint64_t in = my_in;
    if (tid > 0 && active) in = out[-1] & kStateMask;  // the predecessor's run is full: its last checkpoint is its run end
    if (tid == 0) S.changed = 0;
    __syncthreads();
    if (in != my_in) {
      my_in = in;
      S.changed = 1;
      ParState st = par_unpack(in);
      for (int c = 0; c < ncp; c++) {
        const unsigned bnd = run0 + (unsigned)(c + 1) * kSubBits;
        const int nb = par_decode<false>(S, G, cl, Lw, st, bnd < run1 ? bnd : run1, nullptr, nullptr, dummy_b);
        const uint64_t nw = par_pack(st, nb), old = out[c];
        out[c] = nw;
        if (((nw ^ old) & kStateMask) == 0) break;
      }
    }
    __syncthreads();
    const int ch = S.changed;
    __syncthreads();
    if (!ch) break;
  }

  // ---- block index at the start of every run, then the writing pass ----
  int mine = 0;
  for (int c = 0; c < ncp; c++) mine += (int)(out[c] >> 48);
  int incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) S.wsum[warp] = (unsigned)incl;
  __syncthreads();
  int b = incl - mine;
  for (int w = 0; w < warp; w++) b += (int)S.wsum[w];
  ParState st = par_unpack(my_in);
  if (ncp > 0) par_decode<true>(S, G, cl, Lw, st, run1, coef, dcs, b);
  if (active && run1 == total_bits) {
    // The run that reaches the end of the data.  A complete scan has decoded every block by now.  After a premature end
    // (truncated file, stray marker) libjpeg finishes the MCU in which a request for bits ran past the data from zero bits
    // and skips every later MCU, whose coefficients stay zero - grey (jdhuff.c decode_mcu, insufficient_data).
    if (b < G.nblk_scan) {
      if (st.p <= total_bits) par_decode<true, false, true>(S, G, cl, Lw, st, st.p + 1u, coef, dcs, b);  // the request that finds no data
      while (!(st.k == 0 && st.j == 0) && b < G.nblk_scan) par_decode<true, false, true>(S, G, cl, Lw, st, st.p + 1u, coef, dcs, b);
    }
    end_blk[img] = b;  // first block that was never decoded
  }
}

// one CTA per (file, component): DC differences -> DC terms, in scan order (a single restart interval); 256 blocks per
// step, scanned by the warps and stitched through shared memory (one warp per component had left the kernel waiting on
// 167 dependent steps of the card files' luma plane: 0.13 ms)
constexpr int kDcThreads = 256;
__global__ void __launch_bounds__(kDcThreads) k_jpeg_dc(const JpegImg* __restrict__ imgs, int16_t* __restrict__ dcs,
                                                        const int32_t* __restrict__ end_blk) {
  __shared__ int wsum[kDcThreads / 32];
  const JpegImg& im = imgs[blockIdx.x];
  const int c = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (c >= im.ncomp || im.par != 1) return;  // restart-interval files carry their DC terms already
  int nb_mcu = 0, joff = 0;  // blocks per MCU, first block of this component within the MCU
  for (int k = 0; k < im.ncomp; k++) {
    if (k < c) joff += im.ch[k] * im.cv[k];
    nb_mcu += im.ch[k] * im.cv[k];
  }
  const int eb = end_blk[blockIdx.x];  // blocks from here on were skipped (premature end of data): they stay zero
  const int ch = im.ch[c], cv = im.cv[c], bw = im.bw[c], nbc = ch * cv, mcux = im.mcux, total = im.mcux * im.mcuy * nbc;
  int16_t* base = dcs + (im.coef_blk + im.blk0[c]);  // one entry per block, same block order as the coefficient array
  int carry = 0;
  for (int q0 = 0; q0 < total; q0 += kDcThreads) {
    const int q = q0 + tid;
    int16_t* ptr = nullptr;
    int v = 0;
    bool keep = false;
    if (q < total) {
      const int m = q / nbc, wi = q - m * nbc, by = wi / ch, bx = wi - by * ch, my = m / mcux, mx = m - my * mcux;
      ptr = base + ((int64_t)(my * cv + by) * bw + mx * ch + bx);
      v = *ptr;
      keep = m * nb_mcu + joff + wi < eb;
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += t;
    }
    if (lane == 31) wsum[warp] = v;
    __syncthreads();
    int before = carry, all = 0;
#pragma unroll
    for (int w = 0; w < kDcThreads / 32; w++) {
      if (w < warp) before += wsum[w];
      all += wsum[w];
    }
    if (keep) *ptr = (int16_t)(v + before);
    carry += all;
    __syncthreads();
  }
}

// grid (ceil(max blocks / 32), n images), 256 threads = 32 blocks of 8 threads.  Thread t of a block loads coefficient ROW t
// and its quantiser row as one 16-byte word each, the products go through shared memory to the thread of their COLUMN
// (jidctint.c runs the column pass first; the order is part of the rounding), and back for the row pass.
// (Measured and dropped: clearing every coefficient row behind its use instead of a memset per batch - the stores cost
// the kernel what the memset costs.)
__global__ void __launch_bounds__(256) k_jpeg_idct(const JpegImg* __restrict__ imgs, const JpegTables* __restrict__ tbs,
                                                   const int16_t* __restrict__ coef, const int16_t* __restrict__ dcs,
                                                   uint8_t* __restrict__ planes) {
  __shared__ int ws[32][8][9];
  const JpegImg& im = imgs[blockIdx.y];
  const int g = threadIdx.x >> 3, t = threadIdx.x & 7;
  const int b = blockIdx.x * 32 + g;
  const bool live = b < im.nblk;
  int c = 0;
  if (live) {
    while (c + 1 < im.ncomp && b >= im.blk0[c + 1]) c++;
    const uint4 cw = __ldg((const uint4*)(coef + (im.coef_blk + b) * 64) + t);
    const uint4 qw = __ldg((const uint4*)tbs[blockIdx.y].qt[im.tq[c]] + t);
    const uint32_t cv[4] = {cw.x, cw.y, cw.z, cw.w}, qv[4] = {qw.x, qw.y, qw.z, qw.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      ws[g][t][2 * k] = (int)(int16_t)(cv[k] & 0xffffu) * (int)(qv[k] & 0xffffu);
      ws[g][t][2 * k + 1] = (int)(int16_t)(cv[k] >> 16) * (int)(qv[k] >> 16);
    }
    if (im.par && t == 0) ws[g][0][0] = (int)dcs[im.coef_blk + b] * (int)(qv[0] & 0xffffu);  // DC term from the integrated per-block array
  }
  __syncwarp();
  int o[8];
  if (live) {
    int x[8];
#pragma unroll
    for (int r = 0; r < 8; r++) x[r] = ws[g][r][t];
    jpeg_idct8(x, o, kJpegPass1Shift);  // column t
  }
  __syncwarp();
  if (live) {
#pragma unroll
    for (int r = 0; r < 8; r++) ws[g][r][t] = o[r];
  }
  __syncwarp();
  if (live) {
    int x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = ws[g][t][k];
    jpeg_idct8(x, o, kJpegPass2Shift);  // row t
    const int lb = b - im.blk0[c], by = lb / im.bw[c], bx = lb - by * im.bw[c];
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      lo |= (uint32_t)jpeg_clamp255(o[k] + 128) << (8 * k);
      hi |= (uint32_t)jpeg_clamp255(o[k + 4] + 128) << (8 * k);
    }
    // plane offsets are multiples of 8 and the pitch is a multiple of 8: the row segment is 8-byte aligned
    uint2* dst = (uint2*)(planes + im.plane_off[c] + (int64_t)(by * 8 + t) * (im.bw[c] * 8) + bx * 8);
    *dst = make_uint2(lo, hi);
  }
}

// grid (row bands, n images): a CTA walks the rows of its band; a thread converts 4 consecutive pixels of a row and
// stores their 12 bytes as three words when the row start is word aligned.  4:2:0 (h2v2) and 4:4:4 groups take a
// straight-line path (one aligned 4-byte luma load, the vertical chroma blend shared by the four pixels), row ends
// included (clamped chroma columns); narrow images and the other sampling modes go through the generic jpeg_pixel.
// (Measured and dropped: 8 pixels per thread with word loads of the chroma rows - 18 % slower at 64 registers; the two
// chroma words per row funnel-shifted instead of four byte loads - 10 % slower.)
#ifndef MTGV_JPEG_COLOR_ROWS
#define MTGV_JPEG_COLOR_ROWS 32
#endif
constexpr int kColorRows = MTGV_JPEG_COLOR_ROWS;

__device__ __forceinline__ void ycc_store(int Y, int cb, int cr, uint8_t* px) {
  cb -= 128; cr -= 128;
  px[0] = (uint8_t)jpeg_clamp255(Y + ((91881 * cr + 32768) >> 16));
  px[1] = (uint8_t)jpeg_clamp255(Y + ((-22554 * cb + 32768 - 46802 * cr) >> 16));
  px[2] = (uint8_t)jpeg_clamp255(Y + ((116130 * cb + 32768) >> 16));
}

// the 4 (or fewer, at the row end) pixels of row yy from x0 on
__device__ __forceinline__ void color_group4(const JpegImg& im, const uint8_t* __restrict__ planes, uint8_t* __restrict__ out, bool h2v2, bool h1v1,
                                             int yy, int x0) {
  const int W = im.w, H = im.h;
  const uint8_t* PY = planes + im.plane_off[0];
  const uint8_t* PB = planes + im.plane_off[1];
  const uint8_t* PR = planes + im.plane_off[2];
  const int pitchY = im.bw[0] * 8, pitchC = im.ncomp == 3 ? im.bw[1] * 8 : 0;
  uint8_t* drow = out + im.out_off + (int64_t)yy * W * 3;
  const bool aligned = (((uintptr_t)drow) & 3) == 0;
  // vertical neighbours of the chroma triangle filter (edge rows replicated)
  const int inrow = yy >> 1, dh = im.dh[1];
  const int other = (yy & 1) ? (inrow + 1 < dh ? inrow + 1 : dh - 1) : (inrow > 0 ? inrow - 1 : 0);
  const int bias_e = 8, bias_o = 7;
  uint8_t px[12];
  const int cnt = W - x0 < 4 ? W - x0 : 4;
  if (h2v2) {
    // every group of the row, its ends included: h2v2_fancy_upsample's first / last column formula (4 * colsum + bias) is
    // the general one with the missing neighbour column replaced by the column itself, i.e. with clamped column indices;
    // pixels past the width (cnt < 4) are computed from the padded planes and not stored
    const uint32_t y4 = *(const uint32_t*)(PY + (int64_t)yy * pitchY + x0);
    const int j = x0 >> 1, jmax = im.dw[1] - 1;
    int col[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { const int v = j - 1 + q; col[q] = v < 0 ? 0 : (v > jmax ? jmax : v); }
    const uint8_t* b0 = PB + (int64_t)inrow * pitchC;
    const uint8_t* b1 = PB + (int64_t)other * pitchC;
    const uint8_t* r0 = PR + (int64_t)inrow * pitchC;
    const uint8_t* r1 = PR + (int64_t)other * pitchC;
    int cb[4], cr[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { cb[q] = 3 * b0[col[q]] + b1[col[q]]; cr[q] = 3 * r0[col[q]] + r1[col[q]]; }
    ycc_store(y4 & 255, (3 * cb[1] + cb[0] + bias_e) >> 4, (3 * cr[1] + cr[0] + bias_e) >> 4, px);
    ycc_store((y4 >> 8) & 255, (3 * cb[1] + cb[2] + bias_o) >> 4, (3 * cr[1] + cr[2] + bias_o) >> 4, px + 3);
    ycc_store((y4 >> 16) & 255, (3 * cb[2] + cb[1] + bias_e) >> 4, (3 * cr[2] + cr[1] + bias_e) >> 4, px + 6);
    ycc_store(y4 >> 24, (3 * cb[2] + cb[3] + bias_o) >> 4, (3 * cr[2] + cr[3] + bias_o) >> 4, px + 9);
  } else if (h1v1) {
    const uint32_t y4 = *(const uint32_t*)(PY + (int64_t)yy * pitchY + x0);
    const uint32_t b4 = *(const uint32_t*)(PB + (int64_t)yy * pitchC + x0);
    const uint32_t r4 = *(const uint32_t*)(PR + (int64_t)yy * pitchC + x0);
#pragma unroll
    for (int q = 0; q < 4; q++) ycc_store((y4 >> (8 * q)) & 255, (b4 >> (8 * q)) & 255, (r4 >> (8 * q)) & 255, px + 3 * q);
  } else {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      int rgb[3] = {0, 0, 0};
      if (q < cnt) jpeg_pixel(im, planes, yy, x0 + q, rgb);
      px[3 * q] = (uint8_t)rgb[0]; px[3 * q + 1] = (uint8_t)rgb[1]; px[3 * q + 2] = (uint8_t)rgb[2];
    }
  }
  if (im.out_layout == 2) {
    // background pool layout: one RGBX word per pixel; x0 and the row pitch are multiples of 4 pixels -> 16-byte store
    uint32_t* d = (uint32_t*)(out + im.out_off) + (int64_t)yy * im.out_pitch + x0;
    uint32_t wv[4];
#pragma unroll
    for (int q = 0; q < 4; q++) wv[q] = (uint32_t)px[3 * q] | ((uint32_t)px[3 * q + 1] << 8) | ((uint32_t)px[3 * q + 2] << 16);
    if (cnt == 4) *(uint4*)d = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    else for (int q = 0; q < cnt; q++) d[q] = wv[q];
  } else if (im.out_layout == 1) {
    // card pool layout: three planes of rows padded to 16 bytes; 4 pixels -> one word per plane
#pragma unroll
    for (int c = 0; c < 3; c++) {
      uint8_t* d = out + im.out_off + ((int64_t)c * H + yy) * im.out_pitch + x0;
      if (cnt == 4) *(uint32_t*)d = (uint32_t)px[c] | ((uint32_t)px[3 + c] << 8) | ((uint32_t)px[6 + c] << 16) | ((uint32_t)px[9 + c] << 24);
      else for (int q = 0; q < cnt; q++) d[q] = px[3 * q + c];
    }
  } else if (aligned && cnt == 4) {
    uint32_t* d = (uint32_t*)(drow + x0 * 3);
#pragma unroll
    for (int q = 0; q < 3; q++)
      d[q] = (uint32_t)px[4 * q] | ((uint32_t)px[4 * q + 1] << 8) | ((uint32_t)px[4 * q + 2] << 16) | ((uint32_t)px[4 * q + 3] << 24);
  } else {
    for (int q = 0; q < 3 * cnt; q++) drow[x0 * 3 + q] = px[q];
  }
}

__global__ void __launch_bounds__(256) k_jpeg_color(const JpegImg* __restrict__ imgs, const uint8_t* __restrict__ planes,
                                                    uint8_t* __restrict__ out) {
  __shared__ JpegImg im;
  for (int i = threadIdx.x; i < (int)(sizeof(JpegImg) / 4); i += 256) ((uint32_t*)&im)[i] = ((const uint32_t*)&imgs[blockIdx.y])[i];
  __syncthreads();
  const int W = im.w, H = im.h;
  const bool h2v2 = im.ncomp == 3 && im.hf[1] == 2 && im.vf[1] == 2 && im.hf[2] == 2 && im.vf[2] == 2 && im.dw[1] > 2;
  const bool h1v1 = im.ncomp == 3 && im.hf[1] == 1 && im.vf[1] == 1 && im.hf[2] == 1 && im.vf[2] == 1;
  const int groups = (W + 3) >> 2;
  const float inv_groups = 1.0f / (float)groups;
  for (int y = blockIdx.x * kColorRows; y < H; y += gridDim.x * kColorRows) {
    const int yend = y + kColorRows < H ? y + kColorRows : H;
    // the band's (row, pixel group) pairs are dealt to the threads as one list: rows narrower than the CTA
    // would otherwise leave most of it idle
    for (int idx = threadIdx.x; idx < (yend - y) * groups; idx += 256) {
      int r = (int)((float)idx * inv_groups);  // idx / groups (idx < 2^17) without the integer division
      if (r * groups > idx) r--;
      else if ((r + 1) * groups <= idx) r++;
      const int yy = y + r, g = idx - r * groups;
      color_group4(im, planes, out, h2v2, h1v1, yy, g * 4);
    }
  }
}

int jpeg_destroy(mtgv_ctx* ctx) {
  JpegState* st = (JpegState*)ctx->jpeg;
  if (!st) return MTGV_OK;
  cudaFree(st->files); cudaFree(st->coef); cudaFree(st->planes); cudaFree(st->desc); cudaFree(st->clean); cudaFree(st->sync); cudaFree(st->endblk); cudaFree(st->dcs); cudaFree(st->rstpos);
  if (st->desc_host) cudaFreeHost(st->desc_host);
  if (st->prog_host) cudaFreeHost(st->prog_host);
  for (auto& e : st->ev) if (e) cudaEventDestroy(e);
  delete st;
  ctx->jpeg = nullptr;
  return MTGV_OK;
}

int jpeg_last_kernel_ms(mtgv_ctx* ctx, float* ms) {
  JpegState* st = (JpegState*)ctx->jpeg;
  if (!st || !st->timed) return fail(ctx, MTGV_ERR_INVALID, "mtgv_jpeg_last_kernel_ms: no batch was decoded yet");
  MTGV_CUDA_OK(ctx, cudaEventSynchronize(st->ev[3]));
  for (int k = 0; k < 3; k++) MTGV_CUDA_OK(ctx, cudaEventElapsedTime(&ms[k], st->ev[k], st->ev[k + 1]));
  return MTGV_OK;
}

int jpeg_info(mtgv_ctx* ctx, const uint8_t* file, int64_t len, int32_t* hw) {
  JpegImg im;
  JpegTables tb;
  std::string err;
  if (jpeg_parse(file, len, 0, &im, &tb, nullptr, &err) != 0) return fail(ctx, MTGV_ERR_INVALID, "mtgv_jpeg_info: " + err);
  hw[0] = im.h;
  hw[1] = im.w;
  return MTGV_OK;
}

// marker walk of n files on a few host threads; fn(i, im, tb) receives every parsed file (called concurrently for
// different i).  Returns the index of the first failing file (message in *err) or -1.
template <class Fn>
static int jpeg_parse_many(const uint8_t* files, const int64_t* file_off, int n, std::string* err, Fn fn) {
  unsigned hc = std::thread::hardware_concurrency();
  const int nt = n < 64 ? 1 : (int)(hc < 2 ? 1 : (hc > 8 ? 8 : hc));
  std::vector<std::string> terr(nt);
  std::vector<int> tbad(nt, -1);
  auto run = [&](int t) {
    const int i0 = (int)((int64_t)n * t / nt), i1 = (int)((int64_t)n * (t + 1) / nt);
    JpegImg im;
    JpegTables tb;
    for (int i = i0; i < i1; i++) {
      const int64_t len = file_off[i + 1] - file_off[i];
      if (len < 0) { terr[t] = "bad offsets"; tbad[t] = i; return; }
      if (jpeg_parse(files + file_off[i], len, i, &im, &tb, nullptr, &terr[t]) != 0) { tbad[t] = i; return; }
      fn(i, im, tb);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; t++) th.emplace_back(run, t);
  run(0);
  for (auto& x : th) x.join();
  for (int t = 0; t < nt; t++)
    if (tbad[t] >= 0) { *err = "file " + std::to_string(tbad[t]) + ": " + terr[t]; return tbad[t]; }
  return -1;
}

int jpeg_info_batch(mtgv_ctx* ctx, const uint8_t* files, const int64_t* file_off, int n, int32_t* hw) {
  std::string err;
  if (jpeg_parse_many(files, file_off, n, &err, [&](int i, const JpegImg& im, const JpegTables&) { hw[2 * i] = im.h; hw[2 * i + 1] = im.w; }) >= 0)
    return fail(ctx, MTGV_ERR_INVALID, "mtgv_jpeg_info_batch: " + err);
  return MTGV_OK;
}

int jpeg_decode_batch(mtgv_ctx* ctx, const uint8_t* files, const int64_t* file_off, int n, uint8_t* out, const int64_t* out_off,
                      const int32_t* hw, cudaStream_t stream) {
  std::vector<JpegDst> dst(n);
  for (int i = 0; i < n; i++) dst[i] = JpegDst{out_off[i], hw[2 * i], hw[2 * i + 1], 0, 0};
  return jpeg_decode_batch_ex(ctx, files, file_off, n, out, dst.data(), stream);
}

int jpeg_decode_batch_ex(mtgv_ctx* ctx, const uint8_t* files, const int64_t* file_off, int n, uint8_t* out, const JpegDst* dst,
                         cudaStream_t stream) {
  if (!ctx->jpeg) {
    ctx->jpeg = new JpegState();
    MTGV_CUDA_OK(ctx, cudaMemcpyToSymbol(c_zigzag, kJpegZigzag, 64));
    MTGV_CUDA_OK(ctx, cudaDeviceSynchronize());  // once: the table is in place before kernels on the caller's stream read it
    for (auto& e : ((JpegState*)ctx->jpeg)->ev) MTGV_CUDA_OK(ctx, cudaEventCreate(&e));
  }
  JpegState* st = (JpegState*)ctx->jpeg;
  std::vector<JpegImg> imgs(n);
  std::vector<JpegTables> tbs(n);
  std::vector<std::vector<int16_t>> prog(n);  // coefficients of progressive files
  std::mutex prog_mu;
  int prog_bad = -1;
  std::string prog_err;
  {  // marker walk + table build, a few host threads over contiguous file ranges
    std::string err;
    if (jpeg_parse_many(files, file_off, n, &err, [&](int i, const JpegImg& im, const JpegTables& tb) {
          imgs[i] = im; tbs[i] = tb;
          if (im.progressive) {  // host entropy stage of a progressive file (mtgv_jpeg_prog.h), on this worker thread
            prog[i].assign((size_t)im.nblk * 64, 0);
            std::string perr;
            if (jpeg_decode_progressive(files + file_off[i], file_off[i + 1] - file_off[i], im, im.comp_id, prog[i].data(), &perr) != 0) {
              std::lock_guard<std::mutex> lk(prog_mu);
              if (prog_bad < 0 || i < prog_bad) { prog_bad = i; prog_err = perr; }
            }
          }
        }) >= 0)
      return fail(ctx, MTGV_ERR_INVALID, "mtgv_decode_jpeg_batch: " + err);
    if (prog_bad >= 0) return fail(ctx, MTGV_ERR_INVALID, "mtgv_decode_jpeg_batch: file " + std::to_string(prog_bad) + ": " + prog_err);
  }
  int64_t nblk_total = 0, plane_total = 0;
  int max_blk = 0, max_h = 0;
  for (int i = 0; i < n; i++) {
    JpegImg& im = imgs[i];
    if (im.h != dst[i].h || im.w != dst[i].w)
      return fail(ctx, MTGV_ERR_INVALID, "mtgv_decode_jpeg_batch: file " + std::to_string(i) + " is " + std::to_string(im.h) + "x" +
                                             std::to_string(im.w) + ", the destination is " + std::to_string(dst[i].h) + "x" +
                                             std::to_string(dst[i].w));
    im.file_off = file_off[i] - file_off[0];
    im.coef_blk = nblk_total;
    im.out_off = dst[i].off;
    im.out_layout = dst[i].layout;
    im.out_pitch = dst[i].pitch;
    for (int c = 0; c < im.ncomp; c++) {
      im.plane_off[c] = plane_total;
      plane_total += (int64_t)im.bw[c] * im.bh[c] * 64;
    }
    nblk_total += im.nblk;
    if (im.nblk > max_blk) max_blk = im.nblk;
    if (im.h > max_h) max_h = im.h;
  }
  const size_t file_bytes = (size_t)(file_off[n] - file_off[0]);
  // scratch of the entropy kernel (one CTA per file): the byte-unstuffed scan, run checkpoints, restart-interval starts
  int64_t clean_total = 0, sync_total = 0, rst_total = 0;
  for (int i = 0; i < n; i++) {
    JpegImg& im = imgs[i];
    im.par = im.progressive ? 0 : (im.nseg == 1 ? 1 : 2);
    const int64_t scan_bytes = (int64_t)im.file_len - im.scan_off;
    im.clean_off = clean_total;
    im.sync_off = sync_total;
    im.rst_off = rst_total;
    clean_total += (scan_bytes + 32 + 15) / 16 * 16;
    sync_total += scan_bytes * 8 / kSubBits + 2 * kParThreads + 8;
    rst_total += im.nseg;
  }
  const size_t o_tb = (sizeof(JpegImg) * n + 15) / 16 * 16, desc_bytes = o_tb + sizeof(JpegTables) * n;
  int rc;
  if ((rc = grow(ctx, (void**)&st->files, &st->files_cap, file_bytes + 16))) return rc;
  if ((rc = grow(ctx, (void**)&st->coef, &st->coef_cap, (size_t)nblk_total * 64 * sizeof(int16_t)))) return rc;
  if ((rc = grow(ctx, (void**)&st->planes, &st->planes_cap, (size_t)plane_total))) return rc;
  if ((rc = grow(ctx, (void**)&st->desc, &st->desc_cap, desc_bytes))) return rc;
  if ((rc = grow(ctx, (void**)&st->clean, &st->clean_cap, (size_t)clean_total + 16))) return rc;
  if ((rc = grow(ctx, (void**)&st->sync, &st->sync_cap, ((size_t)sync_total + 1) * sizeof(uint64_t)))) return rc;
  if ((rc = grow(ctx, (void**)&st->endblk, &st->endblk_cap, (size_t)n * sizeof(int32_t)))) return rc;
  if ((rc = grow(ctx, (void**)&st->dcs, &st->dcs_cap, ((size_t)nblk_total + 1) * sizeof(int16_t)))) return rc;
  if ((rc = grow(ctx, (void**)&st->rstpos, &st->rstpos_cap, ((size_t)rst_total + 1) * sizeof(uint32_t)))) return rc;
  // an earlier batch (on whatever stream) may still be reading the staging buffer and the scratch arrays
  if (st->timed) MTGV_CUDA_OK(ctx, cudaEventSynchronize(st->ev[3]));
  size_t prog_total = 0;
  for (int i = 0; i < n; i++)
    if (imgs[i].progressive) prog_total += (size_t)imgs[i].nblk * 64;
  if (prog_total > st->prog_host_cap) {
    if (st->prog_host) MTGV_CUDA_OK(ctx, cudaFreeHost(st->prog_host));
    st->prog_host = nullptr; st->prog_host_cap = 0;
    MTGV_CUDA_OK(ctx, cudaMallocHost((void**)&st->prog_host, (prog_total + prog_total / 4) * sizeof(int16_t)));
    st->prog_host_cap = prog_total + prog_total / 4;
  }
  if (desc_bytes > st->desc_host_cap) {
    if (st->desc_host) MTGV_CUDA_OK(ctx, cudaFreeHost(st->desc_host));
    st->desc_host = nullptr; st->desc_host_cap = 0;
    MTGV_CUDA_OK(ctx, cudaMallocHost((void**)&st->desc_host, desc_bytes + desc_bytes / 4));
    st->desc_host_cap = desc_bytes + desc_bytes / 4;
  }
  memcpy(st->desc_host, imgs.data(), sizeof(JpegImg) * n);
  memcpy(st->desc_host + o_tb, tbs.data(), sizeof(JpegTables) * n);
  MTGV_CUDA_OK(ctx, cudaMemcpyAsync(st->desc, st->desc_host, desc_bytes, cudaMemcpyHostToDevice, stream));
  MTGV_CUDA_OK(ctx, cudaMemcpyAsync(st->files, files + file_off[0], file_bytes, cudaMemcpyHostToDevice, stream));
  const JpegImg* d_img = (const JpegImg*)st->desc;
  const JpegTables* d_tb = (const JpegTables*)(st->desc + o_tb);
  MTGV_CUDA_OK(ctx, cudaEventRecord(st->ev[0], stream));
  MTGV_CUDA_OK(ctx, cudaMemsetAsync(st->coef, 0, (size_t)nblk_total * 64 * sizeof(int16_t), stream));
  MTGV_CUDA_OK(ctx, cudaMemsetAsync(st->dcs, 0, (size_t)nblk_total * sizeof(int16_t), stream));
  k_jpeg_entropy_par<<<(unsigned)n, kParThreads, 0, stream>>>(st->files, d_img, d_tb, st->clean, st->sync, st->coef, st->dcs, st->endblk, st->rstpos);
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  if (prog_total) {
    // progressive files: their coefficient blocks (absolute DC terms included) go straight into the scratch the IDCT reads
    size_t o = 0;
    for (int i = 0; i < n; i++)
      if (imgs[i].progressive) {
        const size_t cnt = (size_t)imgs[i].nblk * 64;
        memcpy(st->prog_host + o, prog[i].data(), cnt * sizeof(int16_t));
        MTGV_CUDA_OK(ctx, cudaMemcpyAsync(st->coef + (size_t)imgs[i].coef_blk * 64, st->prog_host + o, cnt * sizeof(int16_t), cudaMemcpyHostToDevice, stream));
        o += cnt;
      }
  }
  k_jpeg_dc<<<dim3((unsigned)n, 3), kDcThreads, 0, stream>>>(d_img, st->dcs, st->endblk);
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  MTGV_CUDA_OK(ctx, cudaEventRecord(st->ev[1], stream));
  k_jpeg_idct<<<dim3((max_blk + 31) / 32, n), 256, 0, stream>>>(d_img, d_tb, st->coef, st->dcs, st->planes);
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  MTGV_CUDA_OK(ctx, cudaEventRecord(st->ev[2], stream));
  const int gx = (max_h + kColorRows - 1) / kColorRows;
  k_jpeg_color<<<dim3(gx < 1024 ? gx : 1024, n), 256, 0, stream>>>(d_img, st->planes, out);
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  MTGV_CUDA_OK(ctx, cudaEventRecord(st->ev[3], stream));
  st->timed = true;
  ctx->launches += 4;
  return MTGV_OK;
}

}  // namespace mtgv
