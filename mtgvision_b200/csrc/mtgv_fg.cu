// mtgv_fg.cu - k_foreground: INTER_AREA resize of one card plane for sm_100a.
//
// crop_to_size(pad=True) of make_virtual (mtgvision/util/image.py:349-377 -> resize :321-334) and
// remove_border_resized of make_cropped (:337-346) are cv2.resize(INTER_AREA) down-scales of the
// 680x488 card to (178,128) / (192,128).  This is the path's HBM-bound kernel: every source byte
// is read exactly once with 16-byte loads.
//
// The separable area filter is evaluated vertical-first.  A destination row's taps are
// [partial row?] full rows* [partial row?] (computeResizeAreaTab); full rows
// all weigh 1/cell, so they are summed as packed 16-bit integers (two byte columns per 32-bit add,
// exact), and only the partial rows are converted to float32.  The weighted column sums of one
// destination row go through a per-warp row buffer and the lanes reduce them horizontally with the
// column taps held in registers.  cv2 filters horizontally first; the real-arithmetic result is the
// same and the float32 roundings differ by ~1e-7, inside the +-1 uint8 LSB bar (the u8 -> [0,1]
// scaling 1/255 is folded into the row weights for the same reason).  Coordinates (tap windows,
// partial-tap thresholds) follow cv2's double arithmetic exactly (mtgv_geom.cuh: area_compact).
//
// One warp owns one (sample, plane, destination-row range) item; there are no block-level barriers
// after the column tables are built.
#include "mtgv_internal.cuh"

namespace mtgv {

constexpr int kFgThreads = 256;
constexpr int kFgWarps = kFgThreads / 32;
constexpr int kFgRows = 32;  // max destination rows per item
constexpr int kFgMaxOW = 256;
[[maybe_unused]] constexpr int kFgPf = 4;  // source rows in flight per warp (cp.async ring, MTGV_FG_BULK == 0)
// MTGV_FG_BULK=1: row streaming through the TMA unit instead - one lane copies kFgStageRows whole source rows (contiguous in the
// planar pool) with a single cp.async.bulk into a per-warp ring of kFgStages stages and arms the stage's mbarrier with the byte
// count; the warp waits on the barrier's phase (UBLKCP + SYNCS in the SASS).  Built, parity-green (tests/test_gpu_encoder.py,
// test_gpu_shapes.py incl. wide cards) and MEASURED on one B200 against the cp.async ring, pipeline ms per 1024 x-samples:
// cp.async ring 4.304 | bulk 4 rows x 3 stages 4.323 | 4x2 4.321 | 2x4 4.349 | 8x3 4.503 | 16x2 4.486
// (profiles/r02_ab_k_background_variants.txt).  The kernel is bound by its arithmetic (70 % issue utilisation), not by the row
// loads; a 2 KB bulk copy per 4 rows and warp has more fixed latency than 31 lanes each posting one 16-byte cp.async.  The
// cp.async ring stays the default.
#ifndef MTGV_FG_BULK
#define MTGV_FG_BULK 0
#endif
#ifndef MTGV_FG_STAGE_ROWS
#define MTGV_FG_STAGE_ROWS 4
#endif
#ifndef MTGV_FG_STAGES
#define MTGV_FG_STAGES 3
#endif
[[maybe_unused]] constexpr int kFgStageRows = MTGV_FG_STAGE_ROWS, kFgStages = MTGV_FG_STAGES;

__device__ __forceinline__ void fg_mbar_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void fg_bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar), d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src), "r"(bytes), "r"(b)
               : "memory");
}
__device__ __forceinline__ void fg_mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  unsigned ok = 0;
  for (int spin = 0; !ok; spin++) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
    if (spin > (1 << 26)) __trap();  // a lost copy must not hang the GPU: fail the launch instead
  }
}

struct FgGeom {  // the two INTER_AREA geometries of a batch: [0] virtual (whole card, padded), [1] cropped
  int src_h, src_w, rh, rw;
};

struct FgRow {  // vertical taps of one destination row: start row, n | left << 8 | right << 9, weights / 255
  int start, n;
  float wl, wm, wr;
};

__device__ __forceinline__ int fg_pos(int x) { return x + ((x >> 5) << 2); }  // skew: conflict-free 16-byte row stores
__device__ __forceinline__ float fg_clip01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }
__device__ __forceinline__ float fg_byte(uint32_t w, int k) {  // exact u8 -> float
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u + k)) - 8388608.f;
}

// NCH: 16-byte chunks per lane and row (pitch <= 512 * NCH).  REGW: horizontal tap weights held in
// registers (<= 5 taps, <= 4 destination columns per lane: scales < 4, out_w <= 128); otherwise they are
// read from the shared tables.
template <int NCH, bool REGW>
__device__ __forceinline__ void fg_item(const mtgv_enc_params* __restrict__ sp, int c, int part, int split, const FgRow* __restrict__ cols,
                                        FgRow* rows, float* rowbuf, uint8_t* ring, uint64_t* bars, unsigned& phases,
                                        const uint8_t* __restrict__ card_planes, int pitch, float* __restrict__ fg_out, int s, int lane) {
  const int kind = sp->kind;
  const int src_h = sp->src_h, rh = sp->fg_rh, rw = sp->fg_rw;
  const int OH = sp->out_h, OW = sp->out_w, card_h = sp->card_h, card_w = sp->card_w;
  const int src_y0 = sp->src_y0, src_x0 = sp->src_x0, fy0 = sp->fg_y0, fx0 = sp->fg_x0;
  const bool flip_src = sp->upsidedown && kind == MTGV_KIND_VIRTUAL;  // rot180 of the card before masking
  const bool flip_dst = sp->upsidedown && kind == MTGV_KIND_CROPPED;  // rot180 of the resized crop
  const int rows_per = (rh + split - 1) / split;
  const int r0 = part * rows_per, r1 = min(rh, r0 + rows_per);
  if (r0 >= r1) return;
  __syncwarp();
  if (lane < r1 - r0) {
    FgRow e;
    area_compact(src_h, rh, r0 + lane, &e.start, &e.n, &e.wl, &e.wm, &e.wr);
    const float k255 = 1.f / 255.f;
    e.wl *= k255; e.wm *= k255; e.wr *= k255;
    rows[lane] = e;
  }
  __syncwarp();
  const uint8_t* plane = card_planes + ((size_t)sp->card * 3 + c) * card_h * pitch;
  float* outp = fg_out + ((size_t)s * 3 + c) * OH * OW;
  const int sy_last = rows[r1 - r0 - 1].start + (rows[r1 - r0 - 1].n & 255) - 1;
#if MTGV_FG_BULK
  // Source rows are consumed in increasing order, kFgStageRows per ring stage.  Stage j holds rows sy_first + j*K ... (+K), one
  // contiguous block of the plane (descending addresses when the card is rotated by 180 degrees: the block then starts at the
  // stage's LAST row).  A stage is refilled, by lane 0 after a warp barrier, when the warp moves on to the next one.
  const int sy_first = rows[0].start;
  const int n_rows = sy_last - sy_first + 1, n_stages = (n_rows + kFgStageRows - 1) / kFgStageRows;
  const int stage_bytes = kFgStageRows * NCH * 512;
  auto issue_stage = [&](int j) {  // lane 0 only
    if (j >= n_stages) return;
    const int a = j * kFgStageRows, cnt = min(kFgStageRows, n_rows - a);
    const int row_lo = flip_src ? card_h - 1 - (src_y0 + sy_first + a + cnt - 1) : src_y0 + sy_first + a;
    fg_bulk_load(ring + (size_t)(j % kFgStages) * stage_bytes, plane + (size_t)row_lo * pitch, (unsigned)(cnt * pitch), bars + (j % kFgStages));
  };
  if (lane == 0)
    for (int j = 0; j < kFgStages; j++) issue_stage(j);
  uint4 cur[NCH];
  int cur_sy = sy_first - 1;
#else
  // Source rows are consumed in increasing order.  kFgPf rows are kept in flight ahead of the one in use as
  // 16-byte cp.async copies into a per-warp ring in shared memory (each lane later reads back only the bytes
  // it copied itself, so no warp barrier is involved); exactly one group is committed per row.
  auto prefetch_row = [&](int sy) {
    if (sy <= sy_last) {
      int row = src_y0 + sy;
      if (flip_src) row = card_h - 1 - row;
      const uint8_t* p = plane + (size_t)row * pitch;
      uint8_t* dst = ring + (size_t)(sy % kFgPf) * (NCH * 512);
#pragma unroll
      for (int h = 0; h < NCH; h++) {
        const int off = (lane + 32 * h) * 16;
        if (off < pitch) {
          const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + off);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(p + off) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  uint4 cur[NCH];
  const int sy_first = rows[0].start;
  int cur_sy = sy_first - 1;
#pragma unroll
  for (int u = 0; u < kFgPf; u++) prefetch_row(sy_first + u);
#endif
  for (int r = r0; r < r1; r++) {
    const FgRow e = rows[r - r0];
    const int ny = e.n & 255;
    float fa[NCH][16];
    uint32_t ia[NCH][8];
#pragma unroll
    for (int h = 0; h < NCH; h++) {
#pragma unroll
      for (int t = 0; t < 16; t++) fa[h][t] = 0.f;
#pragma unroll
      for (int t = 0; t < 8; t++) ia[h][t] = 0u;
    }
    bool any_mid = false;
    for (int j = 0; j < ny; j++) {
      const int sy = e.start + j;
      while (sy > cur_sy) {  // rows shared by two destination rows (a partial tap of each) stay in registers
        cur_sy++;
#if MTGV_FG_BULK
        const int rel = cur_sy - sy_first, j = rel / kFgStageRows, i = rel - j * kFgStageRows, slot = j % kFgStages;
        if (i == 0) {
          if (j > 0) {  // every lane is done with stage j-1 (its rows are in registers or consumed): refill it with stage j-1+S
            __syncwarp();
            if (lane == 0) {
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the warp's reads of the stage are ordered before the copy's writes
              issue_stage(j - 1 + kFgStages);
            }
          }
          fg_mbar_wait(bars + slot, (phases >> slot) & 1u);
          phases ^= 1u << slot;
        }
        const int cnt = min(kFgStageRows, n_rows - j * kFgStageRows);
        const uint8_t* src = ring + (size_t)slot * stage_bytes + (size_t)(flip_src ? cnt - 1 - i : i) * pitch;
#else
        asm volatile("cp.async.wait_group %0;" ::"n"(kFgPf - 1) : "memory");
        const uint8_t* src = ring + (size_t)(cur_sy % kFgPf) * (NCH * 512);
#endif
#pragma unroll
        for (int h = 0; h < NCH; h++) {
          const int off = (lane + 32 * h) * 16;
          cur[h] = off < pitch ? *(const uint4*)(src + off) : make_uint4(0u, 0u, 0u, 0u);
        }
#if !MTGV_FG_BULK
        prefetch_row(cur_sy + kFgPf);
#endif
      }
      const bool left = j == 0 && (e.n & 256), right = j == ny - 1 && (e.n & 512);
      if (left || right) {
        const float w = left ? e.wl : e.wr;
#pragma unroll
        for (int h = 0; h < NCH; h++) {
          const uint32_t w4[4] = {cur[h].x, cur[h].y, cur[h].z, cur[h].w};
#pragma unroll
          for (int t = 0; t < 16; t++) fa[h][t] = __fmaf_rn(fg_byte(w4[t >> 2], t & 3), w, fa[h][t]);
        }
      } else {
        any_mid = true;
#pragma unroll
        for (int h = 0; h < NCH; h++) {
          const uint32_t w4[4] = {cur[h].x, cur[h].y, cur[h].z, cur[h].w};
#pragma unroll
          for (int t = 0; t < 4; t++) {
            ia[h][2 * t] += __byte_perm(w4[t], 0u, 0x4140);      // bytes 0,1 as two 16-bit lanes
            ia[h][2 * t + 1] += __byte_perm(w4[t], 0u, 0x4342);  // bytes 2,3
          }
        }
      }
    }
    if (any_mid) {
#pragma unroll
      for (int h = 0; h < NCH; h++)
#pragma unroll
        for (int t = 0; t < 8; t++) {
          fa[h][2 * t] = __fmaf_rn((float)(ia[h][t] & 0xFFFFu), e.wm, fa[h][2 * t]);
          fa[h][2 * t + 1] = __fmaf_rn((float)(ia[h][t] >> 16), e.wm, fa[h][2 * t + 1]);
        }
    }
    __syncwarp();  // the previous row's horizontal readers are done
#pragma unroll
    for (int h = 0; h < NCH; h++) {
      const int xb = (lane + 32 * h) * 16;
      if (xb < pitch) {
        if (!flip_src) {
#pragma unroll
          for (int t = 0; t < 4; t++)
            *(float4*)(rowbuf + fg_pos(xb + 4 * t)) = make_float4(fa[h][4 * t], fa[h][4 * t + 1], fa[h][4 * t + 2], fa[h][4 * t + 3]);
        } else {
#pragma unroll
          for (int t = 0; t < 16; t++) {
            const int x = card_w - 1 - (xb + t);
            if (x >= 0) rowbuf[fg_pos(x)] = fa[h][t];
          }
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < (REGW ? 4 : kFgMaxOW / 32); q++) {
      const int d = lane + 32 * q;
      if (d < rw) {
        // column taps: first * wl + (sum of the inner taps) * wm + last * wr; the table stores wl = wm / wr = wm
        // for taps that are not partial (REGW: <= 5 taps, unrolled)
        const FgRow ce = cols[d];
        const int nx = ce.n & 255;
        const float* v = rowbuf + fg_pos(src_x0 + ce.start);
        const int wrap = 32 - ((src_x0 + ce.start) & 31);  // taps from this index on sit behind the next skew step
        float hs = v[0] * ce.wl, mid = 0.f;
        if (REGW) {
#pragma unroll
          for (int k = 1; k < 4; k++)
            if (k < nx - 1) mid += v[k + (k >= wrap ? 4 : 0)];
        } else {
          for (int k = 1; k < nx - 1; k++) mid += v[k + (((src_x0 + ce.start + k) >> 5) - ((src_x0 + ce.start) >> 5)) * 4];
        }
        hs = __fmaf_rn(mid, ce.wm, hs);
        if (nx > 1) {
          const int k = nx - 1;
          hs = __fmaf_rn(v[k + (((src_x0 + ce.start + k) >> 5) - ((src_x0 + ce.start) >> 5)) * 4], ce.wr, hs);
        }
        const int y = fy0 + r, x = fx0 + d;
        const int o = flip_dst ? (OH - 1 - y) * OW + (OW - 1 - x) : y * OW + x;
        outp[o] = fg_clip01(hs);  // img_clip after cv2.resize (util/image.py:334)
      }
    }
  }
}

#ifndef MTGV_FG_BLOCKS
#define MTGV_FG_BLOCKS 2
#endif
template <int NCH, bool REGW>
__global__ void __launch_bounds__(kFgThreads, NCH == 1 ? MTGV_FG_BLOCKS : 2)
    k_foreground(const mtgv_enc_params* __restrict__ params, int n, int split, FgGeom g0, FgGeom g1,
                 const uint8_t* __restrict__ card_planes, int pitch, float* __restrict__ fg_out, int* __restrict__ work_counter) {
  extern __shared__ __align__(16) unsigned char fg_smem_raw[];
  // layout: per geometry: cols[256] | per warp: rows[32] | per warp: rowbuf
  FgRow* ctab = (FgRow*)fg_smem_raw;
  FgRow* rtab = ctab + 2 * kFgMaxOW;
  const int rowbuf_len = (pitch + (pitch >> 3) + 8 + 3) & ~3;  // keeps every warp's buffer 16-byte aligned
  float* rowbufs = (float*)(((uintptr_t)(rtab + kFgWarps * kFgRows) + 15) & ~(uintptr_t)15);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int q = tid; q < 2 * kFgMaxOW; q += blockDim.x) {
    const FgGeom& g = q < kFgMaxOW ? g0 : g1;
    const int d = q & (kFgMaxOW - 1);
    FgRow e{0, 0, 0.f, 0.f, 0.f};
    if (d < g.rw && g.rw > 0) {
      area_compact(g.src_w, g.rw, d, &e.start, &e.n, &e.wl, &e.wm, &e.wr);
      if ((e.n & 255) == 1 && (e.n & 512) && !(e.n & 256)) e.wl = e.wr;  // single right-partial tap
      else if (!(e.n & 256)) e.wl = e.wm;
      if (!(e.n & 512)) e.wr = e.wm;
    }
    ctab[q] = e;
  }
  __syncthreads();
  FgRow* rows = rtab + warp * kFgRows;
  float* rowbuf = rowbufs + (size_t)warp * rowbuf_len;
#if MTGV_FG_BULK
  uint8_t* ring0 = (uint8_t*)(rowbufs + (size_t)kFgWarps * rowbuf_len);
  uint8_t* ring = ring0 + (size_t)warp * kFgStages * kFgStageRows * NCH * 512;
  uint64_t* bars = (uint64_t*)(ring0 + (size_t)kFgWarps * kFgStages * kFgStageRows * NCH * 512) + warp * kFgStages;
  if (lane == 0)
    for (int j = 0; j < kFgStages; j++) fg_mbar_init(bars + j);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // the barriers are visible to the async proxy before the first copy
  __syncwarp();
  unsigned phases = 0u;  // parity to wait for, one bit per ring stage
#else
  uint8_t* ring = (uint8_t*)(rowbufs + (size_t)kFgWarps * rowbuf_len) + (size_t)warp * kFgPf * NCH * 512;
  uint64_t* bars = nullptr;
  unsigned phases = 0u;
#endif
  const int n_items = n * 3 * split;
  for (;;) {  // warps pull items from a global queue: skipped (aliased, cropped-geometry) items cost nothing
    int item = 0;
    if (lane == 0) item = atomicAdd(work_counter, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_items) break;
    const int s = item / (3 * split), rem = item - s * 3 * split, c = rem / split, part = rem - c * split;
    const mtgv_enc_params* sp = params + s;
    const int kind = sp->kind;
    if (sp->status != 0 || kind == MTGV_KIND_BG_ONLY) continue;
    if (fg_plane_owner(params, n, s) != s) continue;  // the pair partner's planes are identical and are reused
    const int gi = kind == MTGV_KIND_CROPPED ? 1 : 0;
    const FgGeom& g = gi ? g1 : g0;
    if (sp->src_h != g.src_h || sp->src_w != g.src_w || sp->fg_rh != g.rh || sp->fg_rw != g.rw) continue;  // host invariant
    fg_item<NCH, REGW>(sp, c, part, split, ctab + gi * kFgMaxOW, rows, rowbuf, ring, bars, phases, card_planes, pitch, fg_out, s, lane);
  }
}

static size_t fg_smem_bytes(int pitch) {
  size_t b = 2 * kFgMaxOW * sizeof(FgRow);
  b += (size_t)kFgWarps * kFgRows * sizeof(FgRow) + 16;
  b += (size_t)kFgWarps * ((pitch + (pitch >> 3) + 8 + 3) & ~3) * 4;
#if MTGV_FG_BULK
  b += (size_t)kFgWarps * kFgStages * kFgStageRows * (pitch > 512 ? 2 : 1) * 512 + (size_t)kFgWarps * kFgStages * 8;
#else
  b += (size_t)kFgWarps * kFgPf * (pitch > 512 ? 2 : 1) * 512;
#endif
  return (b + 15) & ~(size_t)15;
}

int fg_launch(mtgv_ctx* ctx, const mtgv_enc_params* params, int m, int OH, int OW, float* fg_out, cudaStream_t st) {
  if (OW > kFgMaxOW) return fail(ctx, MTGV_ERR_LIMIT, "x_size_hw width > 256");
  if (ctx->card_pitch > 1024) return fail(ctx, MTGV_ERR_LIMIT, "card width > 1024");
  // the two INTER_AREA geometries a batch can contain (make_virtual / make_cropped)
  FgGeom g0{0, 0, 0, 0}, g1{0, 0, 0, 0};
  g0.src_h = ctx->card_h; g0.src_w = ctx->card_w;
  if (ctx->card_h == OH && ctx->card_w == OW) { g0.rh = OH; g0.rw = OW; }
  else { int y0, x0; crop_geometry(ctx->card_h, ctx->card_w, OH, OW, true, &g0.rh, &g0.rw, &y0, &x0); }
  const int border = (int)ceil(fmax(0.02 * ctx->card_h, 0.02 * ctx->card_w));
  g1.src_h = ctx->card_h - 2 * border; g1.src_w = ctx->card_w - 2 * border; g1.rh = OH; g1.rw = OW;
  const int split = (OH + kFgRows - 1) / kFgRows > 8 ? (OH + kFgRows - 1) / kFgRows : 8;
  const size_t smem = fg_smem_bytes(ctx->card_pitch);
  const bool wide = ctx->card_pitch > 512;
  // register-resident column taps need <= 5 taps (scale < 4) and <= 4 destination columns per lane
  int max_taps = 0;
  for (int gi = 0; gi < 2; gi++) {
    const FgGeom& g = gi ? g1 : g0;
    for (int d = 0; d < g.rw; d++) {
      int st;
      float ww[kAreaMaxTaps];
      const int nn = area_taps(g.src_w, g.rw, d, &st, ww);
      if (nn > max_taps) max_taps = nn;
    }
  }
  const bool regw = max_taps <= 5 && OW <= 128;
  void (*kern)(const mtgv_enc_params*, int, int, FgGeom, FgGeom, const uint8_t*, int, float*, int*) =
      wide ? (regw ? k_foreground<2, true> : k_foreground<2, false>) : (regw ? k_foreground<1, true> : k_foreground<1, false>);
  MTGV_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  int blocks = 0;
  MTGV_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, kFgThreads, smem));
  int grid = ctx->sm_count * (blocks > 0 ? blocks : 1);
  const int items = m * 3 * split;
  if (grid > (items + kFgWarps - 1) / kFgWarps) grid = (items + kFgWarps - 1) / kFgWarps;
  if (!ctx->fg_counter) MTGV_CUDA_OK(ctx, cudaMalloc(&ctx->fg_counter, 4));
  MTGV_CUDA_OK(ctx, cudaMemsetAsync(ctx->fg_counter, 0, 4, st));
  kern<<<grid, kFgThreads, smem, st>>>(params, m, split, g0, g1, ctx->card_planes, ctx->card_pitch, fg_out, ctx->fg_counter);
  ctx->launches++;
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  return MTGV_OK;
}

}  // namespace mtgv
