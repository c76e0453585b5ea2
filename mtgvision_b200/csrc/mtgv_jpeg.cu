// mtgv_jpeg.cu - batched baseline JPEG decode into the uint8 pool format.
// Replaces the cv2.imread behind imread_float (mtgvision/util/image.py:107-114) for the images that feed
// mtgv_set_bg_pool / mtgv_set_card_pool; arithmetic in mtgv_jpeg.cuh, bit-exact with cv2 (libjpeg-turbo ISLOW +
// fancy upsampling).  Three kernels per batch, all images of the batch in each launch:
//   k_jpeg_entropy  one decoder thread per restart interval (a file without DRI is one interval).  Huffman decode
//                   is a serial bit walk, so the parallelism is files x intervals; when there are fewer intervals
//                   than warps the machine can hold, every decoder gets a warp to itself (lane 0 walks, no
//                   divergence partners), else `lanes` decoders share a warp.
//   k_jpeg_idct     8 threads per 8x8 block: dequantise, column pass, row pass through shared memory, one 8-byte
//                   store per thread into the component sample plane.
//   k_jpeg_color    one thread per output pixel: triangle-filter chroma upsampling + fixed-point YCbCr->RGB,
//                   written as HWC uint8 at the caller's offsets (= the layout mtgv_set_bg_pool ingests).
// The container (markers, tables) is parsed on the host; file bytes, tables and descriptors go up in one
// staging copy each.
#include "mtgv_internal.cuh"
#include "mtgv_jpeg.cuh"

namespace mtgv {

struct JpegState {
  uint8_t* files = nullptr;   size_t files_cap = 0;
  int16_t* coef = nullptr;    size_t coef_cap = 0;    // int16 elements
  uint8_t* planes = nullptr;  size_t planes_cap = 0;
  uint8_t* desc = nullptr;    size_t desc_cap = 0;    // JpegImg[n] | JpegTables[n] | JpegSeg[nseg]
  uint8_t* desc_host = nullptr; size_t desc_host_cap = 0;  // pinned
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

__global__ void k_jpeg_entropy(const uint8_t* __restrict__ files, const JpegImg* __restrict__ imgs, const JpegTables* __restrict__ tbs,
                               const JpegSeg* __restrict__ segs, int nseg, int stride, int16_t* __restrict__ coef) {
  __shared__ uint8_t zz[64];
  if (threadIdx.x < 64) zz[threadIdx.x] = c_zigzag[threadIdx.x];
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t % stride) return;
  const int s = t / stride;
  if (s >= nseg) return;
  const JpegSeg sg = segs[s];
  const JpegImg im = imgs[sg.img];
  jpeg_decode_segment(files + im.file_off, im, tbs + sg.img, sg, coef, zz);
}

// grid (ceil(max blocks / 32), n images), 256 threads = 32 blocks of 8 threads
__global__ void __launch_bounds__(256) k_jpeg_idct(const JpegImg* __restrict__ imgs, const JpegTables* __restrict__ tbs,
                                                   const int16_t* __restrict__ coef, uint8_t* __restrict__ planes) {
  __shared__ int ws[32][8][9];
  const JpegImg& im = imgs[blockIdx.y];
  const int g = threadIdx.x >> 3, t = threadIdx.x & 7;
  const int b = blockIdx.x * 32 + g;
  const bool live = b < im.nblk;
  int c = 0;
  if (live) {
    while (c + 1 < im.ncomp && b >= im.blk0[c + 1]) c++;
    const int16_t* blk = coef + (im.coef_blk + b) * 64;
    const uint16_t* q = tbs[blockIdx.y].qt[im.tq[c]];
    int x[8], o[8];
#pragma unroll
    for (int r = 0; r < 8; r++) x[r] = (int)blk[r * 8 + t] * (int)q[r * 8 + t];
    jpeg_idct8(x, o, kJpegPass1Shift);  // column t
#pragma unroll
    for (int r = 0; r < 8; r++) ws[g][r][t] = o[r];
  }
  __syncwarp();
  if (live) {
    int x[8], o[8];
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

// grid (ceil(max pixels / 256), n images)
__global__ void __launch_bounds__(256) k_jpeg_color(const JpegImg* __restrict__ imgs, const uint8_t* __restrict__ planes,
                                                    uint8_t* __restrict__ out) {
  const JpegImg& im = imgs[blockIdx.y];
  const int npix = im.h * im.w;
  uint8_t* dst = out + im.out_off;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < npix; i += gridDim.x * 256) {
    const int y = i / im.w, x = i - y * im.w;
    int rgb[3];
    jpeg_pixel(im, planes, y, x, rgb);
    dst[(int64_t)i * 3 + 0] = (uint8_t)rgb[0];
    dst[(int64_t)i * 3 + 1] = (uint8_t)rgb[1];
    dst[(int64_t)i * 3 + 2] = (uint8_t)rgb[2];
  }
}

int jpeg_destroy(mtgv_ctx* ctx) {
  JpegState* st = (JpegState*)ctx->jpeg;
  if (!st) return MTGV_OK;
  cudaFree(st->files); cudaFree(st->coef); cudaFree(st->planes); cudaFree(st->desc);
  if (st->desc_host) cudaFreeHost(st->desc_host);
  delete st;
  ctx->jpeg = nullptr;
  return MTGV_OK;
}

int jpeg_info(mtgv_ctx* ctx, const uint8_t* file, int64_t len, int32_t* hw) {
  JpegImg im;
  JpegTables tb;
  std::vector<JpegSeg> segs;
  std::string err;
  if (jpeg_parse(file, len, 0, &im, &tb, &segs, &err) != 0) return fail(ctx, MTGV_ERR_INVALID, "mtgv_jpeg_info: " + err);
  hw[0] = im.h;
  hw[1] = im.w;
  return MTGV_OK;
}

int jpeg_decode_batch(mtgv_ctx* ctx, const uint8_t* files, const int64_t* file_off, int n, uint8_t* out, const int64_t* out_off,
                      const int32_t* hw, cudaStream_t stream) {
  if (!ctx->jpeg) {
    ctx->jpeg = new JpegState();
    MTGV_CUDA_OK(ctx, cudaMemcpyToSymbol(c_zigzag, kJpegZigzag, 64));
  }
  JpegState* st = (JpegState*)ctx->jpeg;
  std::vector<JpegImg> imgs(n);
  std::vector<JpegTables> tbs(n);
  std::vector<JpegSeg> segs;
  int64_t nblk_total = 0, plane_total = 0;
  int max_blk = 0, max_pix = 0;
  for (int i = 0; i < n; i++) {
    std::string err;
    const int64_t len = file_off[i + 1] - file_off[i];
    if (len < 0 || jpeg_parse(files + file_off[i], len, i, &imgs[i], &tbs[i], &segs, &err) != 0)
      return fail(ctx, MTGV_ERR_INVALID, "mtgv_decode_jpeg_batch: file " + std::to_string(i) + ": " + (len < 0 ? "bad offsets" : err));
    JpegImg& im = imgs[i];
    if (im.h != hw[2 * i] || im.w != hw[2 * i + 1])
      return fail(ctx, MTGV_ERR_INVALID, "mtgv_decode_jpeg_batch: file " + std::to_string(i) + " is " + std::to_string(im.h) + "x" +
                                             std::to_string(im.w) + ", the caller's hw says otherwise");
    im.file_off = file_off[i] - file_off[0];
    im.coef_blk = nblk_total;
    im.out_off = out_off[i];
    for (int c = 0; c < im.ncomp; c++) {
      im.plane_off[c] = plane_total;
      plane_total += (int64_t)im.bw[c] * im.bh[c] * 64;
    }
    nblk_total += im.nblk;
    if (im.nblk > max_blk) max_blk = im.nblk;
    if (im.h * im.w > max_pix) max_pix = im.h * im.w;
  }
  const size_t file_bytes = (size_t)(file_off[n] - file_off[0]);
  const size_t nseg = segs.size();
  const size_t o_tb = sizeof(JpegImg) * n, o_sg = o_tb + sizeof(JpegTables) * n, desc_bytes = o_sg + sizeof(JpegSeg) * nseg;
  int rc;
  if ((rc = grow(ctx, (void**)&st->files, &st->files_cap, file_bytes + 16))) return rc;
  if ((rc = grow(ctx, (void**)&st->coef, &st->coef_cap, (size_t)nblk_total * 64 * sizeof(int16_t)))) return rc;
  if ((rc = grow(ctx, (void**)&st->planes, &st->planes_cap, (size_t)plane_total))) return rc;
  if ((rc = grow(ctx, (void**)&st->desc, &st->desc_cap, desc_bytes))) return rc;
  MTGV_CUDA_OK(ctx, cudaStreamSynchronize(stream));  // an earlier batch may still be reading the staging buffer
  if (desc_bytes > st->desc_host_cap) {
    if (st->desc_host) MTGV_CUDA_OK(ctx, cudaFreeHost(st->desc_host));
    st->desc_host = nullptr; st->desc_host_cap = 0;
    MTGV_CUDA_OK(ctx, cudaMallocHost((void**)&st->desc_host, desc_bytes + desc_bytes / 4));
    st->desc_host_cap = desc_bytes + desc_bytes / 4;
  }
  memcpy(st->desc_host, imgs.data(), o_tb);
  memcpy(st->desc_host + o_tb, tbs.data(), sizeof(JpegTables) * n);
  memcpy(st->desc_host + o_sg, segs.data(), sizeof(JpegSeg) * nseg);
  MTGV_CUDA_OK(ctx, cudaMemcpyAsync(st->desc, st->desc_host, desc_bytes, cudaMemcpyHostToDevice, stream));
  MTGV_CUDA_OK(ctx, cudaMemcpyAsync(st->files, files + file_off[0], file_bytes, cudaMemcpyHostToDevice, stream));
  MTGV_CUDA_OK(ctx, cudaMemsetAsync(st->coef, 0, (size_t)nblk_total * 64 * sizeof(int16_t), stream));
  const JpegImg* d_img = (const JpegImg*)st->desc;
  const JpegTables* d_tb = (const JpegTables*)(st->desc + o_tb);
  const JpegSeg* d_sg = (const JpegSeg*)(st->desc + o_sg);
  // decoders per warp: alone while the intervals fit the machine as whole warps (16 warps per SM), else shared
  int lanes = 1;
  while (lanes < 32 && (nseg + lanes - 1) / lanes > (size_t)ctx->sm_count * 16) lanes *= 2;
  const int stride = 32 / lanes;
  const long long threads = (long long)nseg * stride;
  k_jpeg_entropy<<<(unsigned)((threads + 63) / 64), 64, 0, stream>>>(st->files, d_img, d_tb, d_sg, (int)nseg, stride, st->coef);
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  k_jpeg_idct<<<dim3((max_blk + 31) / 32, n), 256, 0, stream>>>(d_img, d_tb, st->coef, st->planes);
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  const int gx = (max_pix + 255) / 256;
  k_jpeg_color<<<dim3(gx < 1024 ? gx : 1024, n), 256, 0, stream>>>(d_img, st->planes, out);
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 3;
  return MTGV_OK;
}

}  // namespace mtgv
