// mtgv_dewarp.cu - serving-side card dewarp: InstanceSeg.extract_dewarped
// (mtgvision/od_export.py:95-111) = cv2.getPerspectiveTransform(detected quad -> expanded output
// rectangle) + cv2.warpPerspective(uint8 frame, M, (w, h)).  One launch dewarps every detected card of a
// frame.  Bit-exact: the matrix uses the shared fp64 restatement (mtgv_geom.cuh), the coordinates the
// guarded fast path of mtgv_persp.cuh, and the pixels remapBilinear's uint8 fixed point - weights
// 32*(32-ax|ax)*(32-ay|ay) (= saturate_cast<short>(w * 2^15), summing to 2^15), (sum + 2^14) >> 15,
// zero outside the frame.
#include "mtgv_internal.cuh"
#include "mtgv_persp.cuh"

namespace mtgv {

__global__ void k_dewarp_u8(const uint8_t* __restrict__ frame, int fh, int fw, int fc, const float* __restrict__ quads,
                            const float* __restrict__ dst_rect, uint8_t* __restrict__ out, int oh, int ow) {
  __shared__ double Mi[9];
  __shared__ int ok;
  const int card = blockIdx.y;
  if (threadIdx.x == 0) {
    double M[9];
    ok = get_perspective_transform(quads + (size_t)card * 8, dst_rect, M) ? 1 : 0;
    invert3x3(M, Mi);  // cv::invert inside cv::warpPerspective
  }
  __syncthreads();
  const int bw0 = persp_block_w(oh, ow);
  uint8_t* dst = out + (size_t)card * oh * ow * fc;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < oh * ow; i += gridDim.x * blockDim.x) {
    const int y = i / ow, x = i - y * ow;
    const int bxi = (x / bw0) * bw0;
    double o[3];
    persp_origin(Mi, (double)bxi, (double)y, o);
    const int2 XY = persp_xy(o[0], o[1], o[2], Mi[0], Mi[3], Mi[6], (double)(x - bxi));
    const int sx = sat_short(XY.x >> 5), sy = sat_short(XY.y >> 5), ax = XY.x & 31, ay = XY.y & 31;
    const int w00 = 32 * (32 - ax) * (32 - ay), w01 = 32 * ax * (32 - ay), w10 = 32 * (32 - ax) * ay, w11 = 32 * ax * ay;
    const bool x0 = (unsigned)sx < (unsigned)fw, x1 = (unsigned)(sx + 1) < (unsigned)fw;
    const bool y0 = (unsigned)sy < (unsigned)fh, y1 = (unsigned)(sy + 1) < (unsigned)fh;
    const uint8_t* p = frame + ((size_t)sy * fw + sx) * fc;
    for (int c = 0; c < fc; c++) {
      int acc = 1 << 14;
      if (y0 && x0) acc += w00 * (int)__ldg(p + c);
      if (y0 && x1) acc += w01 * (int)__ldg(p + fc + c);
      if (y1 && x0) acc += w10 * (int)__ldg(p + (size_t)fw * fc + c);
      if (y1 && x1) acc += w11 * (int)__ldg(p + (size_t)fw * fc + fc + c);
      dst[(size_t)i * fc + c] = ok ? (uint8_t)(acc >> 15) : 0;
    }
  }
}

int dewarp_u8(mtgv_ctx* ctx, const uint8_t* frame, int fh, int fw, int fc, const float* quads, int n, const float* dst_rect,
              uint8_t* out, int oh, int ow, cudaStream_t st) {
  if (n <= 0) return MTGV_OK;
  const int per = (oh * ow + 255) / 256;
  dim3 grid(per < 64 ? per : 64, n);
  k_dewarp_u8<<<grid, 256, 0, st>>>(frame, fh, fw, fc, quads, dst_rect, out, oh, ow);
  ctx->launches++;
  MTGV_CUDA_OK(ctx, cudaGetLastError());
  return MTGV_OK;
}

}  // namespace mtgv
