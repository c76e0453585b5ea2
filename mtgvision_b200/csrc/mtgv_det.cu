// mtgv_det.cu - detection-scene generator kernels (placement rejection sampler, scene
// compositor, photometrics).  Filled in after the encoder path; see DESIGN.md.
#include "mtgv_internal.cuh"

namespace mtgv {
int det_destroy(mtgv_ctx* ctx) {
  (void)ctx;
  return MTGV_OK;
}
}  // namespace mtgv
