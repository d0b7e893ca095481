// beam.cu -- batched beam search (placeholder until the device-side search lands).
#include "common.cuh"
#include "kernels.cuh"

namespace capdec {
size_t beam_workspace_bytes(const CapdecDims&, int, int, int) { return 0; }
int beam_search(const CapdecDims&, const CapdecParams&, const float*, const float*, int, int, int, int32_t,
                int32_t, int32_t*, int32_t*, float*, int32_t*, float*, int32_t*, int32_t*, float*, void*,
                size_t, cudaStream_t) {
  set_error("capdec_beam_search: not built yet");
  return CAPDEC_ERR_UNSUPPORTED;
}
}  // namespace capdec
