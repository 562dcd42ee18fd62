// placeholder until the rotated-basis kernels land
#include "common.cuh"
namespace qcb {
size_t cwf_workspace_bytes(int, int, int64_t) { return 256; }
size_t dm_workspace_bytes(int, int, int, int64_t) { return 256; }
int run_cwf_rotated(const StackedParams&, const StackedParams&, const double*, int64_t, const uint8_t*, const double*, int,
                    double*, double*, double*, void*, size_t, cudaStream_t) {
  return fail(QCB_ERR_UNSUPPORTED, "cwf_rotated: not built yet");
}
int run_dm_rotated(const StackedParams&, const StackedParams&, const double*, int64_t, const uint8_t*, const double*, int,
                   double*, double*, double*, void*, size_t, cudaStream_t) {
  return fail(QCB_ERR_UNSUPPORTED, "dm_rotated: not built yet");
}
}  // namespace qcb
