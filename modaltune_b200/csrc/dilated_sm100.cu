// placeholder until the tcgen05 kernels land
#include "mt_common.cuh"
namespace mt {
int dilated_attn_fwd_sm100(const mt_dilated_geometry*, const void*, int64_t, int64_t, void*, float*, cudaStream_t) {
  set_error("tcgen05 forward not built yet");
  return MT_E_UNSUPPORTED;
}
int dilated_attn_bwd_sm100(const mt_dilated_geometry*, const void*, int64_t, int64_t, const void*, const float*,
                           const float*, float*, cudaStream_t) {
  set_error("tcgen05 backward not built yet");
  return MT_E_UNSUPPORTED;
}
}  // namespace mt
