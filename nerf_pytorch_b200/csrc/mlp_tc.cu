// mlp_tc.cu -- tcgen05 (5th-gen tensor core) implementation of the fused MLP forward.
// Placeholder until the TMEM/UMMA pipeline lands: reports "unsupported" so that callers fail loudly.
#include "common.cuh"

namespace nerfb200 {

int launch_mlp_fwd_tc(const Plan&, const float*, const float*, int, const float*, int64_t, int, float*, float*,
                      cudaStream_t) {
  set_error("mlp_fwd impl=1 (tcgen05) is not built into this library yet");
  return NERFB200_ERR_UNSUPPORTED;
}

}  // namespace nerfb200
