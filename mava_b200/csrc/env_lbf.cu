// Level-Based Foraging env-step (placeholder until the LBF kernel lands: every entry point
// reports MAVA_E_UNSUPPORTED so callers fail loudly instead of silently falling back).
#include "env.cuh"

namespace mava {

int lbf_create(const mava_lbf_config*, mava_env_s*) { return MAVA_E_UNSUPPORTED; }
int lbf_reset(const mava_env_s*, const uint32_t*, uint8_t*, int8_t*, uint8_t*, int, cudaStream_t) {
  return MAVA_E_UNSUPPORTED;
}
int lbf_step(const mava_env_s*, uint8_t*, const int8_t*, int8_t*, uint8_t*, float*, uint8_t*,
             float*, int32_t*, int, int, cudaStream_t) {
  return MAVA_E_UNSUPPORTED;
}
int lbf_peek(const mava_env_s*, const uint8_t*, int, int32_t*, int, cudaStream_t) {
  return MAVA_E_UNSUPPORTED;
}

}  // namespace mava
