// Stage 1, tensor-core kernel (tcgen05 / TMEM / TMA, 3xTF32).  PLACEHOLDER until
// the kernel lands: reports "unsupported" so sb_itq_hash uses the FFMA kernel.
#include "common.cuh"

namespace sb {

int itq_hash_tc_supported(int64_t, int32_t, int64_t, int32_t, const float*, const float*) { return 0; }

int itq_hash_tc_launch(const float*, int64_t, int32_t, int64_t, const float*, const float*, int32_t, int32_t, float,
                       uint32_t*, int32_t, float*, cudaStream_t) {
  set_error("sb_itq_hash: tensor-core variant not built");
  return SB_ERR_UNSUPPORTED;
}

}  // namespace sb
