// K4 tensor-core path (tcgen05 / TMEM) -- placeholder until the kernel lands; the fp32 SIMT
// path in sage_gemm.cu is the parity mode.
#include "common.cuh"

int gs_sage_gemm_fwd_tc(const float*, int64_t, const int32_t*, const float*, int64_t, int32_t, const float*, int64_t,
                        int32_t, int32_t, const int32_t*, int32_t, float*, int64_t, int32_t, int32_t, gs_stream_t) {
  return GS_ERR_UNSUPPORTED;
}
