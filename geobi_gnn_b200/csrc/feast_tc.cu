// bf16 tensor-core (tcgen05 / TMEM) projection path for FeaSt conv and the FC head.
// Placeholder until the tcgen05 kernels land: reports a clear error instead of falling back.
#include "common.cuh"

namespace geobi {
size_t feast_fwd_tc_ws_bytes(int64_t, int, int) { return 256; }
int feast_fwd_tc(const float*, int64_t, int64_t, int, const int32_t*, const int32_t*, const float*, const float*, const float*, const float*, int,
                 float, float*, int64_t, void*, size_t, cudaStream_t) {
  set_error("feast_fwd: GEOBI_PREC_BF16 path is not built in this version of libgeobi");
  return GEOBI_ERR_INVALID;
}
int fc_head_fwd_tc(const float*, int64_t, int64_t, int, const float*, const float*, int, const float*, const float*, int, int, const float*,
                   int64_t, const float*, int64_t, float*, int64_t, cudaStream_t) {
  set_error("fc_head_fwd: GEOBI_PREC_BF16 path is not built in this version of libgeobi");
  return GEOBI_ERR_INVALID;
}
}  // namespace geobi
