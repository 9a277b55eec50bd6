#include "flow_tc.cuh"

namespace tw {

bool tc_supported(const tw_flow_config* c) { (void)c; return false; }
void tc_carve(const tw_flow_config*, int64_t, int64_t, int64_t, Arena&, TcScratch* out) { out->packed = nullptr; out->scores_op = nullptr; out->packed_bytes = 0; }
int tc_begin_pass(const tw_flow_config*, const ParamView&, TcScratch&, const float*, const uint8_t*, int64_t, int64_t, int, cudaStream_t) {
  return fail(TW_ERR_UNSUPPORTED, "tensor-core path not built");
}
int tc_conditioner(const tw_flow_config*, const ParamView&, int, TcScratch&, const int64_t*, const float*, const float*, const float*,
                   const float*, float* const*, float* const*, float* const*, int64_t, int64_t, int, cudaStream_t) {
  return fail(TW_ERR_UNSUPPORTED, "tensor-core path not built");
}

}  // namespace tw
