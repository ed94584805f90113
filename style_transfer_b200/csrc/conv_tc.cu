// Placeholder until the tcgen05 kernel lands: the bf16 mode runs the SIMT convolution.
#include "style_b200.h"
#include "common.cuh"
#include "conv_tc.h"

namespace st {
int tc_init(TcContext& tc, int sm_count) { tc.enabled = false; tc.sm_count = sm_count; return ST_OK; }
void tc_destroy(TcContext&) {}
int tc_pack_weights(TcContext&, TcWeights&, const float*, int, int) { return ST_OK; }
void tc_free_weights(TcWeights&) {}
bool tc_shape_ok(const TcContext& tc, const TcWeights&, int, int) { return tc.enabled; }
int conv3x3_tc(TcContext&, const TcWeights&, const __nv_bfloat16*, __nv_bfloat16*, int, int, int,
               int, bool, const float*, const __nv_bfloat16*, const __nv_bfloat16*, cudaStream_t) {
  set_error("tcgen05 convolution not built");
  return ST_ERR_STATE;
}
}  // namespace st
