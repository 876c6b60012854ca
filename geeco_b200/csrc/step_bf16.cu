// bf16 / tcgen05 path of the conv encoders (placeholder until conv_tc.cu lands).
#include "plan.cuh"
int plan_bf16(geeco_ctx*, size_t*, char*) {
  geeco_set_error("precision GEECO_BF16 is not available in this build");
  return GEECO_ERR_INVALID;
}
int encoders_fwd_bf16(geeco_ctx*, cudaStream_t) { geeco_set_error("bf16 path not built"); return GEECO_ERR_INVALID; }
int encoders_bwd_bf16(geeco_ctx*, int, int, cudaStream_t) { geeco_set_error("bf16 path not built"); return GEECO_ERR_INVALID; }
