// gemm_tc.cu — tcgen05 / TMEM / TMA implicit-GEMM kernels (placeholder: not yet eligible for any shape)
#include "common.cuh"
extern "C" int mopoe_tc_available(void) { return 0; }
int mopoe_tc_fwd_eligible(const mopoe_window_t*, const mopoe_rows_t*) { return 0; }
int mopoe_conv_gemm_tc(const mopoe_window_t*, const void*, const float*, const mopoe_rows_t*, void*) { MOPOE_FAIL("tc: not built"); }
int mopoe_tc_wgrad_eligible(const mopoe_window_t*, const mopoe_rows_t*) { return 0; }
size_t mopoe_conv_wgrad_ws_tc(const mopoe_window_t*, const mopoe_rows_t*) { return 0; }
int mopoe_conv_wgrad_tc(const mopoe_window_t*, const mopoe_rows_t*, float*, int, void*, size_t, void*) { MOPOE_FAIL("tc: not built"); }
