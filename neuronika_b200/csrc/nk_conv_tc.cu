// Tensor-core implicit-GEMM convolution engine (stride 1, dilation 1, groups 1, bf16).
// Placeholder entry points: return NK_ERR_UNSUPPORTED so callers use the direct kernels.
#include "nk_internal.cuh"

int nk_conv2d_fwd_tc(nk_ctx*, void*, const void*, const void*, const void*, int, int64_t, int64_t, int64_t, int64_t,
                     int64_t, int64_t, int64_t) {
  return NK_ERR_UNSUPPORTED;
}
int nk_conv2d_bwd_kernel_tc(nk_ctx*, void*, int, void*, const void*, const void*, int64_t, int64_t, int64_t, int64_t,
                            int64_t, int64_t, int64_t, float) {
  return NK_ERR_UNSUPPORTED;
}
int nk_conv2d_bwd_input_tc(nk_ctx*, void*, const void*, const void*, int64_t, int64_t, int64_t, int64_t, int64_t,
                           int64_t, int64_t, float) {
  return NK_ERR_UNSUPPORTED;
}
