// Softmax / log-softmax forward and backward along one axis of an (outer, len, inner) view.
// Reference: neuronika-variable/src/node/softmax/mod.rs:37-53, 84-104 and
// node/logsoftmax/mod.rs:37-53, 84-102 -- same operation order per lane: max, exp(x - max),
// sum, divide (softmax) / x - ln(sum) - max (log-softmax).
// One warp per lane; lanes are short on the hot path (10 classes in config 4), so the three
// passes hit L1.  inner == 1 gives unit-stride lanes (axis = last).
#include "nk_internal.cuh"

namespace {

constexpr int kThreads = 256;

template <typename T, bool LOG>
__global__ void __launch_bounds__(kThreads) softmax_fwd_kernel(T* __restrict__ y, const T* __restrict__ x,
                                                              int64_t lanes, int64_t len, int64_t inner) {
  const int lane_id = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t l = warp; l < lanes; l += nwarps) {
    const int64_t o = l / inner, i = l - o * inner;
    const T* xp = x + o * len * inner + i;
    T* yp = y + o * len * inner + i;
    float m = -3.402823466e+38f;  // f32::MIN, softmax/mod.rs:41
    for (int64_t k = lane_id; k < len; k += 32) m = fmaxf(m, nk_to_f32<T>(xp[k * inner]));
    m = nk_warp_max(m);
    float s = 0.f;
    for (int64_t k = lane_id; k < len; k += 32) s += expf(nk_to_f32<T>(xp[k * inner]) - m);
    s = nk_warp_sum(s);
    if (LOG) {
      const float lse = logf(s);
      for (int64_t k = lane_id; k < len; k += 32)
        yp[k * inner] = nk_from_f32<T>(nk_to_f32<T>(xp[k * inner]) - lse - m);
    } else {
      for (int64_t k = lane_id; k < len; k += 32)
        yp[k * inner] = nk_from_f32<T>(expf(nk_to_f32<T>(xp[k * inner]) - m) / s);
    }
  }
}

// softmax:     dx += y * (g - sum(g*y))
// log-softmax: dx += g - exp(y) * sum(g)
template <typename T, bool LOG>
__global__ void __launch_bounds__(kThreads) softmax_bwd_kernel(T* __restrict__ dx, const T* __restrict__ y,
                                                              const T* __restrict__ g, int64_t lanes, int64_t len,
                                                              int64_t inner, float beta) {
  const int lane_id = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t l = warp; l < lanes; l += nwarps) {
    const int64_t o = l / inner, i = l - o * inner;
    const int64_t base = o * len * inner + i;
    float s = 0.f;
    for (int64_t k = lane_id; k < len; k += 32) {
      const float gv = nk_to_f32<T>(g[base + k * inner]);
      s += LOG ? gv : gv * nk_to_f32<T>(y[base + k * inner]);
    }
    s = nk_warp_sum(s);
    for (int64_t k = lane_id; k < len; k += 32) {
      const int64_t idx = base + k * inner;
      const float gv = nk_to_f32<T>(g[idx]), yv = nk_to_f32<T>(y[idx]);
      float v = LOG ? gv - expf(yv) * s : yv * (gv - s);
      if (beta != 0.f) v += beta * nk_to_f32<T>(dx[idx]);
      dx[idx] = nk_from_f32<T>(v);
    }
  }
}

template <bool LOG>
int softmax_fwd(nk_ctx* ctx, void* y, const void* x, int64_t outer, int64_t len, int64_t inner, int dtype) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, nk_dtype_ok(dtype), "softmax: bad dtype %d", dtype);
  NK_REQUIRE(ctx, outer >= 0 && len >= 0 && inner >= 0, "softmax: negative size");
  const int64_t lanes = outer * inner;
  if (lanes == 0 || len == 0) return NK_OK;
  NK_REQUIRE(ctx, y && x, "softmax: NULL pointer");
  int64_t blocks = (lanes * 32 + kThreads - 1) / kThreads;
  const int64_t cap = int64_t(ctx->sm_count) * 8;
  if (blocks > cap) blocks = cap;
  if (dtype == NK_BF16)
    softmax_fwd_kernel<__nv_bfloat16, LOG><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>((__nv_bfloat16*)y, (const __nv_bfloat16*)x, lanes, len, inner);
  else
    softmax_fwd_kernel<float, LOG><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>((float*)y, (const float*)x, lanes, len, inner);
  NK_LAUNCHED(ctx, LOG ? "log_softmax_fwd" : "softmax_fwd");
  return NK_OK;
}

template <bool LOG>
int softmax_bwd(nk_ctx* ctx, void* dx, const void* y, const void* g, int64_t outer, int64_t len, int64_t inner,
                int dtype, float beta) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, nk_dtype_ok(dtype), "softmax bwd: bad dtype %d", dtype);
  NK_REQUIRE(ctx, outer >= 0 && len >= 0 && inner >= 0, "softmax bwd: negative size");
  const int64_t lanes = outer * inner;
  if (lanes == 0 || len == 0) return NK_OK;
  NK_REQUIRE(ctx, dx && y && g, "softmax bwd: NULL pointer");
  int64_t blocks = (lanes * 32 + kThreads - 1) / kThreads;
  const int64_t cap = int64_t(ctx->sm_count) * 8;
  if (blocks > cap) blocks = cap;
  if (dtype == NK_BF16)
    softmax_bwd_kernel<__nv_bfloat16, LOG><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>((__nv_bfloat16*)dx, (const __nv_bfloat16*)y, (const __nv_bfloat16*)g, lanes, len, inner, beta);
  else
    softmax_bwd_kernel<float, LOG><<<(unsigned)blocks, kThreads, 0, ctx->stream>>>((float*)dx, (const float*)y, (const float*)g, lanes, len, inner, beta);
  NK_LAUNCHED(ctx, LOG ? "log_softmax_bwd" : "softmax_bwd");
  return NK_OK;
}

}  // namespace

extern "C" {
int nk_softmax_fwd(nk_ctx* ctx, void* y, const void* x, int64_t outer, int64_t len, int64_t inner, int dtype) {
  return softmax_fwd<false>(ctx, y, x, outer, len, inner, dtype);
}
int nk_log_softmax_fwd(nk_ctx* ctx, void* y, const void* x, int64_t outer, int64_t len, int64_t inner, int dtype) {
  return softmax_fwd<true>(ctx, y, x, outer, len, inner, dtype);
}
int nk_softmax_bwd(nk_ctx* ctx, void* dx, const void* y, const void* g, int64_t outer, int64_t len, int64_t inner,
                   int dtype, float beta) {
  return softmax_bwd<false>(ctx, dx, y, g, outer, len, inner, dtype, beta);
}
int nk_log_softmax_bwd(nk_ctx* ctx, void* dx, const void* y, const void* g, int64_t outer, int64_t len, int64_t inner,
                       int dtype, float beta) {
  return softmax_bwd<true>(ctx, dx, y, g, outer, len, inner, dtype, beta);
}
}
