// 1-D and 3-D convolution: the reference's convolution is generic over the number of sample
// dimensions (convolution/mod.rs:85-123 fwd, 146-189 dX, 191-226 dW, grouped :125-144, 228-294; goldens
// convolution/test.rs:144-239 (1-D), 306-444 (3-D) and their strided / dilated / grouped siblings).  The 2-D case has
// its own engines (nk_conv_tc.cu, nk_conv_direct.cu); this file is the CUDA-core gather engine for x (N, C, s0, s1, s2)
// with leading sample dims of extent 1 when there are fewer than three.  Same un-padded cross-correlation, same
// accumulate protocol (beta) and argument checks (utils.rs:427-496) as the 2-D entry points.
#include "nk_internal.cuh"

namespace {

constexpr int kThreads = 256;

struct NdDims {
  int64_t n, cin, cout, groups;
  int64_t in[3], k[3], s[3], d[3], out[3];
};

template <typename T>
__global__ void __launch_bounds__(kThreads) convnd_fwd_kernel(T* __restrict__ y, const T* __restrict__ x,
                                                              const T* __restrict__ wt, NdDims d) {
  const int64_t L = d.out[0] * d.out[1] * d.out[2];
  const int64_t total = d.n * d.cout * L;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  const int64_t cin_g = d.cin / d.groups, cout_g = d.cout / d.groups;
  const int64_t isz = d.in[0] * d.in[1] * d.in[2], ksz = d.k[0] * d.k[1] * d.k[2];
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    int64_t rem = idx;
    const int64_t p2 = rem % d.out[2];
    rem /= d.out[2];
    const int64_t p1 = rem % d.out[1];
    rem /= d.out[1];
    const int64_t p0 = rem % d.out[0];
    rem /= d.out[0];
    const int64_t o = rem % d.cout, n = rem / d.cout;
    const int64_t g = o / cout_g;
    float acc = 0.f;
    for (int64_t c = 0; c < cin_g; ++c) {
      const T* xp = x + (n * d.cin + g * cin_g + c) * isz;
      const T* wp = wt + (o * cin_g + c) * ksz;
      for (int64_t i0 = 0; i0 < d.k[0]; ++i0)
        for (int64_t i1 = 0; i1 < d.k[1]; ++i1)
          for (int64_t i2 = 0; i2 < d.k[2]; ++i2) {
            const int64_t u0 = p0 * d.s[0] + i0 * d.d[0], u1 = p1 * d.s[1] + i1 * d.d[1], u2 = p2 * d.s[2] + i2 * d.d[2];
            acc = fmaf(nk_to_f32<T>(wp[(i0 * d.k[1] + i1) * d.k[2] + i2]),
                       nk_to_f32<T>(xp[(u0 * d.in[1] + u1) * d.in[2] + u2]), acc);
          }
    }
    y[idx] = nk_from_f32<T>(acc);
  }
}

// output position along one axis that reads input coordinate u through tap i, or -1
__device__ __forceinline__ int64_t out_pos(int64_t u, int64_t i, int64_t s, int64_t dil, int64_t out) {
  const int64_t pu = u - i * dil;
  if (pu < 0 || pu % s != 0) return -1;
  const int64_t p = pu / s;
  return p < out ? p : -1;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) convnd_bwd_input_kernel(T* __restrict__ dx, const T* __restrict__ g,
                                                                    const T* __restrict__ wt, NdDims d, float beta) {
  const int64_t isz = d.in[0] * d.in[1] * d.in[2], L = d.out[0] * d.out[1] * d.out[2], ksz = d.k[0] * d.k[1] * d.k[2];
  const int64_t total = d.n * d.cin * isz;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  const int64_t cin_g = d.cin / d.groups, cout_g = d.cout / d.groups;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    int64_t rem = idx;
    const int64_t u2 = rem % d.in[2];
    rem /= d.in[2];
    const int64_t u1 = rem % d.in[1];
    rem /= d.in[1];
    const int64_t u0 = rem % d.in[0];
    rem /= d.in[0];
    const int64_t c = rem % d.cin, n = rem / d.cin;
    const int64_t grp = c / cin_g, cl = c - grp * cin_g;
    float acc = 0.f;
    for (int64_t i0 = 0; i0 < d.k[0]; ++i0) {
      const int64_t p0 = out_pos(u0, i0, d.s[0], d.d[0], d.out[0]);
      if (p0 < 0) continue;
      for (int64_t i1 = 0; i1 < d.k[1]; ++i1) {
        const int64_t p1 = out_pos(u1, i1, d.s[1], d.d[1], d.out[1]);
        if (p1 < 0) continue;
        for (int64_t i2 = 0; i2 < d.k[2]; ++i2) {
          const int64_t p2 = out_pos(u2, i2, d.s[2], d.d[2], d.out[2]);
          if (p2 < 0) continue;
          const int64_t l = (p0 * d.out[1] + p1) * d.out[2] + p2, kidx = (i0 * d.k[1] + i1) * d.k[2] + i2;
          for (int64_t ol = 0; ol < cout_g; ++ol) {
            const int64_t o = grp * cout_g + ol;
            acc = fmaf(nk_to_f32<T>(g[(n * d.cout + o) * L + l]), nk_to_f32<T>(wt[(o * cin_g + cl) * ksz + kidx]), acc);
          }
        }
      }
    }
    if (beta != 0.f) acc += beta * nk_to_f32<T>(dx[idx]);
    dx[idx] = nk_from_f32<T>(acc);
  }
}

// one block per kernel element (x a chunk of the batch): reduce over (n, output positions), f32 atomics into scratch
template <typename T>
__global__ void __launch_bounds__(kThreads) convnd_bwd_kernel_kernel(float* __restrict__ scratch, const T* __restrict__ g,
                                                                     const T* __restrict__ x, NdDims d,
                                                                     int64_t n_per_block) {
  const int64_t cin_g = d.cin / d.groups, cout_g = d.cout / d.groups;
  const int64_t isz = d.in[0] * d.in[1] * d.in[2], L = d.out[0] * d.out[1] * d.out[2], ksz = d.k[0] * d.k[1] * d.k[2];
  const int64_t widx = blockIdx.x;
  int64_t rem = widx;
  const int64_t kidx = rem % ksz;
  rem /= ksz;
  const int64_t c = rem % cin_g, o = rem / cin_g;
  const int64_t i2 = kidx % d.k[2], i1 = (kidx / d.k[2]) % d.k[1], i0 = kidx / (d.k[2] * d.k[1]);
  const int64_t grp = o / cout_g;
  const int64_t n_begin = int64_t(blockIdx.y) * n_per_block;
  int64_t n_end = n_begin + n_per_block;
  if (n_end > d.n) n_end = d.n;
  float acc = 0.f;
  for (int64_t n = n_begin; n < n_end; ++n) {
    const T* gp = g + (n * d.cout + o) * L;
    const T* xp = x + (n * d.cin + grp * cin_g + c) * isz;
    for (int64_t l = threadIdx.x; l < L; l += blockDim.x) {
      const int64_t p2 = l % d.out[2], p1 = (l / d.out[2]) % d.out[1], p0 = l / (d.out[2] * d.out[1]);
      const int64_t u0 = p0 * d.s[0] + i0 * d.d[0], u1 = p1 * d.s[1] + i1 * d.d[1], u2 = p2 * d.s[2] + i2 * d.d[2];
      acc = fmaf(nk_to_f32<T>(gp[l]), nk_to_f32<T>(xp[(u0 * d.in[1] + u1) * d.in[2] + u2]), acc);
    }
  }
  acc = nk_warp_sum(acc);
  __shared__ float sm[kThreads / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < kThreads / 32; ++k) s += sm[k];
    atomicAdd(&scratch[widx], s);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) finalize_dwnd(T* __restrict__ dst, const float* __restrict__ scratch,
                                                          int64_t n, float beta) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = scratch[i];
  if (beta != 0.f) v += beta * nk_to_f32<T>(dst[i]);
  dst[i] = nk_from_f32<T>(v);
}

int make_dims(nk_ctx* ctx, const char* who, int nsp, int64_t n, int64_t cin, const int64_t* in_sp, int64_t cout,
              const int64_t* k, const int64_t* s, const int64_t* dil, int64_t groups, NdDims* d) {
  NK_REQUIRE(ctx, nsp >= 1 && nsp <= 3, "%s: 1 to 3 sample dimensions (got %d)", who, nsp);
  NK_REQUIRE(ctx, in_sp && k && s && dil, "%s: NULL shape pointer", who);
  NK_REQUIRE(ctx, n >= 0 && cin > 0 && cout > 0, "%s: bad sizes", who);
  NK_REQUIRE(ctx, groups >= 1, "%s: groups must be >= 1", who);
  NK_REQUIRE(ctx, cin % groups == 0, "In channels %lld is not divisible by groups %lld", (long long)cin, (long long)groups);
  NK_REQUIRE(ctx, cout % groups == 0, "Out channels %lld is not divisible by groups %lld", (long long)cout, (long long)groups);
  d->n = n, d->cin = cin, d->cout = cout, d->groups = groups;
  for (int a = 0; a < 3; ++a) d->in[a] = d->k[a] = d->s[a] = d->d[a] = d->out[a] = 1;
  for (int a = 0; a < nsp; ++a) {
    const int j = 3 - nsp + a;
    NK_REQUIRE(ctx, k[a] > 0 && s[a] > 0 && dil[a] > 0, "%s: kernel, stride and dilation must be positive", who);
    NK_REQUIRE(ctx, in_sp[a] >= (k[a] - 1) * dil[a] + 1, "The kernel size can't be greater than actual input size.");
    d->in[j] = in_sp[a], d->k[j] = k[a], d->s[j] = s[a], d->d[j] = dil[a];
    d->out[j] = (in_sp[a] - dil[a] * (k[a] - 1) - 1) / s[a] + 1;  // conv_out_shape, utils.rs:207-237
  }
  return NK_OK;
}

inline int nd_blocks(nk_ctx* ctx, int64_t total) {
  int64_t b = (total + kThreads - 1) / kThreads;
  const int64_t cap = int64_t(ctx->sm_count) * 16;
  if (b > cap) b = cap;
  return int(b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

int nk_convnd_fwd(nk_ctx* ctx, void* y, const void* x, const void* w, int nsp, int64_t n, int64_t cin,
                  const int64_t* in_sp, int64_t cout, const int64_t* k, const int64_t* stride, const int64_t* dilation,
                  int64_t groups, int dtype) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, nk_dtype_ok(dtype), "nk_convnd_fwd: bad dtype %d", dtype);
  NdDims d;
  int rc = make_dims(ctx, "nk_convnd_fwd", nsp, n, cin, in_sp, cout, k, stride, dilation, groups, &d);
  if (rc) return rc;
  const int64_t total = d.n * d.cout * d.out[0] * d.out[1] * d.out[2];
  if (total == 0) return NK_OK;
  NK_REQUIRE(ctx, y && x && w, "nk_convnd_fwd: NULL pointer");
  const int blocks = nd_blocks(ctx, total);
  if (dtype == NK_BF16)
    convnd_fwd_kernel<__nv_bfloat16><<<blocks, kThreads, 0, ctx->stream>>>((__nv_bfloat16*)y, (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, d);
  else
    convnd_fwd_kernel<float><<<blocks, kThreads, 0, ctx->stream>>>((float*)y, (const float*)x, (const float*)w, d);
  NK_LAUNCHED(ctx, "convnd_fwd");
  ctx->last_conv_kernel = "direct_nd_fwd";
  return NK_OK;
}

int nk_convnd_bwd_input(nk_ctx* ctx, void* dx, const void* g, const void* w, int nsp, int64_t n, int64_t cin,
                        const int64_t* in_sp, int64_t cout, const int64_t* k, const int64_t* stride,
                        const int64_t* dilation, int64_t groups, int dtype, float beta) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, nk_dtype_ok(dtype), "nk_convnd_bwd_input: bad dtype %d", dtype);
  NdDims d;
  int rc = make_dims(ctx, "nk_convnd_bwd_input", nsp, n, cin, in_sp, cout, k, stride, dilation, groups, &d);
  if (rc) return rc;
  const int64_t total = d.n * d.cin * d.in[0] * d.in[1] * d.in[2];
  if (total == 0) return NK_OK;
  NK_REQUIRE(ctx, dx && g && w, "nk_convnd_bwd_input: NULL pointer");
  const int blocks = nd_blocks(ctx, total);
  if (dtype == NK_BF16)
    convnd_bwd_input_kernel<__nv_bfloat16><<<blocks, kThreads, 0, ctx->stream>>>((__nv_bfloat16*)dx, (const __nv_bfloat16*)g, (const __nv_bfloat16*)w, d, beta);
  else
    convnd_bwd_input_kernel<float><<<blocks, kThreads, 0, ctx->stream>>>((float*)dx, (const float*)g, (const float*)w, d, beta);
  NK_LAUNCHED(ctx, "convnd_bwd_input");
  ctx->last_conv_kernel = "direct_nd_dx";
  return NK_OK;
}

int nk_convnd_bwd_kernel(nk_ctx* ctx, void* dwt, int dw_dtype, const void* g, const void* x, int nsp, int64_t n,
                         int64_t cin, const int64_t* in_sp, int64_t cout, const int64_t* k, const int64_t* stride,
                         const int64_t* dilation, int64_t groups, int dtype, float beta) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, nk_dtype_ok(dtype) && nk_dtype_ok(dw_dtype), "nk_convnd_bwd_kernel: bad dtype");
  NdDims d;
  int rc = make_dims(ctx, "nk_convnd_bwd_kernel", nsp, n, cin, in_sp, cout, k, stride, dilation, groups, &d);
  if (rc) return rc;
  const int64_t welems = d.cout * (d.cin / d.groups) * d.k[0] * d.k[1] * d.k[2];
  NK_REQUIRE(ctx, dwt && (d.n == 0 || (g && x)), "nk_convnd_bwd_kernel: NULL pointer");
  float* scratch;
  rc = nk_workspace(ctx, size_t(welems) * sizeof(float), (void**)&scratch);
  if (rc) return rc;
  NK_CUDA(ctx, cudaMemsetAsync(scratch, 0, size_t(welems) * sizeof(float), ctx->stream));
  if (d.n > 0) {
    int64_t want_y = (int64_t(ctx->sm_count) * 4 + welems - 1) / welems;
    if (want_y > d.n) want_y = d.n;
    if (want_y < 1) want_y = 1;
    if (want_y > 65535) want_y = 65535;
    const int64_t n_per_block = (d.n + want_y - 1) / want_y;
    const int64_t gy = (d.n + n_per_block - 1) / n_per_block;
    dim3 grid((unsigned)welems, (unsigned)gy);
    if (dtype == NK_BF16)
      convnd_bwd_kernel_kernel<__nv_bfloat16><<<grid, kThreads, 0, ctx->stream>>>(scratch, (const __nv_bfloat16*)g, (const __nv_bfloat16*)x, d, n_per_block);
    else
      convnd_bwd_kernel_kernel<float><<<grid, kThreads, 0, ctx->stream>>>(scratch, (const float*)g, (const float*)x, d, n_per_block);
    NK_LAUNCHED(ctx, "convnd_bwd_kernel");
  }
  const int fb = int((welems + kThreads - 1) / kThreads);
  if (dw_dtype == NK_BF16)
    finalize_dwnd<__nv_bfloat16><<<fb, kThreads, 0, ctx->stream>>>((__nv_bfloat16*)dwt, scratch, welems, beta);
  else
    finalize_dwnd<float><<<fb, kThreads, 0, ctx->stream>>>((float*)dwt, scratch, welems, beta);
  NK_LAUNCHED(ctx, "convnd_finalize");
  ctx->last_conv_kernel = "direct_nd_dw";
  return NK_OK;
}

}  // extern "C"
