// Matrix-vector, vector-matrix and vector-vector products: HBM-bound streaming kernels, one
// pass over the matrix.
//   mv  y = A.v        matrix_vector_mul/mod.rs:32-40;  dA += g (x) v  :64-69;  dv += A^T.g  :93-101
//   vm  y = v.A        vector_matrix_mul/mod.rs:32-40;  dv += A.g      :64-72;  dA += v (x) g :96-101
//   vv  s = <l, r>     vector_vector_mul/mod.rs:32-34;  dl += r*g, dr += l*g (g a 0-d tensor) :58-63
// A is (rows, cols) row-major.  nk_gemv: trans = 0 -> y[rows] = A.x[cols] (one warp per row, 16-byte loads, shuffle
// reduction); trans = 1 -> y[cols] = A^T.x[rows] (threads own columns, coalesced rows, row blocks combined through
// an f32 workspace with atomics).  Algorithmic bytes = the matrix once.
#include "nk_internal.cuh"

namespace {

constexpr int kThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kThreads) gemv_n_kernel(void* __restrict__ y, int y_bf16, const T* __restrict__ A,
                                                          const T* __restrict__ x, int64_t rows, int64_t cols,
                                                          float beta, bool vec) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = int64_t(gridDim.x) * (kThreads / 32);
  for (int64_t r = int64_t(blockIdx.x) * (kThreads / 32) + (threadIdx.x >> 5); r < rows; r += warps) {
    const T* row = A + r * cols;
    float acc = 0.f;
    int64_t done = 0;
    if (vec) {
      constexpr int V = NkVec<T>::N;
      const int64_t nv = cols / V;
      for (int64_t v = lane; v < nv; v += 32) {
        NkVec<T> a, b;
        a.load(row + v * V);
        b.load(x + v * V);
#pragma unroll
        for (int i = 0; i < V; ++i) acc = fmaf(a.get(i), b.get(i), acc);
      }
      done = nv * V;
    }
    for (int64_t c = done + lane; c < cols; c += 32) acc = fmaf(nk_to_f32<T>(row[c]), nk_to_f32<T>(x[c]), acc);
    acc = nk_warp_sum(acc);
    if (lane == 0) {
      if (y_bf16) {
        __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(y) + r;
        *yp = __float2bfloat16_rn(beta != 0.f ? beta * __bfloat162float(*yp) + acc : acc);
      } else {
        float* yp = static_cast<float*>(y) + r;
        *yp = beta != 0.f ? beta * (*yp) + acc : acc;
      }
    }
  }
}

// scratch[c] += sum over this block's rows of A[r][c] * x[r]
template <typename T>
__global__ void __launch_bounds__(kThreads) gemv_t_kernel(float* __restrict__ scratch, const T* __restrict__ A,
                                                          const T* __restrict__ x, int64_t rows, int64_t cols,
                                                          int64_t rows_per_block) {
  const int64_t c = int64_t(blockIdx.x) * kThreads + threadIdx.x;
  const int64_t r0 = int64_t(blockIdx.y) * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  if (c >= cols) return;
  float acc = 0.f;
  for (int64_t r = r0; r < r1; ++r) acc = fmaf(nk_to_f32<T>(A[r * cols + c]), nk_to_f32<T>(x[r]), acc);
  atomicAdd(&scratch[c], acc);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) finalize_vec(T* __restrict__ dst, const float* __restrict__ scratch, int64_t n,
                                                         float beta) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = scratch[i];
  if (beta != 0.f) v += beta * nk_to_f32<T>(dst[i]);
  dst[i] = nk_from_f32<T>(v);
}

// A[r][c] = beta*A[r][c] + u[r]*v[c]
template <typename TD, typename T>
__global__ void __launch_bounds__(kThreads) outer_kernel(TD* __restrict__ A, const T* __restrict__ u, const T* __restrict__ v,
                                                         int64_t rows, int64_t cols, float beta) {
  const int64_t n = rows * cols;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = i / cols, c = i - r * cols;
    float val = nk_to_f32<T>(u[r]) * nk_to_f32<T>(v[c]);
    if (beta != 0.f) val += beta * nk_to_f32<TD>(A[i]);
    A[i] = nk_from_f32<TD>(val);
  }
}

// dot product: per-block f64 partials, then one warp (deterministic order)
template <typename T>
__global__ void __launch_bounds__(kThreads) dot_stage1(double* __restrict__ partials, const T* __restrict__ a,
                                                       const T* __restrict__ b, size_t n) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  float acc = 0.f;
  double dacc = 0.0;
  int cnt = 0;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    acc = fmaf(nk_to_f32<T>(a[i]), nk_to_f32<T>(b[i]), acc);
    if (++cnt == 64) {
      dacc += double(acc);
      acc = 0.f;
      cnt = 0;
    }
  }
  dacc += double(acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
  __shared__ double sm[kThreads / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = dacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < kThreads / 32; ++i) s += sm[i];
    partials[blockIdx.x] = s;
  }
}
__global__ void dot_stage2(float* __restrict__ out, const double* __restrict__ partials, int nparts) {
  double s = 0.0;
  for (int i = threadIdx.x; i < nparts; i += 32) s += partials[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) *out = float(s);
}

// dst = beta*dst + x * (*s)
template <typename TD, typename T>
__global__ void __launch_bounds__(kThreads) scale_acc_kernel(TD* __restrict__ dst, const T* __restrict__ x,
                                                             const float* __restrict__ s, size_t n, float beta) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  const float sv = *s;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = nk_to_f32<T>(x[i]) * sv;
    if (beta != 0.f) v += beta * nk_to_f32<TD>(dst[i]);
    dst[i] = nk_from_f32<TD>(v);
  }
}

inline int gv_blocks(nk_ctx* ctx, size_t items) {
  size_t b = (items + kThreads - 1) / kThreads, cap = size_t(ctx->sm_count) * 8;
  if (b > cap) b = cap;
  return int(b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

int nk_gemv(nk_ctx* ctx, int trans, int64_t rows, int64_t cols, const void* A, const void* x, float beta, void* y,
            int ax_dtype, int y_dtype) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, nk_dtype_ok(ax_dtype) && nk_dtype_ok(y_dtype), "nk_gemv: bad dtype");
  NK_REQUIRE(ctx, rows >= 0 && cols >= 0, "nk_gemv: negative dimension");
  const int64_t ylen = trans ? cols : rows;
  if (ylen == 0) return NK_OK;
  NK_REQUIRE(ctx, y && (rows * cols == 0 || (A && x)), "nk_gemv: NULL pointer");
  if (!trans) {
    const bool vec = (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                     cols % (ax_dtype == NK_BF16 ? 8 : 4) == 0;
    const int blocks = gv_blocks(ctx, size_t(rows) * 32);
    if (ax_dtype == NK_BF16)
      gemv_n_kernel<__nv_bfloat16><<<blocks, kThreads, 0, ctx->stream>>>(y, y_dtype == NK_BF16, (const __nv_bfloat16*)A, (const __nv_bfloat16*)x, rows, cols, beta, vec);
    else
      gemv_n_kernel<float><<<blocks, kThreads, 0, ctx->stream>>>(y, y_dtype == NK_BF16, (const float*)A, (const float*)x, rows, cols, beta, vec);
    NK_LAUNCHED(ctx, "gemv_n");
    return NK_OK;
  }
  float* scratch;
  int rc = nk_workspace(ctx, size_t(cols) * sizeof(float), (void**)&scratch);
  if (rc) return rc;
  NK_CUDA(ctx, cudaMemsetAsync(scratch, 0, size_t(cols) * sizeof(float), ctx->stream));
  if (rows > 0) {
    const int64_t col_blocks = (cols + kThreads - 1) / kThreads;
    int64_t want_y = (int64_t(ctx->sm_count) * 8 + col_blocks - 1) / col_blocks;
    int64_t rows_per_block = (rows + want_y - 1) / want_y;
    if (rows_per_block < 32) rows_per_block = 32;
    const int64_t gy = (rows + rows_per_block - 1) / rows_per_block;
    NK_REQUIRE(ctx, gy <= 65535, "nk_gemv: grid too large");
    dim3 grid((unsigned)col_blocks, (unsigned)gy);
    if (ax_dtype == NK_BF16)
      gemv_t_kernel<__nv_bfloat16><<<grid, kThreads, 0, ctx->stream>>>(scratch, (const __nv_bfloat16*)A, (const __nv_bfloat16*)x, rows, cols, rows_per_block);
    else
      gemv_t_kernel<float><<<grid, kThreads, 0, ctx->stream>>>(scratch, (const float*)A, (const float*)x, rows, cols, rows_per_block);
    NK_LAUNCHED(ctx, "gemv_t");
  }
  const int fb = int((cols + kThreads - 1) / kThreads);
  if (y_dtype == NK_BF16)
    finalize_vec<__nv_bfloat16><<<fb, kThreads, 0, ctx->stream>>>((__nv_bfloat16*)y, scratch, cols, beta);
  else
    finalize_vec<float><<<fb, kThreads, 0, ctx->stream>>>((float*)y, scratch, cols, beta);
  NK_LAUNCHED(ctx, "gemv_finalize");
  return NK_OK;
}

int nk_outer_acc(nk_ctx* ctx, void* A, int a_dtype, const void* u, const void* v, int64_t rows, int64_t cols,
                 int uv_dtype, float beta) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, nk_dtype_ok(a_dtype) && nk_dtype_ok(uv_dtype), "nk_outer_acc: bad dtype");
  NK_REQUIRE(ctx, rows >= 0 && cols >= 0, "nk_outer_acc: negative dimension");
  if (rows * cols == 0) return NK_OK;
  NK_REQUIRE(ctx, A && u && v, "nk_outer_acc: NULL pointer");
  const int blocks = gv_blocks(ctx, size_t(rows) * size_t(cols));
  if (a_dtype == NK_F32 && uv_dtype == NK_F32)
    outer_kernel<float, float><<<blocks, kThreads, 0, ctx->stream>>>((float*)A, (const float*)u, (const float*)v, rows, cols, beta);
  else if (a_dtype == NK_BF16 && uv_dtype == NK_BF16)
    outer_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, kThreads, 0, ctx->stream>>>((__nv_bfloat16*)A, (const __nv_bfloat16*)u, (const __nv_bfloat16*)v, rows, cols, beta);
  else if (a_dtype == NK_F32)
    outer_kernel<float, __nv_bfloat16><<<blocks, kThreads, 0, ctx->stream>>>((float*)A, (const __nv_bfloat16*)u, (const __nv_bfloat16*)v, rows, cols, beta);
  else
    outer_kernel<__nv_bfloat16, float><<<blocks, kThreads, 0, ctx->stream>>>((__nv_bfloat16*)A, (const float*)u, (const float*)v, rows, cols, beta);
  NK_LAUNCHED(ctx, "outer_acc");
  return NK_OK;
}

int nk_dot(nk_ctx* ctx, float* out, const void* a, const void* b, size_t n, int dtype) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, nk_dtype_ok(dtype), "nk_dot: bad dtype %d", dtype);
  NK_REQUIRE(ctx, out && (n == 0 || (a && b)), "nk_dot: NULL pointer");
  const int blocks = gv_blocks(ctx, n ? n : 1);
  double* partials;
  int rc = nk_workspace(ctx, size_t(blocks) * sizeof(double), (void**)&partials);
  if (rc) return rc;
  if (dtype == NK_BF16)
    dot_stage1<__nv_bfloat16><<<blocks, kThreads, 0, ctx->stream>>>(partials, (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, n);
  else
    dot_stage1<float><<<blocks, kThreads, 0, ctx->stream>>>(partials, (const float*)a, (const float*)b, n);
  NK_LAUNCHED(ctx, "dot_stage1");
  dot_stage2<<<1, 32, 0, ctx->stream>>>(out, partials, blocks);
  NK_LAUNCHED(ctx, "dot_stage2");
  return NK_OK;
}

int nk_scale_acc(nk_ctx* ctx, void* dst, int dst_dtype, const void* x, int x_dtype, const float* scalar, size_t n,
                 float beta) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, nk_dtype_ok(dst_dtype) && nk_dtype_ok(x_dtype), "nk_scale_acc: bad dtype");
  if (n == 0) return NK_OK;
  NK_REQUIRE(ctx, dst && x && scalar, "nk_scale_acc: NULL pointer");
  const int blocks = gv_blocks(ctx, n);
  if (dst_dtype == NK_F32 && x_dtype == NK_F32)
    scale_acc_kernel<float, float><<<blocks, kThreads, 0, ctx->stream>>>((float*)dst, (const float*)x, scalar, n, beta);
  else if (dst_dtype == NK_BF16 && x_dtype == NK_BF16)
    scale_acc_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, kThreads, 0, ctx->stream>>>((__nv_bfloat16*)dst, (const __nv_bfloat16*)x, scalar, n, beta);
  else if (dst_dtype == NK_F32)
    scale_acc_kernel<float, __nv_bfloat16><<<blocks, kThreads, 0, ctx->stream>>>((float*)dst, (const __nv_bfloat16*)x, scalar, n, beta);
  else
    scale_acc_kernel<__nv_bfloat16, float><<<blocks, kThreads, 0, ctx->stream>>>((__nv_bfloat16*)dst, (const float*)x, scalar, n, beta);
  NK_LAUNCHED(ctx, "scale_acc");
  return NK_OK;
}

}  // extern "C"
