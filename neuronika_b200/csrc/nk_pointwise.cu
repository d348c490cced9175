// The rest of the broadcast / elementwise / shape family, same HBM-bound template as
// ReLU / add in nk_elementwise.cu: 16-byte vector accesses on aligned contiguous operands, grids of 8 CTAs per SM,
// f32 arithmetic whatever the storage type.
//   binary, broadcasting   sub / mul / div (and add)   subtraction/mod.rs:44-49,87-140, multiplication/mod.rs:44-49,90-148,
//                                                       division/mod.rs:44-49,90-148
//   unary                  neg, exp, ln, sqrt, sigmoid, tanh, softplus, leaky_relu, powi
//                          negation/mod.rs:32-36,66-68; exp/mod.rs:32-36,69-74; logn/mod.rs:32-36,69-74;
//                          sqrt/mod.rs:32-36,69-74; sigmoid/mod.rs:32-36,69-76; tanh/mod.rs:32-36,69-76;
//                          softplus/mod.rs:32-36,69-76; leaky_relu/mod.rs:33-39,73-81; power/mod.rs:41-45,81-88
//   transpose              transpose/mod.rs:32-36 (`.t()` reverses every axis), 66-68 (dX += G^T)
//   n-d padding            Constant / Zero / Reflective / Replicative over 1..3 sample dims
//                          pad/mod.rs:97-129 (forward), 157-182 (backward = interior slice, whatever the mode),
//                          pad/reflective/mod.rs:9-140, pad/replicative/mod.rs:9-130, pad/constant/mod.rs:14-39
#include "nk_internal.cuh"

namespace {

constexpr int kThreads = 256;

inline int pw_blocks(nk_ctx* ctx, size_t work_items) {
  size_t b = (work_items + kThreads - 1) / kThreads;
  size_t cap = size_t(ctx->sm_count) * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return int(b);
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------------------------------- functors
// binary forward value and the per-element factor of each operand's gradient (before un-broadcasting)
struct BinFwd {
  int op;
  __device__ __forceinline__ float operator()(float l, float r) const {
    switch (op) {
      case NK_BIN_ADD: return l + r;
      case NK_BIN_SUB: return l - r;
      case NK_BIN_MUL: return l * r;
      default: return l / r;
    }
  }
};
struct BinBwd {
  int op, side;  // side 0 = left operand, 1 = right operand
  __device__ __forceinline__ float operator()(float g, float l, float r) const {
    switch (op) {
      case NK_BIN_ADD: return g;
      case NK_BIN_SUB: return side ? -g : g;                      // subtraction/mod.rs:87-92, 130-135
      case NK_BIN_MUL: return side ? g * l : g * r;               // multiplication/mod.rs:90-101, 139-148
      default: return side ? (-g * l) / (r * r) : g / r;          // division/mod.rs:90-100 (g / r), 142-151 (-g*l / r.powi(2))
    }
  }
};

__device__ __forceinline__ float powi_f32(float x, int e) {  // f32::powi: repeated squaring, negative -> reciprocal
  unsigned n = e < 0 ? unsigned(-(long long)e) : unsigned(e);
  float r = 1.f, b = x;
  while (n) {
    if (n & 1u) r *= b;
    b *= b;
    n >>= 1;
  }
  return e < 0 ? 1.f / r : r;
}

struct UnFwd {
  int op, ip;
  __device__ __forceinline__ float operator()(float x) const {
    switch (op) {
      case NK_UN_NEG: return -x;
      case NK_UN_EXP: return expf(x);
      case NK_UN_LN: return logf(x);
      case NK_UN_SQRT: return sqrtf(x);
      case NK_UN_SIGMOID: return 1.f / (1.f + expf(-x));
      case NK_UN_TANH: return tanhf(x);
      case NK_UN_SOFTPLUS: return logf(1.f + expf(x));            // softplus/mod.rs:35 (1 + e^x).ln(), no threshold
      case NK_UN_LEAKY_RELU: return x > 0.f ? x : 0.01f * x;      // leaky_relu/mod.rs:36-38
      default: return powi_f32(x, ip);
    }
  }
};
// s = the tensor the reference's Backward node keeps: the node's OUTPUT for exp / sqrt / sigmoid / tanh, its INPUT
// for ln / softplus / leaky_relu / powi; unused for neg
struct UnBwd {
  int op, ip;
  __device__ __forceinline__ float operator()(float g, float s) const {
    switch (op) {
      case NK_UN_NEG: return -g;
      case NK_UN_EXP: return g * s;                               // exp/mod.rs:73
      case NK_UN_LN: return g / s;                                // logn/mod.rs:73
      case NK_UN_SQRT: return g / (s * 2.f);                      // sqrt/mod.rs:73
      case NK_UN_SIGMOID: return g * s * (1.f - s);               // sigmoid/mod.rs:74
      case NK_UN_TANH: return g * (1.f - s * s);                  // tanh/mod.rs:74
      case NK_UN_SOFTPLUS: return g / (1.f + expf(-s));           // softplus/mod.rs:74
      // leaky_relu/mod.rs:78-79 adds the constant 0.01 (not 0.01*g) on the negative side: a defect that its own
      // test (g = 1, leaky_relu/test.rs:152-172) cannot see; the intended slope*g is implemented
      case NK_UN_LEAKY_RELU: return s > 0.f ? g : 0.01f * g;
      default: return g * powi_f32(s, ip - 1) * float(ip);        // power/mod.rs:86
    }
  }
};

// ---------------------------------------------------------------------------------------- contiguous kernels
// out = beta*out + f(a[, b[, c]]) over n contiguous elements; TO may differ from TI only in the scalar tail path
template <typename T, int NIN, bool VEC, typename F>
__global__ void __launch_bounds__(kThreads) map_kernel(T* __restrict__ out, const T* __restrict__ a,
                                                       const T* __restrict__ b, const T* __restrict__ c, size_t n,
                                                       float beta, F f) {
  const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  size_t done = 0;
  if (VEC) {
    constexpr int V = NkVec<T>::N;
    const size_t nvec = n / V;
    for (size_t v = tid; v < nvec; v += stride) {
      NkVec<T> va, vb, vc, vo;
      va.load(a + v * V);
      if (NIN > 1) vb.load(b + v * V);
      if (NIN > 2) vc.load(c + v * V);
      if (beta != 0.f) vo.load(out + v * V);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float r = f(va.get(i), NIN > 1 ? vb.get(i) : 0.f, NIN > 2 ? vc.get(i) : 0.f);
        if (beta != 0.f) r += beta * vo.get(i);
        vo.set(i, r);
      }
      vo.store(out + v * V);
    }
    done = nvec * V;
  }
  for (size_t i = done + tid; i < n; i += stride) {
    float r = f(nk_to_f32<T>(a[i]), NIN > 1 ? nk_to_f32<T>(b[i]) : 0.f, NIN > 2 ? nk_to_f32<T>(c[i]) : 0.f);
    if (beta != 0.f) r += beta * nk_to_f32<T>(out[i]);
    out[i] = nk_from_f32<T>(r);
  }
}

template <typename T, int NIN, typename F>
int launch_map(nk_ctx* ctx, const char* name, void* out, const void* a, const void* b, const void* c, size_t n,
               float beta, F f) {
  if (n == 0) return NK_OK;
  const bool vec = aligned16(out) && aligned16(a) && (NIN < 2 || aligned16(b)) && (NIN < 3 || aligned16(c));
  const int blocks = pw_blocks(ctx, vec ? n / NkVec<T>::N + 1 : n);
  if (vec)
    map_kernel<T, NIN, true, F><<<blocks, kThreads, 0, ctx->stream>>>((T*)out, (const T*)a, (const T*)b, (const T*)c, n, beta, f);
  else
    map_kernel<T, NIN, false, F><<<blocks, kThreads, 0, ctx->stream>>>((T*)out, (const T*)a, (const T*)b, (const T*)c, n, beta, f);
  NK_LAUNCHED(ctx, name);
  return NK_OK;
}

struct Un1 {
  UnFwd f;
  __device__ __forceinline__ float operator()(float x, float, float) const { return f(x); }
};
struct Un2 {
  UnBwd f;
  __device__ __forceinline__ float operator()(float g, float s, float) const { return f(g, s); }
};
struct Bin2 {
  BinFwd f;
  __device__ __forceinline__ float operator()(float l, float r, float) const { return f(l, r); }
};
struct Bin3 {
  BinBwd f;
  __device__ __forceinline__ float operator()(float g, float l, float r) const { return f(g, l, r); }
};

// ---------------------------------------------------------------------------------------- broadcasting kernels
struct BDims {
  int ndim;
  int64_t shape[NK_MAX_DIMS];
  int64_t ls[NK_MAX_DIMS];  // element strides of the operands in the broadcast shape (0 on broadcast axes)
  int64_t rs[NK_MAX_DIMS];
};

__device__ __forceinline__ void bcast_offsets(const BDims& d, size_t i, int64_t& lo, int64_t& ro) {
  size_t rem = i;
  lo = ro = 0;
#pragma unroll
  for (int k = NK_MAX_DIMS - 1; k >= 0; --k) {
    if (k < d.ndim) {
      const int64_t c = int64_t(rem % size_t(d.shape[k]));
      rem /= size_t(d.shape[k]);
      lo += c * d.ls[k];
      ro += c * d.rs[k];
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) bin_bcast_fwd_kernel(T* __restrict__ y, const T* __restrict__ l,
                                                                 const T* __restrict__ r, size_t n, BDims d, BinFwd f) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    int64_t lo, ro;
    bcast_offsets(d, i, lo, ro);
    y[i] = nk_from_f32<T>(f(nk_to_f32<T>(l[lo]), nk_to_f32<T>(r[ro])));
  }
}

// buffer (f32, the broadcast shape) = factor(g, l, r): the reference's BufferedGradient step before `accumulate`
template <typename T>
__global__ void __launch_bounds__(kThreads) bin_bcast_bwd_kernel(float* __restrict__ buf, const T* __restrict__ g,
                                                                 const T* __restrict__ l, const T* __restrict__ r,
                                                                 size_t n, BDims d, BinBwd f) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    int64_t lo, ro;
    bcast_offsets(d, i, lo, ro);
    buf[i] = f(nk_to_f32<T>(g[i]), l ? nk_to_f32<T>(l[lo]) : 0.f, r ? nk_to_f32<T>(r[ro]) : 0.f);
  }
}

// co-broadcast rule of utils.rs:97-125 + operand strides
int make_bdims(nk_ctx* ctx, int l_ndim, const int64_t* ls, int r_ndim, const int64_t* rs, BDims* d, size_t* n,
               size_t* nl, size_t* nr) {
  NK_REQUIRE(ctx, l_ndim >= 0 && l_ndim <= NK_MAX_DIMS && r_ndim >= 0 && r_ndim <= NK_MAX_DIMS,
             "broadcast: at most %d dims", NK_MAX_DIMS);
  const int nd = l_ndim > r_ndim ? l_ndim : r_ndim;
  d->ndim = nd;
  int64_t lstride = 1, rstride = 1;
  *n = *nl = *nr = 1;
  for (int k = nd - 1; k >= 0; --k) {
    const int lk = k - (nd - l_ndim), rk = k - (nd - r_ndim);
    const int64_t a = lk >= 0 ? ls[lk] : 1, b = rk >= 0 ? rs[rk] : 1;
    NK_REQUIRE(ctx, a >= 0 && b >= 0, "broadcast: negative dimension");
    NK_REQUIRE(ctx, a == b || a == 1 || b == 1, "The two tensors have incompatible shape.");
    d->shape[k] = a == 1 ? b : a;
    d->ls[k] = (a == 1 && d->shape[k] != 1) ? 0 : lstride;
    d->rs[k] = (b == 1 && d->shape[k] != 1) ? 0 : rstride;
    lstride *= a;
    rstride *= b;
    *n *= size_t(d->shape[k]);
    *nl *= size_t(a);
    *nr *= size_t(b);
  }
  for (int k = nd; k < NK_MAX_DIMS; ++k) d->shape[k] = 1, d->ls[k] = 0, d->rs[k] = 0;
  return NK_OK;
}

// ---------------------------------------------------------------------------------------- transpose
// 2-D: 32x32 tiles through shared memory (coalesced on both sides); dst (cols, rows) = beta*dst + src(rows, cols)^T
template <typename TD, typename TS>
__global__ void __launch_bounds__(256) transpose2d_kernel(TD* __restrict__ dst, const TS* __restrict__ src, int64_t rows,
                                                          int64_t cols, float beta) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int64_t tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32;
  for (int64_t t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
    const int64_t r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const int64_t r = r0 + ty + j, c = c0 + tx;
      if (r < rows && c < cols) tile[ty + j][tx] = nk_to_f32<TS>(src[r * cols + c]);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const int64_t c = c0 + ty + j, r = r0 + tx;  // dst row = c, dst col = r
      if (r < rows && c < cols) {
        float v = tile[tx][ty + j];
        if (beta != 0.f) v += beta * nk_to_f32<TD>(dst[c * rows + r]);
        dst[c * rows + r] = nk_from_f32<TD>(v);
      }
    }
    __syncthreads();
  }
}

// n-d reversal of all axes (ndarray `.t()`): dst[i_{n-1},...,i_0] = src[i_0,...,i_{n-1}]
struct TDims {
  int ndim;
  int64_t dshape[NK_MAX_DIMS];   // dst shape = reversed src shape
  int64_t sstride[NK_MAX_DIMS];  // src element stride of the axis that dst axis k walks
};
template <typename TD, typename TS>
__global__ void __launch_bounds__(kThreads) transpose_nd_kernel(TD* __restrict__ dst, const TS* __restrict__ src, size_t n,
                                                                TDims d, float beta) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    size_t rem = i;
    int64_t so = 0;
    for (int k = d.ndim - 1; k >= 0; --k) {
      so += int64_t(rem % size_t(d.dshape[k])) * d.sstride[k];
      rem /= size_t(d.dshape[k]);
    }
    float v = nk_to_f32<TS>(src[so]);
    if (beta != 0.f) v += beta * nk_to_f32<TD>(dst[i]);
    dst[i] = nk_from_f32<TD>(v);
  }
}

// ---------------------------------------------------------------------------------------- n-d padding
struct PadDims {
  int64_t in[3], pad[3], out[3];  // sample dims (leading ones are 1 for fewer than 3 dims)
};
// source coordinate of padded coordinate o along an axis of length len padded by p on both sides
__device__ __forceinline__ int64_t pad_src(int64_t o, int64_t len, int64_t p, int mode) {
  if (o >= p && o < len + p) return o - p;
  if (mode == NK_PAD_REFLECTIVE) return (o < p ? 2 * p - o : 2 * (len + p - 1) - o) - p;  // reflective/mod.rs:22-31
  if (mode == NK_PAD_REPLICATIVE) return o < p ? 0 : len - 1;                              // replicative/mod.rs:22-31
  return -1;                                                                               // constant: fill
}
template <typename T>
__global__ void __launch_bounds__(kThreads) padnd_fwd_kernel(T* __restrict__ y, const T* __restrict__ x, int64_t planes,
                                                             PadDims d, int mode, float value) {
  const int64_t osz = d.out[0] * d.out[1] * d.out[2], isz = d.in[0] * d.in[1] * d.in[2];
  const int64_t total = planes * osz;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  const T fillv = nk_from_f32<T>(value);
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t pl = i / osz;
    int64_t rem = i - pl * osz;
    const int64_t c2 = rem % d.out[2];
    rem /= d.out[2];
    const int64_t c1 = rem % d.out[1], c0 = rem / d.out[1];
    const int64_t s0 = pad_src(c0, d.in[0], d.pad[0], mode), s1 = pad_src(c1, d.in[1], d.pad[1], mode),
                  s2 = pad_src(c2, d.in[2], d.pad[2], mode);
    y[i] = (s0 < 0 || s1 < 0 || s2 < 0) ? fillv : x[pl * isz + (s0 * d.in[1] + s1) * d.in[2] + s2];  // bit-exact copy
  }
}
// dx += g[interior]  (pad/mod.rs:157-182: the same slice for every mode)
template <typename T>
__global__ void __launch_bounds__(kThreads) padnd_bwd_kernel(T* __restrict__ dx, const T* __restrict__ g, int64_t planes,
                                                             PadDims d, float beta) {
  const int64_t osz = d.out[0] * d.out[1] * d.out[2], isz = d.in[0] * d.in[1] * d.in[2];
  const int64_t total = planes * isz;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t pl = i / isz;
    int64_t rem = i - pl * isz;
    const int64_t c2 = rem % d.in[2];
    rem /= d.in[2];
    const int64_t c1 = rem % d.in[1], c0 = rem / d.in[1];
    const T gv = g[pl * osz + ((c0 + d.pad[0]) * d.out[1] + c1 + d.pad[1]) * d.out[2] + c2 + d.pad[2]];
    if (beta != 0.f)
      dx[i] = nk_from_f32<T>(beta * nk_to_f32<T>(dx[i]) + nk_to_f32<T>(gv));
    else
      dx[i] = gv;
  }
}

int make_pad_dims(nk_ctx* ctx, int nsp, const int64_t* in_sp, const int64_t* pad, int mode, PadDims* d) {
  NK_REQUIRE(ctx, nsp >= 1 && nsp <= 3, "pad: 1 to 3 sample dimensions (got %d)", nsp);
  NK_REQUIRE(ctx, mode >= NK_PAD_CONSTANT && mode <= NK_PAD_REPLICATIVE, "pad: bad mode %d", mode);
  for (int k = 0; k < 3; ++k) d->in[k] = 1, d->pad[k] = 0, d->out[k] = 1;
  for (int k = 0; k < nsp; ++k) {
    const int j = 3 - nsp + k;
    NK_REQUIRE(ctx, in_sp[k] >= 0 && pad[k] >= 0, "pad: negative size");
    // a reflection needs pad < len (the reference indexes out of bounds otherwise)
    NK_REQUIRE(ctx, mode != NK_PAD_REFLECTIVE || pad[k] == 0 || pad[k] < in_sp[k],
               "pad: reflective padding %lld must be smaller than the dimension %lld", (long long)pad[k], (long long)in_sp[k]);
    NK_REQUIRE(ctx, mode != NK_PAD_REPLICATIVE || pad[k] == 0 || in_sp[k] > 0, "pad: replicative padding of an empty dimension");
    d->in[j] = in_sp[k];
    d->pad[j] = pad[k];
    d->out[j] = in_sp[k] + 2 * pad[k];
  }
  return NK_OK;
}

}  // namespace

extern "C" {

int nk_unary_fwd(nk_ctx* ctx, int op, void* y, const void* x, size_t n, int dtype, int iparam) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, nk_dtype_ok(dtype), "nk_unary_fwd: bad dtype %d", dtype);
  NK_REQUIRE(ctx, op >= NK_UN_NEG && op <= NK_UN_POWI, "nk_unary_fwd: bad op %d", op);
  NK_REQUIRE(ctx, (y && x) || n == 0, "nk_unary_fwd: NULL pointer");
  Un1 f{UnFwd{op, iparam}};
  NK_DISPATCH_DTYPE(dtype, T, return (launch_map<T, 1>(ctx, "unary_fwd", y, x, nullptr, nullptr, n, 0.f, f)));
}

int nk_unary_bwd(nk_ctx* ctx, int op, void* dx, const void* saved, const void* g, size_t n, int dtype, int iparam,
                 float beta) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, nk_dtype_ok(dtype), "nk_unary_bwd: bad dtype %d", dtype);
  NK_REQUIRE(ctx, op >= NK_UN_NEG && op <= NK_UN_POWI, "nk_unary_bwd: bad op %d", op);
  NK_REQUIRE(ctx, (dx && g && (saved || op == NK_UN_NEG)) || n == 0, "nk_unary_bwd: NULL pointer");
  Un2 f{UnBwd{op, iparam}};
  if (op == NK_UN_NEG) saved = g;  // unused by the functor; keeps the two-input kernel
  NK_DISPATCH_DTYPE(dtype, T, return (launch_map<T, 2>(ctx, "unary_bwd", dx, g, saved, nullptr, n, beta, f)));
}

int nk_binary_bcast_fwd(nk_ctx* ctx, int op, void* y, const void* l, const void* r, int dtype, int y_ndim,
                        const int64_t* y_shape, int l_ndim, const int64_t* l_shape, int r_ndim,
                        const int64_t* r_shape) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, nk_dtype_ok(dtype), "nk_binary_bcast_fwd: bad dtype %d", dtype);
  NK_REQUIRE(ctx, op >= NK_BIN_ADD && op <= NK_BIN_DIV, "nk_binary_bcast_fwd: bad op %d", op);
  if (op == NK_BIN_ADD)
    return nk_add_bcast_fwd(ctx, y, l, r, dtype, y_ndim, y_shape, l_ndim, l_shape, r_ndim, r_shape);
  BDims d;
  size_t n, nl, nr;
  int rc = make_bdims(ctx, l_ndim, l_shape, r_ndim, r_shape, &d, &n, &nl, &nr);
  if (rc) return rc;
  NK_REQUIRE(ctx, d.ndim == y_ndim, "nk_binary_bcast_fwd: output rank %d != broadcast rank %d", y_ndim, d.ndim);
  for (int k = 0; k < d.ndim; ++k)
    NK_REQUIRE(ctx, y_shape[k] == d.shape[k], "nk_binary_bcast_fwd: output dim %d is %lld, expected %lld", k,
               (long long)y_shape[k], (long long)d.shape[k]);
  if (n == 0) return NK_OK;
  NK_REQUIRE(ctx, y && l && r, "nk_binary_bcast_fwd: NULL pointer");
  if (nl == n && nr == n) {
    Bin2 f{BinFwd{op}};
    NK_DISPATCH_DTYPE(dtype, T, return (launch_map<T, 2>(ctx, "binary_fwd", y, l, r, nullptr, n, 0.f, f)));
  }
  const int blocks = pw_blocks(ctx, n);
  if (dtype == NK_BF16)
    bin_bcast_fwd_kernel<__nv_bfloat16><<<blocks, kThreads, 0, ctx->stream>>>((__nv_bfloat16*)y, (const __nv_bfloat16*)l, (const __nv_bfloat16*)r, n, d, BinFwd{op});
  else
    bin_bcast_fwd_kernel<float><<<blocks, kThreads, 0, ctx->stream>>>((float*)y, (const float*)l, (const float*)r, n, d, BinFwd{op});
  NK_LAUNCHED(ctx, "binary_bcast_fwd");
  return NK_OK;
}

int nk_binary_bcast_bwd(nk_ctx* ctx, int op, int side, void* dst, int dst_dtype, const void* g, const void* l,
                        const void* r, int dtype, int l_ndim, const int64_t* l_shape, int r_ndim,
                        const int64_t* r_shape, float beta) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, nk_dtype_ok(dtype) && nk_dtype_ok(dst_dtype), "nk_binary_bcast_bwd: bad dtype");
  NK_REQUIRE(ctx, op >= NK_BIN_ADD && op <= NK_BIN_DIV && (side == 0 || side == 1), "nk_binary_bcast_bwd: bad op/side");
  BDims d;
  size_t n, nl, nr;
  int rc = make_bdims(ctx, l_ndim, l_shape, r_ndim, r_shape, &d, &n, &nl, &nr);
  if (rc) return rc;
  if (n == 0) return NK_OK;
  const bool need_l = (op == NK_BIN_MUL && side == 1) || (op == NK_BIN_DIV && side == 1);
  const bool need_r = (op == NK_BIN_MUL && side == 0) || op == NK_BIN_DIV;
  NK_REQUIRE(ctx, dst && g && (!need_l || l) && (!need_r || r), "nk_binary_bcast_bwd: NULL pointer");
  const int dst_ndim = side ? r_ndim : l_ndim;
  const int64_t* dst_shape = side ? r_shape : l_shape;
  const size_t n_dst = side ? nr : nl;
  if (n_dst == n && nl == n && nr == n && dst_dtype == dtype) {  // no broadcasting at all: one fused pass
    Bin3 f{BinBwd{op, side}};
    const void* ll = l ? l : g;
    const void* rr = r ? r : g;
    NK_DISPATCH_DTYPE(dtype, T, return (launch_map<T, 3>(ctx, "binary_bwd", dst, g, ll, rr, n, beta, f)));
  }
  if ((op == NK_BIN_ADD || (op == NK_BIN_SUB && side == 0)))  // the factor is the gradient itself
    return nk_unbroadcast_acc(ctx, dst, dst_dtype, dst_ndim, dst_shape, g, dtype, d.ndim, d.shape, beta);
  // buffer = factor over the broadcast shape (f32), then un-broadcast into the operand gradient
  void* buf = nullptr;
  NK_CUDA(ctx, cudaMallocAsync(&buf, n * sizeof(float), ctx->stream));
  const int blocks = pw_blocks(ctx, n);
  if (dtype == NK_BF16)
    bin_bcast_bwd_kernel<__nv_bfloat16><<<blocks, kThreads, 0, ctx->stream>>>((float*)buf, (const __nv_bfloat16*)g, (const __nv_bfloat16*)l, (const __nv_bfloat16*)r, n, d, BinBwd{op, side});
  else
    bin_bcast_bwd_kernel<float><<<blocks, kThreads, 0, ctx->stream>>>((float*)buf, (const float*)g, (const float*)l, (const float*)r, n, d, BinBwd{op, side});
  ctx->launches++;
  rc = cudaGetLastError() == cudaSuccess
           ? nk_unbroadcast_acc(ctx, dst, dst_dtype, dst_ndim, dst_shape, buf, NK_F32, d.ndim, d.shape, beta)
           : nk_set_error(ctx, NK_ERR_CUDA, "launch of binary_bcast_bwd failed");
  cudaFreeAsync(buf, ctx->stream);
  return rc;
}

int nk_transpose(nk_ctx* ctx, void* dst, int dst_dtype, const void* src, int src_dtype, int ndim,
                 const int64_t* src_shape, float beta) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, nk_dtype_ok(dst_dtype) && nk_dtype_ok(src_dtype), "nk_transpose: bad dtype");
  NK_REQUIRE(ctx, ndim >= 0 && ndim <= NK_MAX_DIMS, "nk_transpose: at most %d dims", NK_MAX_DIMS);
  size_t n = 1;
  for (int k = 0; k < ndim; ++k) {
    NK_REQUIRE(ctx, src_shape[k] >= 0, "nk_transpose: negative dimension");
    n *= size_t(src_shape[k]);
  }
  if (n == 0) return NK_OK;
  NK_REQUIRE(ctx, dst && src, "nk_transpose: NULL pointer");
#define NK_TR_DISPATCH(KERNEL, ...)                                                                                   \
  do {                                                                                                                \
    if (dst_dtype == NK_F32 && src_dtype == NK_F32)                                                                   \
      KERNEL<float, float><<<blocks, kThreads, 0, ctx->stream>>>((float*)dst, (const float*)src, __VA_ARGS__);        \
    else if (dst_dtype == NK_BF16 && src_dtype == NK_BF16)                                                            \
      KERNEL<__nv_bfloat16, __nv_bfloat16><<<blocks, kThreads, 0, ctx->stream>>>((__nv_bfloat16*)dst, (const __nv_bfloat16*)src, __VA_ARGS__); \
    else if (dst_dtype == NK_F32)                                                                                     \
      KERNEL<float, __nv_bfloat16><<<blocks, kThreads, 0, ctx->stream>>>((float*)dst, (const __nv_bfloat16*)src, __VA_ARGS__); \
    else                                                                                                              \
      KERNEL<__nv_bfloat16, float><<<blocks, kThreads, 0, ctx->stream>>>((__nv_bfloat16*)dst, (const float*)src, __VA_ARGS__); \
  } while (0)
  if (ndim == 2) {
    const int64_t rows = src_shape[0], cols = src_shape[1];
    const int64_t tiles = ((rows + 31) / 32) * ((cols + 31) / 32);
    const int blocks = int(tiles < int64_t(ctx->sm_count) * 8 ? tiles : int64_t(ctx->sm_count) * 8);
    NK_TR_DISPATCH(transpose2d_kernel, rows, cols, beta);
    NK_LAUNCHED(ctx, "transpose2d");
    return NK_OK;
  }
  TDims d;
  d.ndim = ndim;
  int64_t sstr[NK_MAX_DIMS];
  int64_t acc = 1;
  for (int k = ndim - 1; k >= 0; --k) {
    sstr[k] = acc;
    acc *= src_shape[k];
  }
  for (int k = 0; k < NK_MAX_DIMS; ++k) {
    d.dshape[k] = k < ndim ? src_shape[ndim - 1 - k] : 1;
    d.sstride[k] = k < ndim ? sstr[ndim - 1 - k] : 0;
  }
  const int blocks = pw_blocks(ctx, n);
  NK_TR_DISPATCH(transpose_nd_kernel, n, d, beta);
#undef NK_TR_DISPATCH
  NK_LAUNCHED(ctx, "transpose_nd");
  return NK_OK;
}

int nk_padnd_fwd(nk_ctx* ctx, void* y, const void* x, int64_t planes, int nsp, const int64_t* in_sp,
                 const int64_t* pad, int mode, float value, int dtype) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, nk_dtype_ok(dtype), "nk_padnd_fwd: bad dtype %d", dtype);
  NK_REQUIRE(ctx, planes >= 0 && in_sp && pad, "nk_padnd_fwd: bad arguments");
  PadDims d;
  int rc = make_pad_dims(ctx, nsp, in_sp, pad, mode, &d);
  if (rc) return rc;
  const size_t total = size_t(planes) * size_t(d.out[0] * d.out[1] * d.out[2]);
  if (total == 0) return NK_OK;
  NK_REQUIRE(ctx, y && x, "nk_padnd_fwd: NULL pointer");
  const int blocks = pw_blocks(ctx, total);
  if (dtype == NK_BF16)
    padnd_fwd_kernel<__nv_bfloat16><<<blocks, kThreads, 0, ctx->stream>>>((__nv_bfloat16*)y, (const __nv_bfloat16*)x, planes, d, mode, value);
  else
    padnd_fwd_kernel<float><<<blocks, kThreads, 0, ctx->stream>>>((float*)y, (const float*)x, planes, d, mode, value);
  NK_LAUNCHED(ctx, "padnd_fwd");
  return NK_OK;
}

int nk_padnd_bwd(nk_ctx* ctx, void* dx, const void* g, int64_t planes, int nsp, const int64_t* in_sp,
                 const int64_t* pad, int dtype, float beta) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, nk_dtype_ok(dtype), "nk_padnd_bwd: bad dtype %d", dtype);
  NK_REQUIRE(ctx, planes >= 0 && in_sp && pad, "nk_padnd_bwd: bad arguments");
  PadDims d;
  int rc = make_pad_dims(ctx, nsp, in_sp, pad, NK_PAD_CONSTANT, &d);
  if (rc) return rc;
  const size_t total = size_t(planes) * size_t(d.in[0] * d.in[1] * d.in[2]);
  if (total == 0) return NK_OK;
  NK_REQUIRE(ctx, dx && g, "nk_padnd_bwd: NULL pointer");
  const int blocks = pw_blocks(ctx, total);
  if (dtype == NK_BF16)
    padnd_bwd_kernel<__nv_bfloat16><<<blocks, kThreads, 0, ctx->stream>>>((__nv_bfloat16*)dx, (const __nv_bfloat16*)g, planes, d, beta);
  else
    padnd_bwd_kernel<float><<<blocks, kThreads, 0, ctx->stream>>>((float*)dx, (const float*)g, planes, d, beta);
  NK_LAUNCHED(ctx, "padnd_bwd");
  return NK_OK;
}

}  // extern "C"
