// General 2-D convolution on the tensor cores: im2col + batched tcgen05 GEMM (bf16, groups = 1, any stride / dilation /
// channel count).  This is the reference's own formulation -- convolution/mod.rs:85-123 (out[n] = Wflat . cols[n]^T),
// :146-189 (dX = col2im(G[n]^T . Wflat)), :191-226 (dW += G[n] . cols[n]) -- with the GEMMs on tcgen05 instead of one
// sgemm per sample: ONE batched launch (3-D TMA maps over (k, row, sample)) per product, and the kernel gradient as a
// split reduction over the samples inside the GEMM's k loop.  It takes every bf16 shape the two specialised engines
// (nk_conv_tz.cu: thin inputs; nk_conv_tc.cu: 3x3 with Cin <= 3 backward) do not, e.g. config 5's 32 -> 64 layer;
// only shapes TMA cannot address (Ho*Wo not a multiple of 8) fall through to the CUDA-core kernels.
//   colsT[n][k][l]  k = (c, i, j) as in the reference (utils.rs:332-353), l = output pixel: the TRANSPOSE of the
//   reference's (l, k) column matrix, so that a row is a shifted copy of image rows (im2col / col2im move contiguous
//   runs) and is the MN-major UMMA operand as it lies; the buffer lives in HBM for the duration of the call (sample
//   chunks of <= 4 GB).  The (Cout, K) kernel is copied once into rows of Kp = ceil8(K) elements (TMA row pitch).
// Compute bound for Cin >= 16 (K >= 144); the column buffer adds 2 x |cols| of HBM traffic.
#include "nk_internal.cuh"

int nk_gemm_tcgen05_batched(nk_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, float alpha, const void* A,
                            int64_t lda, int64_t strideA, const void* B, int64_t ldb, int64_t strideB, void* C, int64_t ldc,
                            int64_t strideC, int64_t batch, int c_dtype, const void* row_bias, int bias_dtype, int relu,
                            int reduce);

namespace {

constexpr int kThreads = 256;
constexpr int64_t kChunkBytes = int64_t(4) << 30;

struct CgDims {
  int64_t n, cin, h, w, cout, kh, kw, sh, sw, dh, dw, ho, wo, K, Kp, L;
};

// colsT[ns][k][l0 .. l0+7]: one thread per 16-byte vector of eight consecutive output pixels of one im2col row
// k = (c, i, j).  With a unit horizontal stride the eight values are a contiguous (2-byte aligned) run of the image row:
// aligned 4-byte loads + a funnel shift when the run starts on an odd element; every store is a full 16-byte vector and a
// warp writes 512 contiguous bytes, so the kernel runs at copy speed (the first version gathered element by element into
// a k-contiguous buffer: 6.4 ms for config 5's 2.35 GB, profiles/r02_launches.md).
__global__ void __launch_bounds__(kThreads) im2col_kernel(__nv_bfloat16* __restrict__ cols, const __nv_bfloat16* __restrict__ x,
                                                          CgDims d, int64_t n0, int64_t nn) {
  const uint32_t lv_n = uint32_t(d.L / 8), K = uint32_t(d.K), wo = uint32_t(d.wo), kw = uint32_t(d.kw), kh = uint32_t(d.kh);
  const int64_t total = nn * int64_t(K) * lv_n;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  const int64_t x_elems = d.n * d.cin * d.h * d.w;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const uint32_t row = uint32_t(idx / lv_n);          // (ns, k)
    const uint32_t lv = uint32_t(idx - int64_t(row) * lv_n);
    const uint32_t ns = row / K, k = row - ns * K;
    const uint32_t j = k % kw, ci = k / kw, i = ci % kh, c = ci / kh;
    const uint32_t l0 = lv * 8, p = l0 / wo, q = l0 - p * wo;
    const int64_t plane = ((n0 + ns) * d.cin + c) * d.h;
    uint4 out;
    if (d.sw == 1 && q + 8 <= wo) {
      const int64_t off = (plane + p * d.sh + i * d.dh) * d.w + q + j * d.dw;   // first of 8 consecutive source elements
      const uint32_t* xw = reinterpret_cast<const uint32_t*>(x);
      if ((off & 1) == 0) {
        const uint32_t* s = xw + (off >> 1);
        out = make_uint4(__ldg(s), __ldg(s + 1), __ldg(s + 2), __ldg(s + 3));
      } else {
        const uint32_t* s = xw + ((off - 1) >> 1);
        const uint32_t w0 = __ldg(s), w1 = __ldg(s + 1), w2 = __ldg(s + 2), w3 = __ldg(s + 3);
        // the fifth word holds source element off+7 in its low half; its high half may lie past the end of x
        const uint32_t w4 = (off + 9 <= x_elems) ? __ldg(s + 4)
                                                 : uint32_t(reinterpret_cast<const unsigned short*>(x)[off + 7]);
        out = make_uint4(__funnelshift_r(w0, w1, 16), __funnelshift_r(w1, w2, 16), __funnelshift_r(w2, w3, 16),
                         __funnelshift_r(w3, w4, 16));
      }
    } else {
      __align__(16) unsigned short e[8];
      const unsigned short* xs = reinterpret_cast<const unsigned short*>(x);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const uint32_t l = l0 + t, pp = l / wo, qq = l - pp * wo;
        e[t] = __ldg(xs + (plane + pp * d.sh + i * d.dh) * d.w + qq * d.sw + j * d.dw);
      }
      out = *reinterpret_cast<const uint4*>(e);
    }
    reinterpret_cast<uint4*>(cols)[idx] = out;
  }
}

// dx[n,c,u,v] = beta*dx + sum_{i,j : u = p*sh + i*dh, v = q*sw + j*dw} dcolsT[ns][(c,i,j)][p*wo+q]
// one thread per dx element; for a fixed tap the reads of a warp are consecutive pixels of one dcolsT row (coalesced)
__global__ void __launch_bounds__(kThreads) col2im_kernel(__nv_bfloat16* __restrict__ dx, const __nv_bfloat16* __restrict__ dcols,
                                                          CgDims d, int64_t n0, int64_t nn, float beta) {
  const int64_t total = nn * d.cin * d.h * d.w;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  const int w = int(d.w), h = int(d.h), cin = int(d.cin), kh = int(d.kh), kw = int(d.kw), sh = int(d.sh), sw = int(d.sw),
            dh = int(d.dh), dw = int(d.dw), ho = int(d.ho), wo = int(d.wo);
  const uint32_t hw = uint32_t(h) * uint32_t(w);
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const uint32_t pl = uint32_t(idx / hw);             // (ns, c)
    const uint32_t uv = uint32_t(idx - int64_t(pl) * hw);
    const int u = int(uv / uint32_t(w)), v = int(uv) - u * w;
    const uint32_t ns = pl / uint32_t(cin), c = pl - ns * uint32_t(cin);
    const __nv_bfloat16* dc = dcols + (int64_t(ns) * d.K + int64_t(c) * kh * kw) * d.L;
    float acc = 0.f;
    for (int i = 0; i < kh; ++i) {
      const int pu = u - i * dh;
      if (pu < 0) break;
      const int p = pu / sh;
      if (p * sh != pu || p >= ho) continue;
      for (int j = 0; j < kw; ++j) {
        const int qv = v - j * dw;
        if (qv < 0) break;
        const int q = qv / sw;
        if (q * sw != qv || q >= wo) continue;
        acc += __bfloat162float(dc[int64_t(i * kw + j) * d.L + p * wo + q]);
      }
    }
    __nv_bfloat16* o = dx + n0 * d.cin * d.h * d.w + idx;
    if (beta != 0.f) acc += beta * __bfloat162float(*o);
    *o = __float2bfloat16_rn(acc);
  }
}

// ---- unit stride / dilation: one block per (sample, channel) plane.  The kh*kw rows (c, i, j) of a channel are CONSECUTIVE
// rows of colsT, i.e. one contiguous block of kh*kw*L elements, and all of them are shifted views of ONE image plane:
// stage the small side in shared memory, stream the large side with full 16-byte vectors.
//   im2col: x plane (h*w) -> shared, then kh*kw*L elements out;   col2im: kh*kw*L elements -> shared, then h*w out.
constexpr int kPlaneThreads = 256;

__global__ void __launch_bounds__(kPlaneThreads) im2col_plane_kernel(__nv_bfloat16* __restrict__ cols, const __nv_bfloat16* __restrict__ x,
                                                                     CgDims d, int64_t n0, int64_t nn) {
  extern __shared__ __align__(16) unsigned short plane[];   // h*w (+ slack) input pixels
  const int hw = int(d.h * d.w), kk = int(d.kh * d.kw), L = int(d.L), wo = int(d.wo), w = int(d.w), kw = int(d.kw);
  const int lv_n = L / 8;
  const int64_t planes = nn * d.cin;
  for (int64_t pl = blockIdx.x; pl < planes; pl += gridDim.x) {
    const int64_t ns = pl / d.cin, c = pl - ns * d.cin;
    const unsigned short* src = reinterpret_cast<const unsigned short*>(x) + ((n0 + ns) * d.cin + c) * hw;
    __syncthreads();
    for (int i = threadIdx.x; i < hw; i += kPlaneThreads) plane[i] = __ldg(src + i);
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(cols + (ns * d.K + c * kk) * int64_t(L));
    for (int idx = threadIdx.x; idx < kk * lv_n; idx += kPlaneThreads) {
      const int t = idx / lv_n, lv = idx - t * lv_n;
      const int i = t / kw, j = t - i * kw;
      const int l0 = lv * 8, p = l0 / wo, q = l0 - p * wo;
      __align__(16) unsigned short e[8];
      if (q + 8 <= wo) {
        const unsigned short* s0 = plane + (p + i) * w + q + j;
#pragma unroll
        for (int k = 0; k < 8; ++k) e[k] = s0[k];
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int l = l0 + k, pp = l / wo, qq = l - pp * wo;
          e[k] = plane[(pp + i) * w + qq + j];
        }
      }
      dst[idx] = *reinterpret_cast<const uint4*>(e);
    }
  }
}

__global__ void __launch_bounds__(kPlaneThreads) col2im_plane_kernel(__nv_bfloat16* __restrict__ dx, const __nv_bfloat16* __restrict__ dcols,
                                                                     CgDims d, int64_t n0, int64_t nn, float beta) {
  extern __shared__ __align__(16) unsigned short taps[];    // kh*kw rows of L column gradients
  const int hw = int(d.h * d.w), kk = int(d.kh * d.kw), L = int(d.L), wo = int(d.wo), ho = int(d.ho), w = int(d.w),
            kw = int(d.kw), kh = int(d.kh);
  const int64_t planes = nn * d.cin;
  for (int64_t pl = blockIdx.x; pl < planes; pl += gridDim.x) {
    const int64_t ns = pl / d.cin, c = pl - ns * d.cin;
    const uint4* src = reinterpret_cast<const uint4*>(dcols + (ns * d.K + c * kk) * int64_t(L));
    __syncthreads();
    for (int i = threadIdx.x; i < kk * L / 8; i += kPlaneThreads) reinterpret_cast<uint4*>(taps)[i] = __ldcs(src + i);
    __syncthreads();
    __nv_bfloat16* out = dx + ((n0 + ns) * d.cin + c) * hw;
    for (int idx = threadIdx.x; idx < hw; idx += kPlaneThreads) {
      const int u = idx / w, v = idx - u * w;
      float acc = 0.f;
      for (int i = 0; i < kh; ++i) {
        const int p = u - i;
        if (p < 0 || p >= ho) continue;
        for (int j = 0; j < kw; ++j) {
          const int q = v - j;
          if (q < 0 || q >= wo) continue;
          acc += __uint_as_float(uint32_t(taps[(i * kw + j) * L + p * wo + q]) << 16);
        }
      }
      if (beta != 0.f) acc += beta * __bfloat162float(out[idx]);
      out[idx] = __float2bfloat16_rn(acc);
    }
  }
}

// The same with the kh*kw tap planes staged ZERO-PADDED ((ho + 2(kh-1)) x (wo + 2(kw-1)) each), so that the sum over the
// taps needs no bounds checks: 9 shared loads + 9 adds per dx element instead of ~110 instructions (the checked version ran
// at 1.3 TB/s on config 5, instruction bound).  Needs wo % 8 == 0 (a 16-byte vector of column gradients stays in one row).
__global__ void __launch_bounds__(kPlaneThreads) col2im_plane_padded_kernel(__nv_bfloat16* __restrict__ dx,
                                                                            const __nv_bfloat16* __restrict__ dcols, CgDims d,
                                                                            int64_t n0, int64_t nn, float beta) {
  extern __shared__ __align__(16) unsigned short taps[];    // kh*kw padded planes
  const int hw = int(d.h * d.w), kk = int(d.kh * d.kw), L = int(d.L), wo = int(d.wo), ho = int(d.ho), w = int(d.w),
            kw = int(d.kw), kh = int(d.kh);
  const int wp = wo + 2 * (kw - 1), hp = ho + 2 * (kh - 1), plane = wp * hp;
  for (int i = threadIdx.x; i < kk * plane / 2; i += kPlaneThreads) reinterpret_cast<uint32_t*>(taps)[i] = 0u;   // borders stay zero
  const int vec_per_row = wo / 8;
  const int64_t planes = nn * d.cin;
  for (int64_t pl = blockIdx.x; pl < planes; pl += gridDim.x) {
    const int64_t ns = pl / d.cin, c = pl - ns * d.cin;
    const uint4* src = reinterpret_cast<const uint4*>(dcols + (ns * d.K + c * kk) * int64_t(L));
    __syncthreads();
    for (int i = threadIdx.x; i < kk * L / 8; i += kPlaneThreads) {
      const uint4 v = __ldcs(src + i);
      const int t = i / (L / 8), lv = i - t * (L / 8);
      const int p = lv / vec_per_row, q = (lv - p * vec_per_row) * 8;
      uint32_t* dst = reinterpret_cast<uint32_t*>(taps + t * plane + (p + kh - 1) * wp + (kw - 1) + q);   // 4-byte aligned: wp, kw-1+q even
      dst[0] = v.x, dst[1] = v.y, dst[2] = v.z, dst[3] = v.w;
    }
    __syncthreads();
    __nv_bfloat16* out = dx + ((n0 + ns) * d.cin + c) * hw;
    for (int idx = threadIdx.x; idx < hw; idx += kPlaneThreads) {
      const int u = idx / w, v = idx - u * w;
      // dx[u][v] = sum_{i,j} T[i][j][u - i][v - j], zero outside: padded coordinates (u - i + kh - 1, v - j + kw - 1)
      const unsigned short* base = taps + (u + kh - 1) * wp + (v + kw - 1);
      float acc = 0.f;
      for (int i = 0; i < kh; ++i)
        for (int j = 0; j < kw; ++j)
          acc += __uint_as_float(uint32_t(base[(i * kw + j) * plane - i * wp - j]) << 16);
      if (beta != 0.f) acc += beta * __bfloat162float(out[idx]);
      out[idx] = __float2bfloat16_rn(acc);
    }
  }
}

// (Cout, Kp) bf16 copy of the (Cout, K) kernel, zero padded (the GEMM operand rows must be 16-byte multiples)
__global__ void __launch_bounds__(kThreads) pad_kernel_rows(__nv_bfloat16* __restrict__ wp, const __nv_bfloat16* __restrict__ w,
                                                            int64_t cout, int64_t K, int64_t Kp) {
  const int64_t total = cout * Kp;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t k = idx % Kp, o = idx / Kp;
    wp[idx] = k < K ? w[o * K + k] : __float2bfloat16_rn(0.f);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) finalize_dw_padded(T* __restrict__ dw, const float* __restrict__ scratch, int64_t cout,
                                                               int64_t K, int64_t Kp, float beta) {
  const int64_t total = cout * K;
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t k = idx % K, o = idx / K;
  float v = scratch[o * Kp + k];
  if (beta != 0.f) v += beta * nk_to_f32<T>(dw[idx]);
  dw[idx] = nk_from_f32<T>(v);
}

inline int cg_blocks(nk_ctx* ctx, int64_t items) {
  int64_t b = (items + kThreads - 1) / kThreads;
  const int64_t cap = int64_t(ctx->sm_count) * 16;
  if (b > cap) b = cap;
  return int(b < 1 ? 1 : b);
}

bool make_dims(CgDims& d, int64_t n, int64_t cin, int64_t h, int64_t w, int64_t cout, int64_t kh, int64_t kw, int64_t sh,
               int64_t sw, int64_t dh, int64_t dw) {
  d = CgDims{n, cin, h, w, cout, kh, kw, sh, sw, dh, dw, 0, 0, 0, 0, 0};
  d.ho = (h - dh * (kh - 1) - 1) / sh + 1;
  d.wo = (w - dw * (kw - 1) - 1) / sw + 1;
  d.K = cin * kh * kw;
  d.Kp = (d.K + 7) & ~int64_t(7);
  d.L = d.ho * d.wo;
  // TMA: every leading dimension / batch stride a multiple of 8 elements; a useful amount of work per GEMM tile
  return n > 0 && d.L > 0 && d.L % 8 == 0 && d.Kp >= 16 && d.L * d.Kp < (int64_t(1) << 31);
}

struct Scratch {   // stream-ordered temporaries released on scope exit
  nk_ctx* ctx;
  void* p[3] = {nullptr, nullptr, nullptr};
  explicit Scratch(nk_ctx* c) : ctx(c) {}
  ~Scratch() {
    for (void* q : p)
      if (q) nk_free(ctx, q);
  }
};

constexpr size_t kPlaneSmemMax = 96 * 1024;
inline bool unit_steps(const CgDims& d) { return d.sh == 1 && d.sw == 1 && d.dh == 1 && d.dw == 1; }
inline int plane_blocks(nk_ctx* ctx, int64_t planes) {
  const int64_t cap = int64_t(ctx->sm_count) * 8;
  return int(planes < cap ? planes : cap);
}

int launch_im2col(nk_ctx* ctx, __nv_bfloat16* cols, const __nv_bfloat16* x, const CgDims& d, int64_t n0, int64_t nn) {
  const size_t smem = size_t(d.h * d.w + 8) * 2;
  if (unit_steps(d) && smem <= 48 * 1024) {
    im2col_plane_kernel<<<plane_blocks(ctx, nn * d.cin), kPlaneThreads, smem, ctx->stream>>>(cols, x, d, n0, nn);
  } else {
    im2col_kernel<<<cg_blocks(ctx, nn * d.K * (d.L / 8)), kThreads, 0, ctx->stream>>>(cols, x, d, n0, nn);
  }
  NK_LAUNCHED(ctx, "im2col");
  return NK_OK;
}

int launch_col2im(nk_ctx* ctx, __nv_bfloat16* dx, const __nv_bfloat16* dcols, const CgDims& d, int64_t n0, int64_t nn, float beta) {
  const size_t smem = size_t(d.kh * d.kw * d.L) * 2;
  static bool attr_done[64] = {};
  const int64_t wp = d.wo + 2 * (d.kw - 1), hp = d.ho + 2 * (d.kh - 1);
  const size_t smem_p = size_t(d.kh * d.kw * wp * hp) * 2;
  if (unit_steps(d) && d.wo % 8 == 0 && d.kw % 2 == 1 && smem_p <= kPlaneSmemMax) {   // (kw odd: kw - 1 even, 4-byte aligned rows)
    static bool attr_done_p[64] = {};
    if (!attr_done_p[ctx->device & 63]) {
      cudaFuncSetAttribute(col2im_plane_padded_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPlaneSmemMax);
      attr_done_p[ctx->device & 63] = true;
    }
    col2im_plane_padded_kernel<<<plane_blocks(ctx, nn * d.cin), kPlaneThreads, smem_p, ctx->stream>>>(dx, dcols, d, n0, nn, beta);
  } else if (unit_steps(d) && smem <= kPlaneSmemMax) {
    if (!attr_done[ctx->device & 63]) {
      cudaFuncSetAttribute(col2im_plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPlaneSmemMax);
      attr_done[ctx->device & 63] = true;
    }
    col2im_plane_kernel<<<plane_blocks(ctx, nn * d.cin), kPlaneThreads, smem, ctx->stream>>>(dx, dcols, d, n0, nn, beta);
  } else {
    col2im_kernel<<<cg_blocks(ctx, nn * d.cin * d.h * d.w), kThreads, 0, ctx->stream>>>(dx, dcols, d, n0, nn, beta);
  }
  NK_LAUNCHED(ctx, "col2im");
  return NK_OK;
}

int64_t chunk_samples(const CgDims& d) {
  int64_t per = d.L * d.Kp * 2;
  int64_t c = kChunkBytes / per;
  if (c < 1) c = 1;
  return c < d.n ? c : d.n;
}

}  // namespace

// all three return NK_ERR_UNSUPPORTED (last_error untouched) when the shape is outside this engine

int nk_conv_gemm_fwd(nk_ctx* ctx, void* y, const void* x, const void* w, const void* bias, int relu, int64_t n, int64_t cin,
                     int64_t h, int64_t wd, int64_t cout, int64_t kh, int64_t kw, int64_t sh, int64_t sw, int64_t dh, int64_t dw) {
  CgDims d;
  if (!make_dims(d, n, cin, h, wd, cout, kh, kw, sh, sw, dh, dw) || !ctx->encode_tiled) return NK_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(y) & 15) || (reinterpret_cast<uintptr_t>(x) & 3) || cout < 8) return NK_ERR_UNSUPPORTED;
  Scratch s(ctx);
  int rc = nk_alloc_uninit(ctx, size_t(cout * d.Kp * 2), &s.p[0]);
  if (rc) return rc;
  pad_kernel_rows<<<cg_blocks(ctx, cout * d.Kp), kThreads, 0, ctx->stream>>>((__nv_bfloat16*)s.p[0], (const __nv_bfloat16*)w, cout, d.K, d.Kp);
  NK_LAUNCHED(ctx, "conv_pad_kernel");
  const int64_t cs = chunk_samples(d);
  rc = nk_alloc_uninit(ctx, size_t(cs * d.L * d.Kp * 2), &s.p[1]);
  if (rc) return rc;
  for (int64_t n0 = 0; n0 < n; n0 += cs) {
    const int64_t nn = n - n0 < cs ? n - n0 : cs;
    rc = launch_im2col(ctx, (__nv_bfloat16*)s.p[1], (const __nv_bfloat16*)x, d, n0, nn);
    if (rc) return rc;
    // y[n] (Cout x L) = Wp (Cout x K) . colsT[n] (K x L) : NN, A shared by every sample (TMA zero-fills k >= K)
    rc = nk_gemm_tcgen05_batched(ctx, 0, 0, cout, d.L, d.K, 1.f, s.p[0], d.Kp, 0, s.p[1], d.L, d.K * d.L,
                                 static_cast<__nv_bfloat16*>(y) + n0 * cout * d.L, d.L, cout * d.L, nn, NK_BF16, bias, NK_BF16,
                                 relu, 0);
    if (rc) return rc;
  }
  ctx->last_conv_kernel = "tcgen05_im2col_gemm_fwd";
  return NK_OK;
}

int nk_conv_gemm_bwd_input(nk_ctx* ctx, void* dx, const void* g, const void* w, int64_t n, int64_t cin, int64_t h, int64_t wd,
                           int64_t cout, int64_t kh, int64_t kw, int64_t sh, int64_t sw, int64_t dh, int64_t dw, float beta) {
  CgDims d;
  if (!make_dims(d, n, cin, h, wd, cout, kh, kw, sh, sw, dh, dw) || !ctx->encode_tiled) return NK_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(g) & 15) || cout % 8 != 0) return NK_ERR_UNSUPPORTED;
  Scratch s(ctx);
  int rc = nk_alloc_uninit(ctx, size_t(cout * d.Kp * 2), &s.p[0]);
  if (rc) return rc;
  pad_kernel_rows<<<cg_blocks(ctx, cout * d.Kp), kThreads, 0, ctx->stream>>>((__nv_bfloat16*)s.p[0], (const __nv_bfloat16*)w, cout, d.K, d.Kp);
  NK_LAUNCHED(ctx, "conv_pad_kernel");
  const int64_t cs = chunk_samples(d);
  rc = nk_alloc_uninit(ctx, size_t(cs * d.L * d.Kp * 2), &s.p[1]);
  if (rc) return rc;
  for (int64_t n0 = 0; n0 < n; n0 += cs) {
    const int64_t nn = n - n0 < cs ? n - n0 : cs;
    // dcolsT[n] (K x L) = Wp^T (K x Cout) . G[n] (Cout x L) : TN (A = Wp stored (Cout, K), shared), B = G[n] as stored
    rc = nk_gemm_tcgen05_batched(ctx, 1, 0, d.K, d.L, cout, 1.f, s.p[0], d.Kp, 0,
                                 static_cast<const __nv_bfloat16*>(g) + n0 * cout * d.L, d.L, cout * d.L, s.p[1], d.L, d.K * d.L, nn,
                                 NK_BF16, nullptr, NK_BF16, 0, 0);
    if (rc) return rc;
    rc = launch_col2im(ctx, (__nv_bfloat16*)dx, (const __nv_bfloat16*)s.p[1], d, n0, nn, beta);
    if (rc) return rc;
  }
  ctx->last_conv_kernel = "tcgen05_im2col_gemm_dx";
  return NK_OK;
}

int nk_conv_gemm_bwd_kernel(nk_ctx* ctx, void* dwt, int dw_dtype, const void* g, const void* x, int64_t n, int64_t cin, int64_t h,
                            int64_t wd, int64_t cout, int64_t kh, int64_t kw, int64_t sh, int64_t sw, int64_t dh, int64_t dw,
                            float beta) {
  CgDims d;
  if (!make_dims(d, n, cin, h, wd, cout, kh, kw, sh, sw, dh, dw) || !ctx->encode_tiled) return NK_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(g) & 15) || (reinterpret_cast<uintptr_t>(x) & 3)) return NK_ERR_UNSUPPORTED;
  Scratch s(ctx);
  const int64_t cs = chunk_samples(d);
  int rc = nk_alloc_uninit(ctx, size_t(cs * d.L * d.Kp * 2), &s.p[1]);
  if (rc) return rc;
  rc = nk_alloc(ctx, size_t(cout * d.Kp * 4), &s.p[2]);   // f32 accumulator of the split reduction, zero filled
  if (rc) return rc;
  for (int64_t n0 = 0; n0 < n; n0 += cs) {
    const int64_t nn = n - n0 < cs ? n - n0 : cs;
    rc = launch_im2col(ctx, (__nv_bfloat16*)s.p[1], (const __nv_bfloat16*)x, d, n0, nn);
    if (rc) return rc;
    // acc (Cout x K) += sum_n G[n] (Cout x L) . colsT[n]^T (L x K) : NT, reduced over the samples inside the k loop
    rc = nk_gemm_tcgen05_batched(ctx, 0, 1, cout, d.K, d.L, 1.f, static_cast<const __nv_bfloat16*>(g) + n0 * cout * d.L, d.L,
                                 cout * d.L, s.p[1], d.L, d.K * d.L, s.p[2], d.Kp, 0, nn, NK_F32, nullptr, NK_F32, 0, 1);
    if (rc) return rc;
  }
  const int fb = int((cout * d.K + kThreads - 1) / kThreads);
  if (dw_dtype == NK_BF16)
    finalize_dw_padded<__nv_bfloat16><<<fb, kThreads, 0, ctx->stream>>>((__nv_bfloat16*)dwt, (const float*)s.p[2], cout, d.K, d.Kp, beta);
  else
    finalize_dw_padded<float><<<fb, kThreads, 0, ctx->stream>>>((float*)dwt, (const float*)s.p[2], cout, d.K, d.Kp, beta);
  NK_LAUNCHED(ctx, "conv_dw_finalize_padded");
  ctx->last_conv_kernel = "tcgen05_im2col_gemm_dw";
  return NK_OK;
}
