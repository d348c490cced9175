// Tensor-core implicit-GEMM convolution engine (sm_100a): stride 1, dilation 1, groups 1, bf16, NCHW.
//
// Reference semantics: convolution/mod.rs:85-123 (out[n] = Wflat . im2col(x[n])^T, no padding, beta = 0).
// The reference materialises the (N, L, K) column matrix (1.36 GB at config 3) and runs 256 skinny
// sgemms; here the columns never exist.  The op is HBM-bound at config 3 (K = 27: ~26 flop/byte), so the
// design goal is ONE pass over x and y:
//
//   D[co][px] = sum_k Wt[co][k] * Xwin[k][px]         ("swap-AB": M = Cout, N = pixels)
//
//   * B operand (pixels): for a run of 64 consecutive output pixels of one output row, the im2col row
//     k = (j, c, i) is the 64-element window x[n, c, p+i, q0+j .. q0+j+63] -- contiguous in memory.  One TMA
//     box {64 (W), kh (H), cpg (C)} therefore lands 16 K-rows of 128 bytes: exactly the MN-major
//     SWIZZLE_128B UMMA layout (no im2col copy, no shared-memory shuffling).  The kw horizontal taps are
//     three boxes whose W coordinate is shifted by j; overlapping windows are served by L2.
//   * A operand (weights): (Cout x K) re-ordered to k = (j, group, c, i), written once per CTA into
//     K-major SWIZZLE_128B tiles that stay resident in shared memory for the whole kernel.
//   * D lives in TMEM (128 lanes = Cout, 256 columns = pixels), double buffered; the epilogue warps read
//     it with tcgen05.ld (thread = output channel, registers = consecutive pixels), add bias / ReLU,
//     stage bf16 rows in swizzled shared memory and write NCHW rows with fully coalesced stores.
//   A tile is 4 chunks of 64 pixels (256 pixels); one CTA per SM loops over tiles (persistent).
#include <stdlib.h>

#include "nk_internal.cuh"
#include "nk_ptx.cuh"

int nk_conv2d_bwd_kernel_tc(nk_ctx*, void*, int, void*, const void*, const void*, int64_t, int64_t, int64_t, int64_t,
                            int64_t, int64_t, int64_t, float) {
  return NK_ERR_UNSUPPORTED;
}
int nk_conv2d_bwd_input_tc(nk_ctx*, void*, const void*, const void*, int64_t, int64_t, int64_t, int64_t, int64_t,
                           int64_t, int64_t, float) {
  return NK_ERR_UNSUPPORTED;
}

namespace {

constexpr int kThreads = 320;      // warp 0: TMA, warp 1: MMA + TMEM, warps 2..9: epilogue (2 per TMEM lane quarter)
constexpr int kEpiThreads = 256;
constexpr int kChunk = 64;         // pixels per chunk (128 bytes of bf16)
constexpr int kChunksPerTile = 4;  // UMMA N = 256
constexpr int kStageBytes = kChunksPerTile * 16 * 128;  // 4 chunks x 16 K-rows x 128 B = 8 KB
constexpr int kStages = 8;
constexpr int kMaxKBlocks = 5;     // resident weight tiles of 16 KB (K <= 320)

struct ConvP {
  int n, cin, h, w, cout, kh, kw, ho, wo;
  int cpr, chunks_per_img, tiles_per_img, num_tiles;
  int cpg, ng, ksteps, kblocks;
  int vec;  // store vector width in elements (1, 2, 4 or 8), from the alignment of Wo
  const __nv_bfloat16* wt;
  const __nv_bfloat16* bias;
  __nv_bfloat16* y;
  int relu;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(kThreads, 1)
conv_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const ConvP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  // layout: [weights kblocks x 16 KB][stages x 8 KB][staging 2 x 4 chunks x rows x 128 B][barriers]
  const int rows = (p.cout + 31) & ~31;                     // staged output rows (multiple of 32)
  const uint32_t w_off = 0;
  const uint32_t st_off = w_off + p.kblocks * 16384;
  const uint32_t sg_off = st_off + kStages * kStageBytes;
  const uint32_t sg_buf_bytes = kChunksPerTile * rows * 128;
  const uint32_t bar_off = sg_off + 2 * sg_buf_bytes;
  const uint32_t bar_base = base + bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + bar_off + 8u * (2 * kStages + 4));

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup: zero the stage ring (K-rows the boxes never write must be 0), weights -> smem
  {
    uint4* z = reinterpret_cast<uint4*>(base_ptr + st_off);
    for (int i = threadIdx.x; i < kStages * kStageBytes / 16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
    uint4* wz = reinterpret_cast<uint4*>(base_ptr + w_off);
    for (int i = threadIdx.x; i < p.kblocks * 16384 / 16; i += kThreads) wz[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  {
    // A[co][kk], kk = ks*16 + (cl*kh + i), ks = j*ng + grp, channel c = grp*cpg + cl
    const int per_co = p.cin * p.kh * p.kw;
    for (int idx = threadIdx.x; idx < p.cout * per_co; idx += kThreads) {
      const int co = idx / per_co, r = idx - co * per_co;
      const int c = r / (p.kh * p.kw), i = (r / p.kw) % p.kh, j = r % p.kw;
      const int grp = c / p.cpg, cl = c - grp * p.cpg;
      const int kk = (j * p.ng + grp) * 16 + cl * p.kh + i;
      const int blk = kk >> 6, col = kk & 63;
      const uint32_t off = blk * 16384 + co * 128 + (((col >> 3) ^ (co & 7)) << 4) + (col & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>(base_ptr + w_off + off) = p.wt[idx];
    }
  }
  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tmem_full_bar(s), 1);
      ptx::mbar_init(tmem_empty_bar(s), kEpiThreads / 32);
    }
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();  // generic-proxy smem writes (weights, zeros) -> visible to the async proxy (UMMA/TMA)
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const uint32_t box_bytes = 64u * 2u * p.kh * p.cpg;

  if (warp_idx == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n = tile / p.tiles_per_img;
        const int g0 = (tile - n * p.tiles_per_img) * kChunksPerTile;
        for (int ks = 0; ks < p.ksteps; ++ks) {
          const int j = ks / p.ng, grp = ks - j * p.ng;
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          ptx::mbar_expect_tx(full_bar(stage), kChunksPerTile * box_bytes);
          const uint32_t sb = base + st_off + stage * kStageBytes;
#pragma unroll
          for (int c = 0; c < kChunksPerTile; ++c) {
            const int g = g0 + c;
            int prow = p.h, qcol = 0;  // fully out of bounds -> zero fill, still counts the bytes
            if (g < p.chunks_per_img) {
              prow = g / p.cpr;
              qcol = (g - prow * p.cpr) * kChunk;
            }
            ptx::tma_load_4d(sb + c * 2048, &tmap_x, full_bar(stage), qcol + j, prow, grp * p.cpg, n);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, 256, /*A MN-major*/ false, /*B MN-major*/ true);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(tmem_empty_bar(as), aphase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(as * 256);
        for (int ks = 0; ks < p.ksteps; ++ks) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint64_t adesc =
              ptx::make_smem_desc_sw128(base + w_off + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
          const uint64_t bdesc = ptx::make_smem_desc_sw128(base + st_off + stage * kStageBytes, 2048, 1024);
          ptx::mma_f16_ss(tmem_d, adesc, bdesc, idesc, ks != 0 ? 1u : 0u);
          ptx::mma_commit(empty_bar(stage));
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        ptx::mma_commit(tmem_full_bar(as));
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else {
    // ===================================================== epilogue: TMEM -> regs -> swizzled smem -> NCHW rows
    const int q = warp_idx & 3;          // TMEM lane quarter this warp may read
    const int ew = warp_idx - 2;         // 0..7 rank among epilogue warps (row striping of the copy-out)
    const int c_begin = (ew >> 2) * 2;   // the two warps of a quarter split the tile's chunks: {0,1} and {2,3}
    const int co = q * 32 + lane;
    const bool row_active = co < rows;
    const float bias_v = (p.bias && co < p.cout) ? __bfloat162float(p.bias[co]) : 0.f;
    int as = 0;
    uint32_t aphase = 0;
    int buf = 0;
    const int64_t plane = int64_t(p.ho) * p.wo;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int n = tile / p.tiles_per_img;
      const int g0 = (tile - n * p.tiles_per_img) * kChunksPerTile;
      uint8_t* sg = base_ptr + sg_off + buf * sg_buf_bytes;
      // the copy-out of the tile that used this buffer two tiles ago finished before the barrier at its end
      ptx::mbar_wait(tmem_full_bar(as), aphase);
      ptx::tc_fence_after();
      if (q * 32 < rows) {
#pragma unroll 1
        for (int c = c_begin; c < c_begin + 2; ++c) {
          if (g0 + c >= p.chunks_per_img) break;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * 256 + c * 64 + half * 32), r);
            ptx::tmem_ld_wait();
            if (row_active) {
              uint8_t* rowp = sg + (c * rows + co) * 128;
#pragma unroll
              for (int v = 0; v < 4; ++v) {  // 4 x 16 bytes = 32 pixels
                uint32_t w4[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float a = __uint_as_float(r[v * 8 + e * 2]) + bias_v;
                  float b = __uint_as_float(r[v * 8 + e * 2 + 1]) + bias_v;
                  if (p.relu) {
                    a = a > 0.f ? a : 0.f;
                    b = b > 0.f ? b : 0.f;
                  }
                  w4[e] = pack_bf16(a, b);
                }
                const int chunk16 = (half * 4 + v) ^ (co & 7);
                *reinterpret_cast<uint4*>(rowp + chunk16 * 16) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tmem_empty_bar(as));
      asm volatile("bar.sync 1, 256;" ::: "memory");  // staging complete (epilogue warps only)
      // ---- coalesced copy-out: each warp takes rows ew, ew+4, ...; a row is 64 pixels = 128 B
      for (int c = 0; c < kChunksPerTile; ++c) {
        const int g = g0 + c;
        if (g >= p.chunks_per_img) break;
        const int prow = g / p.cpr, q0 = (g - prow * p.cpr) * kChunk;
        const int valid = min(kChunk, p.wo - q0);
        for (int rr = ew; rr < p.cout; rr += kEpiThreads / 32) {
          const uint8_t* rowp = sg + (c * rows + rr) * 128;
          __nv_bfloat16* dst = p.y + (int64_t(n) * p.cout + rr) * plane + int64_t(prow) * p.wo + q0;
          if (p.vec >= 2) {
            // lane handles pixels 2*lane, 2*lane+1 (4 bytes); dst is 4-byte aligned because Wo is even
            const int px = lane * 2;
            const int c16 = (px >> 3) ^ (rr & 7);
            const uint32_t v = *reinterpret_cast<const uint32_t*>(rowp + c16 * 16 + (px & 7) * 2);
            if (px + 1 < valid)
              *reinterpret_cast<uint32_t*>(dst + px) = v;
            else if (px < valid)
              dst[px] = *reinterpret_cast<const __nv_bfloat16*>(&v);
          } else {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int px = lane + hh * 32;
              const int c16 = (px >> 3) ^ (rr & 7);
              if (px < valid) dst[px] = *reinterpret_cast<const __nv_bfloat16*>(rowp + c16 * 16 + (px & 7) * 2);
            }
          }
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // buffer may be overwritten two tiles from now
      buf ^= 1;
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

int nk_conv2d_fwd_tc(nk_ctx* ctx, void* y, const void* x, const void* w, const void* bias, int relu, int64_t n,
                     int64_t cin, int64_t h, int64_t wd, int64_t cout, int64_t kh, int64_t kw) {
  if (getenv("NK_CONV_DIRECT")) return NK_ERR_UNSUPPORTED;
  if (!ctx->encode_tiled) return NK_ERR_UNSUPPORTED;
  // TMA addressing: 16-byte aligned base, every global stride a multiple of 16 bytes  => W % 8 == 0
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (wd % 8) != 0) return NK_ERR_UNSUPPORTED;
  if (kh > 16 || kh < 1 || kw < 1 || cout > 128 || cout < 1) return NK_ERR_UNSUPPORTED;
  if (n > 65535 || h > (1 << 20) || wd > (1 << 20)) return NK_ERR_UNSUPPORTED;
  ConvP p;
  p.n = (int)n, p.cin = (int)cin, p.h = (int)h, p.w = (int)wd, p.cout = (int)cout, p.kh = (int)kh, p.kw = (int)kw;
  p.ho = int(h - kh + 1), p.wo = int(wd - kw + 1);
  p.cpg = 16 / p.kh;
  if (p.cpg > p.cin) p.cpg = p.cin;
  p.ng = (p.cin + p.cpg - 1) / p.cpg;
  p.ksteps = p.kw * p.ng;
  p.kblocks = (p.ksteps * 16 + 63) / 64;
  if (p.kblocks > kMaxKBlocks) return NK_ERR_UNSUPPORTED;
  p.cpr = (p.wo + kChunk - 1) / kChunk;
  p.chunks_per_img = p.ho * p.cpr;
  p.tiles_per_img = (p.chunks_per_img + kChunksPerTile - 1) / kChunksPerTile;
  const int64_t nt = int64_t(p.n) * p.tiles_per_img;
  if (nt > (int64_t(1) << 30)) return NK_ERR_UNSUPPORTED;
  p.num_tiles = (int)nt;
  p.vec = (p.wo % 2 == 0 && (reinterpret_cast<uintptr_t>(y) & 3) == 0) ? 2 : 1;
  p.wt = static_cast<const __nv_bfloat16*>(w);
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.y = static_cast<__nv_bfloat16*>(y);
  p.relu = relu;
  const int rows = (p.cout + 31) & ~31;
  const size_t smem = 1024 + size_t(p.kblocks) * 16384 + size_t(kStages) * kStageBytes +
                      2 * size_t(kChunksPerTile) * rows * 128 + 512;
  if (smem > 232448) return NK_ERR_UNSUPPORTED;

  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)wd, (cuuint64_t)h, (cuuint64_t)cin, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)wd * 2, (cuuint64_t)wd * h * 2, (cuuint64_t)wd * h * cin * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)p.kh, (cuuint32_t)p.cpg, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      &tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return NK_ERR_UNSUPPORTED;

  static bool attr_done = false;
  if (!attr_done) {
    NK_CUDA(ctx, cudaFuncSetAttribute(conv_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_done = true;
  }
  const int grid = p.num_tiles < ctx->sm_count ? p.num_tiles : ctx->sm_count;
  conv_fwd_tc_kernel<<<grid, kThreads, smem, ctx->stream>>>(tm, p);
  NK_LAUNCHED(ctx, "conv_fwd_tc");
  ctx->last_conv_kernel = "tcgen05_implicit_gemm_fwd";
  return NK_OK;
}
