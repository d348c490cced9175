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


namespace {

// warp roles of the forward kernel
constexpr int kThreads = 448;      // warp 0: TMA, warp 1: MMA + TMEM, warps 2..5: shift, warps 6..13: epilogue
constexpr int kShiftWarp0 = 2, kShiftThreads = 128;
constexpr int kEpiWarp0 = 6, kEpiThreads = 256;
constexpr int kChunk = 64;         // pixels per chunk (128 bytes of bf16)
constexpr int kChunksPerTile = 4;  // UMMA N = 256
constexpr int kSlotBytes = kChunksPerTile * 16 * 128;   // one tap: 4 chunks x 16 K-rows x 128 B = 8 KB
constexpr int kHaloBytes = kChunksPerTile * 256;        // 4 chunks x 16 rows x 16 B
constexpr int kMaxKBlocks = 6;     // resident weight tiles of 16 KB (K <= 384)

struct ConvP {
  int n, cin, h, w, cout, kh, kw, ho, wo;
  int cpr, chunks_per_img, tiles_per_img, num_tiles;
  int cpg, ng, ksteps, kblocks, R, stages;
  int vec;  // store vector width in elements (1 or 2), from the alignment of Wo
  const __nv_bfloat16* wt;
  const __nv_bfloat16* bias;
  __nv_bfloat16* y;
  int relu;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// byte offset of 32-bit word `w` of row `r` in a [rows][128 B] SWIZZLE_128B tile (1024-byte aligned base)
__device__ __forceinline__ uint32_t sw128_word(int r, int w) {
  return uint32_t(r) * 128u + (uint32_t(((w >> 2) ^ (r & 7))) << 4) + uint32_t(w & 3) * 4u;
}

// TMA can only start a box on a 16-byte boundary (profiles/r01_tma_alignment_probe.md), so the horizontal taps
// j >= 1 are produced here: out[px] = in[px + j] for the R live K-rows of each chunk, reading the tap-0 window
// (swizzled, written by TMA) plus the 8-pixel halo box, writing the swizzled tap-j slot.
// kw <= 3 fast path: one warp instruction moves TWO K-rows (lanes 0-15 / 16-31), 8 bytes per lane
__device__ __forceinline__ void shift_taps_kw3(uint8_t* stage_ptr, int kw, int R, int t) {
  const uint8_t* halo = stage_ptr + kw * kSlotBytes;
  const int lane = t & 31, wrp = t >> 5, l16 = lane & 15, sub = lane >> 4;
  const int pairs = (R + 1) >> 1;
  for (int c = 0; c < kChunksPerTile; ++c)
    for (int pr = wrp; pr < pairs; pr += kShiftThreads / 32) {
      const int r = pr * 2 + sub;
      const bool live = r < R;
      const int rr = live ? r : 0;
      // words 2*l16, 2*l16+1 of the row: same 16-byte chunk, so the pair stays contiguous under the swizzle
      const uint32_t off = c * 2048 + sw128_word(rr, 2 * l16);
      const uint2 own = *reinterpret_cast<const uint2*>(stage_ptr + off);
      const uint32_t h0 = *reinterpret_cast<const uint32_t*>(halo + c * 256 + rr * 16);
      uint32_t n0 = __shfl_down_sync(0xffffffffu, own.x, 1);  // word 2*l16 + 2 lives in the next lane
      if (l16 == 15) n0 = h0;
      if (live) {
        uint2 o1;
        o1.x = __funnelshift_r(own.x, own.y, 16);
        o1.y = __funnelshift_r(own.y, n0, 16);
        *reinterpret_cast<uint2*>(stage_ptr + kSlotBytes + off) = o1;
        if (kw > 2) *reinterpret_cast<uint2*>(stage_ptr + 2 * kSlotBytes + off) = make_uint2(own.y, n0);
      }
    }
}

__device__ __forceinline__ void shift_taps(uint8_t* stage_ptr, int kw, int R, int t) {
  const uint8_t* halo = stage_ptr + kw * kSlotBytes;
  const int lane = t & 31, wrp = t >> 5;
  // one warp per (chunk, row): lane = 32-bit word of the 128-byte row; word 32.. come from the halo box
  for (int c = 0; c < kChunksPerTile; ++c)
  for (int r = wrp; r < R; r += kShiftThreads / 32) {
    const uint8_t* rawrow = stage_ptr + c * 2048;
    const uint32_t own = *reinterpret_cast<const uint32_t*>(rawrow + sw128_word(r, lane));
    const uint32_t hl = *reinterpret_cast<const uint32_t*>(halo + c * 256 + r * 16 + (lane & 3) * 4);  // halo word lane&3
    const uint32_t dst_off = c * 2048 + sw128_word(r, lane);
    for (int j = 1; j < kw; ++j) {
      const int ws = j >> 1;
      // word(lane + ws) and word(lane + ws + 1) of the 36-word extended row
      const int i0 = lane + ws, i1 = i0 + 1;
      uint32_t lo = __shfl_sync(0xffffffffu, own, i0 & 31);
      uint32_t hi = __shfl_sync(0xffffffffu, own, i1 & 31);
      const uint32_t h0 = __shfl_sync(0xffffffffu, hl, i0 & 3), h1 = __shfl_sync(0xffffffffu, hl, i1 & 3);
      if (i0 >= 32) lo = h0;
      if (i1 >= 32) hi = h1;
      const uint32_t out = (j & 1) ? __funnelshift_r(lo, hi, 16) : lo;
      *reinterpret_cast<uint32_t*>(stage_ptr + j * kSlotBytes + dst_off) = out;
    }
  }
}

template <bool kBias, bool kRelu>
__global__ void __launch_bounds__(kThreads, 1)
conv_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_halo,
                   const ConvP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  // layout: [weights kblocks x 16 KB][stages x (kw tap slots + halo)][staging 4 chunks x rows x 128 B][barriers]
  const int rows = (p.cout + 31) & ~31;
  const uint32_t stage_bytes = p.kw * kSlotBytes + kHaloBytes;
  const uint32_t st_off = p.kblocks * 16384;
  const uint32_t sg_off = st_off + p.stages * stage_bytes;
  const uint32_t sg_bytes = kChunksPerTile * rows * 128;
  const uint32_t bar_off = sg_off + 2 * sg_bytes;
  const uint32_t bar_base = base + bar_off;
  const int S = p.stages;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto ready_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * S + s); };
  auto tmem_full_bar = [&](int s) { return bar_base + 8u * (3 * S + s); };
  auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (3 * S + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * S + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + bar_off + 8u * (3 * S + 4));

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup: zero weights + stage ring (K-rows no box ever writes must be 0), then weights -> smem
  {
    uint4* z = reinterpret_cast<uint4*>(base_ptr);
    for (uint32_t i = threadIdx.x; i < sg_off / 16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
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
      *reinterpret_cast<__nv_bfloat16*>(base_ptr + off) = p.wt[idx];
    }
  }
  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_halo);
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(ready_bar(s), kShiftThreads / 32);
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

  if (warp_idx == 0) {
    // ===================================================== TMA producer: tap-0 windows + 8-pixel halos
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = kChunksPerTile * (64u + 8u) * 2u * p.R;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n = tile / p.tiles_per_img;
        const int g0 = (tile - n * p.tiles_per_img) * kChunksPerTile;
        for (int grp = 0; grp < p.ng; ++grp) {
          ptx::mbar_wait_relaxed(empty_bar(stage), phase ^ 1u);
          ptx::mbar_expect_tx(full_bar(stage), tx);
          const uint32_t sb = base + st_off + stage * stage_bytes;
#pragma unroll
          for (int c = 0; c < kChunksPerTile; ++c) {
            const int g = g0 + c;
            int prow = p.h, qcol = 0;  // fully out of bounds -> zero fill, still counts the bytes
            if (g < p.chunks_per_img) {
              prow = g / p.cpr;
              qcol = (g - prow * p.cpr) * kChunk;
            }
            ptx::tma_load_4d(sb + c * 2048, &tmap_x, full_bar(stage), qcol, prow, grp * p.cpg, n);
            ptx::tma_load_4d(sb + p.kw * kSlotBytes + c * 256, &tmap_halo, full_bar(stage), qcol + kChunk, prow,
                             grp * p.cpg, n);
          }
          if (++stage == S) {
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
        for (int grp = 0; grp < p.ng; ++grp) {
          ptx::mbar_wait(ready_bar(stage), phase);
          ptx::tc_fence_after();
          const uint32_t sb = base + st_off + stage * stage_bytes;
          for (int j = 0; j < p.kw; ++j) {
            const int ks = j * p.ng + grp;
            const uint64_t adesc = ptx::make_smem_desc_sw128(base + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
            const uint64_t bdesc = ptx::make_smem_desc_sw128(sb + j * kSlotBytes, 2048, 1024);
            ptx::mma_f16_ss(tmem_d, adesc, bdesc, idesc, (grp | j) != 0 ? 1u : 0u);
          }
          ptx::mma_commit(empty_bar(stage));
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
        ptx::mma_commit(tmem_full_bar(as));
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else if (warp_idx < kEpiWarp0) {
    // ===================================================== shift warps: taps j >= 1 from the tap-0 window
    const int t = threadIdx.x - kShiftWarp0 * 32;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      for (int grp = 0; grp < p.ng; ++grp) {
        ptx::mbar_wait(full_bar(stage), phase);
        if (p.kw <= 3)
          shift_taps_kw3(base_ptr + st_off + stage * stage_bytes, p.kw, p.R, t);
        else
          shift_taps(base_ptr + st_off + stage * stage_bytes, p.kw, p.R, t);
        ptx::fence_proxy_async();  // these generic-proxy writes are read by the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ready_bar(stage));
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    // ===================================================== epilogue: TMEM -> regs -> swizzled smem -> NCHW rows
    const int q = warp_idx & 3;          // TMEM lane quarter this warp may read
    const int ew = warp_idx - kEpiWarp0; // 0..7 rank among epilogue warps (row striping of the copy-out)
    const int c_begin = (ew >> 2) * 2;   // the two warps of a quarter split the tile's chunks: {0,1} and {2,3}
    const int co = q * 32 + lane;
    const bool row_active = co < rows;
    const float bias_v = (kBias && co < p.cout) ? __bfloat162float(p.bias[co]) : 0.f;
    int as = 0;
    uint32_t aphase = 0;
    const int64_t plane = int64_t(p.ho) * p.wo;
    int buf = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int n = tile / p.tiles_per_img;
      const int g0 = (tile - n * p.tiles_per_img) * kChunksPerTile;
      // double-buffered staging: the barrier of the previous tile separates this buffer's last readers (two
      // tiles ago) from the writes below
      uint8_t* sg = base_ptr + sg_off + buf * sg_bytes;
      buf ^= 1;
      ptx::mbar_wait_relaxed(tmem_full_bar(as), aphase);
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
                  float a = __uint_as_float(r[v * 8 + e * 2]);
                  float b = __uint_as_float(r[v * 8 + e * 2 + 1]);
                  if (kBias) {
                    a += bias_v;
                    b += bias_v;
                  }
                  if (kRelu) {
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
      if (lane == 0) ptx::mbar_arrive(tmem_empty_bar(as));  // accumulator drained: the next tile's MMAs may start
      asm volatile("bar.sync 1, 256;" ::: "memory");          // staging complete (epilogue warps only)
      // ---- coalesced copy-out: each warp takes rows ew, ew+8, ...; a row is 64 pixels = 128 B.
      // rr & 7 == ew & 7 for every row of this warp, so the swizzled lane offset is loop invariant.
      int prow = g0 / p.cpr, qc = g0 - prow * p.cpr;
      for (int c = 0; c < kChunksPerTile; ++c, ++qc) {
        if (g0 + c >= p.chunks_per_img) break;
        if (qc == p.cpr) {
          qc = 0;
          ++prow;
        }
        const int q0 = qc * kChunk;
        const int valid = min(kChunk, p.wo - q0);
        __nv_bfloat16* dst = p.y + (int64_t(n) * p.cout + ew) * plane + int64_t(prow) * p.wo + q0;
        const uint8_t* rowp = sg + (c * rows + ew) * 128;
        const int64_t dstep = 8 * plane;
        const bool row8 = p.vec >= 2 && (((int64_t(prow) * p.wo + q0) & 3) == 0) && ((plane & 3) == 0) &&
                          ((reinterpret_cast<uintptr_t>(p.y) & 7) == 0) && valid >= 4;
        if (row8) {
          // 8-byte aligned rows (every other output row at Wo = 222): lanes 0-15 / 16-31 move 4 pixels each of TWO
          // rows (ew + 8k, ew + 8(k+1)) per instruction
          const int l16 = lane & 15, sub = lane >> 4, px = l16 * 4;
          const uint8_t* src = rowp + sub * 1024 + ((((px >> 3) ^ (ew & 7))) << 4) + (px & 7) * 2;
          __nv_bfloat16* d4 = dst + sub * dstep + px;
          const int nrows = (p.cout - ew + 7) >> 3;
          for (int k = sub; k < nrows; k += 2, src += 2048, d4 += 2 * dstep) {
            const uint2 v2 = *reinterpret_cast<const uint2*>(src);
            if (px + 3 < valid) {
              *reinterpret_cast<uint2*>(d4) = v2;
            } else if (px < valid) {
              const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&v2);
              for (int i = 0; i < 4; ++i)
                if (px + i < valid) d4[i] = e[i];
            }
          }
        } else if (p.vec >= 2) {
          const int px = lane * 2;
          const uint8_t* src = rowp + ((((px >> 3) ^ (ew & 7))) << 4) + (px & 7) * 2;
          __nv_bfloat16* d2 = dst + px;
          if (px + 1 < valid) {
            const int nrows = (p.cout - ew + 7) >> 3;
            int k = 0;
            for (; k + 4 <= nrows; k += 4) {  // 4 independent load/store pairs in flight
              const uint32_t v0 = *reinterpret_cast<const uint32_t*>(src + (k + 0) * 1024);
              const uint32_t v1 = *reinterpret_cast<const uint32_t*>(src + (k + 1) * 1024);
              const uint32_t v2 = *reinterpret_cast<const uint32_t*>(src + (k + 2) * 1024);
              const uint32_t v3 = *reinterpret_cast<const uint32_t*>(src + (k + 3) * 1024);
              *reinterpret_cast<uint32_t*>(d2 + (k + 0) * dstep) = v0;
              *reinterpret_cast<uint32_t*>(d2 + (k + 1) * dstep) = v1;
              *reinterpret_cast<uint32_t*>(d2 + (k + 2) * dstep) = v2;
              *reinterpret_cast<uint32_t*>(d2 + (k + 3) * dstep) = v3;
            }
            for (; k < nrows; ++k)
              *reinterpret_cast<uint32_t*>(d2 + k * dstep) = *reinterpret_cast<const uint32_t*>(src + k * 1024);
          } else if (px < valid) {
            for (int rr = ew; rr < p.cout; rr += 8, src += 1024, d2 += dstep)
              *d2 = *reinterpret_cast<const __nv_bfloat16*>(src);
          }
        } else {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int px = lane + hh * 32;
            if (px < valid) {
              const uint8_t* src = rowp + ((((px >> 3) ^ (ew & 7))) << 4) + (px & 7) * 2;
              __nv_bfloat16* d1 = dst + px;
              for (int rr = ew; rr < p.cout; rr += 8, src += 1024, d1 += dstep)
                *d1 = *reinterpret_cast<const __nv_bfloat16*>(src);
            }
          }
        }
      }
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

// x as a 4-D (W, H, C, N) bf16 tensor; box {box_w, kh, cpg, 1}
int make_x_tmap(nk_ctx* ctx, CUtensorMap* tm, const void* x, int64_t n, int64_t cin, int64_t h, int64_t wd, int box_w,
                int kh, int cpg, bool swizzle) {
  cuuint64_t dims[4] = {(cuuint64_t)wd, (cuuint64_t)h, (cuuint64_t)cin, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)wd * 2, (cuuint64_t)wd * h * 2, (cuuint64_t)wd * h * cin * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_w, (cuuint32_t)kh, (cuuint32_t)cpg, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NK_OK : NK_ERR_UNSUPPORTED;
}

}  // namespace

int nk_conv2d_fwd_tc(nk_ctx* ctx, void* y, const void* x, const void* w, const void* bias, int relu, int64_t n,
                     int64_t cin, int64_t h, int64_t wd, int64_t cout, int64_t kh, int64_t kw) {
  if (ctx->conv_engine == NK_CONV_DIRECT) return NK_ERR_UNSUPPORTED;
  if (!ctx->encode_tiled) return NK_ERR_UNSUPPORTED;
  // TMA addressing: 16-byte aligned base, every global stride a multiple of 16 bytes  => W % 8 == 0
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (wd % 8) != 0) return NK_ERR_UNSUPPORTED;
  if (kh > 16 || kh < 1 || kw < 1 || kw > 8 || cout > 128 || cout < 1) return NK_ERR_UNSUPPORTED;
  if (n > 65535 || h > (1 << 20) || wd > (1 << 20)) return NK_ERR_UNSUPPORTED;
  ConvP p;
  p.n = (int)n, p.cin = (int)cin, p.h = (int)h, p.w = (int)wd, p.cout = (int)cout, p.kh = (int)kh, p.kw = (int)kw;
  p.ho = int(h - kh + 1), p.wo = int(wd - kw + 1);
  p.cpg = 16 / p.kh;
  if (p.cpg > p.cin) p.cpg = p.cin;
  p.R = p.cpg * p.kh;
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
  const size_t stage_bytes = size_t(p.kw) * kSlotBytes + kHaloBytes;
  const size_t fixed = 1024 + size_t(p.kblocks) * 16384 + 2 * size_t(kChunksPerTile) * rows * 128 + 512;
  if (fixed + 2 * stage_bytes > 232448) return NK_ERR_UNSUPPORTED;
  p.stages = int((232448 - fixed) / stage_bytes);
  if (p.stages > 6) p.stages = 6;
  const size_t smem = fixed + size_t(p.stages) * stage_bytes;

  CUtensorMap tm, tmh;
  if (make_x_tmap(ctx, &tm, x, n, cin, h, wd, 64, p.kh, p.cpg, true) != NK_OK) return NK_ERR_UNSUPPORTED;
  if (make_x_tmap(ctx, &tmh, x, n, cin, h, wd, 8, p.kh, p.cpg, false) != NK_OK) return NK_ERR_UNSUPPORTED;

  static bool attr_done[64] = {};  // the attribute is per device
  if (!attr_done[ctx->device & 63]) {
    NK_CUDA(ctx, cudaFuncSetAttribute(conv_fwd_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    NK_CUDA(ctx, cudaFuncSetAttribute(conv_fwd_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    NK_CUDA(ctx, cudaFuncSetAttribute(conv_fwd_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    NK_CUDA(ctx, cudaFuncSetAttribute(conv_fwd_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_done[ctx->device & 63] = true;
  }
  const int grid = p.num_tiles < ctx->sm_count ? p.num_tiles : ctx->sm_count;
  if (bias && relu)
    conv_fwd_tc_kernel<true, true><<<grid, kThreads, smem, ctx->stream>>>(tm, tmh, p);
  else if (bias)
    conv_fwd_tc_kernel<true, false><<<grid, kThreads, smem, ctx->stream>>>(tm, tmh, p);
  else if (relu)
    conv_fwd_tc_kernel<false, true><<<grid, kThreads, smem, ctx->stream>>>(tm, tmh, p);
  else
    conv_fwd_tc_kernel<false, false><<<grid, kThreads, smem, ctx->stream>>>(tm, tmh, p);
  NK_LAUNCHED(ctx, "conv_fwd_tc");
  ctx->last_conv_kernel = "tcgen05_implicit_gemm_fwd";
  return NK_OK;
}

// =====================================================================================================
// dW (+ dbias): dW[co][k] += sum_px G[co][px] * Xwin[k][px]      (convolution/mod.rs:191-226, beta = 1)
//
// The reference streams the 1.36 GB column matrix once per output channel (64 GEMVs, ~87 GB of reads at
// config 3).  Here G and x are each read once: per 64-pixel chunk
//   A = G chunk   [Cout rows][64 px]  K-major SWIZZLE_128B, brought in by cp.async (G rows have a 2*Wo-byte
//                 pitch, not a multiple of 16 bytes when Wo = 222, so TMA cannot address them)
//   B = X windows [(j,grp,c,i) rows][64 px] K-major SWIZZLE_128B -- the same TMA boxes as the forward pass
//   D[co][k] accumulates in TMEM across ALL chunks a CTA owns; one atomicAdd pass per CTA at the end.
// dbias rides along: row 15 of the first B slot is all ones, so D[co][15] = sum_px G[co][px].
// =====================================================================================================
namespace {

constexpr int kWThreads = 448;  // warp 0: TMA (x windows), warp 1: MMA + TMEM, warps 2..5: shift taps + final epilogue,
                                // warps 6..13: cp.async (G rows)
constexpr int kWProducers = 256;
constexpr int kWShiftThreads = 128;

struct ConvWP {
  int n, cin, h, w, cout, kh, kw, ho, wo;
  int cpr, chunks_per_img, tiles_per_img;
  long long total_tiles, tiles_per_cta;
  int cpg, ng, ksteps, ncols, stages, fuse_dbias, R, cout_pad, nacc;
  uint32_t tmem_cols;
  const __nv_bfloat16* g;
  float* scratch;
};

__global__ void __launch_bounds__(kWThreads, 1)
conv_dw_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_halo,
                  const ConvWP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  // a tile = 4 chunks of 64 pixels; per chunk: [G slot cout_pad x 128 B][X slots ksteps x 16 x 128 B][halo ng x 256 B]
  const uint32_t g_bytes = p.cout_pad * 128;
  const uint32_t x_bytes = p.ksteps * 2048;
  const uint32_t chunk_bytes = g_bytes + x_bytes + 1024;
  const uint32_t stage_bytes = kChunksPerTile * chunk_bytes;
  const uint32_t bar_off = p.stages * stage_bytes;
  const uint32_t bar_base = base + bar_off;
  const int S = p.stages;
  auto fullx_bar = [&](int s) { return bar_base + 8u * s; };
  auto ready_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * S + s); };
  const uint32_t done_bar = bar_base + 8u * (3 * S);
  const uint32_t tmem_slot = bar_base + 8u * (3 * S + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + bar_off + 8u * (3 * S + 1));
  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;

  {
    uint4* z = reinterpret_cast<uint4*>(base_ptr);
    for (uint32_t i = threadIdx.x; i < S * stage_bytes / 16; i += kWThreads) z[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  if (p.fuse_dbias) {  // ones row: X slot 0, row 15 of every chunk (no box and no shift writes rows >= R, R <= 15)
    for (int i = threadIdx.x; i < S * kChunksPerTile * 32; i += kWThreads) {
      const int sc = i / 32, wq = i % 32;
      const int s = sc / kChunksPerTile, c = sc % kChunksPerTile;
      *reinterpret_cast<uint32_t*>(base_ptr + s * stage_bytes + c * chunk_bytes + g_bytes + 15 * 128 + wq * 4) =
          0x3F803F80u;
    }
  }
  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_halo);
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(fullx_bar(s), 1);
      ptx::mbar_init(ready_bar(s), kWProducers + kWShiftThreads / 32);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(done_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) {
    ptx::tmem_alloc(tmem_slot, p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const long long t_begin = (long long)blockIdx.x * p.tiles_per_cta;
  long long t_end = t_begin + p.tiles_per_cta;
  if (t_end > p.total_tiles) t_end = p.total_tiles;

  if (warp_idx == 0) {
    if (lane == 0) {  // TMA: tap-0 windows of every channel group + their 8-pixel halos, 4 chunks per stage
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = uint32_t(kChunksPerTile) * p.ng * (64u + 8u) * 2u * p.R;
      for (long long tl = t_begin; tl < t_end; ++tl) {
        const int n = int(tl / p.tiles_per_img);
        const int g0 = int(tl - (long long)n * p.tiles_per_img) * kChunksPerTile;
        ptx::mbar_wait_relaxed(empty_bar(stage), phase ^ 1u);
        ptx::mbar_expect_tx(fullx_bar(stage), tx);
        for (int c = 0; c < kChunksPerTile; ++c) {
          const int g = g0 + c;
          int prow = p.h, q0 = 0;  // out of range chunk: zero fill
          if (g < p.chunks_per_img) {
            prow = g / p.cpr;
            q0 = (g - prow * p.cpr) * kChunk;
          }
          const uint32_t sc = base + stage * stage_bytes + c * chunk_bytes;
          for (int grp = 0; grp < p.ng; ++grp) {
            ptx::tma_load_4d(sc + g_bytes + grp * 2048, &tmap_x, fullx_bar(stage), q0, prow, grp * p.cpg, n);
            ptx::tma_load_4d(sc + g_bytes + x_bytes + grp * 256, &tmap_halo, fullx_bar(stage), q0 + kChunk, prow,
                             grp * p.cpg, n);
          }
        }
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp_idx == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(128, p.ncols, false, false);
      int stage = 0;
      uint32_t phase = 0;
      // Each UMMA here is tiny (N = ncols <= 64 columns, ~24 tensor cycles) and they all accumulate: issued into ONE
      // accumulator they serialise on the accumulate latency (measured 1.35 ms at config 3).  Round-robin over
      // `nacc` independent accumulators (summed in the epilogue) keeps the tensor pipe busy.
      uint32_t started = 0;  // bit a set once accumulator a has been written
      for (long long tl = t_begin; tl < t_end; ++tl) {
        ptx::mbar_wait(ready_bar(stage), phase);
        ptx::fence_proxy_async();  // cp.async / st.shared (generic proxy) writes -> async proxy (UMMA)
        ptx::tc_fence_after();
        for (int c = 0; c < kChunksPerTile; ++c) {
          const uint32_t sc = base + stage * stage_bytes + c * chunk_bytes;
          const uint64_t adesc = ptx::make_smem_desc_sw128(sc, 16, 1024);
          const uint64_t bdesc = ptx::make_smem_desc_sw128(sc + g_bytes, 16, 1024);
#pragma unroll
          for (int kq = 0; kq < 4; ++kq) {
            const uint32_t a = uint32_t(c * 4 + kq) & uint32_t(p.nacc - 1);
            ptx::mma_f16_ss(tmem_base + a * uint32_t(p.ncols), adesc + uint64_t(kq * 2), bdesc + uint64_t(kq * 2), idesc,
                            (started >> a) & 1u);
            started |= 1u << a;
          }
        }
        ptx::mma_commit(empty_bar(stage));
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
      ptx::mma_commit(done_bar);
    }
  } else if (warp_idx >= 6) {
    // ---- G loaders: cp.async 4-byte pieces of the G rows into the swizzled K-major A slots (8 warps)
    const int t = threadIdx.x - 6 * 32;  // 0..255
    const int piece = t & 31, co0 = t >> 5;   // rows co0, co0 + 8, ...  ((co0 + 8k) & 7 == co0 & 7)
    const uint32_t sw = (((piece >> 2) ^ (co0 & 7)) << 4) + (piece & 3) * 4;
    int stage = 0;
    uint32_t phase = 0;
    const long long plane = (long long)p.ho * p.wo;
    for (long long tl = t_begin; tl < t_end; ++tl) {
      const int n = int(tl / p.tiles_per_img);
      const int g0 = int(tl - (long long)n * p.tiles_per_img) * kChunksPerTile;
      ptx::mbar_wait_relaxed(empty_bar(stage), phase ^ 1u);
      const uint32_t sa = base + stage * stage_bytes;
      int prow = g0 / p.cpr, qc = g0 - prow * p.cpr;
      for (int c = 0; c < kChunksPerTile; ++c, ++qc) {
        if (qc == p.cpr) {
          qc = 0;
          ++prow;
        }
        const bool live = g0 + c < p.chunks_per_img;
        const int q0 = qc * kChunk;
        const int valid = live ? min(kChunk, p.wo - q0) : 0;
        int nbytes = (valid - piece * 2) * 2;
        nbytes = nbytes < 0 ? 0 : (nbytes > 4 ? 4 : nbytes);
        const __nv_bfloat16* src = p.g + (long long)n * p.cout * plane + (long long)co0 * plane +
                                   (nbytes ? (long long)prow * p.wo + q0 + piece * 2 : 0);
        uint32_t dst = sa + c * chunk_bytes + co0 * 128 + sw;
        const long long sstep = 8 * plane;
#pragma unroll 8
        for (int co = co0; co < p.cout; co += 8, src += sstep, dst += 1024)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(ready_bar(stage)) : "memory");
      if (++stage == S) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else {
    // ---- shift warps: taps slot(j*ng + grp)[r][px] = slot(grp)[r][px + j] (tap-0 window + 8-pixel halo, see shift_taps)
    const int wrp = warp_idx - 2;
    int stage = 0;
    uint32_t phase = 0;
    for (long long tl = t_begin; tl < t_end; ++tl) {
      uint8_t* sp = base_ptr + stage * stage_bytes;
      ptx::mbar_wait_relaxed(fullx_bar(stage), phase);
      for (int c = 0; c < kChunksPerTile; ++c) {
        uint8_t* xs = sp + c * chunk_bytes + g_bytes;
        const uint8_t* halo = xs + x_bytes;
        for (int grp = 0; grp < p.ng; ++grp)
          for (int r = wrp; r < p.R; r += kWShiftThreads / 32) {
            const uint32_t own = *reinterpret_cast<const uint32_t*>(xs + grp * 2048 + sw128_word(r, lane));
            const uint32_t hl = *reinterpret_cast<const uint32_t*>(halo + grp * 256 + r * 16 + (lane & 3) * 4);
            const uint32_t doff = sw128_word(r, lane);
            for (int j = 1; j < p.kw; ++j) {
              const int i0 = lane + (j >> 1), i1 = i0 + 1;
              uint32_t lo = __shfl_sync(0xffffffffu, own, i0 & 31);
              uint32_t hi = __shfl_sync(0xffffffffu, own, i1 & 31);
              const uint32_t h0 = __shfl_sync(0xffffffffu, hl, i0 & 3), h1 = __shfl_sync(0xffffffffu, hl, i1 & 3);
              if (i0 >= 32) lo = h0;
              if (i1 >= 32) hi = h1;
              const uint32_t out = (j & 1) ? __funnelshift_r(lo, hi, 16) : lo;
              *reinterpret_cast<uint32_t*>(xs + (j * p.ng + grp) * 2048 + doff) = out;
            }
          }
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ready_bar(stage));
      if (++stage == S) {
        stage = 0;
        phase ^= 1u;
      }
    }
    // ---- epilogue: one pass of atomics per CTA
    ptx::mbar_wait(done_bar, 0);
    ptx::tc_fence_after();
    if (t_end > t_begin) {
      const int q = warp_idx & 3;
      const int co = q * 32 + lane;
      if (q * 32 < p.cout) {
        const long long ntiles = t_end - t_begin;
        const int live_acc = ntiles * 16 < p.nacc ? int(ntiles * 16) : p.nacc;  // accumulators that were written
        for (int c0 = 0; c0 < p.ncols; c0 += 16) {
          float sum[16];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) sum[jj] = 0.f;
          for (int a = 0; a < live_acc; ++a) {
            uint32_t r[16];
            ptx::tmem_ld_32x32b_x16(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(a * p.ncols + c0), r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) sum[jj] += __uint_as_float(r[jj]);
          }
          if (co < p.cout) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) atomicAdd(&p.scratch[co * p.ncols + c0 + jj], sum[jj]);
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

template <typename T>
__global__ void conv_dw_finalize(T* __restrict__ dw, T* __restrict__ dbias, const float* __restrict__ scratch, int cout,
                                 int cin, int kh, int kw, int cpg, int ng, int ncols, float beta, int fuse_dbias) {
  const int per_co = cin * kh * kw;
  const int total = cout * per_co;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total + cout; idx += gridDim.x * blockDim.x) {
    if (idx < total) {
      const int co = idx / per_co, r = idx - co * per_co;
      const int c = r / (kh * kw), i = (r / kw) % kh, j = r % kw;
      const int grp = c / cpg, cl = c - grp * cpg;
      float v = scratch[co * ncols + (j * ng + grp) * 16 + cl * kh + i];
      if (beta != 0.f) v += beta * nk_to_f32<T>(dw[idx]);
      dw[idx] = nk_from_f32<T>(v);
    } else if (fuse_dbias && dbias) {
      const int co = idx - total;
      float v = scratch[co * ncols + 15];
      if (beta != 0.f) v += beta * nk_to_f32<T>(dbias[co]);
      dbias[co] = nk_from_f32<T>(v);
    }
  }
}

}  // namespace

int nk_conv2d_bwd_kernel_tc(nk_ctx* ctx, void* dwt, int dw_dtype, void* dbias, const void* g, const void* x,
                            int64_t n, int64_t cin, int64_t h, int64_t wd, int64_t cout, int64_t kh, int64_t kw,
                            float beta) {
  if (ctx->conv_engine == NK_CONV_DIRECT) return NK_ERR_UNSUPPORTED;
  if (!ctx->encode_tiled) return NK_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (wd % 8) != 0) return NK_ERR_UNSUPPORTED;
  if (kh > 16 || kh < 1 || kw < 1 || kw > 8 || cout > 128 || cout < 1 || n > 65535) return NK_ERR_UNSUPPORTED;
  ConvWP p;
  p.n = (int)n, p.cin = (int)cin, p.h = (int)h, p.w = (int)wd, p.cout = (int)cout, p.kh = (int)kh, p.kw = (int)kw;
  p.ho = int(h - kh + 1), p.wo = int(wd - kw + 1);
  // cp.async moves 4-byte pieces of G rows: rows must start 4-byte aligned
  if ((p.wo & 1) || (reinterpret_cast<uintptr_t>(g) & 3)) return NK_ERR_UNSUPPORTED;
  p.cpg = 16 / p.kh;
  if (p.cpg > p.cin) p.cpg = p.cin;
  p.R = p.cpg * p.kh;
  p.ng = (p.cin + p.cpg - 1) / p.cpg;
  p.ksteps = p.kw * p.ng;
  p.ncols = p.ksteps * 16;
  if (p.ncols > 256 || p.ng > 4) return NK_ERR_UNSUPPORTED;
  p.fuse_dbias = (dbias != nullptr && p.kh * p.cpg < 16) ? 1 : 0;
  // ONE accumulator: tools/umma_probe.cu (profiles/r01_umma_issue_probe.txt) measures 77 cycles per UMMA when consecutive
  // instructions accumulate into the same TMEM columns and 219 when they alternate between accumulators -- the
  // round-robin scheme tried first made every instruction pay the switch
  p.nacc = 1;
  p.tmem_cols = 512;
  p.cpr = (p.wo + kChunk - 1) / kChunk;
  p.chunks_per_img = p.ho * p.cpr;
  p.tiles_per_img = (p.chunks_per_img + kChunksPerTile - 1) / kChunksPerTile;
  p.total_tiles = (long long)p.n * p.tiles_per_img;
  p.cout_pad = (p.cout + 7) & ~7;
  const uint32_t stage_bytes = kChunksPerTile * (p.cout_pad * 128 + p.ksteps * 2048 + 1024);
  // + 16 KB tail: the M = 128 UMMA reads 128 A rows although only cout_pad are loaded (lanes >= Cout are unused)
  p.stages = int((232448 - 2048 - 16384) / stage_bytes);
  if (p.stages > 4) p.stages = 4;
  if (p.stages < 2) return NK_ERR_UNSUPPORTED;
  int grid = ctx->sm_count;
  if (p.total_tiles < grid) grid = (int)p.total_tiles;
  p.tiles_per_cta = (p.total_tiles + grid - 1) / grid;
  grid = int((p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta);
  p.g = static_cast<const __nv_bfloat16*>(g);
  float* scratch;
  int rc = nk_workspace(ctx, size_t(128) * p.ncols * sizeof(float), (void**)&scratch);
  if (rc) return rc;
  p.scratch = scratch;

  CUtensorMap tm, tmh;
  if (make_x_tmap(ctx, &tm, x, n, cin, h, wd, 64, p.kh, p.cpg, true) != NK_OK) return NK_ERR_UNSUPPORTED;
  if (make_x_tmap(ctx, &tmh, x, n, cin, h, wd, 8, p.kh, p.cpg, false) != NK_OK) return NK_ERR_UNSUPPORTED;

  if (dbias && !p.fuse_dbias) {
    int64_t dshape[3] = {cout, 1, 1};
    int64_t gshape[4] = {n, cout, p.ho, p.wo};
    rc = nk_unbroadcast_acc(ctx, dbias, dw_dtype, 3, dshape, g, NK_BF16, 4, gshape, beta);
    if (rc) return rc;
    rc = nk_workspace(ctx, size_t(128) * p.ncols * sizeof(float), (void**)&scratch);
    if (rc) return rc;
    p.scratch = scratch;
  }
  NK_CUDA(ctx, cudaMemsetAsync(scratch, 0, size_t(128) * p.ncols * sizeof(float), ctx->stream));
  static bool attr_done[64] = {};  // the attribute is per device
  if (!attr_done[ctx->device & 63]) {
    NK_CUDA(ctx, cudaFuncSetAttribute(conv_dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_done[ctx->device & 63] = true;
  }
  const size_t smem = 1024 + size_t(p.stages) * stage_bytes + 512 + 16384;
  conv_dw_tc_kernel<<<grid, kWThreads, smem, ctx->stream>>>(tm, tmh, p);
  NK_LAUNCHED(ctx, "conv_dw_tc");
  const int total = int(cout * cin * kh * kw + cout);
  const int blocks = (total + 255) / 256;
  if (dw_dtype == NK_BF16)
    conv_dw_finalize<__nv_bfloat16><<<blocks, 256, 0, ctx->stream>>>((__nv_bfloat16*)dwt, (__nv_bfloat16*)dbias, scratch, p.cout, p.cin, p.kh, p.kw, p.cpg, p.ng, p.ncols, beta, p.fuse_dbias);
  else
    conv_dw_finalize<float><<<blocks, 256, 0, ctx->stream>>>((float*)dwt, (float*)dbias, scratch, p.cout, p.cin, p.kh, p.kw, p.cpg, p.ng, p.ncols, beta, p.fuse_dbias);
  NK_LAUNCHED(ctx, "conv_dw_finalize");
  ctx->last_conv_kernel = "tcgen05_implicit_gemm_dw";
  return NK_OK;
}


// =====================================================================================================
// dX: dx[n,c,u,v] += sum_{o,i,j} G[n,o,u-i,v-j] * W[o,c,i,j]      (convolution/mod.rs:146-189, intended maths)
//
// The reference builds the (N, K, L) column buffer with one sgemm per sample and scatters it back with
// overlapping strided adds (col2im).  Here, per output-gradient row (n, p):
//   D[m = (c,i,j)][px] = sum_o Wt[m][o] * G[o][px]        one UMMA chain, M = Cin*kh*kw (<= 32), N = 256 pixels,
//                                                          K = Cout; B = the G row tile exactly as it lies in memory
//   col2im in registers: thread v (one dx column) sums D[(c,i,j)][v-j] over j and keeps a kh-deep ring of partial
//   dx rows; after G row p, dx row u = p is complete and is written once (coalesced along W).
// G is read exactly once (cp.async: its 2*Wo-byte row pitch is not TMA-addressable), dx is written exactly once.
// A CTA owns blocks of dx rows of one image and recomputes the kh-1 halo rows of G at a block boundary.
// =====================================================================================================
namespace {

constexpr int kXThreads = 576;  // warp 0: TMEM reader, warp 1: MMA + TMEM alloc, warps 2..9: G loaders (cp.async), warps 10..17: pixels
constexpr int kXProducers = 256;  // cp.async loader threads
constexpr int kXPixelWarp0 = 10;
constexpr int kSRowFloats = 260;  // padded row of the f32 exchange buffer (conflict-free 16-byte writes per lane)

struct ConvXP {
  int n, cin, h, w, cout, ho, wo;
  int rb, blocks_per_img, num_units, stages, kblocks;
  const __nv_bfloat16* g;
  const __nv_bfloat16* wt;
  __nv_bfloat16* dx;
  float beta;
};

template <int CIN, int KH, int KW>
__global__ void __launch_bounds__(kXThreads, 1) conv_dx_tc_kernel(const ConvXP p) {
  constexpr int M = CIN * KH * KW;
  static_assert(M <= 32, "the (c,i,j) rows must fit one TMEM lane quarter");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  // layout: [weights kblocks x 16 KB][stages x (4 chunks x Cout rows x 128 B)][S: 2 x 32 rows x 1040 B][barriers]
  const uint32_t chunk_bytes = p.cout * 128;
  const uint32_t stage_bytes = kChunksPerTile * chunk_bytes;
  const uint32_t st_off = p.kblocks * 16384;
  const uint32_t s_off = st_off + p.stages * stage_bytes;
  const uint32_t s_bytes = 32 * kSRowFloats * 4;
  const uint32_t bar_off = s_off + 2 * s_bytes;
  const uint32_t bar_base = base + bar_off;
  const int S = p.stages;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * S + s); };
  auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * S + 2 + s); };
  auto s_full_bar = [&](int s) { return bar_base + 8u * (2 * S + 4 + s); };
  auto s_empty_bar = [&](int s) { return bar_base + 8u * (2 * S + 6 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 8);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + bar_off + 8u * (2 * S + 8));
  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;

  {
    uint4* z = reinterpret_cast<uint4*>(base_ptr);
    for (uint32_t i = threadIdx.x; i < s_off / 16; i += kXThreads) z[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  {
    // A[m][o] = W[o][c][i][j], m = (c*KH + i)*KW + j ; K-major SWIZZLE_128B tiles of 64 output channels
    for (int idx = threadIdx.x; idx < p.cout * M; idx += kXThreads) {
      const int o = idx / M, m = idx - o * M;
      const int blk = o >> 6, col = o & 63;
      const uint32_t off = blk * 16384 + m * 128 + (((col >> 3) ^ (m & 7)) << 4) + (col & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>(base_ptr + off) = p.wt[idx];
    }
  }
  if (warp_idx == 0 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(full_bar(s), kXProducers);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tmem_full_bar(s), 1);
      ptx::mbar_init(tmem_empty_bar(s), 1);
      ptx::mbar_init(s_full_bar(s), 1);
      ptx::mbar_init(s_empty_bar(s), 8);
    }
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // every role walks the same sequence of (unit, G row) tiles
  auto unit_rows = [&](int unit, int& n, int& u0, int& u1, int& p_lo, int& p_hi) {
    n = unit / p.blocks_per_img;
    const int b = unit - n * p.blocks_per_img;
    u0 = b * p.rb;
    u1 = min(u0 + p.rb, p.h);
    p_lo = max(0, u0 - (KH - 1));
    p_hi = min(p.ho, u1) - 1;
  };

  if (warp_idx == 1) {
    if (lane == 0) {  // ===================================================== MMA issuer
      const uint32_t idesc = ptx::make_idesc_bf16(128, 256, false, true);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
        int n, u0, u1, p_lo, p_hi;
        unit_rows(unit, n, u0, u1, p_lo, p_hi);
        for (int pr = p_lo; pr <= p_hi; ++pr) {
          ptx::mbar_wait(tmem_empty_bar(as), aphase ^ 1u);
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::fence_proxy_async();
          ptx::tc_fence_after();
          const uint32_t sb = base + st_off + stage * stage_bytes;
          const uint32_t tmem_d = tmem_base + uint32_t(as * 256);
          for (int ks = 0; ks < p.cout / 16; ++ks) {
            const uint64_t adesc = ptx::make_smem_desc_sw128(base + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
            const uint64_t bdesc = ptx::make_smem_desc_sw128(sb + ks * 2048, chunk_bytes, 1024);
            ptx::mma_f16_ss(tmem_d, adesc, bdesc, idesc, ks != 0 ? 1u : 0u);
          }
          ptx::mma_commit(empty_bar(stage));
          ptx::mma_commit(tmem_full_bar(as));
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
          as ^= 1;
          if (as == 0) aphase ^= 1u;
        }
      }
    }
  } else if (warp_idx == 0) {
    // ===================================================== TMEM reader: D rows (c,i,j) -> f32 exchange buffer
    int as = 0, sbuf = 0;
    uint32_t aphase = 0, sphase = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int n, u0, u1, p_lo, p_hi;
      unit_rows(unit, n, u0, u1, p_lo, p_hi);
      for (int pr = p_lo; pr <= p_hi; ++pr) {
        ptx::mbar_wait(tmem_full_bar(as), aphase);
        ptx::tc_fence_after();
        ptx::mbar_wait(s_empty_bar(sbuf), sphase ^ 1u);
        float* srow = reinterpret_cast<float*>(base_ptr + s_off + sbuf * s_bytes) + lane * kSRowFloats;
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += 32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(tmem_base + uint32_t(as * 256 + c0), r);
          ptx::tmem_ld_wait();
          if (lane < M) {
#pragma unroll
            for (int v = 0; v < 8; ++v)
              *reinterpret_cast<uint4*>(srow + c0 + v * 4) = make_uint4(r[v * 4], r[v * 4 + 1], r[v * 4 + 2], r[v * 4 + 3]);
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive(tmem_empty_bar(as));
          ptx::mbar_arrive(s_full_bar(sbuf));
        }
        as ^= 1;
        if (as == 0) aphase ^= 1u;
        sbuf ^= 1;
        if (sbuf == 0) sphase ^= 1u;
      }
    }
  } else if (warp_idx < kXPixelWarp0) {
    // ===================================================== G loaders: cp.async the (Cout x Wo) row tile into the B tile.
    // G rows have a 2*Wo-byte pitch (4-byte aligned only): neither TMA nor 16-byte cp.async can move them.  A plain
    // LDG.32 -> STS.32 loader was tried and is 2x slower (1.24 ms vs 0.59 ms at config 3: it is latency bound), so
    // the rows go through 4-byte cp.async, spread over 8 warps.
    const int t = threadIdx.x - 64;              // 0..255
    const int piece = t & 31, o0 = t >> 5;       // rows o0, o0 + 8, ...
    int stage = 0;
    uint32_t phase = 0;
    const long long plane = (long long)p.ho * p.wo;
    const uint32_t sw = (((piece >> 2) ^ (o0 & 7)) << 4) + (piece & 3) * 4;   // (o0 + 8k) & 7 == o0 & 7
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int n, u0, u1, p_lo, p_hi;
      unit_rows(unit, n, u0, u1, p_lo, p_hi);
      for (int pr = p_lo; pr <= p_hi; ++pr) {
        ptx::mbar_wait_relaxed(empty_bar(stage), phase ^ 1u);
        const uint32_t sb = base + st_off + stage * stage_bytes;
        const __nv_bfloat16* grow = p.g + (long long)n * p.cout * plane + (long long)pr * p.wo;
#pragma unroll
        for (int c = 0; c < kChunksPerTile; ++c) {
          int nbytes = (p.wo - c * kChunk - piece * 2) * 2;
          nbytes = nbytes < 0 ? 0 : (nbytes > 4 ? 4 : nbytes);
          const __nv_bfloat16* src = grow + (long long)o0 * plane + (nbytes ? c * kChunk + piece * 2 : 0);
          uint32_t dst = sb + c * chunk_bytes + o0 * 128 + sw;
          const long long sstep = 8 * plane;
#pragma unroll 8
          for (int o = o0; o < p.cout; o += 8, src += sstep, dst += 1024)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full_bar(stage)) : "memory");
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    // ===================================================== pixel warps: col2im in registers, one dx column each
    const int v = threadIdx.x - kXPixelWarp0 * 32;  // 0..255
    const bool col_live = v < p.w;
    int sbuf = 0;
    uint32_t sphase = 0;
    const long long img = (long long)p.h * p.w;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int n, u0, u1, p_lo, p_hi;
      unit_rows(unit, n, u0, u1, p_lo, p_hi);
      float ring[CIN][KH];
#pragma unroll
      for (int c = 0; c < CIN; ++c)
#pragma unroll
        for (int k = 0; k < KH; ++k) ring[c][k] = 0.f;
      __nv_bfloat16* dxn = p.dx + (long long)n * CIN * img + v;
      auto emit = [&](int u) {  // ring[.][0] holds the finished dx row u of every channel
        if (u >= u0 && u < u1 && col_live) {
#pragma unroll
          for (int c = 0; c < CIN; ++c) {
            __nv_bfloat16* d = dxn + (long long)c * img + (long long)u * p.w;
            float val = ring[c][0];
            if (p.beta != 0.f) val += p.beta * __bfloat162float(*d);
            *d = __float2bfloat16_rn(val);
          }
        }
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
#pragma unroll
          for (int k = 0; k + 1 < KH; ++k) ring[c][k] = ring[c][k + 1];
          ring[c][KH - 1] = 0.f;
        }
      };
      for (int pr = p_lo; pr <= p_hi; ++pr) {
        ptx::mbar_wait_relaxed(s_full_bar(sbuf), sphase);
        const float* Sb = reinterpret_cast<const float*>(base_ptr + s_off + sbuf * s_bytes);
#pragma unroll
        for (int c = 0; c < CIN; ++c)
#pragma unroll
          for (int i = 0; i < KH; ++i) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < KW; ++j) {
              const int q = v - j;
              if (q >= 0) acc += Sb[((c * KH + i) * KW + j) * kSRowFloats + q];
            }
            ring[c][i] += acc;
          }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(s_empty_bar(sbuf));
        sbuf ^= 1;
        if (sbuf == 0) sphase ^= 1u;
        emit(pr);
      }
      // rows below the last G row (only the last block of an image has them)
      for (int u = p_hi + 1; u < u1; ++u) emit(u);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

template <int CIN>
int launch_dx(nk_ctx* ctx, const ConvXP& p, int grid, size_t smem) {
  auto kern = conv_dx_tc_kernel<CIN, 3, 3>;
  static bool attr_done[64] = {};  // the attribute is per device
  if (!attr_done[ctx->device & 63]) {
    NK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_done[ctx->device & 63] = true;
  }
  kern<<<grid, kXThreads, smem, ctx->stream>>>(p);
  NK_LAUNCHED(ctx, "conv_dx_tc");
  return NK_OK;
}

}  // namespace

int nk_conv2d_bwd_input_tc(nk_ctx* ctx, void* dx, const void* g, const void* w, int64_t n, int64_t cin, int64_t h,
                           int64_t wd, int64_t cout, int64_t kh, int64_t kw, float beta) {
  if (ctx->conv_engine == NK_CONV_DIRECT) return NK_ERR_UNSUPPORTED;
  if (kh != 3 || kw != 3 || cin < 1 || cin > 3) return NK_ERR_UNSUPPORTED;       // (c,i,j) rows <= 32 TMEM lanes
  if (cout % 16 != 0 || cout > 128 || wd > 256 || n > (1 << 20)) return NK_ERR_UNSUPPORTED;
  ConvXP p;
  p.n = (int)n, p.cin = (int)cin, p.h = (int)h, p.w = (int)wd, p.cout = (int)cout;
  p.ho = int(h - kh + 1), p.wo = int(wd - kw + 1);
  if ((p.wo & 1) || (reinterpret_cast<uintptr_t>(g) & 3)) return NK_ERR_UNSUPPORTED;  // 4-byte cp.async pieces
  p.kblocks = (p.cout + 63) / 64;
  const size_t stage_bytes = size_t(kChunksPerTile) * p.cout * 128;
  const size_t fixed = 1024 + size_t(p.kblocks) * 16384 + 2 * 32 * kSRowFloats * 4 + 512;
  if (fixed + 2 * stage_bytes > 232448) return NK_ERR_UNSUPPORTED;
  p.stages = int((232448 - fixed) / stage_bytes);
  if (p.stages > 4) p.stages = 4;
  // blocks of dx rows: enough units to balance 148 SMs, large enough to amortise the kh-1 recomputed halo rows
  p.rb = 56;
  if (p.h < 2 * p.rb) p.rb = p.h;
  p.blocks_per_img = (p.h + p.rb - 1) / p.rb;
  p.num_units = p.n * p.blocks_per_img;
  p.g = static_cast<const __nv_bfloat16*>(g);
  p.wt = static_cast<const __nv_bfloat16*>(w);
  p.dx = static_cast<__nv_bfloat16*>(dx);
  p.beta = beta;
  const size_t smem = fixed + size_t(p.stages) * stage_bytes;
  const int grid = p.num_units < ctx->sm_count ? p.num_units : ctx->sm_count;
  int rc = cin == 1 ? launch_dx<1>(ctx, p, grid, smem) : cin == 2 ? launch_dx<2>(ctx, p, grid, smem) : launch_dx<3>(ctx, p, grid, smem);
  if (rc) return rc;
  ctx->last_conv_kernel = "tcgen05_implicit_gemm_dx";
  return NK_OK;
}


// =====================================================================================================
// Fused ConvolutionBackward: dW (+ dbias) and dX in ONE pass over the output gradient
//   (convolution/mod.rs:146-226: the reference's backward() runs both halves back to back, each streaming G)
//
// One cp.async pass over G feeds both UMMA chains, because the same swizzled bytes are a valid operand for both:
//   dW:  Da[co][k]        += sum_px G[co][px] * Xwin[k][px]       A = G chunk read K-major  [Cout rows][64 px]
//   dX:  Dt[px][(c,i,j)]   = sum_o  G[o][px]  * Wt[(c,i,j)][o]    A = G tile read MN-major  [M = 128 px][K = Cout]
// The dX product is computed TRANSPOSED -- pixels on the 128 TMEM lanes, the 27 (c,i,j) taps on 32 columns -- so
// that each of the 256 "pixel" threads reads its own pixel's 27 values with one tcgen05.ld and does the col2im in
// registers (neighbour pixels v-1, v-2 by warp shuffle, the two pixels across a warp boundary through 256 bytes of
// shared memory).  The first fused version kept the untransposed D[(c,i,j)][256 px] of the stand-alone dX kernel,
// needed a reader warp + an f32 exchange buffer, and with the dW accumulators resident could not double-buffer its
// 256 TMEM columns: reader and UMMAs ping-ponged (profiles/r01_conv_ncu.md).  Transposed, a row tile is 2 x 32
// columns, double buffered.
// A CTA owns blocks of dx rows of one image; the kh-1 halo rows of G it recomputes for dX are NOT added to dW (and
// their x windows are not fetched).  G is read ~1.04x, x ~1x, dx written once.
// =====================================================================================================
namespace {

constexpr int kFThreads = 768;  // warp 0: MMA + TMEM alloc, 1: TMA (x windows), 2..3: idle, 4..7: shift taps + dW epilogue,
                                // 8..15: dX pixel warps (warp 8 + w reads TMEM lane quarter w % 4), 16..23: cp.async G loaders
constexpr int kFShiftWarp0 = 4, kFPixelWarp0 = 8, kFLoadWarp0 = 16;
constexpr int kFDxCols = 64;    // TMEM columns of one dX row tile: 2 pixel halves x 32 tap columns
constexpr int kFDxBufs = 4;     // dX row tiles in flight in TMEM (4 x 64 columns; the dW accumulators sit behind them)

struct ConvBP {
  int n, cin, h, w, cout, ho, wo;
  int rb, blocks_per_img, num_units, stages, kblocks;
  int cpr, cpg, ng, ksteps, ncols, R, nacc, fuse_dbias;
  const __nv_bfloat16* g;
  const __nv_bfloat16* wt;
  __nv_bfloat16* dx;
  float beta_dx;
  float* scratch;
  int dbg;              // development only (NK_CONV_DBG bit mask, read once): 1 no dW UMMAs, 2 no dX UMMAs, 4 no col2im,
                        // 8 no tap shifting, 16 no x TMA -- wrong results, used to time the parts of the pipeline
  int use_const;        // the output gradient is one value everywhere (backward(seed) on the convolution's own output):
  uint32_t const_bits;  // the loaders synthesise the G tiles (value packed twice as bf16) instead of reading 2|G| bytes
};

template <int CIN>
__global__ void __launch_bounds__(kFThreads, 1)
conv_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_halo,
                      const ConvBP p) {
  constexpr int KH = 3, KW = 3, M = CIN * KH * KW;
  static_assert(M <= 32, "the (c,i,j) taps must fit 32 TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  // layout: [dX weights kblocks x 4 KB: B operand, 32 tap rows x 64 output channels, K-major]
  //         [stages x 4 chunks x (G slot Cout x 128 B | X slots ksteps x 2 KB | halo 1 KB)]
  //         [boundary exchange: 2 parities x 8 warps x 2 pixels x 32 floats][barriers]
  // uniform output gradient: ONE constant G tile (4 chunks) shared by every stage instead of a G slot per stage, so a
  // stage is only the x windows (28 KB instead of 61 KB at config 3: 7 rows in flight instead of 3)
  const uint32_t g_bytes = p.cout * 128;
  const uint32_t x_bytes = p.ksteps * 2048;
  const uint32_t g_in_chunk = p.use_const ? 0u : g_bytes;          // bytes of G in front of the x slots of a chunk
  const uint32_t chunk_bytes = g_in_chunk + x_bytes + 1024;
  const uint32_t stage_bytes = kChunksPerTile * chunk_bytes;
  const uint32_t gconst_off = p.kblocks * 4096;
  const uint32_t st_off = gconst_off + (p.use_const ? kChunksPerTile * g_bytes : 0u);
  const uint32_t g_stride = p.use_const ? g_bytes : chunk_bytes;   // distance between the G tiles of consecutive chunks
  const uint32_t h_off = st_off + p.stages * stage_bytes;
  const uint32_t h_bytes = 2 * 8 * 2 * 32 * 4;
  const uint32_t bar_off = h_off + h_bytes;
  const uint32_t bar_base = base + bar_off;
  const int S = p.stages;
  auto fullx_bar = [&](int s) { return bar_base + 8u * s; };
  auto ready_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * S + s); };
  auto d_full_bar = [&](int d) { return bar_base + 8u * (3 * S + d); };
  auto d_empty_bar = [&](int d) { return bar_base + 8u * (3 * S + kFDxBufs + d); };
  const uint32_t done_bar = bar_base + 8u * (3 * S + 2 * kFDxBufs), tmem_slot = done_bar + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + bar_off + 8u * (3 * S + 2 * kFDxBufs + 1));
  // waits of roles with stages of slack back off between polls (issue slots for the working warps); NK_CONV_DBG bit 32
  // times the kernel with plain hardware-suspended waits everywhere
  auto wait_slack = [&](uint32_t bar, uint32_t parity) {
    if (p.dbg & 32) ptx::mbar_wait(bar, parity);
    else ptx::mbar_wait_relaxed(bar, parity);
  };
  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;

  {
    uint4* z = reinterpret_cast<uint4*>(base_ptr);
    for (uint32_t i = threadIdx.x; i < bar_off / 16; i += kFThreads) z[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  // dX B operand: B[m][o] = W[o][c][i][j], m = (c*KH + i)*KW + j ; K-major SWIZZLE_128B, 64 output channels per block
  // (N = 32 rows; rows >= M stay zero)
  for (int idx = threadIdx.x; idx < p.cout * M; idx += kFThreads) {
    const int o = idx / M, m = idx - o * M;
    const int blk = o >> 6, col = o & 63;
    const uint32_t off = blk * 4096 + m * 128 + (((col >> 3) ^ (m & 7)) << 4) + (col & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(base_ptr + off) = p.wt[idx];
  }
  if (p.fuse_dbias) {  // ones row: X slot 0, row 15 of every chunk (no box and no shift writes rows >= R, R <= 15)
    for (int i = threadIdx.x; i < S * kChunksPerTile * 32; i += kFThreads) {
      const int sc = i / 32, wq = i % 32;
      const int s = sc / kChunksPerTile, c = sc % kChunksPerTile;
      *reinterpret_cast<uint32_t*>(base_ptr + st_off + s * stage_bytes + c * chunk_bytes + g_in_chunk + 15 * 128 + wq * 4) =
          0x3F803F80u;
    }
  }
  if (p.use_const) {
    // uniform output gradient: the G tile is the same for every row -- written ONCE into every stage (value inside the
    // row, zero beyond Wo), never fetched and never rewritten; the loader warps have nothing to do per row
    for (int i = threadIdx.x; i < kChunksPerTile * p.cout * 32; i += kFThreads) {
      const int wq = i & 31, o = (i >> 5) % p.cout, c = (i >> 5) / p.cout;
      const uint32_t val = (p.wo - c * kChunk - wq * 2) > 0 ? p.const_bits : 0u;
      *reinterpret_cast<uint32_t*>(base_ptr + gconst_off + c * g_bytes + o * 128 +
                                   ((((wq >> 2) ^ (o & 7)) << 4) + (wq & 3) * 4)) = val;
    }
  }
  if (warp_idx == 1 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_halo);
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(fullx_bar(s), 1);
      ptx::mbar_init(ready_bar(s), p.use_const ? 4 : 256 + 4);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int d = 0; d < kFDxBufs; ++d) {
      ptx::mbar_init(d_full_bar(d), 1);
      ptx::mbar_init(d_empty_bar(d), 8);   // the 8 pixel warps
    }
    ptx::mbar_init(done_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp_idx == 0) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tmem_dw = tmem_base + uint32_t(kFDxBufs * kFDxCols);   // dW accumulators behind the dX buffers

  // every role walks the same sequence of (unit, G row) tiles; G row pr feeds dW only when the unit owns it (pr >= u0)
  auto unit_rows = [&](int unit, int& n, int& u0, int& u1, int& p_lo, int& p_hi) {
    n = unit / p.blocks_per_img;
    const int b = unit - n * p.blocks_per_img;
    u0 = b * p.rb;
    u1 = min(u0 + p.rb, p.h);
    p_lo = max(0, u0 - (KH - 1));
    p_hi = min(p.ho, u1) - 1;
  };
  bool any_owned = false;

  if (warp_idx == 0) {
    if (lane == 0) {  // ===================================================== MMA issuer
      const uint32_t idesc_dx = ptx::make_idesc_bf16(128, 32, true, false);
      const uint32_t idesc_dw = ptx::make_idesc_bf16(128, p.ncols, false, false);
      int stage = 0, db = 0;
      uint32_t phase = 0, dphase = 0, started = 0, cnt = 0;
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
        int n, u0, u1, p_lo, p_hi;
        unit_rows(unit, n, u0, u1, p_lo, p_hi);
        for (int pr = p_lo; pr <= p_hi; ++pr) {
          ptx::mbar_wait(ready_bar(stage), phase);
          ptx::fence_proxy_async();  // cp.async / st.shared (generic proxy) writes -> async proxy (UMMA)
          ptx::mbar_wait(d_empty_bar(db), dphase ^ 1u);
          ptx::tc_fence_after();
          const uint32_t sb = base + st_off + stage * stage_bytes;
          const uint32_t gb = p.use_const ? base + gconst_off : sb;   // G tile of chunk c at gb + c * g_stride
          // dX (transposed): pixel half hf = chunks 2hf, 2hf+1 as the MN-major A operand (64-px atoms g_stride apart)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const uint32_t tmem_d = tmem_base + uint32_t(db * kFDxCols + hf * 32);
            for (int ks = 0; ks < ((p.dbg & 2) ? 0 : p.cout / 16); ++ks) {
              const uint64_t adesc = ptx::make_smem_desc_sw128(gb + 2 * hf * g_stride + ks * 2048, g_stride, 1024);
              const uint64_t bdesc = ptx::make_smem_desc_sw128(base + (ks >> 2) * 4096 + (ks & 3) * 32, 16, 1024);
              ptx::mma_f16_ss(tmem_d, adesc, bdesc, idesc_dx, ks != 0 ? 1u : 0u);
            }
          }
          ptx::mma_commit(d_full_bar(db));
          db = (db + 1) & (kFDxBufs - 1);
          if (db == 0) dphase ^= 1u;
          if (pr >= u0 && !(p.dbg & 1)) {
            for (int c = 0; c < p.cpr; ++c) {
              const uint64_t adesc = ptx::make_smem_desc_sw128(gb + c * g_stride, 16, 1024);
              const uint64_t bdesc = ptx::make_smem_desc_sw128(sb + c * chunk_bytes + g_in_chunk, 16, 1024);
#pragma unroll
              for (int kq = 0; kq < 4; ++kq) {
                const uint32_t a = (cnt++) & uint32_t(p.nacc - 1);
                ptx::mma_f16_ss(tmem_dw + a * uint32_t(p.ncols), adesc + uint64_t(kq * 2), bdesc + uint64_t(kq * 2),
                                idesc_dw, (started >> a) & 1u);
                started |= 1u << a;
              }
            }
          }
          ptx::mma_commit(empty_bar(stage));
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      ptx::mma_commit(done_bar);
    }
  } else if (warp_idx == 1) {
    if (lane == 0) {  // ===================================================== TMA: tap-0 x windows + halos of owned rows
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = uint32_t(p.cpr) * p.ng * (64u + 8u) * 2u * p.R;
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
        int n, u0, u1, p_lo, p_hi;
        unit_rows(unit, n, u0, u1, p_lo, p_hi);
        for (int pr = p_lo; pr <= p_hi; ++pr) {
          wait_slack(empty_bar(stage), phase ^ 1u);
          if (pr >= u0 && !(p.dbg & 16)) {
            ptx::mbar_expect_tx(fullx_bar(stage), tx);
            for (int c = 0; c < p.cpr; ++c) {
              const uint32_t sc = base + st_off + stage * stage_bytes + c * chunk_bytes + g_in_chunk;
              for (int grp = 0; grp < p.ng; ++grp) {
                ptx::tma_load_4d(sc + grp * 2048, &tmap_x, fullx_bar(stage), c * kChunk, pr, grp * p.cpg, n);
                ptx::tma_load_4d(sc + x_bytes + grp * 256, &tmap_halo, fullx_bar(stage), (c + 1) * kChunk, pr,
                                 grp * p.cpg, n);
              }
            }
          } else {
            ptx::mbar_arrive(fullx_bar(stage));
          }
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp_idx >= kFShiftWarp0 && warp_idx < kFPixelWarp0) {
    // ===================================================== shift warps: taps j = 1, 2 of the x windows (two K-rows per
    // instruction, see shift_taps_kw3), then the dW epilogue
    const int wrp = warp_idx - kFShiftWarp0;
    const int l16 = lane & 15, sub = lane >> 4;
    const int pairs = (p.R + 1) >> 1;
    int stage = 0;
    uint32_t phase = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int n, u0, u1, p_lo, p_hi;
      unit_rows(unit, n, u0, u1, p_lo, p_hi);
      for (int pr = p_lo; pr <= p_hi; ++pr) {
        wait_slack(fullx_bar(stage), phase);
        if (pr >= u0 && !(p.dbg & 8)) {
          any_owned = true;  // at least one dW chain ran: every accumulator has been written (>= 4 UMMAs >= nacc)
          uint8_t* sp = base_ptr + st_off + stage * stage_bytes;
          for (int c = 0; c < p.cpr; ++c) {
            uint8_t* xs = sp + c * chunk_bytes + g_in_chunk;
            const uint8_t* halo = xs + x_bytes;
            for (int grp = 0; grp < p.ng; ++grp)
              for (int q = wrp; q < pairs; q += 4) {
                const int r = q * 2 + sub;
                const bool live = r < p.R;
                const int rr = live ? r : 0;
                const uint32_t off = sw128_word(rr, 2 * l16);
                const uint2 own = *reinterpret_cast<const uint2*>(xs + grp * 2048 + off);
                const uint32_t h0 = *reinterpret_cast<const uint32_t*>(halo + grp * 256 + rr * 16);
                uint32_t n0 = __shfl_down_sync(0xffffffffu, own.x, 1);
                if (l16 == 15) n0 = h0;
                if (live) {
                  uint2 o1;
                  o1.x = __funnelshift_r(own.x, own.y, 16);
                  o1.y = __funnelshift_r(own.y, n0, 16);
                  *reinterpret_cast<uint2*>(xs + (p.ng + grp) * 2048 + off) = o1;
                  *reinterpret_cast<uint2*>(xs + (2 * p.ng + grp) * 2048 + off) = make_uint2(own.y, n0);
                }
              }
          }
          ptx::fence_proxy_async();
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ready_bar(stage));
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    // ---- dW epilogue: one pass of atomics per CTA
    ptx::mbar_wait(done_bar, 0);
    ptx::tc_fence_after();
    if (any_owned) {
      const int q = warp_idx & 3;
      const int co = q * 32 + lane;
      if (q * 32 < p.cout) {
        for (int c0 = 0; c0 < p.ncols; c0 += 16) {
          float sum[16];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) sum[jj] = 0.f;
          for (int a = 0; a < p.nacc; ++a) {
            uint32_t r[16];
            ptx::tmem_ld_32x32b_x16(tmem_dw + (uint32_t(q * 32) << 16) + uint32_t(a * p.ncols + c0), r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) sum[jj] += __uint_as_float(r[jj]);
          }
          if (co < p.cout) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) atomicAdd(&p.scratch[co * p.ncols + c0 + jj], sum[jj]);
          }
        }
      }
    }
  } else if (warp_idx >= kFLoadWarp0) {
    // ===================================================== G loaders: cp.async pieces of the (Cout x Wo) row tile.
    // A G row starts 4-byte aligned in general (2*Wo-byte pitch), but when (pr*Wo) % 4 == 0 -- every other row at
    // Wo = 222 -- all Cout rows of the tile start 8-byte aligned and 8-byte pieces halve the instruction count.
    const int t = threadIdx.x - kFLoadWarp0 * 32;  // 0..255
    const int piece = t & 31, o0 = t >> 5;         // 4-byte path: rows o0, o0 + 8, ...
    const int piece8 = t & 15, o8 = t >> 4;        // 8-byte path: rows o8, o8 + 16, ...
    int stage = 0;
    uint32_t phase = 0;
    const long long plane = (long long)p.ho * p.wo;
    const uint32_t sw = (((piece >> 2) ^ (o0 & 7)) << 4) + (piece & 3) * 4;   // (o0 + 8k) & 7 == o0 & 7
    const uint32_t sw8 = (((piece8 >> 1) ^ (o8 & 7)) << 4) + (piece8 & 1) * 8;  // (o8 + 16k) & 7 == o8 & 7
    const bool base8 = ((reinterpret_cast<uintptr_t>(p.g) & 7) == 0) && ((plane & 3) == 0);
    for (int unit = p.use_const ? p.num_units : int(blockIdx.x); unit < p.num_units; unit += gridDim.x) {
      int n, u0, u1, p_lo, p_hi;
      unit_rows(unit, n, u0, u1, p_lo, p_hi);
      for (int pr = p_lo; pr <= p_hi; ++pr) {
        wait_slack(empty_bar(stage), phase ^ 1u);
        const uint32_t sb = base + st_off + stage * stage_bytes;
        const __nv_bfloat16* grow = p.g + (long long)n * p.cout * plane + (long long)pr * p.wo;
        if (base8 && (((long long)pr * p.wo) & 3) == 0) {
#pragma unroll
          for (int c = 0; c < kChunksPerTile; ++c) {
            int nbytes = (p.wo - c * kChunk - piece8 * 4) * 2;
            nbytes = nbytes < 0 ? 0 : (nbytes > 8 ? 8 : nbytes);
            const __nv_bfloat16* src = grow + (long long)o8 * plane + (nbytes ? c * kChunk + piece8 * 4 : 0);
            uint32_t dst = sb + c * chunk_bytes + o8 * 128 + sw8;
            const long long sstep = 16 * plane;
#pragma unroll 4
            for (int o = o8; o < p.cout; o += 16, src += sstep, dst += 2048)
              asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
          }
        } else {
#pragma unroll
          for (int c = 0; c < kChunksPerTile; ++c) {
            int nbytes = (p.wo - c * kChunk - piece * 2) * 2;
            nbytes = nbytes < 0 ? 0 : (nbytes > 4 ? 4 : nbytes);
            const __nv_bfloat16* src = grow + (long long)o0 * plane + (nbytes ? c * kChunk + piece * 2 : 0);
            uint32_t dst = sb + c * chunk_bytes + o0 * 128 + sw;
            const long long sstep = 8 * plane;
#pragma unroll 8
            for (int o = o0; o < p.cout; o += 8, src += sstep, dst += 1024)
              asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
          }
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(ready_bar(stage)) : "memory");
        if (++stage == S) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp_idx >= kFPixelWarp0 && warp_idx < kFLoadWarp0) {
    // ===================================================== pixel warps: thread = dx column v = TMEM lane of half v / 128.
    // One tcgen05.ld brings this pixel's 27 tap products; dx[c][p+i][v] += sum_j Dt[v-j][(c,i,j)]: the j = 1, 2 terms
    // come from lanes v-1, v-2 (shuffle), across a warp boundary from the previous warp's lanes 30, 31 (shared memory)
    const int pw = warp_idx - kFPixelWarp0;          // 0..7 ; (warp_idx & 3) == (pw & 3): the TMEM lane quarter
    const int v = pw * 32 + lane;
    const bool col_live = v < p.w;
    int db = 0;
    uint32_t dphase = 0, parity = 0;
    const long long img = (long long)p.h * p.w;
    float* Hx = reinterpret_cast<float*>(base_ptr + h_off);
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int n, u0, u1, p_lo, p_hi;
      unit_rows(unit, n, u0, u1, p_lo, p_hi);
      float ring[CIN][KH];
#pragma unroll
      for (int c = 0; c < CIN; ++c)
#pragma unroll
        for (int k = 0; k < KH; ++k) ring[c][k] = 0.f;
      __nv_bfloat16* dxn = p.dx + (long long)n * CIN * img + v;
      auto emit = [&](int u) {  // ring[.][0] holds the finished dx row u of every channel
        if (u >= u0 && u < u1 && col_live) {
#pragma unroll
          for (int c = 0; c < CIN; ++c) {
            __nv_bfloat16* d = dxn + (long long)c * img + (long long)u * p.w;
            float val = ring[c][0];
            if (p.beta_dx != 0.f) val += p.beta_dx * __bfloat162float(*d);
            *d = __float2bfloat16_rn(val);
          }
        }
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
#pragma unroll
          for (int k = 0; k + 1 < KH; ++k) ring[c][k] = ring[c][k + 1];
          ring[c][KH - 1] = 0.f;
        }
      };
      for (int pr = p_lo; pr <= p_hi; ++pr) {
        wait_slack(d_full_bar(db), dphase);
        ptx::tc_fence_after();
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + (uint32_t((pw & 3) * 32) << 16) + uint32_t(db * kFDxCols + (pw >> 2) * 32), r);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(d_empty_bar(db));
        db = (db + 1) & (kFDxBufs - 1);
        if (db == 0) dphase ^= 1u;
        if (p.dbg & 4) continue;
        // publish the two last pixels of this warp for the next warp's lanes 0 and 1
        float* mine = Hx + (parity * 8 + pw) * 64;
        if (lane >= 30) {
#pragma unroll
          for (int m = 0; m < M; ++m) mine[(lane - 30) * 32 + m] = __uint_as_float(r[m]);
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        const float* prev = Hx + (parity * 8 + (pw > 0 ? pw - 1 : 0)) * 64;   // [0]: pixel 32pw-2, [1]: pixel 32pw-1
        parity ^= 1u;
#pragma unroll
        for (int c = 0; c < CIN; ++c)
#pragma unroll
          for (int i = 0; i < KH; ++i) {
            const int m0 = (c * KH + i) * KW;
            const float a0 = __uint_as_float(r[m0]);
            float a1 = __shfl_up_sync(0xffffffffu, __uint_as_float(r[m0 + 1]), 1);
            float a2 = __shfl_up_sync(0xffffffffu, __uint_as_float(r[m0 + 2]), 2);
            if (lane == 0) a1 = pw > 0 ? prev[32 + m0 + 1] : 0.f;
            if (lane < 2) a2 = pw > 0 ? prev[lane * 32 + m0 + 2] : 0.f;
            float acc = a0;       // same order as the stand-alone kernel: j = 0, 1, 2
            acc += a1;
            acc += a2;
            ring[c][i] += acc;
          }
        emit(pr);
      }
      for (int u = p_hi + 1; u < u1; ++u) emit(u);  // rows below the last G row (last block of an image)
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

template <int CIN>
int launch_bwd_fused(nk_ctx* ctx, const CUtensorMap& tm, const CUtensorMap& tmh, const ConvBP& p, int grid, size_t smem) {
  auto kern = conv_bwd_fused_kernel<CIN>;
  static bool attr_done[64] = {};  // the attribute is per device
  if (!attr_done[ctx->device & 63]) {
    NK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_done[ctx->device & 63] = true;
  }
  kern<<<grid, kFThreads, smem, ctx->stream>>>(tm, tmh, p);
  NK_LAUNCHED(ctx, "conv_bwd_fused_tc");
  return NK_OK;
}

}  // namespace

int nk_conv2d_bwd_fused_tc(nk_ctx* ctx, void* dx, float beta_dx, void* dwt, int dw_dtype, void* dbias, float beta_dw,
                           const void* g, const void* x, const void* w, int64_t n, int64_t cin, int64_t h, int64_t wd,
                           int64_t cout, int64_t kh, int64_t kw, const float* g_const) {
  if (ctx->conv_engine != NK_CONV_AUTO) return NK_ERR_UNSUPPORTED;
  if (!ctx->encode_tiled) return NK_ERR_UNSUPPORTED;
  if (kh != 3 || kw != 3 || cin < 1 || cin > 3) return NK_ERR_UNSUPPORTED;
  if (cout % 16 != 0 || cout > 128 || wd > 256 || (wd % 8) != 0 || n > 65535) return NK_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(x) & 15) return NK_ERR_UNSUPPORTED;
  ConvBP p;
  p.n = (int)n, p.cin = (int)cin, p.h = (int)h, p.w = (int)wd, p.cout = (int)cout;
  p.ho = int(h - kh + 1), p.wo = int(wd - kw + 1);
  if ((p.wo & 1) || (reinterpret_cast<uintptr_t>(g) & 3)) return NK_ERR_UNSUPPORTED;  // 4-byte cp.async pieces
  p.cpg = 16 / 3;
  if (p.cpg > p.cin) p.cpg = p.cin;
  p.R = p.cpg * 3;
  p.ng = (p.cin + p.cpg - 1) / p.cpg;
  p.ksteps = 3 * p.ng;
  p.ncols = p.ksteps * 16;
  p.fuse_dbias = (dbias != nullptr && p.R < 16) ? 1 : 0;
  p.use_const = g_const != nullptr;
  p.const_bits = 0;
  static const int dbg_mask = getenv("NK_CONV_DBG") ? atoi(getenv("NK_CONV_DBG")) : 0;
  p.dbg = dbg_mask;
  if (g_const) {
    if (dbias && !p.fuse_dbias) return NK_ERR_UNSUPPORTED;   // the separate bias-gradient pass reads G from memory
    const __nv_bfloat16 hv = __float2bfloat16_rn(*g_const);
    const uint16_t hb = *reinterpret_cast<const uint16_t*>(&hv);
    p.const_bits = uint32_t(hb) | (uint32_t(hb) << 16);
  }
  // ONE dW accumulator: consecutive UMMAs into the same TMEM columns issue every 77 cycles, alternating accumulators
  // costs 219 cycles per instruction (tools/umma_probe.cu, profiles/r01_umma_issue_probe.txt)
  p.nacc = 1;
  p.cpr = (p.wo + kChunk - 1) / kChunk;
  while (p.nacc > p.cpr * 4) p.nacc /= 2;   // one owned row (4 UMMAs per live chunk) must touch every accumulator
  p.kblocks = (p.cout + 63) / 64;
  // uniform gradient: one constant G tile for all stages (a stage holds x windows only)
  const size_t g_tile = size_t(kChunksPerTile) * p.cout * 128;
  const size_t stage_bytes = size_t(kChunksPerTile) * ((p.use_const ? 0 : p.cout * 128) + p.ksteps * 2048 + 1024);
  const size_t h_bytes = 2 * 8 * 2 * 32 * 4;   // boundary exchange of the pixel warps
  const size_t fixed = 1024 + size_t(p.kblocks) * 4096 + (p.use_const ? g_tile : 0) + h_bytes + 512;
  if (fixed + 2 * stage_bytes > 232448) return NK_ERR_UNSUPPORTED;
  p.stages = int((232448 - fixed) / stage_bytes);
  if (p.stages > (p.use_const ? 6 : 4)) p.stages = p.use_const ? 6 : 4;
  p.rb = 56;
  if (p.h < 2 * p.rb) p.rb = p.h;
  p.blocks_per_img = (p.h + p.rb - 1) / p.rb;
  p.num_units = p.n * p.blocks_per_img;
  p.g = static_cast<const __nv_bfloat16*>(g);
  p.wt = static_cast<const __nv_bfloat16*>(w);
  p.dx = static_cast<__nv_bfloat16*>(dx);
  p.beta_dx = beta_dx;
  float* scratch;
  int rc = nk_workspace(ctx, size_t(128) * p.ncols * sizeof(float), (void**)&scratch);
  if (rc) return rc;
  CUtensorMap tm, tmh;
  if (make_x_tmap(ctx, &tm, x, n, cin, h, wd, 64, 3, p.cpg, true) != NK_OK) return NK_ERR_UNSUPPORTED;
  if (make_x_tmap(ctx, &tmh, x, n, cin, h, wd, 8, 3, p.cpg, false) != NK_OK) return NK_ERR_UNSUPPORTED;
  if (dbias && !p.fuse_dbias) {
    int64_t dshape[3] = {cout, 1, 1};
    int64_t gshape[4] = {n, cout, p.ho, p.wo};
    rc = nk_unbroadcast_acc(ctx, dbias, dw_dtype, 3, dshape, g, NK_BF16, 4, gshape, beta_dw);
    if (rc) return rc;
    rc = nk_workspace(ctx, size_t(128) * p.ncols * sizeof(float), (void**)&scratch);
    if (rc) return rc;
  }
  p.scratch = scratch;
  NK_CUDA(ctx, cudaMemsetAsync(scratch, 0, size_t(128) * p.ncols * sizeof(float), ctx->stream));
  const size_t smem = fixed + size_t(p.stages) * stage_bytes;
  const int grid = p.num_units < ctx->sm_count ? p.num_units : ctx->sm_count;
  rc = cin == 1 ? launch_bwd_fused<1>(ctx, tm, tmh, p, grid, smem)
       : cin == 2 ? launch_bwd_fused<2>(ctx, tm, tmh, p, grid, smem)
                  : launch_bwd_fused<3>(ctx, tm, tmh, p, grid, smem);
  if (rc) return rc;
  const int total = int(cout * cin * 9 + cout);
  const int blocks = (total + 255) / 256;
  if (dw_dtype == NK_BF16)
    conv_dw_finalize<__nv_bfloat16><<<blocks, 256, 0, ctx->stream>>>((__nv_bfloat16*)dwt, (__nv_bfloat16*)dbias, scratch, p.cout, p.cin, 3, 3, p.cpg, p.ng, p.ncols, beta_dw, p.fuse_dbias);
  else
    conv_dw_finalize<float><<<blocks, 256, 0, ctx->stream>>>((float*)dwt, (float*)dbias, scratch, p.cout, p.cin, 3, 3, p.cpg, p.ng, p.ncols, beta_dw, p.fuse_dbias);
  NK_LAUNCHED(ctx, "conv_dw_finalize");
  ctx->last_conv_kernel = "tcgen05_implicit_gemm_bwd_fused";
  return NK_OK;
}
