// tcgen05 GEMM engine (sm_100a): C = alpha * op(A).op(B) + beta*C (+bias[n], ReLU)
// bf16 operands, f32 accumulation in tensor memory, f32 or bf16 output.
//
// Serves the three GEMM forms of every matmul node:
//   NT  Y  = X.W^T   (MatrixMatrixMulT::forward, matrix_matrix_mul_t/mod.rs:31-41;  mm dA :63-73)
//   NN  dX = G.W     (MatrixMatrixMulTBackwardLeft :63-73;  mm forward matrix_matrix_mul/mod.rs:31-41)
//   TN  dW = G^T.X   (MatrixMatrixMulTBackwardRight :95-105;  mm dB :95-105)
// A "transposed" operand is never copied: TMA loads it as stored and the UMMA shared-memory
// descriptor is MN-major instead of K-major.
//
// Structure (persistent over output tiles; one CTA per SM, or -- for outputs wider and taller than 128 -- one CTA PAIR per
// TPC running cta_group::2 UMMAs of M = 256, see the comment on gemm_tc_kernel):
//   warp 0      : TMA producer  -- cp.async.bulk.tensor 128B-swizzled boxes into a kStages smem ring
//   warp 1      : tcgen05.mma issuer (one lane; in a pair: the even CTA's) + TMEM allocation; accumulators
//                 128 x BLOCK_N f32 per CTA, double buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 2..5  : epilogue -- tcgen05.ld (32 lanes x 32 columns per warp), alpha/bias/ReLU in registers, then either
//                 128B-swizzled staging tiles in shared memory + TMA bulk stores (plain outputs), or direct 16-byte
//                 global stores (beta != 0, ReLU-backward mask, reduce-scatter over NVLink, batched launches)
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), tmem full/empty mbarriers (MMA <-> epilogue).
#include <stdlib.h>

#include "nk_internal.cuh"
#include "nk_ptx.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 192;
constexpr int kRsPitch = 36;                        // floats per staged row (conflict-free 16-byte accesses)
constexpr uint32_t kRsStageBytes = 4 * 32 * kRsPitch * 4;  // 4 epilogue warps x 32 rows
constexpr uint32_t kSmemLimit = 232448;  // 227 KB
constexpr uint32_t kStoreTileBytes = 4096;                       // 32 rows x 128 B (32 f32 or 64 bf16 columns), SWIZZLE_128B
constexpr uint32_t kEpiStageBytes = 4 * 2 * kStoreTileBytes;     // 4 epilogue warps x 2 buffers = 32 KB (>= kRsStageBytes)

struct GemmParams {
  int64_t M, N, K;
  int64_t ldc;
  void* C;
  const void* bias;
  float alpha, beta;
  int bias_bf16;
  int bias_per_row;  // the bias is indexed by the output ROW (convolution: C rows are output channels) instead of the column
  int relu;
  const void* mask;    // optional (M, N) tensor of C's element type and leading dimension: v = mask > 0 ? v : 0 before the
                       // beta accumulate -- the ReLU backward of the layer below fused into the dX GEMM (relu/mod.rs:71-78)
  float* colsum;       // optional (N floats, accumulated with atomics; beta must be 0): column sums of the values the
                       // epilogue stores -- the bias gradient of the layer below (un-broadcast of its Addition,
                       // addition/mod.rs:81-135) without a separate pass over the (M, N) gradient
  int num_m_blocks, num_n_blocks, num_k_blocks;
  // UMMA descriptor parameters (bytes)
  uint32_t a_lbo, a_sbo, a_kstep;
  uint32_t b_lbo, b_sbo, b_kstep;
  // fused reduce-scatter epilogue (data parallel dW): rows [o*rs_rows, (o+1)*rs_rows) go to rs_dst[o]; m_rot = rank
  // staggers the owner order between ranks (tile_m_block)
  int rs_world, m_rot;
  int64_t rs_rows;
  void* rs_dst[8];
  // batched operation (im2col convolution, nk_gemm_batched): `batch` independent products whose operands are 3-D
  // tensor maps (k, rows, batch); C of product b starts c_batch_stride elements further.  batch_reduce: ONE output,
  // the products of all batches are summed (the k loop runs over (batch, k)); the batch range is split over
  // `splits` CTAs per tile, which add their partial sums into C with f32 atomics (C zeroed / scaled by the host).
  int batch, a_batched, b_batched, batch_reduce, splits;
  int64_t c_batch_stride;
  // epilogue through shared memory + TMA bulk stores (tmap_c): every store instruction of the direct path writes 32
  // separate 16-byte row pieces; the staged tile leaves as full 128-byte rows.  Set by the host when beta == 0, no mask,
  // no reduce-scatter, not batched, and C is TMA-addressable.
  int tma_store;
  int cta_pairs;   // host-side: launch the cta_group::2 variant (the B tensor map then has 128-row boxes)
  // tail split (CTA pairs only): the tiles of the last, partial wave are cut in two k halves, each half on its own
  // pair.  The first half writes its raw f32 accumulator rows to `split_ws` and raises a per-warp flag; the second half
  // adds them to its own in the epilogue.  full_waves whole waves of tiles come first; split_tiles = 0: off.
  int split_tiles, full_waves, allow_split;
  float* split_ws;
  uint32_t* split_flags;
};

// one unit of work of a persistent CTA (pair): a whole tile, or one k half of a tile of the split tail
struct Work {
  int tile, kb0, kb1, kind, split_index;   // kind 0 = whole tile, 1 = first half (publishes), 2 = second half (adds + stores)
};
__device__ __forceinline__ bool get_work(const GemmParams& p, int it, int unit, int units, int num_tiles, int kblocks, Work& w) {
  w.kb0 = 0, w.kb1 = kblocks, w.kind = 0, w.split_index = 0;
  if (p.split_tiles == 0) {
    w.tile = unit + it * units;
    return w.tile < num_tiles;
  }
  if (it < p.full_waves) {
    w.tile = unit + it * units;
    return true;
  }
  if (it > p.full_waves || unit >= 2 * p.split_tiles) return false;
  w.split_index = unit >> 1;
  w.tile = p.full_waves * units + w.split_index;
  const int half = kblocks / 2;
  if (unit & 1) w.kb0 = half, w.kind = 2;
  else w.kb1 = half, w.kind = 1;
  return true;
}

__device__ __forceinline__ int tile_m_block(const GemmParams& p, int tile) {
  const int t = tile % p.num_m_blocks;
  if (p.rs_world == 0) return t;
  // consecutive tiles go to consecutive owners (starting with a different one on every rank), so that the remote
  // stores are spread evenly over the whole kernel and over all links instead of bunching up at the end
  const int owner = (t + p.m_rot) % p.rs_world;
  return owner * (p.num_m_blocks / p.rs_world) + t / p.rs_world;
}

// tile index -> (m block, n block).  Tiles that run concurrently are consecutive indices (persistent CTAs stride by the
// grid), so consecutive indices walk kGroupM m-blocks before moving to the next n-block: a wave of ~128 tiles then
// covers a near-square 16 x 8 patch of the output and re-reads far less of A and B than an m-fastest order
// (8192-row GEMMs of config 4 moved 2.7x their algorithmic DRAM bytes, profiles/r01_launches.md).
constexpr int kGroupM = 16;
__device__ __forceinline__ void tile_coords(const GemmParams& p, int tile, int& m_blk, int& n_blk) {
  if (p.rs_world) {
    m_blk = tile_m_block(p, tile);
    n_blk = tile / p.num_m_blocks;
    return;
  }
  const int group_size = kGroupM * p.num_n_blocks;
  const int group = tile / group_size, in_group = tile - group * group_size;
  const int first_m = group * kGroupM;
  const int gm = min(p.num_m_blocks - first_m, kGroupM);
  m_blk = first_m + in_group % gm;
  n_blk = in_group / gm;
}

template <int BLOCK_N, int CG = 1>
struct Cfg {
  static constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
  static constexpr uint32_t B_BYTES = (BLOCK_N / CG) * BLOCK_K * 2;   // a CTA of a pair stages half of the B tile
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  // after the ring: 1 KB of barriers, then kEpiStageBytes of epilogue staging (TMA-store tiles / reduce-scatter rows)
  static constexpr int kStagesMax = (kSmemLimit - 2048 - kEpiStageBytes) / STAGE_BYTES;
  static constexpr int kStages = kStagesMax > 8 ? 8 : kStagesMax;
  static constexpr uint32_t SMEM_BYTES = kStages * STAGE_BYTES + 2048 + kEpiStageBytes;  // + alignment slack + barriers
  static constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                        : (2 * BLOCK_N <= 256) ? 256 : 512;
};

// v[j] = alpha * acc[j] + bias (row- or column-indexed) for one 32-column chunk of one output row
__device__ __forceinline__ void scale_and_bias(const GemmParams& p, int64_t row, int64_t col0, const uint32_t* r, bool full,
                                               float* v) {
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = p.alpha * __uint_as_float(r[j]);
  if (p.bias && p.bias_per_row) {
    const float b = p.bias_bf16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(p.bias)[row]) : static_cast<const float*>(p.bias)[row];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += b;
  } else if (p.bias) {
    // 32 scalar loads here serialise on L1 latency and made the epilogue slower than a K = 1024 main loop
    // (profiles/r01_launches.md): fetch the 32 bias values of a full chunk with 16-byte loads
    if (full && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
      if (p.bias_bf16) {
        const uint4* bp = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.bias) + col0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 w = __ldg(bp + q);
          const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[q * 8 + 2 * i] += __uint_as_float(ww[i] << 16);
            v[q * 8 + 2 * i + 1] += __uint_as_float(ww[i] & 0xffff0000u);
          }
        }
      } else {
        const float4* bp = reinterpret_cast<const float4*>(static_cast<const float*>(p.bias) + col0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 w = __ldg(bp + q);
          v[q * 4] += w.x, v[q * 4 + 1] += w.y, v[q * 4 + 2] += w.z, v[q * 4 + 3] += w.w;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N)
          v[j] += p.bias_bf16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(p.bias)[col0 + j])
                              : static_cast<const float*>(p.bias)[col0 + j];
    }
  }
}

// column sums of one 32-row x 32-column chunk (all 32 lanes of the warp take part; lane = row): the values are exactly the
// ones epilogue_store_chunk32 stores (alpha, mask, rounding to the output type), rows / columns outside the matrix count as
// zero.  A butterfly of 31 shuffles leaves the sum of column j on lane j; one atomic per column and chunk.
template <typename TC>
__device__ __forceinline__ void epilogue_colsum_chunk32(const GemmParams& p, int64_t row, int64_t col0, const uint32_t* r,
                                                        int lane, bool vec_ok) {
  float v[32];
  const bool live = row < p.M;
  const TC* mrow = (p.mask && live) ? static_cast<const TC*>(p.mask) + row * p.ldc + col0 : nullptr;
  const bool full = col0 + 32 <= p.N;
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = (live && col0 + j < p.N) ? p.alpha * __uint_as_float(r[j]) : 0.f;
  if (mrow) {
    if (full && vec_ok) {
      // the same four 16-byte loads per row the store path issues right after (then L1 hits): 32 scalar loads per row,
      // each a first touch of the mask, made the epilogue of an 8192 x 4096 GEMM longer than its main loop
      constexpr int V = 16 / sizeof(TC);
#pragma unroll
      for (int q = 0; q < 32 / V; ++q) {
        NkVec<TC> mk;
        mk.load(mrow + q * V);
#pragma unroll
        for (int i = 0; i < V; ++i) v[q * V + i] = mk.get(i) > 0.f ? v[q * V + i] : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) v[j] = nk_to_f32<TC>(mrow[j]) > 0.f ? v[j] : 0.f;
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = nk_to_f32<TC>(nk_from_f32<TC>(v[j]));   // what the output tensor will hold
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  if (col0 + lane < p.N) atomicAdd(p.colsum + col0 + lane, v[0]);
}

template <typename TC>
__device__ __forceinline__ void epilogue_store_chunk32(const GemmParams& p, int64_t row, int64_t col0, const uint32_t* r,
                                                       int ncols, bool vec_ok, int64_t c_off = 0, bool atomic = false) {
  TC* crow = static_cast<TC*>(p.C) + c_off + row * p.ldc + col0;
  if (atomic) {  // partial sum of a split reduction: f32 atomics (alpha applied, beta handled by the host)
    if constexpr (sizeof(TC) == 4) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols && col0 + j < p.N) atomicAdd(reinterpret_cast<float*>(crow) + j, p.alpha * __uint_as_float(r[j]));
    }
    return;
  }
  if (p.rs_world) {
    const int owner = int(row / p.rs_rows);
    crow = static_cast<TC*>(p.rs_dst[owner]) + (row - owner * p.rs_rows) * p.ldc + col0;
  }
  float v[32];
  const bool full = (col0 + 32 <= p.N) && ncols == 32;
  scale_and_bias(p, row, col0, r, full, v);
  const TC* mrow = p.mask ? static_cast<const TC*>(p.mask) + row * p.ldc + col0 : nullptr;
  if (full && vec_ok) {
    constexpr int V = 16 / sizeof(TC);
    if (mrow) {
#pragma unroll
      for (int q = 0; q < 32 / V; ++q) {
        NkVec<TC> mk;
        mk.load(mrow + q * V);
#pragma unroll
        for (int i = 0; i < V; ++i) v[q * V + i] = mk.get(i) > 0.f ? v[q * V + i] : 0.f;
      }
    }
    if (p.beta != 0.f) {
#pragma unroll
      for (int q = 0; q < 32 / V; ++q) {
        NkVec<TC> c;
        c.load(crow + q * V);
#pragma unroll
        for (int i = 0; i < V; ++i) v[q * V + i] += p.beta * c.get(i);
      }
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 32 / V; ++q) {
      NkVec<TC> o;
#pragma unroll
      for (int i = 0; i < V; ++i) o.set(i, v[q * V + i]);
      o.store(crow + q * V);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < ncols && col0 + j < p.N) {
        float x = v[j];
        if (mrow) x = nk_to_f32<TC>(mrow[j]) > 0.f ? x : 0.f;
        if (p.beta != 0.f) x += p.beta * nk_to_f32<TC>(crow[j]);
        if (p.relu) x = x > 0.f ? x : 0.f;
        crow[j] = nk_from_f32<TC>(x);
      }
    }
  }
}

// CG = 2: CTA pairs (cta_group::2).  A cluster of two CTAs owns a 256 x BLOCK_N tile: CTA r loads rows [128r, 128r + 128) of
// the A tile and rows [BLOCK_N/2 r, ...) of the B tile into its own shared memory (32 KB per stage instead of 48 KB: six
// stages in flight instead of four, and a third less L2 -> SM traffic per flop), the even CTA issues one M = 256 UMMA per
// k step for both, and each CTA drains its own 128 accumulator rows from its own TMEM.
template <int BLOCK_N, bool A_MN, bool B_MN, typename TC, bool BATCH = false, int CG = 1>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_c, const GemmParams p) {
  using C_ = Cfg<BLOCK_N, CG>;
  constexpr int kStages = C_::kStages;
  constexpr int LOAD_N = BLOCK_N / CG;
  static_assert(CG == 1 || (CG == 2 && !BATCH && BLOCK_N == 256), "CTA pairs: plain 256-wide tiles only");
  const uint32_t cta_rank = CG == 2 ? ptx::cluster_ctarank() : 0u;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms: 1024 B aligned
  const uint32_t smem_a0 = smem_base;
  const uint32_t smem_b0 = smem_base + kStages * C_::A_BYTES;
  const uint32_t bar_base = smem_base + kStages * C_::STAGE_BYTES;  // 8-byte aligned
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  // generic pointer to the slot, to read the allocated TMEM address back
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    if (p.tma_store) ptx::prefetch_tmap(&tmap_c);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tmem_full_bar(s), 1);
      ptx::mbar_init(tmem_empty_bar(s), 4 * CG);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) {
    if constexpr (CG == 2) {
      ptx::tmem_alloc_2sm(tmem_slot, C_::TMEM_COLS);
      ptx::tmem_relinquish_2sm();
    } else {
      ptx::tmem_alloc(tmem_slot, C_::TMEM_COLS);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if constexpr (CG == 2)
    ptx::cluster_sync();   // the peer's barriers are initialised before anything signals them
  else
    __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int tiles_mn = p.num_m_blocks * p.num_n_blocks;
  // BATCH: a "tile" is (batch or split, m block, n block); the k loop of a reducing launch walks its share of the batches
  const int num_tiles = BATCH ? tiles_mn * (p.batch_reduce ? p.splits : p.batch) : tiles_mn;
  auto batch_range = [&](int outer, int& b0, int& b1) {   // batches whose products tile `outer` accumulates
    if (!p.batch_reduce) {
      b0 = outer, b1 = outer + 1;
    } else {
      const int per = (p.batch + p.splits - 1) / p.splits;
      b0 = outer * per;
      b1 = min(p.batch, b0 + per);
    }
  };

  if (warp_idx == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      Work wk;
      for (int it = 0; get_work(p, it, int(blockIdx.x) / CG, int(gridDim.x) / CG, num_tiles, p.num_k_blocks, wk); ++it) {
        const int tile = wk.tile;
        int m_blk, n_blk, b0 = 0, b1 = 1;
        tile_coords(p, BATCH ? tile % tiles_mn : tile, m_blk, n_blk);
        if (BATCH) batch_range(tile / tiles_mn, b0, b1);
        const int m0 = (m_blk * CG + int(cta_rank)) * BLOCK_M, n0 = n_blk * BLOCK_N + int(cta_rank) * LOAD_N;
        for (int bb = b0; bb < b1; ++bb)
        for (int kb = wk.kb0; kb < wk.kb1; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          if constexpr (CG == 2) {
            // both CTAs' loads complete on the leader's barrier, which expects the bytes of both
            if (cta_rank == 0) ptx::mbar_expect_tx(full_bar(stage), 2 * C_::STAGE_BYTES);
          } else {
            ptx::mbar_expect_tx(full_bar(stage), C_::STAGE_BYTES);
          }
          const int k0 = kb * BLOCK_K;
          const uint32_t sa = smem_a0 + stage * C_::A_BYTES;
          const uint32_t sb = smem_b0 + stage * C_::B_BYTES;
          if (A_MN) {  // stored (K, M): boxes of 64 (m) x 64 (k)
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c) {
              if (BATCH && p.a_batched)
                ptx::tma_load_3d(sa + c * (64 * BLOCK_K * 2), &tmap_a, full_bar(stage), m0 + c * 64, k0, bb);
              else if (CG == 2)
                ptx::tma_load_2d_2sm(sa + c * (64 * BLOCK_K * 2), &tmap_a, full_bar(stage), m0 + c * 64, k0);
              else
                ptx::tma_load_2d(sa + c * (64 * BLOCK_K * 2), &tmap_a, full_bar(stage), m0 + c * 64, k0);
            }
          } else {  // stored (M, K): one box of 64 (k) x 128 (m)
            if (BATCH && p.a_batched)
              ptx::tma_load_3d(sa, &tmap_a, full_bar(stage), k0, m0, bb);
            else if (CG == 2)
              ptx::tma_load_2d_2sm(sa, &tmap_a, full_bar(stage), k0, m0);
            else
              ptx::tma_load_2d(sa, &tmap_a, full_bar(stage), k0, m0);
          }
          if (B_MN) {  // stored (K, N)
#pragma unroll
            for (int c = 0; c < LOAD_N / 64; ++c) {
              if (BATCH && p.b_batched)
                ptx::tma_load_3d(sb + c * (64 * BLOCK_K * 2), &tmap_b, full_bar(stage), n0 + c * 64, k0, bb);
              else if (CG == 2)
                ptx::tma_load_2d_2sm(sb + c * (64 * BLOCK_K * 2), &tmap_b, full_bar(stage), n0 + c * 64, k0);
              else
                ptx::tma_load_2d(sb + c * (64 * BLOCK_K * 2), &tmap_b, full_bar(stage), n0 + c * 64, k0);
            }
          } else {  // stored (N, K)
            if (BATCH && p.b_batched)
              ptx::tma_load_3d(sb, &tmap_b, full_bar(stage), k0, n0, bb);
            else if (CG == 2)
              ptx::tma_load_2d_2sm(sb, &tmap_b, full_bar(stage), k0, n0);
            else
              ptx::tma_load_2d(sb, &tmap_b, full_bar(stage), k0, n0);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================================================== MMA issuer (CTA pairs: the even CTA issues for both)
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(BLOCK_M * CG, BLOCK_N, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      Work wk;
      for (int it = 0; get_work(p, it, int(blockIdx.x) / CG, int(gridDim.x) / CG, num_tiles, p.num_k_blocks, wk); ++it) {
        const int tile = wk.tile;
        ptx::mbar_wait(tmem_empty_bar(as), aphase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(as * BLOCK_N);
        int k_total = wk.kb1 - wk.kb0;
        if (BATCH) {
          int b0, b1;
          batch_range(tile / tiles_mn, b0, b1);
          k_total *= (b1 - b0);
        }
        for (int kb = 0; kb < k_total; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint64_t adesc = ptx::make_smem_desc_sw128(smem_a0 + stage * C_::A_BYTES, p.a_lbo, p.a_sbo);
          const uint64_t bdesc = ptx::make_smem_desc_sw128(smem_b0 + stage * C_::B_BYTES, p.b_lbo, p.b_sbo);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            if constexpr (CG == 2)
              ptx::mma_f16_ss_2sm(tmem_d, adesc + uint64_t((k * p.a_kstep) >> 4), bdesc + uint64_t((k * p.b_kstep) >> 4),
                                  idesc, (kb | k) != 0 ? 1u : 0u);
            else
              ptx::mma_f16_ss(tmem_d, adesc + uint64_t((k * p.a_kstep) >> 4), bdesc + uint64_t((k * p.b_kstep) >> 4),
                              idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (CG == 2)
            ptx::mma_commit_2sm(empty_bar(stage));  // the slot of BOTH CTAs is reusable once these MMAs retire
          else
            ptx::mma_commit(empty_bar(stage));  // smem slot reusable once these MMAs retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if constexpr (CG == 2)
          ptx::mma_commit_2sm(tmem_full_bar(as));  // accumulator complete -> the epilogue warps of both CTAs
        else
          ptx::mma_commit(tmem_full_bar(as));  // accumulator complete -> epilogue
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else {
    // ===================================================== epilogue (4 warps)
    const int q = warp_idx & 3;  // TMEM lane quarter this warp may access
    const bool vec_ok = ((p.ldc * int64_t(sizeof(TC))) % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.mask) & 15) == 0);
    const uint32_t epi_stage = bar_base + 1024;   // 1024-byte aligned (SWIZZLE_128B store tiles)
    float* rs_stage = reinterpret_cast<float*>(smem_raw + (epi_stage - ptx::smem_u32(smem_raw)));
    (void)rs_stage;
    uint32_t store_buf = 0;                      // which of this warp's two store tiles the next chunk uses
    int as = 0;
    uint32_t aphase = 0;
    Work wk;
    for (int it = 0; get_work(p, it, int(blockIdx.x) / CG, int(gridDim.x) / CG, num_tiles, p.num_k_blocks, wk); ++it) {
      const int tile = wk.tile;
      int m_blk, n_blk;
      tile_coords(p, BATCH ? tile % tiles_mn : tile, m_blk, n_blk);
      const int64_t c_off = (BATCH && !p.batch_reduce) ? int64_t(tile / tiles_mn) * p.c_batch_stride : 0;
      const bool atomic = BATCH && p.batch_reduce;
      const int64_t tile_row0 = (int64_t(m_blk) * CG + cta_rank) * BLOCK_M;   // first output row of this CTA's half
      const int64_t row = tile_row0 + q * 32 + lane;
      ptx::mbar_wait(tmem_full_bar(as), aphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BLOCK_N);
      // ---- split tail: this warp's 32 rows of the other k half's accumulator live at `part` as [chunk][lane][32 floats]
      const float4* part = nullptr;
      uint32_t* split_flag = nullptr;
      if constexpr (CG == 2 && !BATCH) {
        if (wk.kind != 0) {
          const size_t wi = (size_t(wk.split_index) * 2 + cta_rank) * 4 + q;
          float4* mine = reinterpret_cast<float4*>(p.split_ws + wi * (size_t(BLOCK_N) * 32));
          split_flag = p.split_flags + wi;
          if (wk.kind == 1) {
            // first k half: publish the raw accumulator rows (a warp writes 4 KB contiguous per chunk), raise the flag
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
              uint32_t r[32];
              ptx::tmem_ld_32x32b_x32(taddr + c * 32, r);
              ptx::tmem_ld_wait();
              float4* dst = mine + (size_t(c) * 32 + lane) * 8;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                     __uint_as_float(r[4 * j + 3]));
            }
            __threadfence();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              ptx::mbar_arrive_leader(tmem_empty_bar(as));
              asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(split_flag), "r"(1u) : "memory");
            }
            as ^= 1;
            if (as == 0) aphase ^= 1u;
            continue;
          }
          // second k half: wait for the first one's rows (its pair runs concurrently and waits for nobody)
          if (lane == 0) {
            const long long t0 = clock64();
            for (;;) {
              uint32_t v;
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(split_flag) : "memory");
              if (v != 0u) break;
              if (clock64() - t0 > (1ll << 32)) {
                printf("nk_b200: gemm split-tail watchdog: block %d warp %d tile %d\n", blockIdx.x, warp_idx, wk.tile);
                __trap();
              }
              __nanosleep(64);
            }
          }
          __syncwarp();
          part = mine;
        }
      }
      // adds the other half's partial sums of 32-column chunk c32 to the freshly loaded accumulator values
      auto add_partial = [&](uint32_t* r, int c32) {
        if constexpr (CG == 2 && !BATCH) {
          if (part) {
            const float4* src = part + (size_t(c32) * 32 + lane) * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 v = __ldcg(src + j);   // written by another SM: read through L2
              r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + v.x);
              r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + v.y);
              r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + v.z);
              r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + v.w);
            }
          }
        }
      };
      auto split_done = [&]() {   // the partial rows have been consumed: the flag is zero again for the next launch
        if constexpr (CG == 2 && !BATCH) {
          if (part) {
            __syncwarp();
            if (lane == 0) *reinterpret_cast<volatile uint32_t*>(split_flag) = 0u;
          }
        }
      };
      if constexpr (BLOCK_N >= 64 && !BATCH) {
        if (p.tma_store) {
          // ---- staged epilogue: 32 rows x 128 bytes per tile (32 f32 / 64 bf16 columns), written 128B-swizzled so that
          // the eight lanes of a store phase hit distinct banks, then ONE bulk tensor store per tile (TMA clips M/N tails)
          constexpr int kColsPerTile = 128 / int(sizeof(TC));
          const int64_t row0 = tile_row0 + q * 32;
#pragma unroll 1
          for (int c = 0; c < BLOCK_N / kColsPerTile; ++c) {
            const int64_t col0 = int64_t(n_blk) * BLOCK_N + c * kColsPerTile;
            if (col0 >= p.N) break;
            const uint32_t tile = epi_stage + (uint32_t(warp_idx - 2) * 2u + store_buf) * kStoreTileBytes;
            if (lane == 0) ptx::tma_store_wait_read<1>();   // the store that last read this buffer has finished reading
            __syncwarp();
            const uint32_t rowaddr = tile + uint32_t(lane) * 128u;
            const uint32_t sx = uint32_t(lane & 7);
            if constexpr (sizeof(TC) == 4) {
              uint32_t r[32];
              ptx::tmem_ld_32x32b_x32(taddr + c * 32, r);
              ptx::tmem_ld_wait();
              add_partial(r, c);
              float v[32];
              scale_and_bias(p, row, col0, r, col0 + 32 <= p.N, v);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                if (p.relu) o = make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + ((uint32_t(j) ^ sx) << 4)), "f"(o.x),
                             "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
              }
            } else {
#pragma unroll
              for (int hc = 0; hc < 2; ++hc) {
                uint32_t r[32];
                ptx::tmem_ld_32x32b_x32(taddr + c * 64 + hc * 32, r);
                ptx::tmem_ld_wait();
                add_partial(r, c * 2 + hc);
                float v[32];
                scale_and_bias(p, row, col0 + hc * 32, r, col0 + hc * 32 + 32 <= p.N, v);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  uint32_t w[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    float a = v[8 * j + 2 * e], b = v[8 * j + 2 * e + 1];
                    if (p.relu) a = fmaxf(a, 0.f), b = fmaxf(b, 0.f);
                    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
                    w[e] = *reinterpret_cast<uint32_t*>(&h);
                  }
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + ((uint32_t(hc * 4 + j) ^ sx) << 4)),
                               "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
                }
              }
            }
            ptx::fence_proxy_async();   // generic-proxy writes -> visible to the bulk-copy (async proxy) read
            __syncwarp();
            if (lane == 0) {
              ptx::tma_store_2d(&tmap_c, tile, int32_t(col0), int32_t(row0));
              ptx::tma_store_commit();
            }
            store_buf ^= 1u;
          }
          split_done();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2)
              ptx::mbar_arrive_leader(tmem_empty_bar(as));   // the MMA issuer lives in the even CTA
            else
              ptx::mbar_arrive(tmem_empty_bar(as));
          }
          as ^= 1;
          if (as == 0) aphase ^= 1u;
          continue;
        }
      }
      if (BLOCK_N >= 32) {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(taddr + c * 32, r);
          ptx::tmem_ld_wait();
          add_partial(r, c);
          const int64_t col0 = int64_t(n_blk) * BLOCK_N + c * 32;
          if constexpr (sizeof(TC) == 4 && BLOCK_N == 256) {
            if (p.rs_world) {
              // reduce-scatter epilogue: the chunk goes to the owner's slot over NVLink.  One thread per row would
              // emit 16-byte packets; transpose the 32x32 chunk through shared memory so that every store
              // instruction writes four full 128-byte row segments.
              float* stg = rs_stage + (warp_idx - 2) * (32 * kRsPitch);
#pragma unroll
              for (int qq = 0; qq < 8; ++qq)
                *reinterpret_cast<float4*>(stg + lane * kRsPitch + qq * 4) =
                    make_float4(p.alpha * __uint_as_float(r[qq * 4]), p.alpha * __uint_as_float(r[qq * 4 + 1]),
                                p.alpha * __uint_as_float(r[qq * 4 + 2]), p.alpha * __uint_as_float(r[qq * 4 + 3]));
              __syncwarp();
              const int64_t row0 = tile_row0 + q * 32;
              const int owner = int(row0 / p.rs_rows);   // a 128-row tile never straddles two shards
              float* dst0 = static_cast<float*>(p.rs_dst[owner]) + (row0 - owner * p.rs_rows) * p.ldc + col0 + (lane & 7) * 4;
              const bool col_ok = col0 + (lane & 7) * 4 < p.N;
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + (lane >> 3);
                const float4 v = *reinterpret_cast<const float4*>(stg + rr * kRsPitch + (lane & 7) * 4);
                if (col_ok && row0 + rr < p.M) *reinterpret_cast<float4*>(dst0 + int64_t(rr) * p.ldc) = v;
              }
              __syncwarp();
              continue;
            }
          }
          if (p.colsum && col0 < p.N) epilogue_colsum_chunk32<TC>(p, row, col0, r, lane, vec_ok);
          if (row < p.M && col0 < p.N) epilogue_store_chunk32<TC>(p, row, col0, r, 32, vec_ok, c_off, atomic);
        }
      } else {
        uint32_t r[32];
        uint32_t r16[16];
        ptx::tmem_ld_32x32b_x16(taddr, r16);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = j < 16 ? r16[j & 15] : 0u;
        const int64_t col0 = int64_t(n_blk) * BLOCK_N;
        if (row < p.M && col0 < p.N) epilogue_store_chunk32<TC>(p, row, col0, r, 16, false, c_off, atomic);
      }
      split_done();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
            if constexpr (CG == 2)
              ptx::mbar_arrive_leader(tmem_empty_bar(as));   // the MMA issuer lives in the even CTA
            else
              ptx::mbar_arrive(tmem_empty_bar(as));
          }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
    if (p.tma_store && lane == 0) ptx::tma_store_wait<0>();   // shared memory must outlive the bulk stores that read it
  }

  ptx::tc_fence_before();
  if constexpr (CG == 2)
    ptx::cluster_sync();   // neither CTA of a pair may free its TMEM / exit while the other still signals it
  else
    __syncthreads();
  if (warp_idx == 1) {
    ptx::tc_fence_after();
    if constexpr (CG == 2)
      ptx::tmem_dealloc_2sm(tmem_base, C_::TMEM_COLS);
    else
      ptx::tmem_dealloc(tmem_base, C_::TMEM_COLS);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D bf16 tensor map: `rows` x `cols` row-major with leading dimension ld, box (box_cols=64, box_rows)
int make_tmap_2d(nk_ctx* ctx, CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld,
                 uint32_t box_cols, uint32_t box_rows) {
  if (!ctx->encode_tiled) return nk_set_error(ctx, NK_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return nk_set_error(ctx, NK_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
                        (long long)rows, (long long)cols, (long long)ld);
  return NK_OK;
}

// output tensor map of the staged epilogue: (M, N) row-major with leading dimension ldc, boxes of 32 rows x 128 bytes
int make_tmap_c(nk_ctx* ctx, CUtensorMap* tm, void* base, int64_t rows, int64_t cols, int64_t ld, int c_dtype) {
  if (!ctx->encode_tiled) return NK_ERR_UNSUPPORTED;
  const size_t es = nk_dtype_size(c_dtype);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * es};
  cuuint32_t box[2] = {(cuuint32_t)(128 / es), 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      tm, c_dtype == NK_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box,
      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NK_OK : NK_ERR_UNSUPPORTED;
}

uint32_t env_u32(const char* name, uint32_t dflt) {
  const char* v = getenv(name);
  return v ? (uint32_t)strtoul(v, nullptr, 0) : dflt;
}

template <int BLOCK_N, bool A_MN, bool B_MN, typename TC, bool BATCH = false, int CG = 1>
int launch_cfg(nk_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, GemmParams& p, const CUtensorMap* tc = nullptr) {
  using C_ = Cfg<BLOCK_N, CG>;
  auto kern = gemm_tc_kernel<BLOCK_N, A_MN, B_MN, TC, BATCH, CG>;
  static bool attr_done[64] = {};  // per template instantiation and device (the attribute is per device)
  if (!attr_done[ctx->device & 63]) {
    static_assert(C_::SMEM_BYTES <= kSmemLimit && kRsStageBytes <= kEpiStageBytes, "shared memory budget");
    NK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_::SMEM_BYTES));
    attr_done[ctx->device & 63] = true;
  }
  p.num_n_blocks = int((p.N + BLOCK_N - 1) / BLOCK_N);
  if (CG == 2) p.num_m_blocks = int((p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M));   // a CTA pair owns 256 rows
  int num_tiles = p.num_m_blocks * p.num_n_blocks;
  if (BATCH) {
    if (p.batch_reduce) {
      // split the batch range so that the launch fills the machine; every split gets at least one batch
      int want = (ctx->sm_count + num_tiles - 1) / num_tiles;
      if (want > p.batch) want = p.batch;
      const int per = (p.batch + want - 1) / want;
      p.splits = (p.batch + per - 1) / per;
      num_tiles *= p.splits;
    } else {
      num_tiles *= p.batch;
    }
  }
  // persistent grid balanced over the waves the tiles need anyway: 512 tiles on 148 SMs take 4 waves whether 148
  // or 128 CTAs run them, and the 20 SMs left free let a concurrent NCCL all-reduce make progress
  const int units = ctx->sm_count / CG;                       // CTAs, or CTA pairs
  const int waves = (num_tiles + units - 1) / units;
  int grid = ((num_tiles + waves - 1) / waves) * CG;
  p.split_tiles = 0, p.full_waves = 0, p.split_ws = nullptr, p.split_flags = nullptr;
  if (CG == 2 && !BATCH && p.rs_world == 0 && p.allow_split) {
    // tail split: when the last wave is at most half full, its tiles are cut in two k halves on two pairs each -- 256
    // tiles on 74 pairs then take 3.5 tile times instead of 4
    const int full = num_tiles / units, rem = num_tiles - full * units;
    constexpr size_t kFlagBytes = 4096;                                       // 2 x 4 words per split tile, <= 74 tiles
    const size_t ws_bytes = size_t(units / 2) * 2 * 4 * (size_t(BLOCK_N) * 32) * sizeof(float);
    if (rem > 0 && 2 * rem <= units && p.num_k_blocks >= 8) {
      if (!ctx->gemm_split_mem && !ctx->capturing) {
        if (cudaMalloc(&ctx->gemm_split_mem, kFlagBytes + ws_bytes) == cudaSuccess) {
          cudaMemsetAsync(ctx->gemm_split_mem, 0, kFlagBytes, ctx->stream);
        } else {
          (void)cudaGetLastError();
          ctx->gemm_split_mem = nullptr;
        }
      }
      if (ctx->gemm_split_mem) {
        p.split_tiles = rem, p.full_waves = full;
        p.split_flags = static_cast<uint32_t*>(ctx->gemm_split_mem);
        p.split_ws = reinterpret_cast<float*>(static_cast<char*>(ctx->gemm_split_mem) + kFlagBytes);
        grid = units * CG;
      }
    }
  }
  size_t smem = C_::SMEM_BYTES;
  if (p.rs_world) {
    if (BLOCK_N != 256 || sizeof(TC) != 4 || p.N % 4 != 0)
      return nk_set_error(ctx, NK_ERR_UNSUPPORTED, "tcgen05 gemm: reduce-scatter epilogue needs 128x256 tiles, f32, N %% 4 == 0");
  }
  if (!tc || BLOCK_N < 64 || BATCH) p.tma_store = 0;
  if constexpr (CG == 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(kNumThreads), cfg.dynamicSmemBytes = smem, cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, p.tma_store ? *tc : ta, p);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      return nk_set_error(ctx, NK_ERR_CUDA, "launch of gemm_tcgen05 (CTA pairs) failed: %s (M=%lld N=%lld K=%lld grid=%d smem=%zu)",
                          cudaGetErrorString(e), (long long)p.M, (long long)p.N, (long long)p.K, grid, smem);
    }
  } else {
    kern<<<grid, kNumThreads, smem, ctx->stream>>>(ta, tb, p.tma_store ? *tc : ta, p);
  }
  ctx->launches++;
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess)
    return nk_set_error(ctx, NK_ERR_CUDA, "launch of gemm_tcgen05 failed: %s (M=%lld N=%lld K=%lld BLOCK_N=%d grid=%d smem=%zu "
                        "a_mn=%d b_mn=%d out=%s)", cudaGetErrorString(le), (long long)p.M, (long long)p.N, (long long)p.K,
                        BLOCK_N, grid, smem, int(A_MN), int(B_MN), sizeof(TC) == 4 ? "f32" : "bf16");
  return NK_OK;
}

template <bool A_MN, bool B_MN, typename TC>
int launch_bn(nk_ctx* ctx, int block_n, const CUtensorMap& ta, const CUtensorMap& tb, GemmParams& p, const CUtensorMap* tc) {
  switch (block_n) {
    case 256:
      if (p.cta_pairs) return launch_cfg<256, A_MN, B_MN, TC, false, 2>(ctx, ta, tb, p, tc);
      return launch_cfg<256, A_MN, B_MN, TC>(ctx, ta, tb, p, tc);
    case 128: return launch_cfg<128, A_MN, B_MN, TC>(ctx, ta, tb, p, tc);
    case 64: return launch_cfg<64, A_MN, B_MN, TC>(ctx, ta, tb, p, tc);
    default:
      if (!B_MN) {
        if (block_n == 32) return launch_cfg<32, A_MN, false, TC>(ctx, ta, tb, p);
        if (block_n == 16) return launch_cfg<16, A_MN, false, TC>(ctx, ta, tb, p);
      }
      return nk_set_error(ctx, NK_ERR_UNSUPPORTED, "tcgen05 gemm: unsupported BLOCK_N %d", block_n);
  }
}

}  // namespace

bool nk_gemm_tcgen05_supported(int transA, int transB, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda,
                               const void* B, int64_t ldb) {
  (void)transA;
  (void)transB;
  if (M <= 0 || N <= 0 || K <= 0) return false;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return false;
  if ((lda * 2) % 16 != 0 || (ldb * 2) % 16 != 0) return false;  // TMA global strides: multiples of 16 bytes
  if (M > (int64_t(1) << 30) || N > (int64_t(1) << 30) || K > (int64_t(1) << 30)) return false;
  return true;
}

int nk_gemm_tcgen05(nk_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, float alpha, const void* A,
                    int64_t lda, const void* B, int64_t ldb, float beta, void* C, int64_t ldc, int c_dtype,
                    const void* bias, int bias_dtype, int relu, const void* mask, float* colsum) {
  if (!nk_gemm_tcgen05_supported(transA, transB, M, N, K, A, lda, B, ldb)) return NK_ERR_UNSUPPORTED;
  if (colsum && (N <= 16 || beta != 0.f)) return NK_ERR_UNSUPPORTED;   // (the narrow-tile epilogue has no column-sum path)
  const bool a_mn = transA != 0;   // op(A) = A^T  -> A stored (K, M), M contiguous
  const bool b_mn = transB == 0;   // op(B) = B    -> B stored (K, N), N contiguous
  // tile width: widest that does not leave most of a tile empty
  int block_n = 256;
  if (N <= 16 && !b_mn)
    block_n = 16;
  else if (N <= 32 && !b_mn)
    block_n = 32;
  else if (N <= 64)
    block_n = 64;
  else if (N <= 128)
    block_n = 128;
  else {
    // 128x256 tiles: one MMA reads 16 KB (A) + 32 KB (B) of smem per 512 tensor cycles (96 B/clk, under the
    // 128 B/clk smem port) -- 128x128 tiles need the full 128 B/clk and measured 938 vs 1342 TFLOP/s at 4096^3
    // (profiles/r01_gemm_tile_sweep.md), so the wide tile wins even when it quantises worse over 148 SMs.
    block_n = 256;
  }
  // development knobs (tile sweeps, descriptor sweeps on hardware): the environment is read ONCE per process
  struct Knobs {
    uint32_t block_n, a_lbo, a_sbo, b_lbo, b_sbo, direct_store, single_cta, no_split;
    Knobs() : block_n(env_u32("NK_GEMM_BLOCK_N", 0)), a_lbo(env_u32("NK_DESC_A_LBO", 0)), a_sbo(env_u32("NK_DESC_A_SBO", 0)),
              b_lbo(env_u32("NK_DESC_B_LBO", 0)), b_sbo(env_u32("NK_DESC_B_SBO", 0)),
              direct_store(env_u32("NK_GEMM_DIRECT_STORE", 0)), single_cta(env_u32("NK_GEMM_SINGLE_CTA", 0)),
              no_split(env_u32("NK_GEMM_NO_SPLIT", 0)) {}
  };
  static const Knobs knobs;
  if (knobs.block_n) block_n = int(knobs.block_n);

  GemmParams p;
  p.M = M;
  p.N = N;
  p.K = K;
  p.ldc = ldc;
  p.C = C;
  p.bias = bias;
  p.alpha = alpha;
  p.beta = beta;
  p.bias_bf16 = bias_dtype == NK_BF16;
  p.bias_per_row = 0;
  p.relu = relu;
  p.mask = mask;
  p.colsum = colsum;
  p.batch = 1, p.a_batched = p.b_batched = p.batch_reduce = 0, p.splits = 1, p.c_batch_stride = 0;
  p.num_m_blocks = int((M + BLOCK_M - 1) / BLOCK_M);
  p.num_n_blocks = 0;
  p.num_k_blocks = int((K + BLOCK_K - 1) / BLOCK_K);
  p.rs_world = 0;
  p.m_rot = 0;
  p.rs_rows = M;
  for (int i = 0; i < 8; ++i) p.rs_dst[i] = nullptr;
  if (ctx->rs_world > 1) {
    if (M % (int64_t(ctx->rs_world) * BLOCK_M) != 0 || beta != 0.f || c_dtype != NK_F32)
      return nk_set_error(ctx, NK_ERR_INVALID_ARG, "tcgen05 gemm: reduce-scatter epilogue needs M %% (world*128) == 0, "
                          "beta == 0 and f32 output");
    p.rs_world = ctx->rs_world;
    p.rs_rows = M / ctx->rs_world;
    p.m_rot = ctx->rs_rank;
    for (int i = 0; i < ctx->rs_world; ++i) p.rs_dst[i] = ctx->rs_dst[i];
  }
  // K-major: 8-row groups 1024 B apart, +32 B per UMMA_K.  MN-major: 64-wide chunks
  // BLOCK_K*128 B apart (LBO), 8-k groups 1024 B apart (SBO), +16 rows * 128 B per UMMA_K.
  p.a_lbo = a_mn ? BLOCK_K * 128 : 16;
  p.a_sbo = 1024;
  p.a_kstep = a_mn ? UMMA_K * 128 : UMMA_K * 2;
  p.b_lbo = b_mn ? BLOCK_K * 128 : 16;
  p.b_sbo = 1024;
  p.b_kstep = b_mn ? UMMA_K * 128 : UMMA_K * 2;
  if (knobs.a_lbo) p.a_lbo = knobs.a_lbo;
  if (knobs.a_sbo) p.a_sbo = knobs.a_sbo;
  if (knobs.b_lbo) p.b_lbo = knobs.b_lbo;
  if (knobs.b_sbo) p.b_sbo = knobs.b_sbo;

  // CTA pairs (256 x 256 tiles over two SMs) wherever a pair has two live row blocks (with the reduce-scatter epilogue:
  // when the 256 rows of a pair belong to one owner); narrow outputs stay on single CTAs
  p.cta_pairs = (block_n == 256 && M > BLOCK_M && (ctx->sm_count % 2) == 0 && knobs.single_cta == 0 &&
                 (p.rs_world == 0 || M % (int64_t(p.rs_world) * 2 * BLOCK_M) == 0)) ? 1 : 0;   // a pair's 256 rows: one owner
  p.allow_split = knobs.no_split == 0 && ctx->gemm_tail_split;
  CUtensorMap ta, tb;
  int rc;
  if (a_mn)
    rc = make_tmap_2d(ctx, &ta, A, K, M, lda, 64, BLOCK_K);
  else
    rc = make_tmap_2d(ctx, &ta, A, M, K, lda, BLOCK_K, BLOCK_M);
  if (rc) return rc;
  if (b_mn)
    rc = make_tmap_2d(ctx, &tb, B, K, N, ldb, 64, BLOCK_K);
  else
    rc = make_tmap_2d(ctx, &tb, B, N, K, ldb, BLOCK_K, (uint32_t)(p.cta_pairs ? block_n / 2 : block_n));
  if (rc) return rc;

  static const char* names[2][2][5] = {
      {{"tcgen05_nt_128x256", "tcgen05_nt_128x128", "tcgen05_nt_128x64", "tcgen05_nt_128x32", "tcgen05_nt_128x16"},
       {"tcgen05_nn_128x256", "tcgen05_nn_128x128", "tcgen05_nn_128x64", "", ""}},
      {{"tcgen05_tt_128x256", "tcgen05_tt_128x128", "tcgen05_tt_128x64", "tcgen05_tt_128x32", "tcgen05_tt_128x16"},
       {"tcgen05_tn_128x256", "tcgen05_tn_128x128", "tcgen05_tn_128x64", "", ""}}};
  const int bi = block_n == 256 ? 0 : block_n == 128 ? 1 : block_n == 64 ? 2 : block_n == 32 ? 3 : 4;
  static const char* pair_names[2][2] = {{"tcgen05_nt_2cta_256x256", "tcgen05_nn_2cta_256x256"},
                                         {"tcgen05_tt_2cta_256x256", "tcgen05_tn_2cta_256x256"}};
  ctx->last_gemm_kernel = p.cta_pairs ? pair_names[a_mn][b_mn] : names[a_mn][b_mn][bi];

  // staged TMA-store epilogue wherever the output is plain (no accumulate, mask or reduce-scatter) and TMA-addressable
  CUtensorMap tc;
  p.tma_store = 0;
  if (beta == 0.f && !mask && !colsum && p.rs_world == 0 && block_n >= 64 && knobs.direct_store == 0 &&
      (reinterpret_cast<uintptr_t>(C) & 15) == 0 && (ldc * int64_t(nk_dtype_size(c_dtype))) % 16 == 0) {
    if (make_tmap_c(ctx, &tc, C, M, N, ldc, c_dtype) == NK_OK) p.tma_store = 1;
  }
#define NK_TC(AM, BM_)                                                                 \
  (c_dtype == NK_BF16 ? launch_bn<AM, BM_, __nv_bfloat16>(ctx, block_n, ta, tb, p, &tc) \
                      : launch_bn<AM, BM_, float>(ctx, block_n, ta, tb, p, &tc))
  if (!a_mn && !b_mn) return NK_TC(false, false);
  if (!a_mn && b_mn) return NK_TC(false, true);
  if (a_mn && !b_mn) return NK_TC(true, false);
  return NK_TC(true, true);
#undef NK_TC
}

// 3-D bf16 tensor map (cols, rows, batch)
static int make_tmap_3d(nk_ctx* ctx, CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int64_t batch,
                        int64_t batch_stride, uint32_t box_cols, uint32_t box_rows) {
  if (!ctx->encode_tiled) return nk_set_error(ctx, NK_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)batch_stride * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return nk_set_error(ctx, NK_ERR_CUDA, "cuTensorMapEncodeTiled (3-d) failed (%d) rows=%lld cols=%lld ld=%lld batch=%lld stride=%lld",
                        (int)r, (long long)rows, (long long)cols, (long long)ld, (long long)batch, (long long)batch_stride);
  return NK_OK;
}

// `batch` products C_b = alpha * op(A_b).op(B_b) (+ row bias, ReLU) with operands batch_stride elements apart (stride 0 = the
// same operand for every batch), or -- reduce != 0 -- ONE product C += alpha * sum_b op(A_b).op(B_b) (C f32, accumulated
// with atomics: the caller zeroes / scales C first).  The engine behind the im2col convolution path (nk_conv_gemm.cu).
// Returns NK_ERR_UNSUPPORTED (last_error untouched) when the operands are not TMA-addressable.
int nk_gemm_tcgen05_batched(nk_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, float alpha, const void* A,
                            int64_t lda, int64_t strideA, const void* B, int64_t ldb, int64_t strideB, void* C, int64_t ldc,
                            int64_t strideC, int64_t batch, int c_dtype, const void* row_bias, int bias_dtype, int relu,
                            int reduce) {
  if (!nk_gemm_tcgen05_supported(transA, transB, M, N, K, A, lda, B, ldb)) return NK_ERR_UNSUPPORTED;
  if (batch < 1 || batch > (int64_t(1) << 30) || (strideA * 2) % 16 != 0 || (strideB * 2) % 16 != 0) return NK_ERR_UNSUPPORTED;
  if (reduce && c_dtype != NK_F32) return NK_ERR_UNSUPPORTED;
  const bool a_mn = transA != 0, b_mn = transB == 0;
  int block_n = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  GemmParams p;
  p.M = M, p.N = N, p.K = K, p.ldc = ldc, p.C = C;
  p.bias = row_bias, p.bias_bf16 = bias_dtype == NK_BF16, p.bias_per_row = 1;
  p.alpha = alpha, p.beta = 0.f, p.relu = relu, p.mask = nullptr, p.colsum = nullptr;
  p.num_m_blocks = int((M + BLOCK_M - 1) / BLOCK_M);
  p.num_n_blocks = 0;
  p.num_k_blocks = int((K + BLOCK_K - 1) / BLOCK_K);
  p.rs_world = 0, p.m_rot = 0, p.rs_rows = M;
  for (int i = 0; i < 8; ++i) p.rs_dst[i] = nullptr;
  p.tma_store = 0, p.cta_pairs = 0, p.allow_split = 0, p.split_tiles = 0, p.full_waves = 0, p.split_ws = nullptr,
  p.split_flags = nullptr;
  p.batch = int(batch), p.a_batched = strideA != 0, p.b_batched = strideB != 0, p.batch_reduce = reduce ? 1 : 0, p.splits = 1;
  p.c_batch_stride = strideC;
  p.a_lbo = a_mn ? BLOCK_K * 128 : 16, p.a_sbo = 1024, p.a_kstep = a_mn ? UMMA_K * 128 : UMMA_K * 2;
  p.b_lbo = b_mn ? BLOCK_K * 128 : 16, p.b_sbo = 1024, p.b_kstep = b_mn ? UMMA_K * 128 : UMMA_K * 2;
  CUtensorMap ta, tb;
  int rc;
  if (p.a_batched)
    rc = a_mn ? make_tmap_3d(ctx, &ta, A, K, M, lda, batch, strideA, 64, BLOCK_K)
              : make_tmap_3d(ctx, &ta, A, M, K, lda, batch, strideA, BLOCK_K, BLOCK_M);
  else
    rc = a_mn ? make_tmap_2d(ctx, &ta, A, K, M, lda, 64, BLOCK_K) : make_tmap_2d(ctx, &ta, A, M, K, lda, BLOCK_K, BLOCK_M);
  if (rc) return rc;
  if (p.b_batched)
    rc = b_mn ? make_tmap_3d(ctx, &tb, B, K, N, ldb, batch, strideB, 64, BLOCK_K)
              : make_tmap_3d(ctx, &tb, B, N, K, ldb, batch, strideB, BLOCK_K, (uint32_t)block_n);
  else
    rc = b_mn ? make_tmap_2d(ctx, &tb, B, K, N, ldb, 64, BLOCK_K) : make_tmap_2d(ctx, &tb, B, N, K, ldb, BLOCK_K, (uint32_t)block_n);
  if (rc) return rc;
  ctx->last_gemm_kernel = "tcgen05_batched";
#define NK_TCB(BN, AM, BM_)                                                                          \
  (c_dtype == NK_BF16 ? launch_cfg<BN, AM, BM_, __nv_bfloat16, true>(ctx, ta, tb, p)                  \
                      : launch_cfg<BN, AM, BM_, float, true>(ctx, ta, tb, p))
#define NK_TCB_BN(AM, BM_) (block_n == 256 ? NK_TCB(256, AM, BM_) : block_n == 128 ? NK_TCB(128, AM, BM_) : NK_TCB(64, AM, BM_))
  if (!a_mn && !b_mn) return NK_TCB_BN(false, false);
  if (!a_mn && b_mn) return NK_TCB_BN(false, true);
  if (a_mn && b_mn) return NK_TCB_BN(true, true);
  return nk_set_error(ctx, NK_ERR_UNSUPPORTED, "batched gemm: the TT form is not instantiated");
#undef NK_TCB_BN
#undef NK_TCB
}

extern "C" {

int nk_gemm_bias_act(nk_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, float alpha,
                     const void* A, int64_t lda, const void* B, int64_t ldb, float beta, void* C, int64_t ldc,
                     int ab_dtype, int c_dtype, const void* bias, int bias_dtype, int relu) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, nk_dtype_ok(ab_dtype) && nk_dtype_ok(c_dtype), "nk_gemm: bad dtype");
  NK_REQUIRE(ctx, M >= 0 && N >= 0 && K >= 0, "nk_gemm: negative dimension");
  if (M == 0 || N == 0) return NK_OK;
  NK_REQUIRE(ctx, C != nullptr && (K == 0 || (A && B)), "nk_gemm: NULL pointer");
  NK_REQUIRE(ctx, lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N,
             "nk_gemm: leading dimension too small (lda=%lld ldb=%lld ldc=%lld for M=%lld N=%lld K=%lld tA=%d tB=%d)",
             (long long)lda, (long long)ldb, (long long)ldc, (long long)M, (long long)N, (long long)K, transA, transB);
  NK_REQUIRE(ctx, !bias || nk_dtype_ok(bias_dtype), "nk_gemm: bad bias dtype");
  const bool want_tc = ab_dtype == NK_BF16 && ctx->gemm_engine != NK_GEMM_SIMT && K > 0;
  if (want_tc) {
    int rc = nk_gemm_tcgen05(ctx, transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, c_dtype, bias,
                             bias_dtype, relu, nullptr, nullptr);
    if (rc != NK_ERR_UNSUPPORTED) return rc;
    if (ctx->gemm_engine == NK_GEMM_TCGEN05)
      return nk_set_error(ctx, NK_ERR_UNSUPPORTED,
                          "nk_gemm: tcgen05 engine forced but operands are not TMA-addressable "
                          "(16-byte aligned base, leading dimension multiple of 8 elements)");
  } else if (ctx->gemm_engine == NK_GEMM_TCGEN05) {
    return nk_set_error(ctx, NK_ERR_UNSUPPORTED, "nk_gemm: tcgen05 engine forced but operands are not bf16");
  }
  return nk_gemm_simt(ctx, transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, ab_dtype, c_dtype, bias,
                      bias_dtype, relu);
}

int nk_gemm_relu_bwd_colsum(nk_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda,
                            const void* B, int64_t ldb, float beta, void* C, int64_t ldc, int ab_dtype, int c_dtype,
                            const void* relu_operand, float* colsum) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, nk_dtype_ok(ab_dtype) && nk_dtype_ok(c_dtype), "nk_gemm_relu_bwd: bad dtype");
  NK_REQUIRE(ctx, M >= 0 && N >= 0 && K > 0, "nk_gemm_relu_bwd: bad dimension");
  if (M == 0 || N == 0) return NK_OK;
  NK_REQUIRE(ctx, A && B && C && relu_operand, "nk_gemm_relu_bwd: NULL pointer");
  NK_REQUIRE(ctx, lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, "nk_gemm_relu_bwd: leading dimension too small");
  NK_REQUIRE(ctx, !colsum || beta == 0.f, "nk_gemm_relu_bwd_colsum: the column sums are those of the product (beta must be 0)");
  if (ab_dtype == NK_BF16 && ctx->gemm_engine != NK_GEMM_SIMT) {
    int rc = nk_gemm_tcgen05(ctx, transA, transB, M, N, K, 1.f, A, lda, B, ldb, beta, C, ldc, c_dtype, nullptr, NK_F32, 0,
                             relu_operand, colsum);
    if (rc != NK_ERR_UNSUPPORTED) return rc;
  }
  // the skinny NN shape (K <= 16: the layer above is 10 wide) masks (and sums) in its own epilogue
  if (!transA && !transB) {
    int rc = nk_gemm_simt_small_k_masked(ctx, M, N, K, A, lda, B, ldb, beta, C, ldc, ab_dtype, c_dtype, relu_operand, colsum);
    if (rc != NK_ERR_UNSUPPORTED) return rc;
  }
  if (colsum) return NK_ERR_UNSUPPORTED;   // nothing done: the caller sums the columns itself after the plain call
  // operands the tensor-core engine cannot take: the product into a temporary, then the ordinary ReLU backward
  void* tmp = nullptr;
  int rc = nk_alloc_uninit(ctx, size_t(M) * size_t(N) * nk_dtype_size(c_dtype), &tmp);
  if (rc) return rc;
  rc = nk_gemm_simt(ctx, transA, transB, M, N, K, 1.f, A, lda, B, ldb, 0.f, tmp, N, ab_dtype, c_dtype, nullptr, NK_F32, 0);
  if (rc == NK_OK) {
    if (ldc == N) {
      rc = nk_relu_bwd(ctx, C, relu_operand, tmp, size_t(M) * size_t(N), c_dtype, beta);
    } else {
      rc = nk_set_error(ctx, NK_ERR_UNSUPPORTED, "nk_gemm_relu_bwd: strided output needs the tensor-core engine");
    }
  }
  nk_free(ctx, tmp);
  return rc;
}

int nk_gemm_relu_bwd(nk_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda,
                     const void* B, int64_t ldb, float beta, void* C, int64_t ldc, int ab_dtype, int c_dtype,
                     const void* relu_operand) {
  return nk_gemm_relu_bwd_colsum(ctx, transA, transB, M, N, K, A, lda, B, ldb, beta, C, ldc, ab_dtype, c_dtype, relu_operand,
                                 nullptr);
}

int nk_gemm(nk_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, float alpha, const void* A,
            int64_t lda, const void* B, int64_t ldb, float beta, void* C, int64_t ldc, int ab_dtype, int c_dtype) {
  return nk_gemm_bias_act(ctx, transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, ab_dtype, c_dtype,
                          nullptr, NK_F32, 0);
}

}  // extern "C"
