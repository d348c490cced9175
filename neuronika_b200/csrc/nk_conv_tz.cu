// Thin-input convolution forward on tcgen05 with Toeplitz-expanded weights (sm_100a): stride 1, dilation 1, groups 1,
// bf16, NCHW, few input channels (config 3: Conv2d 3->64 k3 on 224x224, the HBM-bound case of SURVEY.md 8-d).
// Reference semantics: convolution/mod.rs:85-123 (cross-correlation, no padding).
//
// The im2col rows of a horizontal tap j are the image rows shifted by j pixels; TMA cannot fetch a box that starts on a
// 2-byte boundary, and building the shifted copies in shared memory cost the first kernel 20 % of its instructions
// (profiles/r01_conv_ncu.md).  Here NO shifted copy exists.  Eight consecutive output pixels (q = 8a + dq) share the
// input pixels u = 8a + du, du = 0..kw+6, so with the expanded ("Toeplitz") weights
//        T[(o, dq)][(c, i, du)] = W[o, c, i, du - dq]   (0 <= du - dq < kw, else 0)
//        y[o, p, 8a + dq] = sum_{c,i,du} T[(o,dq)][(c,i,du)] * x[c, p + i, 8a + du]
// the pixel operand for a fixed (c, i) is the image row itself, read as 16-byte groups:
//        A[m = (r, a)][k = du]  =  x[c, p0 + r + i, 8a + du]
// du = 0..7 is group a of the row and du = 8..15 is group a+1 -- the SAME bytes 16 further on.  In the UMMA K-major
// no-swizzle layout a row of the operand is 16 bytes and the two K-halves of a K = 16 step are LBO apart, so the
// descriptor simply says LBO = 16 bytes (the second half overlaps the next row of the first) and SBO = 128 bytes:
// the tensor core reads the raw TMA-delivered rows, with a 512-byte row pitch (224 pixels + zero fill) that makes
// m = r * 32 + a linear over four image rows.  One UMMA (M = 128 = 4 rows x 32 groups, N = 256 = 32 channels x 8 dq,
// K = 16) per (c, i) and channel half; the multiplications by the structural zeros of T are free -- the tensor pipe
// has 2x headroom at this arithmetic intensity.
//
// D[(r,a)][(o,dq)] lives in TMEM (128 lanes x 512 columns: two channel halves, each drained while the other is being
// computed).  A thread of the epilogue owns one pixel group (r, a) and reads, per channel, 8 consecutive columns = 8
// consecutive pixels = 16 bytes of y.  y rows have a Wo*2-byte pitch (444 B at config 3): only 4-byte aligned, so
// 16-byte global stores are illegal on most rows.  The tile's rows of one channel are CONTIGUOUS in memory
// (R x Wo pixels), so the groups are staged in shared memory at the same misalignment as their global address
// (4-byte shared stores) and the span is then written with aligned 16-byte stores -- 4x fewer LSU transactions than the
// 4-byte stores of the first kernel (27 % of its instructions).
//
// Warp roles (576 threads): warp 0 = TMA (x rows), warp 1 = MMA issue + TMEM, warps 2..17 = epilogue: four groups of
// four warps, one group per 16 output channels (warp w reads TMEM lanes 32*(w%4)..).  Sixteen epilogue warps because the
// drain is a chain of dependent steps (tcgen05.ld -> convert -> shared store -> barrier -> copy-out): with eight the SM
// sat at 5x the HBM budget per tile (first version: 1.21 ms), the latency of each step exposed.  The tcgen05.ld of the
// next slab is in flight while the current one is copied out.  Algorithmic bytes = 2(|x| + |y|).
#include "nk_internal.cuh"
#include "nk_ptx.cuh"

namespace {

constexpr int kThreads = 576;       // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue (4 channel quarters x 4 TMEM lane quarters)
constexpr int kRowsPerTile = 4;     // output rows per tile (UMMA M = 4 x 32 groups)
constexpr int kGroups = 32;         // 16-byte groups per image row in shared memory (pitch 512 B = 256 pixels)
constexpr int kRowPitch = kGroups * 16;
constexpr int kMaxSteps = 9;        // (c, i) pairs: Cin * kh <= 9
constexpr int kStepBytes = 512 * 32;  // T of one K-step: 512 rows (o, dq) x 16 du x 2 B
constexpr int kStages = 2;
constexpr int kSlabCh = 4;          // channels per staging slab
constexpr int kEpiGroups = 4;       // channel quarters, each with its own pair of slabs and its own named barrier
constexpr int kChPitch = 1808;      // bytes per staged channel: 4 rows x 444 B + 16 B of misalignment, 16-byte multiple
constexpr int kSlabBytes = kSlabCh * kChPitch;

struct TzP {
  int n, cin, h, w, cout, kh, kw, ho, wo;
  int steps;            // cin * kh
  int in_rows;          // kRowsPerTile + kh - 1
  int blocks_per_img;   // ceil(ho / 4)
  int num_tiles;
  uint32_t stage_bytes; // cin * in_rows * 512
  const __nv_bfloat16* wt;
  const __nv_bfloat16* bias;
  __nv_bfloat16* y;
  int relu;
};

// UMMA shared-memory descriptor, no swizzle (layout type 0): [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version 1
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr >> 4) & 0x3FFFu);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= uint64_t(1) << 46;
  return d;
}

__device__ __forceinline__ void named_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// shared -> global bulk copy (dst and src 16-byte aligned, size a multiple of 16), tracked by bulk async-groups
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void tmem_ld_32x32b_x16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

template <bool kBias, bool kRelu>
__global__ void __launch_bounds__(kThreads, 1)
conv_fwd_tz_kernel(const __grid_constant__ CUtensorMap tmap_x, const TzP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 127u) & ~127u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  // layout: [T: steps x 16 KB][x stages][staging: 2 halves x 2 slabs][barriers]
  const uint32_t t_off = 0;
  const uint32_t x_off = t_off + p.steps * kStepBytes;
  const uint32_t sg_off = x_off + kStages * p.stage_bytes + 512;   // + one row of slack: the K-halo of the last group
  const uint32_t bias_off = sg_off + kEpiGroups * 2 * kSlabBytes;   // 64 floats
  const uint32_t bar_off = bias_off + 256;
  const uint32_t bar_base = base + bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tmem_full_bar = [&](int hf) { return bar_base + 8u * (2 * kStages + hf); };
  auto tmem_empty_bar = [&](int hf) { return bar_base + 8u * (2 * kStages + 2 + hf); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + bar_off + 8u * (2 * kStages + 4));

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- expanded weights T, built once per CTA: [step][k-half][row group of 8][row][8 du] (K-major, no swizzle)
  {
    uint4* tz = reinterpret_cast<uint4*>(base_ptr + t_off);
    const int nvec = p.steps * kStepBytes / 16;
    for (int i = threadIdx.x; i < nvec; i += kThreads) tz[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    __nv_bfloat16* t16 = reinterpret_cast<__nv_bfloat16*>(base_ptr + t_off);
    const int total = p.cout * p.steps * p.kw * 8;   // (o, s, j, dq)
    for (int i = threadIdx.x; i < total; i += kThreads) {
      int rem = i;
      const int dq = rem & 7;
      rem >>= 3;
      const int j = rem % p.kw;
      rem /= p.kw;
      const int s = rem % p.steps, o = rem / p.steps;   // s = c * kh + i
      const int du = dq + j, nrow = o * 8 + dq;
      const uint32_t off = uint32_t(s) * kStepBytes + uint32_t(du >> 3) * 8192u + uint32_t(nrow >> 3) * 128u +
                           uint32_t(nrow & 7) * 16u + uint32_t(du & 7) * 2u;
      t16[off >> 1] = p.wt[(int64_t(o) * p.steps + s) * p.kw + j];   // w[o][c][i][j], (c, i) = s
    }
  }
  if (threadIdx.x < 64)
    reinterpret_cast<float*>(base_ptr + bias_off)[threadIdx.x] = p.bias ? __bfloat162float(p.bias[threadIdx.x]) : 0.f;
  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int hf = 0; hf < 2; ++hf) {
      ptx::mbar_init(tmem_full_bar(hf), 1);
      ptx::mbar_init(tmem_empty_bar(hf), 16);  // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();   // T was written with generic stores; the tensor core reads it through the async proxy
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp_idx == 0) {
    // ===================================================== TMA: the input rows of a tile, all channels, one box
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int img = tile / p.blocks_per_img, blk = tile - img * p.blocks_per_img;
        ptx::mbar_wait_relaxed(empty_bar(stage), phase ^ 1u);   // two tiles of slack: do not spin on the issue slots
        ptx::mbar_expect_tx(full_bar(stage), p.stage_bytes);
        // box {256 (W, zero filled beyond w), in_rows (H, zero filled beyond h), cin}
        ptx::tma_load_3d(base + x_off + stage * p.stage_bytes, &tmap_x, full_bar(stage), 0, blk * kRowsPerTile, img * p.cin);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================================================== MMA issue
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, 256, false, false);
      int stage = 0;
      uint32_t phase = 0, aphase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        ptx::mbar_wait_relaxed(full_bar(stage), phase);
        ptx::tc_fence_after();
        const uint32_t xs = base + x_off + stage * p.stage_bytes;
        for (int hf = 0; hf < 2; ++hf) {
          // the drain of the OTHER half (~4 000 cycles) hides this half's wake-up + 9 UMMAs (~1 200): relaxed wait
          ptx::mbar_wait_relaxed(tmem_empty_bar(hf), aphase ^ 1u);
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + uint32_t(hf * 256);
          for (int s = 0; s < p.steps; ++s) {
            const int c = s / p.kh, i = s - c * p.kh;
            // A: rows (r, a) of image rows c, i + r: 16 B per row, 8-row groups 128 B apart, K halves 16 B apart
            const uint64_t adesc = make_desc_nosw(xs + uint32_t(c * p.in_rows + i) * kRowPitch, 16, 128);
            // B: T rows (o, dq) of this half: K halves 8192 B apart, 8-row groups 128 B apart
            const uint64_t bdesc = make_desc_nosw(base + t_off + uint32_t(s) * kStepBytes + uint32_t(hf) * 4096u, 8192, 128);
            ptx::mma_f16_ss(tmem_d, adesc, bdesc, idesc, s != 0 ? 1u : 0u);
          }
          ptx::mma_commit(tmem_full_bar(hf));
        }
        ptx::mma_commit(empty_bar(stage));   // the x rows are free once both halves' MMAs retire
        aphase ^= 1u;
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    // ===================================================== epilogue: 4 channel quarters x 4 TMEM lane quarters
    // All per-element address arithmetic is 32-bit and relative to one 64-bit base per tile: the first version spent 97
    // instructions per 16 bytes of output (24.5 k warp instructions per tile against a budget of ~10 k at the HBM bound).
    const int e = warp_idx - 2;          // 0..15
    const int cq = e >> 2;               // warp group: 8 channels of each TMEM half
    const int q = warp_idx & 3;          // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;         // pixel group (r, a)
    const int r = m >> 5, a = m & 31;
    const int vi = (e & 3) * 32 + lane;  // 0..127 inside the group
    const uint32_t sg_u32 = base + sg_off + cq * 2 * kSlabBytes;
    const float* sbias = reinterpret_cast<const float*>(base_ptr + bias_off);
    const int valid_px = p.wo - 8 * a;       // pixels of this group inside the row (<= 0: none)
    const int words = valid_px >= 8 ? 4 : (valid_px > 0 ? (valid_px + 1) >> 1 : 0);   // wo even: valid_px is even
    const uint32_t ch_bytes = uint32_t(p.ho) * uint32_t(p.wo) * 2u;   // < 2^31 (host check)
    const uint32_t pix_off = uint32_t((r * p.wo + 8 * a) * 2);
    constexpr int kSlabs = 8 / kSlabCh;      // slabs per tile, half and group (8 channels)
    uint32_t aphase = 0;
    uint32_t slab_sel = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int img = tile / p.blocks_per_img, blk = tile - img * p.blocks_per_img;
      const int r0 = blk * kRowsPerTile;
      const int rows_valid = min(kRowsPerTile, p.ho - r0);
      const bool row_ok = r < rows_valid;
      const bool w0 = row_ok && words > 0, w1 = row_ok && words > 1, w2 = row_ok && words > 2, w3 = row_ok && words > 3;
      const uint32_t bytes = uint32_t(rows_valid * p.wo * 2);   // contiguous bytes of one channel in this tile
      // ALL sixteen warps drain half 0 while the tensor core computes half 1 (and vice versa): a group never idles
      // while "its" accumulator is recomputed
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        const int ch0 = hf * 32 + cq * 8;      // first channel of this group in this half
        // the one 64-bit address: channel ch0, first row of the tile
        uint8_t* const g_base = reinterpret_cast<uint8_t*>(p.y) + (((int64_t(img) * p.cout + ch0) * p.ho + r0) * p.wo) * 2;
        const uint32_t shift0 = uint32_t(reinterpret_cast<uintptr_t>(g_base)) & 15u;
        ptx::mbar_wait(tmem_full_bar(hf), aphase);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(hf * 256 + cq * 64);
        uint32_t v[kSlabCh * 8];
        tmem_ld_32x32b_x32_nowait(taddr, v);
#pragma unroll 1
        for (int sl = 0; sl < kSlabs; ++sl) {
          const uint32_t slab = sg_u32 + slab_sel * kSlabBytes;
          auto shift_of = [&](int cc) { return (shift0 + uint32_t(sl * kSlabCh + cc) * ch_bytes) & 15u; };
          ptx::tmem_ld_wait();
#pragma unroll
          for (int cc = 0; cc < kSlabCh; ++cc) {
            const float b = kBias ? sbias[ch0 + sl * kSlabCh + cc] : 0.f;
            // same misalignment as the global address of the channel's span in this tile
            const uint32_t dst = slab + cc * kChPitch + shift_of(cc) + pix_off;
            uint32_t w4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float f0 = __uint_as_float(v[cc * 8 + 2 * k]), f1 = __uint_as_float(v[cc * 8 + 2 * k + 1]);
              if (kBias) f0 += b, f1 += b;
              if (kRelu) f0 = fmaxf(f0, 0.f), f1 = fmaxf(f1, 0.f);
              w4[k] = pack2(f0, f1);
            }
            // four predicated stores, no branches
            asm volatile(
                "{\n\t.reg .pred p0, p1, p2, p3;\n\t"
                "setp.ne.b32 p0, %5, 0;\n\tsetp.ne.b32 p1, %6, 0;\n\tsetp.ne.b32 p2, %7, 0;\n\tsetp.ne.b32 p3, %8, 0;\n\t"
                "@p0 st.shared.b32 [%0], %1;\n\t@p1 st.shared.b32 [%0+4], %2;\n\t@p2 st.shared.b32 [%0+8], %3;\n\t"
                "@p3 st.shared.b32 [%0+12], %4;\n\t}"
                ::"r"(dst), "r"(w4[0]), "r"(w4[1]), "r"(w4[2]), "r"(w4[3]), "r"(int(w0)), "r"(int(w1)), "r"(int(w2)), "r"(int(w3))
                : "memory");
          }
          // the registers are free again: fetch the next slab's columns while this one is copied out
          if (sl + 1 < kSlabs) tmem_ld_32x32b_x32_nowait(taddr + uint32_t((sl + 1) * kSlabCh * 8), v);
          // ---- copy-out.  The body of each channel's span (16-byte aligned start, multiple of 16 bytes) leaves through
          // the bulk-copy engine: ONE instruction per channel instead of 111 LDS.128 + STG.128 pairs; <= 3 head and <= 3
          // tail words per channel go out as words.
          ptx::fence_proxy_async();                 // this thread's slab writes -> visible to the async proxy
          if (vi < kSlabCh) bulk_wait_read_all();   // this thread's bulk reads of the other slab are done: it may be rewritten
          named_barrier(1 + cq, 128);               // the slab is complete
          if (vi < kSlabCh) {                       // threads 0..3: one channel each
            const int cc = vi;
            const uint32_t head = (16u - shift_of(cc)) & 15u;
            const uint32_t body = (bytes - head) & ~15u;
            if (body) bulk_store(g_base + uint32_t(sl * kSlabCh + cc) * ch_bytes + head, slab + cc * kChPitch + shift_of(cc) + head, body);
            bulk_commit();
          } else if (vi >= 32 && vi < 32 + kSlabCh * 8) {   // threads 32..63: (channel, word): words 0..2 head, 4..6 tail
            const int cc = (vi - 32) >> 3, t = (vi - 32) & 7;
            const uint32_t sh = shift_of(cc);
            const uint32_t head = (16u - sh) & 15u;
            const uint32_t body = (bytes - head) & ~15u;
            const uint32_t head_w = head >> 2, tail_w = (bytes - head - body) >> 2;
            uint32_t off = 0xffffffffu;
            if (uint32_t(t) < head_w) off = uint32_t(t) * 4u;
            else if (t >= 4 && uint32_t(t - 4) < tail_w) off = head + body + uint32_t(t - 4) * 4u;
            if (off != 0xffffffffu) {
              uint32_t wv;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wv) : "r"(slab + cc * kChPitch + sh + off));
              *reinterpret_cast<uint32_t*>(g_base + uint32_t(sl * kSlabCh + cc) * ch_bytes + off) = wv;
            }
          }
          slab_sel ^= 1u;
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tmem_empty_bar(hf));
      }
      aphase ^= 1u;
    }
    bulk_wait_all();   // every bulk store this thread issued has been written before the CTA retires
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

// Does the Toeplitz engine take this forward convolution?  (bf16, s1 d1 g1 checked by the caller)
bool nk_conv_tz_supported(int64_t n, int64_t cin, int64_t h, int64_t w, int64_t cout, int64_t kh, int64_t kw, const void* x,
                          const void* y) {
  const int64_t ho = h - kh + 1, wo = w - kw + 1;
  if (n <= 0 || ho <= 0 || wo <= 0) return false;
  if (cin * kh > kMaxSteps || kw > 9) return false;            // T must stay resident: steps x 16 KB
  if (cout != 64) return false;                                // N = cout * 8 = 512 columns = the whole TMEM
  if (w % 8 != 0 || w > 248) return false;                     // TMA row stride multiple of 16 B; row + halo within 256 pixels
  if (wo % 2 != 0) return false;                               // rows of y start 4-byte aligned
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 3)) return false;
  if (int64_t(kRowsPerTile) * wo * 2 + 16 > kChPitch) return false;
  if (ho * wo * 2 >= (int64_t(1) << 31)) return false;
  if (n * cin > (int64_t(1) << 30)) return false;
  const size_t smem = size_t(cin * kh) * kStepBytes + size_t(kStages) * size_t(cin * (kRowsPerTile + kh - 1) * kRowPitch) + 512 +
                      kEpiGroups * 2 * kSlabBytes + 256 + 256 + 128;
  if (smem > 232448) return false;
  return true;
}

int nk_conv_tz_fwd(nk_ctx* ctx, void* y, const void* x, const void* w, const void* bias, int relu, int64_t n, int64_t cin,
                   int64_t h, int64_t wd, int64_t cout, int64_t kh, int64_t kw) {
  if (!ctx->encode_tiled) return nk_set_error(ctx, NK_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  TzP p;
  p.n = int(n), p.cin = int(cin), p.h = int(h), p.w = int(wd), p.cout = int(cout), p.kh = int(kh), p.kw = int(kw);
  p.ho = int(h - kh + 1), p.wo = int(wd - kw + 1);
  p.steps = int(cin * kh);
  p.in_rows = kRowsPerTile + int(kh) - 1;
  p.blocks_per_img = (p.ho + kRowsPerTile - 1) / kRowsPerTile;
  p.num_tiles = p.n * p.blocks_per_img;
  p.stage_bytes = uint32_t(p.cin * p.in_rows * kRowPitch);
  p.wt = static_cast<const __nv_bfloat16*>(w);
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.y = static_cast<__nv_bfloat16*>(y);
  p.relu = relu;
  // x as (W, H, N*C), box {256, in_rows, cin}: columns beyond W and rows beyond H are zero filled by TMA
  CUtensorMap tm;
  cuuint64_t dims[3] = {(cuuint64_t)wd, (cuuint64_t)h, (cuuint64_t)(n * cin)};
  cuuint64_t strides[2] = {(cuuint64_t)wd * 2, (cuuint64_t)wd * h * 2};
  cuuint32_t box[3] = {256, (cuuint32_t)p.in_rows, (cuuint32_t)cin};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      &tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(x), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return nk_set_error(ctx, NK_ERR_CUDA, "cuTensorMapEncodeTiled (conv tz) failed (%d)", (int)r);
  const size_t smem = size_t(p.steps) * kStepBytes + size_t(kStages) * p.stage_bytes + 512 + kEpiGroups * 2 * kSlabBytes + 256 + 256 + 128;
  if (smem > 232448) return nk_set_error(ctx, NK_ERR_UNSUPPORTED, "conv tz: %zu bytes of shared memory", smem);
  static bool attr_done[64] = {};
  if (!attr_done[ctx->device & 63]) {
    NK_CUDA(ctx, cudaFuncSetAttribute(conv_fwd_tz_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    NK_CUDA(ctx, cudaFuncSetAttribute(conv_fwd_tz_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    NK_CUDA(ctx, cudaFuncSetAttribute(conv_fwd_tz_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    NK_CUDA(ctx, cudaFuncSetAttribute(conv_fwd_tz_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_done[ctx->device & 63] = true;
  }
  int grid = ctx->sm_count < p.num_tiles ? ctx->sm_count : p.num_tiles;
  if (bias && relu)
    conv_fwd_tz_kernel<true, true><<<grid, kThreads, smem, ctx->stream>>>(tm, p);
  else if (bias)
    conv_fwd_tz_kernel<true, false><<<grid, kThreads, smem, ctx->stream>>>(tm, p);
  else if (relu)
    conv_fwd_tz_kernel<false, true><<<grid, kThreads, smem, ctx->stream>>>(tm, p);
  else
    conv_fwd_tz_kernel<false, false><<<grid, kThreads, smem, ctx->stream>>>(tm, p);
  NK_LAUNCHED(ctx, "conv_fwd_tz");
  ctx->last_conv_kernel = "tcgen05_toeplitz_fwd";
  return NK_OK;
}
