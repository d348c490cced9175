// Development probe: which 4-D TMA box shapes are accepted by the hardware (run on a B200).
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include <stdlib.h>

__global__ void k4d(const __grid_constant__ CUtensorMap tm, __nv_bfloat16* out, int box_elems, int c0, int c1, int c2,
                    int c3, int rank) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t dst = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
  uint8_t* sm_al = smem + (dst - (uint32_t)__cvta_generic_to_shared(smem));
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(box_elems * 2) : "memory");
    if (rank == 4)
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
          "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
          : "memory");
    else if (rank == 3)
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
          "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"(c0), "r"(c1), "r"(c2)
          : "memory");
    else
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
          "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"(c0), "r"(c1)
          : "memory");
  }
  uint32_t ok = 0;
  int spins = 0;
  while (!ok && spins < 1000000) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(b) : "memory");
    ++spins;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < box_elems; i += blockDim.x) out[i] = reinterpret_cast<__nv_bfloat16*>(sm_al)[i];
  if (threadIdx.x == 0 && !ok) printf("  timeout waiting for TMA\n");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  int only = argc > 1 ? atoi(argv[1]) : -1;
  int Wd = argc > 2 ? atoi(argv[2]) : 72;
  int rank = argc > 3 ? atoi(argv[3]) : 4;
  int c0 = argc > 4 ? atoi(argv[4]) : 2;
  printf("probe only=%d W=%d rank=%d\n", only, Wd, rank);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<EncodeTiledFn>(fn);
  const int W = Wd, H = 64, C = 3, N = 2;
  std::vector<__nv_bfloat16> h(W * H * C * N);
  for (size_t i = 0; i < h.size(); ++i) h[i] = __float2bfloat16(float(i % 251));
  __nv_bfloat16 *d, *out;
  cudaMalloc(&d, h.size() * 2);
  cudaMalloc(&out, 65536);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  struct Case { cuuint32_t b0, b1, b2, b3; CUtensorMapSwizzle sw; CUtensorMapL2promotion l2; const char* name; };
  Case cases[] = {
      {64, 16, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "64x16"},
      {64, 8, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "64x8"},
      {64, 3, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "64x3"},
      {64, 1, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "64x1"},
      {64, 9, 1, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "64x9"},
      {64, 3, 3, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "64x3x3x1"},
      {64, 8, 2, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "64x8x2x1"},
      {64, 16, 1, 1, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, "64x16 none"},
  };
  int idx = -1;
  for (auto& cs : cases) {
    ++idx;
    if (only >= 0 && idx != only) continue;
    CUtensorMap tm;
    cuuint64_t dims[4] = {W, H, C, N};
    cuuint64_t strides[3] = {W * 2, W * H * 2, W * H * C * 2};
    cuuint32_t box[4] = {cs.b0, cs.b1, cs.b2, cs.b3};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     cs.sw, cs.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("%-28s encode=%d ", cs.name, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); continue; }
    int elems = cs.b0 * cs.b1 * cs.b2 * cs.b3;
    cudaMemset(out, 0, 65536);
    if (rank < 4) elems /= cs.b3; if (rank < 3) elems /= cs.b2;
    k4d<<<1, 128, 40960>>>(tm, out, elems, c0, 1, 0, 1, rank);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run=%s", cudaGetErrorString(e));
    if (e == cudaSuccess) {
      std::vector<__nv_bfloat16> o(elems);
      cudaMemcpy(o.data(), out, elems * 2, cudaMemcpyDeviceToHost);
      // expected first element of row 0: x[n=1][c=0][h=1][w=2]
      float want = rank == 4 ? float(((1 * C + 0) * H + 1) * W + c0) : float(1 * W + c0);
      printf(" first=%g (want %g mod 251 = %g)", __bfloat162float(o[0]), want, float(int(want) % 251));
    } else {
      printf("  -- context lost, stopping\n");
      return 1;
    }
    printf("\n");
  }
  return 0;
}
