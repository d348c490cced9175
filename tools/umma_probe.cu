// Development probe: issue interval of back-to-back tcgen05.mma instructions as a function of the instruction shape
// (run on a B200).  One CTA, one issuing thread, operands = zeroed shared memory, `nacc` round-robin accumulators.
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../neuronika_b200/csrc -I../include -o umma_probe umma_probe.cu
//   ./umma_probe
// Prints cycles per UMMA for the shapes the convolution kernels use (M 128 x N {32,48,64,128,256} x K 16, K-major and
// MN-major A) against the tensor-math time of that shape (128*N*16 MACs at 4096 MAC/cycle/SM).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "nk_ptx.cuh"

__global__ void __launch_bounds__(128, 1) probe(int n, int a_mn, int nacc, int iters, int dependent, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - raw);
  for (uint32_t i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(bp)[i] = make_uint4(0, 0, 0, 0);
  const uint32_t bar = base + 96 * 1024, slot = bar + 8;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(bp + 96 * 1024 + 8);
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    ptx::tmem_alloc(slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, n, a_mn != 0, false);
    // A: 16 KB at base (K-major: SBO 1024; MN-major: 64-wide atoms 8 KB apart).  B: at base + 32 KB, K-major.
    const uint64_t adesc = a_mn ? ptx::make_smem_desc_sw128(base, 8192, 1024) : ptx::make_smem_desc_sw128(base, 16, 1024);
    const uint64_t bdesc = ptx::make_smem_desc_sw128(base + 32768, 16, 1024);
    uint32_t phase = 0;
    for (int rep = 0; rep < 2; ++rep) {  // rep 0 warms up
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t a = dependent ? 0u : uint32_t(i % nacc);
        ptx::mma_f16_ss(tmem + a * uint32_t(n), adesc, bdesc, idesc, i >= nacc ? 1u : 0u);
      }
      ptx::mma_commit(bar);
      ptx::mbar_wait(bar, phase);
      phase ^= 1u;
      const long long t1 = clock64();
      out[rep] = t1 - t0;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 4096;
  const int ns[] = {16, 32, 48, 64, 128, 256};
  printf("%6s %6s %5s %5s %12s %12s\n", "N", "A", "nacc", "dep", "cyc/UMMA", "math cyc");
  for (int a_mn = 0; a_mn < 2; ++a_mn)
    for (int n : ns)
      for (int dep = 0; dep < 2; ++dep) {
        int nacc = 512 / n;
        if (nacc > 8) nacc = 8;
        probe<<<1, 128, 100 * 1024>>>(n, a_mn, nacc, iters, dep, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("N=%d a_mn=%d: %s\n", n, a_mn, cudaGetErrorString(e));
          return 1;
        }
        printf("%6d %6s %5d %5d %12.1f %12.1f\n", n, a_mn ? "MN" : "K", nacc, dep, double(out[1]) / iters,
               128.0 * n * 16 / 4096.0);
      }
  return 0;
}
