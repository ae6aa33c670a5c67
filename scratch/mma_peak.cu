// Tensor-pipe microbenchmark: back-to-back tcgen05.mma.kind::i8 (M=128, N in {64,128,192,256}, K=32) from fixed smem tiles.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda.h>
#include "../eda-dm_b200/csrc/common.cuh"
#include "../eda-dm_b200/csrc/tc05.cuh"
using namespace edadm;
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128, 1) peak(int n, int iters, int distinct, int f16, long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 4 * 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); fence_proxy_async(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) {
    // f16 idesc: c_format f32 (1<<4), a/b format f16 = 0, N>>3 <<17, M>>4 <<24
    const uint32_t idesc = f16 ? ((1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24)) : make_idesc_i8(n, 0, 1);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int st = distinct ? (it & 3) : 0;
      const uint64_t ad = make_smem_desc(smem_u32(smem + st * 49152));
      const uint64_t bd = make_smem_desc(smem_u32(smem + st * 49152 + 16384));
      for (int k = 0; k < 4; ++k) {
        if (f16) umma_f16(tbase, ad + 2 * k, bd + 2 * k, idesc, 1u);
        else umma_i8(tbase, ad + 2 * k, bd + 2 * k, idesc, 1u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512)); }
}
int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  long long* cyc; cudaMalloc(&cyc, 148 * 8); long long hc[148];
  cudaFuncSetAttribute(peak, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int f16 = 0; f16 < 2; ++f16)
    for (int n : {64, 128, 192, 256})
      for (int grid : {1, 148}) {
        const int iters = 4096;
        peak<<<grid, 128, 200 * 1024>>>(n, iters, 1, f16, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        peak<<<grid, 128, 200 * 1024>>>(n, iters, 1, f16, cyc);
        cudaEventRecord(b); e = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        if (e != cudaSuccess) { printf("N=%d grid=%d f16=%d failed: %s\n", n, grid, f16, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hc, cyc, 8 * grid, cudaMemcpyDeviceToHost);
        const double ops = 2.0 * 128 * n * (f16 ? 16 : 32) * 4.0 * iters * grid;
        printf("%s N=%3d grid=%3d: %.1f cycles/MMA (clock64), %.3f ms, %.1f T%s/s chip-equivalent %s\n", f16 ? "f16" : "i8 ", n, grid,
               (double)hc[0] / (4.0 * iters), ms, ops / (ms * 1e-3) / 1e12 * (grid == 1 ? 148 : 1), f16 ? "FLOP" : "OP", cudaGetErrorString(e));
      }
  return 0;
}
