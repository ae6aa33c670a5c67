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
__global__ void __launch_bounds__(128, 1) peak(int n, int iters, int distinct, int f16, long long* cycles, int commit_every, int tmem_reader) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint64_t bars2[8];
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 4 * 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars2[i], 1); mbar_init(&bar, 1); fence_barrier_init(); fence_proxy_async(); }
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
      if (commit_every && (it % commit_every) == commit_every - 1) umma_commit(&bars2[it & 7]);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  if (tmem_reader && threadIdx.x >= 32) {
    uint32_t r[16]; uint32_t acc = 0;
    const uint32_t lane_addr = tbase + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + 256;
    for (int it = 0; it < iters * tmem_reader / 4; ++it) {
      tmem_ld16(lane_addr + (it & 7) * 16, r); tmem_ld_wait();
      acc += r[0] + r[15];
    }
    if (acc == 0x12345678u) cycles[147] = acc;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512)); }
}
extern "C" int run_peak(int n, int iters, int f16, int grid, long long* cyc, void* stream, int commit_every, int tmem_reader) {
  cudaFuncSetAttribute(peak, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  peak<<<grid, 128, 200 * 1024, (cudaStream_t)stream>>>(n, iters, 1, f16, cyc, commit_every, tmem_reader);
  return (int)cudaGetLastError();
}

// ---- cta_group::2: one MMA spans a CTA pair (M = 256: 128 rows per CTA, each CTA holds its A tile and HALF of the B tile) ----
__device__ __forceinline__ void umma2_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
peak2(int n, int iters, int f16, long long* cycles, int commit_every) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint64_t bars2[8];
  __shared__ uint32_t tbase;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 4 * 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars2[i], 1); mbar_init(&bar, 1); fence_barrier_init(); fence_proxy_async(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before(); cluster_sync_all(); tc_fence_after();
  if (rank == 0 && threadIdx.x == 0) {
    // M = 256 (field M>>4 = 16)
    const uint32_t idesc = f16 ? ((1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24))
                               : ((2u << 4) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int st = it & 3;
      const uint64_t ad = make_smem_desc(smem_u32(smem + st * 49152));
      const uint64_t bd = make_smem_desc(smem_u32(smem + st * 49152 + 16384));
      for (int k = 0; k < 4; ++k) {
        if (f16) umma2_f16(tbase, ad + 2 * k, bd + 2 * k, idesc, 1u);
        else umma2_i8(tbase, ad + 2 * k, bd + 2 * k, idesc, 1u);
      }
      if (commit_every && (it % commit_every) == commit_every - 1) umma2_commit_mc(&bars2[it & 7], 3);
    }
    umma2_commit_mc(&bar, 1);
    mbar_wait(&bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before(); cluster_sync_all();
  if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512)); }
}
extern "C" int run_peak2(int n, int iters, int f16, int grid, long long* cyc, void* stream, int commit_every) {
  cudaFuncSetAttribute(peak2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  peak2<<<grid, 128, 200 * 1024, (cudaStream_t)stream>>>(n, iters, f16, cyc, commit_every);
  return (int)cudaGetLastError();
}
