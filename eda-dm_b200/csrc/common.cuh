// Shared helpers for the edadm sm_100a kernels: error plumbing for the C ABI, launch sizing
// and small device utilities (warp/block reductions, Philox for QDrop masks).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define EDADM_OK 0
#define EDADM_ERR_ARG (-1)
#define EDADM_ERR_CUDA (-2)
#define EDADM_ERR_UNSUPPORTED (-3)

namespace edadm {

// thread-local last error text, returned by edadm_last_error()
char* last_error_buf();
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);
int sm_count();

constexpr int kThreads = 256;

// grid for an HBM-bound streaming kernel: whole waves of 148 SMs x 8 resident CTAs
inline int stream_grid(long long work_items_per_thread_units) {
  long long blocks = (work_items_per_thread_units + kThreads - 1) / kThreads;
  long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum, result valid in thread 0. blockDim.x must be a multiple of 32 (<=1024).
template <typename T>
__device__ __forceinline__ T block_sum(T v) {
  __shared__ T red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect red[] across repeated calls
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : T(0);
  if (warp == 0) v = warp_sum(v);
  return v;
}

// Philox4x32-10 (Salmon et al.), counter = element-quad index, key = (seed lo, seed hi).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// uniform in [0,1) with 24 bits, same convention as torch (x >> 8) * 2^-24
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

}  // namespace edadm
