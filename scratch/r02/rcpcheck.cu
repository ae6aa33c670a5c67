// brute force: the branch-free reciprocal of csrc/pack.cu (SFU estimate + one Newton step) equals __frcp_rn for EVERY float in
// [1, 2^126] (all 126 binades x 2^23 mantissas), and the quotient built on it equals IEEE v / d on a sweep of numerators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -o rcpcheck rcpcheck.cu && ./rcpcheck
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ float rcp_fast(float d) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return fmaf(r, fmaf(-d, r, 1.0f), r);
}
__global__ void check_rcp(unsigned long long* bad, unsigned long long* first) {
  const uint32_t lo = 0x3F800000u, hi = 0x7E800000u;            // 1.0 .. 2^126
  for (uint64_t i = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= hi; i += (uint64_t)gridDim.x * blockDim.x) {
    const float d = __uint_as_float((uint32_t)i);
    if (rcp_fast(d) != __frcp_rn(d)) { if (atomicAdd(bad, 1ull) == 0) *first = i; }
  }
}
__global__ void check_div(unsigned long long* bad, uint32_t seed) {
  // numerators: 4096 pseudo-random floats in [-90, 20] per denominator sample; denominators 1 + expf(-v') style values
  uint32_t s = seed ^ (blockIdx.x * 9781u + threadIdx.x * 6271u);
  for (int it = 0; it < 4096; ++it) {
    s = s * 1664525u + 1013904223u;
    const float v = -90.0f + 110.0f * (float)(s >> 8) * (1.0f / 16777216.0f);
    s = s * 1664525u + 1013904223u;
    const float w = -90.0f + 110.0f * (float)(s >> 8) * (1.0f / 16777216.0f);
    const float d = fminf(1.0f + expf(-w), 8.507059e37f);
    const float r = rcp_fast(d);
    const float q0 = v * r;
    const float q = fmaf(fmaf(-d, q0, v), r, q0);
    const float ref = __fdiv_rn(v, d);
    if (q != ref && !(fabsf(ref) < 1.2e-38f)) atomicAdd(bad, 1ull);       // denormal quotients excluded (their code is the zero-point)
  }
}
int main() {
  unsigned long long *bad, *first, h[2] = {0, 0};
  cudaMalloc(&bad, 8); cudaMalloc(&first, 8);
  cudaMemset(bad, 0, 8); cudaMemset(first, 0, 8);
  check_rcp<<<148 * 8, 256>>>(bad, first);
  cudaMemcpy(&h[0], bad, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&h[1], first, 8, cudaMemcpyDeviceToHost);
  printf("reciprocal: %llu mismatches over [1, 2^126] (first 0x%llx)\n", h[0], h[1]);
  cudaMemset(bad, 0, 8);
  check_div<<<148 * 64, 256>>>(bad, 12345u);
  cudaMemcpy(&h[0], bad, 8, cudaMemcpyDeviceToHost);
  printf("quotient:   %llu mismatches over %llu (v, d) pairs\n", h[0], 148ull * 64 * 256 * 4096);
  return 0;
}
