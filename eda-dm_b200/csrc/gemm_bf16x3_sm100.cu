// Calibration-path GEMM for sm_100a (SURVEY.md section 8 row N1 / north_star (b); replaces the fp32 library call behind
// `self.fwd_func(input, weight, bias)` at qdiff/quant_layer.py:434 -- and its autograd dgrad / wgrad -- while the fake-quant
// operands carry gradients, i.e. inside block_reconstruction).
//
//   D[m][n] (fp32) (+)= sum_k A[m][k] * B[n][k]  (+ bias[n])
//
// on the bf16 tensor cores with fp32-class accuracy: every fp32 operand x is split into x_hi = bf16(x), x_lo = bf16(x - x_hi)
// (edadm_split_bf16, one pass, optionally also the transposed copies the backward GEMMs need) and the product is evaluated as
// A_hi.B_hi + A_hi.B_lo + A_lo.B_hi with fp32 accumulation in TMEM -- three tcgen05.mma.kind::f16 per K step, relative error
// ~2^-16 per product against 2^-11 for the TF32 path cuDNN / cuBLAS take.  The fake-quantized operands of the reconstruction
// loop are not representable in one bf16 (QDrop passes raw fp32 activations through, soft AdaRound weights have fractional codes),
// hence the split instead of plain bf16.
//
// One kernel serves forward, dgrad and wgrad -- they differ in which (pre-split, K-contiguous) copies are handed in:
//   forward  Y  = X  W^T      A = X  [M][K],   B = W  [N][K]
//   dgrad    dX = dY W        A = dY [M][N],   B = W^T [K][N]
//   wgrad    dW = dY^T X      A = dY^T [N][M], B = X^T [K][M]   (few output tiles, long reduction: split-K, partial tiles are
//                                                               added into the zeroed output by TMA reduce-add)
// Structure: persistent CTAs, warp 0 TMA producer (four 128-byte-swizzled tiles per stage), warp 1 MMA issuer, warps 2..9
// epilogue through shared memory + TMA store / reduce (same as qgemm2_sm100.cu).
#include "tc05.cuh"
#include <cuda_bf16.h>
#include <cstdlib>
#include <algorithm>

namespace edadm {
namespace g3 {

constexpr int BM = 128;
constexpr int BK = 64;                        // bf16 elements per K step (128 bytes)
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + EPI_WARPS * 32;
constexpr int CHUNK = 32;
constexpr int OUT_BUF_BYTES = BM * CHUNK * 4;
constexpr int MAX_BN = 128;                   // 64 KB per stage (A_hi, A_lo, B_hi, B_lo) -> three stages; wider tiles would leave two
constexpr int MAX_STAGES = 6;
constexpr int ACC_STAGES = 2;
constexpr int SMEM_LIMIT = 227 * 1024;

struct Params {
  int M, N, K;
  int block_n, n_tiles, m_tiles, k_steps;     // k_steps per work unit
  int splits;                                 // split-K factor (units = m_tiles * n_tiles * splits)
  int stages, b_tile_bytes;
  int reduce_add;                             // 1: partial tiles are ADDED into the output (TMA reduce), bias only from split 0
  int group_m_tiles;                          // grouped (batched) GEMM: A rows come in groups of group_m_tiles tiles, group g multiplies
  int group_b_rows;                           //   B rows [g * group_b_rows, (g + 1) * group_b_rows); 0 = one shared B
  // convolution mode (conv = 1): A is the halo-padded NHWC tensor [B][Hp][Wp][Cp] read through rank-4 maps, one filter tap x 64
  // channels per K step (implicit GEMM, as qgemm2_sm100.cu); B is [N][taps * C]; the output is NCHW through a rank-3 map
  int conv, taps, S, c_chunks, C, Wo, HoWo, out_hw, pxb, px_shift;
  // convolution wgrad (conv = 2): dW[tap][n][c] = sum over pixels dY[b][n][pixel] * X[b][c][pixel shifted by the tap] -- both operands
  // are pixel-contiguous in NCHW: A = dY through a rank-3 map {HW, N, B}, B = X through a rank-4 map {W, H, C, B} whose box start is
  // moved by the tap offset (out-of-range pixels are zero-filled by TMA = the convolution's zero padding); one K step = 64 pixels
  int wg_cpi, wg_W, wg_pad, wg_B;               // K chunks per image, image width, padding, batch
  const float* bias;
};

struct __align__(8) Barriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tmem_full[ACC_STAGES];
  uint64_t tmem_empty[ACC_STAGES];
  uint32_t tmem_base;
};

// instruction descriptor kind::f16: D fp32, A / B bf16, both K-major
__device__ __forceinline__ uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                   const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
                   const __grid_constant__ CUtensorMap map_out, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int a_tile = BM * BK * 2;                               // 16 KB
  const int stage_bytes = 2 * a_tile + 2 * p.b_tile_bytes;      // A_hi, A_lo, B_hi, B_lo
  uint8_t* out_buf = smem + p.stages * stage_bytes;
  float* epi_bias = reinterpret_cast<float*>(out_buf + 2 * OUT_BUF_BYTES);
  Barriers* bars = reinterpret_cast<Barriers*>(epi_bias + MAX_BN);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int units = p.m_tiles * p.n_tiles * p.splits * (p.conv == 2 ? p.taps : 1);
  const int stages = p.stages;
  const uint32_t stage_tx = (uint32_t)(2 * a_tile + 2 * p.block_n * BK * 2);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_ah); tma_prefetch_desc(&map_al); tma_prefetch_desc(&map_bh); tma_prefetch_desc(&map_bl); tma_prefetch_desc(&map_out);
    for (int i = 0; i < stages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&bars->tmem_full[i], 1); mbar_init(&bars->tmem_empty[i], EPI_WARPS); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  // unit -> (split, m tile, n tile); n fastest so that concurrently running CTAs share the A rows through L2
  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
      const int n_blk = unit % p.n_tiles, rest = unit / p.n_tiles;
      const int m_blk = rest % p.m_tiles, rest2 = rest / p.m_tiles;
      const int sp = rest2 % p.splits, tap_u = rest2 / p.splits;        // (tap_u: wgrad only)
      const int k0 = sp * p.k_steps;
      for (int ks = 0; ks < p.k_steps; ++ks) {
        mbar_wait(&bars->empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* s = smem + stage * stage_bytes;
          mbar_expect_tx(&bars->full[stage], stage_tx);
          int kc = (k0 + ks) * BK;
          if (p.conv == 2) {
            const int kg = k0 + ks, b = kg / p.wg_cpi, p0 = (kg - b * p.wg_cpi) * BK;
            const int kh = tap_u / p.S, kw = tap_u - kh * p.S;
            tma_load_3d(s, &map_ah, &bars->full[stage], p0, m_blk * BM, b);
            tma_load_3d(s + a_tile, &map_al, &bars->full[stage], p0, m_blk * BM, b);
            // X comes as S copies pre-shifted along W (a TMA box must start on a 16-byte boundary of the innermost dimension, which a
            // +-1 pixel offset is not): copy kw sits at batch index kw * B + b; the row shift is a plain (possibly out-of-range) coordinate
            // and the row shift is a whole number of rows on the flattened pixel axis (W % 8 == 0 keeps it 16-byte aligned); pixels before
            // the first / after the last row are out of range = zero-filled = the vertical zero padding
            const int px = p0 + (kh - p.wg_pad) * p.wg_W;
            tma_load_3d(s + 2 * a_tile, &map_bh, &bars->full[stage], px, n_blk * p.block_n, kw * p.wg_B + b);
            tma_load_3d(s + 2 * a_tile + p.b_tile_bytes, &map_bl, &bars->full[stage], px, n_blk * p.block_n, kw * p.wg_B + b);
          } else {
          if (p.conv) {
            const int tap = ks / p.c_chunks, cc = ks - tap * p.c_chunks;
            const int kh = tap / p.S, kw = tap - kh * p.S;
            const int m0 = m_blk * BM, b0 = m0 / p.HoWo, rem = m0 - b0 * p.HoWo, oh0 = rem / p.Wo, ow0 = rem - oh0 * p.Wo;
            tma_load_4d(s, &map_ah, &bars->full[stage], cc * BK, ow0 + kw, oh0 + kh, b0);
            tma_load_4d(s + a_tile, &map_al, &bars->full[stage], cc * BK, ow0 + kw, oh0 + kh, b0);
            kc = tap * p.C + cc * BK;           // weights are [N][tap][channel]
          } else {
            tma_load_2d(s, &map_ah, &bars->full[stage], kc, m_blk * BM);
            tma_load_2d(s + a_tile, &map_al, &bars->full[stage], kc, m_blk * BM);
          }
          const int brow = n_blk * p.block_n + (p.group_m_tiles ? (m_blk / p.group_m_tiles) * p.group_b_rows : 0);
          tma_load_2d(s + 2 * a_tile, &map_bh, &bars->full[stage], kc, brow);
          tma_load_2d(s + 2 * a_tile + p.b_tile_bytes, &map_bl, &bars->full[stage], kc, brow);
          }
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(BM, p.block_n);
    const uint32_t base = smem_u32(smem);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
      mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)acc * MAX_BN;
      for (int ks = 0; ks < p.k_steps; ++ks) {
        mbar_wait(&bars->full[stage], phase);
        tc_fence_after();
        const uint32_t s = base + stage * stage_bytes;
        const uint64_t ah = make_smem_desc(s), al = make_smem_desc(s + a_tile);
        const uint64_t bh = make_smem_desc(s + 2 * a_tile), bl = make_smem_desc(s + 2 * a_tile + p.b_tile_bytes);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {      // 16 bf16 = 32 bytes per MMA: descriptor start advances by 2 (x16 B)
            umma_bf16(tmem_d, ah + 2 * kk, bh + 2 * kk, idesc, (ks | kk) ? 1u : 0u);
            umma_bf16(tmem_d, ah + 2 * kk, bl + 2 * kk, idesc, 1u);
            umma_bf16(tmem_d, al + 2 * kk, bh + 2 * kk, idesc, 1u);
          }
          umma_commit(&bars->empty[stage]);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&bars->tmem_full[acc]);
      __syncwarp();
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;
    const int et = threadIdx.x - 64;
    const bool issuer = et == 0;
    uint32_t st_off[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) st_off[k] = (uint32_t)r * 128u + ((uint32_t)((4 * half + k) ^ (r & 7)) << 4);
    // convolution mode: NCHW staging [image][column][pixel] (see qgemm2_sm100.cu)
    const uint32_t nchw_cstride = (uint32_t)p.pxb * 4u;
    const uint32_t nchw_off = p.conv == 1 ? ((uint32_t)((r >> p.px_shift) * CHUNK + 16 * half) * p.pxb + (r & (p.pxb - 1))) * 4u : 0u;
    const int n_chunks = p.block_n / CHUNK;
    int acc = 0; uint32_t acc_phase = 0; uint32_t gchunk = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
      const int n_blk = unit % p.n_tiles, rest = unit / p.n_tiles;
      const int m_blk = rest % p.m_tiles, rest2 = rest / p.m_tiles;
      const int sp = rest2 % p.splits, tap_u = rest2 / p.splits;
      const int n0 = n_blk * p.block_n;
      for (int j = et; j < p.block_n; j += EPI_WARPS * 32)
        epi_bias[j] = (p.bias && sp == 0 && n0 + j < p.N) ? __ldg(p.bias + n0 + j) : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * MAX_BN;
      mbar_wait(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();
      for (int ci = 0; ci < n_chunks; ++ci, ++gchunk) {
        const int c0 = ci * CHUNK + 16 * half;
        uint32_t a[16];
        tmem_ld16(taddr + c0, a);
        tmem_ld_wait();
        if (ci == n_chunks - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
        }
        uint8_t* ob = out_buf + (gchunk & 1u) * OUT_BUF_BYTES;
        if (issuer) bulk_wait_read<1>();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (p.conv == 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) *reinterpret_cast<float*>(ob + nchw_off + j * nchw_cstride) = __uint_as_float(a[j]) + epi_bias[c0 + j];
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<float4*>(ob + st_off[k]) = make_float4(__uint_as_float(a[4 * k + 0]) + epi_bias[c0 + 4 * k + 0],
                                                                     __uint_as_float(a[4 * k + 1]) + epi_bias[c0 + 4 * k + 1],
                                                                     __uint_as_float(a[4 * k + 2]) + epi_bias[c0 + 4 * k + 2],
                                                                     __uint_as_float(a[4 * k + 3]) + epi_bias[c0 + 4 * k + 3]);
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (issuer) {
          if (p.conv == 2) {
            if (p.reduce_add) tma_reduce_add_3d(&map_out, ob, n0 + ci * CHUNK, m_blk * BM, tap_u);
            else tma_store_3d(&map_out, ob, n0 + ci * CHUNK, m_blk * BM, tap_u);
          } else if (p.conv) {
            const int m0 = m_blk * BM, img0 = m0 / p.out_hw;
            tma_store_3d(&map_out, ob, m0 - img0 * p.out_hw, n0 + ci * CHUNK, img0);
          } else if (p.reduce_add) tma_reduce_add_2d(&map_out, ob, n0 + ci * CHUNK, m_blk * BM);
          else tma_store_2d(&map_out, ob, n0 + ci * CHUNK, m_blk * BM);
          bulk_commit();
        }
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    if (issuer) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// x fp32 [rows][cols] -> hi = bf16(x), lo = bf16(x - hi) as [rows][cols_p] (nullable) and transposed [cols][rows_p] (nullable);
// pitches in elements, multiples of 8 (16-byte TMA strides); padding columns are written as 0.
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ x, long long rows, long long cols, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                  long long cols_p, __nv_bfloat16* __restrict__ hi_t, __nv_bfloat16* __restrict__ lo_t, long long rows_p) {
  __shared__ float tile[32][33];
  // blockIdx.y = matrix of a batch: [batch][rows][cols] -> [batch][rows][cols_p] and / or [batch][cols][rows_p]
  x += (long long)blockIdx.y * rows * cols;
  if (hi) { hi += (long long)blockIdx.y * rows * cols_p; lo += (long long)blockIdx.y * rows * cols_p; }
  if (hi_t) { hi_t += (long long)blockIdx.y * cols * rows_p; lo_t += (long long)blockIdx.y * cols * rows_p; }
  const long long tiles_c = (cols_p + 31) / 32, tiles_r = (rows_p + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  for (long long t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
    const long long r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long r = r0 + ty + 8 * i, c = c0 + tx;
      const float v = (r < rows && c < cols) ? __ldg(x + r * cols + c) : 0.f;
      tile[ty + 8 * i][tx] = v;
      if (hi && r < rows && c < cols_p) {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[r * cols_p + c] = h;
        lo[r * cols_p + c] = __float2bfloat16_rn(v - __bfloat162float(h));
      }
    }
    __syncthreads();
    if (hi_t) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long c = c0 + ty + 8 * i, r = r0 + tx;      // output row = source column
        if (c < cols && r < rows_p) {
          const float v = tile[tx][ty + 8 * i];
          const __nv_bfloat16 h = __float2bfloat16_rn(v);
          hi_t[c * rows_p + r] = h;
          lo_t[c * rows_p + r] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace g3
}  // namespace edadm

using namespace edadm;

extern "C" int edadm_split_bf16_batched(const float* x, int64_t batch, int64_t rows, int64_t cols, void* hi, void* lo, int64_t cols_p,
                                        void* hi_t, void* lo_t, int64_t rows_p, void* stream);

extern "C" int edadm_split_shift_bf16(const float* x, void* hi, void* lo, int64_t rows, int W, int S, int pad, void* stream);

extern "C" int edadm_split_bf16(const float* x, int64_t rows, int64_t cols, void* hi, void* lo, int64_t cols_p, void* hi_t, void* lo_t,
                                int64_t rows_p, void* stream) {
  return edadm_split_bf16_batched(x, 1, rows, cols, hi, lo, cols_p, hi_t, lo_t, rows_p, stream);
}

extern "C" int edadm_split_bf16_batched(const float* x, int64_t batch, int64_t rows, int64_t cols, void* hi, void* lo, int64_t cols_p,
                                        void* hi_t, void* lo_t, int64_t rows_p, void* stream) {
  if (!x || batch < 1 || batch > 65535 || rows < 1 || cols < 1 || (!hi && !hi_t) || (hi && !lo) || (hi_t && !lo_t)) return fail(EDADM_ERR_ARG, "split_bf16: bad arguments");
  if ((hi && (cols_p < cols || (cols_p & 7))) || (hi_t && (rows_p < rows || (rows_p & 7)))) return fail(EDADM_ERR_ARG, "split_bf16: pitches must cover the data and be multiples of 8");
  if (hi && !hi_t && cols_p == cols && (cols & 7) == 0 &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0)
    return edadm_split_shift_bf16(x, hi, lo, batch * rows, (int)cols, 1, 0, stream);     // straight copy only: the 16-byte streaming form
  if (!hi) cols_p = cols;
  if (!hi_t) rows_p = rows;
  const long long tiles = ((cols_p + 31) / 32) * ((rows_p + 31) / 32);
  const int gx = (int)std::min<long long>(tiles, std::max<long long>(1, (long long)sm_count() * 16 / batch));
  dim3 grid(gx, (unsigned)batch);
  g3::split_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, cols_p,
                                                               (__nv_bfloat16*)hi_t, (__nv_bfloat16*)lo_t, rows_p);
  return check_launch("split_bf16");
}

// out[M][N] fp32 (row pitch N) = (or +=, with reduce_add) A.B^T (+ bias) from the split operands: a_hi / a_lo bf16 [M][Kp], b_hi / b_lo
// bf16 [N][Kp] (Kp: row pitch in elements, multiple of 8; K real reduction length; elements past K must be 0).  splits > 1 cuts the
// reduction into `splits` work units per output tile that are ADDED into `out` (zero it first).
extern "C" int edadm_gemm_bf16x3_grouped(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, int64_t groups, int64_t M,
                                         int N, int64_t K, int64_t Kp, const float* bias, float* out, int splits, void* stream);

extern "C" int edadm_gemm_bf16x3(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, int64_t M, int N, int64_t K,
                                 int64_t Kp, const float* bias, float* out, int splits, void* stream) {
  return edadm_gemm_bf16x3_grouped(a_hi, a_lo, b_hi, b_lo, 1, M, N, K, Kp, bias, out, splits, stream);
}

// groups > 1: batched product out[g][M][N] = a[g] . b[g]^T with a_* [groups*M][Kp], b_* [groups*N][Kp], out [groups*M][N]; M must be a
// multiple of 128 so that no tile straddles two groups (the attention matmuls of the reconstruction loop: M = tokens).
extern "C" int edadm_gemm_bf16x3_grouped(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, int64_t groups, int64_t M,
                                         int N, int64_t K, int64_t Kp, const float* bias, float* out, int splits, void* stream) {
  using namespace g3;
  if (groups < 1 || (groups > 1 && (M % BM))) return fail(EDADM_ERR_ARG, "gemm_bf16x3: grouped product needs M %% 128 == 0 (M=%lld)", (long long)M);
  const int64_t M_group = M;
  M = M * groups;
  if (!a_hi || !a_lo || !b_hi || !b_lo || !out) return fail(EDADM_ERR_ARG, "gemm_bf16x3: null pointer");
  if (M < 1 || N < 1 || K < 1 || Kp < K || (Kp & 7) || (N & 3) || M > 0x7fffffffLL || splits < 1 || splits > 64)
    return fail(EDADM_ERR_ARG, "gemm_bf16x3: bad sizes M=%lld N=%d K=%lld Kp=%lld splits=%d", (long long)M, N, (long long)K, (long long)Kp, splits);
  if ((((uintptr_t)a_hi | (uintptr_t)a_lo | (uintptr_t)b_hi | (uintptr_t)b_lo | (uintptr_t)out) & 15)) return fail(EDADM_ERR_ARG, "gemm_bf16x3: operands must be 16-byte aligned");
  Params p;
  memset(&p, 0, sizeof(p));
  p.M = (int)M; p.N = N; p.K = (int)K;
  p.m_tiles = (int)((M + BM - 1) / BM);
  // N tile: multiple of 32, fewest waves then widest (stage size grows with it: 2 x 16 KB + 2 x block_n x 128 B)
  int best = 0; long long best_cost = 0;
  for (int bn = 32; bn <= MAX_BN; bn += 32) {
    const long long tiles = (long long)p.m_tiles * ((N + bn - 1) / bn) * splits;
    const long long waves = (tiles + sm_count() - 1) / sm_count();
    const long long cost = waves * (bn + 128);
    if (!best || cost <= best_cost) { best = bn; best_cost = cost; }
    if (bn >= N) break;
  }
  p.block_n = best;
  p.n_tiles = (N + best - 1) / best;
  const int k_steps_total = (int)((K + BK - 1) / BK);
  p.splits = std::min(splits, k_steps_total);
  p.k_steps = (k_steps_total + p.splits - 1) / p.splits;
  p.splits = (k_steps_total + p.k_steps - 1) / p.k_steps;       // no empty split (tiles past K would read zero-filled data anyway)
  p.reduce_add = p.splits > 1 ? 1 : 0;
  p.bias = bias;
  p.group_m_tiles = groups > 1 ? (int)(M_group / BM) : 0;
  p.group_b_rows = groups > 1 ? N : 0;
  p.b_tile_bytes = (best * BK * 2 + 1023) & ~1023;
  const int stage_bytes = 2 * BM * BK * 2 + 2 * p.b_tile_bytes;
  const int fixed = 1024 + 2 * OUT_BUF_BYTES + MAX_BN * 4 + 1024;
  p.stages = std::min(MAX_STAGES, (SMEM_LIMIT - fixed) / stage_bytes);
  if (p.stages < 2) return fail(EDADM_ERR_UNSUPPORTED, "gemm_bf16x3: tile does not fit");
  const int smem_bytes = fixed + p.stages * stage_bytes;

  CUtensorMap m_ah, m_al, m_bh, m_bl, m_out;
  auto enc = [&](CUtensorMap* m, const void* ptr, long long rows, int box_rows, const char* what) {
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    return encode_map(m, ptr, 2, dims, strides, box, what, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  int rc;
  if ((rc = enc(&m_ah, a_hi, M, BM, "A hi")) || (rc = enc(&m_al, a_lo, M, BM, "A lo")) || (rc = enc(&m_bh, b_hi, (long long)N * groups, best, "B hi")) ||
      (rc = enc(&m_bl, b_lo, (long long)N * groups, best, "B lo")))
    return rc;
  {
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)N * 4};
    cuuint32_t box[2] = {(cuuint32_t)CHUNK, (cuuint32_t)BM};
    if ((rc = encode_map(&m_out, out, 2, dims, strides, box, "output", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  static int attr_dev[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_dev[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return fail(EDADM_ERR_CUDA, "gemm_bf16x3: cannot opt in to %d B shared memory: %s", SMEM_LIMIT, cudaGetErrorString(e));
    attr_dev[dev] = 1;
  }
  const int units = p.m_tiles * p.n_tiles * p.splits;
  const int grid = std::min(units, sm_count());
  gemm_bf16x3_kernel<<<grid, THREADS, smem_bytes, (cudaStream_t)stream>>>(m_ah, m_al, m_bh, m_bl, m_out, p);
  return check_launch("gemm_bf16x3");
}

namespace edadm {
namespace g3 {
// x fp32 [B][C][H][W] -> hi / lo bf16 [B][H+2p][W+2p][Cp] (halo and padded channels = 0): NHWC operands of the implicit-GEMM
// convolution.  One block per (image, row): [C][W] -> [W][C] through shared memory.
__global__ void __launch_bounds__(256)
split_nhwc_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int C, int H, int W, int Cp,
                       int pad) {
  __shared__ float tile[32][33];
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const int b = blockIdx.y, hp = blockIdx.x;                    // hp over padded rows
  const int h = hp - pad;
  __nv_bfloat16* hrow = hi + ((size_t)b * Hp + hp) * Wp * Cp;
  __nv_bfloat16* lrow = lo + ((size_t)b * Hp + hp) * Wp * Cp;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  if (h < 0 || h >= H) {                                        // halo row
    for (int i = threadIdx.x; i < Wp * Cp; i += blockDim.x) { hrow[i] = zero; lrow[i] = zero; }
    return;
  }
  for (int i = threadIdx.x; i < pad * Cp; i += blockDim.x) {    // halo columns
    hrow[i] = zero; lrow[i] = zero;
    hrow[(size_t)(W + pad) * Cp + i] = zero; lrow[(size_t)(W + pad) * Cp + i] = zero;
  }
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* src = x + ((size_t)b * C * H + h) * W;           // + c * H * W + w
  for (int c0 = 0; c0 < Cp; c0 += 32)
    for (int w0 = 0; w0 < W; w0 += 32) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, w = w0 + tx;
        tile[ty + 8 * i][tx] = (c < C && w < W) ? __ldg(src + (size_t)c * H * W + w) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int w = w0 + ty + 8 * i, c = c0 + tx;
        if (w < W && c < Cp) {
          const float v = tile[tx][ty + 8 * i];
          const __nv_bfloat16 hv = __float2bfloat16_rn(v);
          const size_t o = (size_t)(w + pad) * Cp + c;
          hrow[o] = hv;
          lrow[o] = __float2bfloat16_rn(v - __bfloat162float(hv));
        }
      }
      __syncthreads();
    }
}
}  // namespace g3
}  // namespace edadm

namespace edadm {
namespace g3 {
// filter w fp32 [N][C][RS] -> forward operand f_* bf16 [N][RS*C (pitch fp)] (tap-major, channel-minor) and / or dgrad operand
// d_* bf16 [C][RS*N (pitch dp)] with the taps reversed (d[c][t][n] = w[n][c][RS-1-t]).  Blocks [0, N) write the forward rows,
// blocks [N, N + C) the dgrad rows.
__global__ void __launch_bounds__(256)
split_filter_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ f_hi, __nv_bfloat16* __restrict__ f_lo,
                         __nv_bfloat16* __restrict__ d_hi, __nv_bfloat16* __restrict__ d_lo, int N, int C, int RS, int fp, int dp,
                         int n_fwd_blocks) {
  const int b = blockIdx.x;
  if (b < n_fwd_blocks) {
    const int n = b;
    const float* src = w + (size_t)n * C * RS;
    for (int i = threadIdx.x; i < fp; i += blockDim.x) {       // i = t * C + c
      float v = 0.f;
      if (i < RS * C) { const int t = i / C, c = i - t * C; v = __ldg(src + (size_t)c * RS + t); }
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      f_hi[(size_t)n * fp + i] = h;
      f_lo[(size_t)n * fp + i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  } else {
    const int c = b - n_fwd_blocks;
    for (int i = threadIdx.x; i < dp; i += blockDim.x) {       // i = t * N + n
      float v = 0.f;
      if (i < RS * N) { const int t = i / N, n = i - t * N; v = __ldg(w + ((size_t)n * C + c) * RS + (RS - 1 - t)); }
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      d_hi[(size_t)c * dp + i] = h;
      d_lo[(size_t)c * dp + i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}
}  // namespace g3
}  // namespace edadm

extern "C" int edadm_split_filter_bf16(const float* w, int N, int C, int R, int S, void* f_hi, void* f_lo, int64_t f_pitch, void* d_hi,
                                       void* d_lo, int64_t d_pitch, void* stream) {
  const int RS = R * S;
  const bool fwd = f_hi && f_lo, dg = d_hi && d_lo;
  if (!w || N < 1 || C < 1 || RS < 1 || (!fwd && !dg) || (fwd && (f_pitch < (int64_t)RS * C || (f_pitch & 7))) ||
      (dg && (d_pitch < (int64_t)RS * N || (d_pitch & 7))))
    return fail(EDADM_ERR_ARG, "split_filter_bf16: bad arguments");
  const int nf = fwd ? N : 0;
  g3::split_filter_bf16_kernel<<<nf + (dg ? C : 0), 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)f_hi, (__nv_bfloat16*)f_lo,
                                                                                   (__nv_bfloat16*)d_hi, (__nv_bfloat16*)d_lo, N, C, RS,
                                                                                   (int)f_pitch, (int)d_pitch, nf);
  return check_launch("split_filter_bf16");
}

extern "C" int edadm_split_nhwc_bf16(const float* x, void* hi, void* lo, int B, int C, int H, int W, int Cp, int pad, void* stream) {
  if (!x || !hi || !lo || B < 1 || B > 65535 || C < 1 || H < 1 || W < 1 || Cp < C || (Cp & 7) || pad < 0)
    return fail(EDADM_ERR_ARG, "split_nhwc_bf16: bad arguments");
  dim3 grid((unsigned)(H + 2 * pad), (unsigned)B);
  g3::split_nhwc_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, C, H, W, Cp, pad);
  return check_launch("split_nhwc_bf16");
}

// Stride-1 convolution forward on the bf16 x 3 GEMM: a_hi / a_lo NHWC bf16 [B][Hp][Wp][Cp] (halo included, edadm_split_nhwc_bf16),
// w_hi / w_lo bf16 [N][R*S*C] (tap-major, channel-minor; row pitch Kp elements), out fp32 NCHW [B][N][Ho][Wo].
extern "C" int edadm_conv_bf16x3(const void* a_hi, const void* a_lo, int B, int Hp, int Wp, int Cp, const void* w_hi, const void* w_lo, int N,
                                 int R, int S, int C, int64_t Kp, const float* bias, float* out, void* stream) {
  using namespace g3;
  if (!a_hi || !a_lo || !w_hi || !w_lo || !out) return fail(EDADM_ERR_ARG, "conv_bf16x3: null pointer");
  const int Ho = Hp - R + 1, Wo = Wp - S + 1;
  if (B < 1 || Ho < 1 || Wo < 1 || N < 1 || (Cp & 7) || Cp < C || (C & 7) || Kp < (int64_t)R * S * C || (Kp & 7))
    return fail(EDADM_ERR_ARG, "conv_bf16x3: bad geometry");
  const long long M = (long long)B * Ho * Wo;
  const int out_hw = Ho * Wo;
  int box_w, box_h, box_b;
  if (Wo >= BM) { if (Wo % BM) return fail(EDADM_ERR_UNSUPPORTED, "conv_bf16x3: output width"); box_w = BM; box_h = 1; box_b = 1; }
  else {
    if (BM % Wo) return fail(EDADM_ERR_UNSUPPORTED, "conv_bf16x3: output width %d does not divide 128", Wo);
    box_w = Wo;
    const int rows = BM / Wo;
    if (rows <= Ho) { if (Ho % rows) return fail(EDADM_ERR_UNSUPPORTED, "conv_bf16x3: output height"); box_h = rows; box_b = 1; }
    else { if (rows % Ho) return fail(EDADM_ERR_UNSUPPORTED, "conv_bf16x3: output height"); box_h = Ho; box_b = rows / Ho; }
  }
  if (out_hw % 4) return fail(EDADM_ERR_UNSUPPORTED, "conv_bf16x3: output pitch");
  const int pxb = out_hw >= BM ? BM : out_hw;
  if ((out_hw >= BM ? out_hw % BM : BM % out_hw) || (pxb & (pxb - 1)) || M % BM) return fail(EDADM_ERR_UNSUPPORTED, "conv_bf16x3: tile geometry");
  Params p;
  memset(&p, 0, sizeof(p));
  p.M = (int)M; p.N = N; p.K = R * S * C;
  p.m_tiles = (int)(M / BM);
  int best = 0; long long best_cost = 0;
  for (int bn = 32; bn <= MAX_BN; bn += 32) {
    const long long tiles = (long long)p.m_tiles * ((N + bn - 1) / bn);
    const long long waves = (tiles + sm_count() - 1) / sm_count();
    const long long cost = waves * (bn + 128);
    if (!best || cost <= best_cost) { best = bn; best_cost = cost; }
    if (bn >= N) break;
  }
  p.block_n = best; p.n_tiles = (N + best - 1) / best;
  p.conv = 1; p.taps = R * S; p.S = S; p.C = C; p.c_chunks = (C + BK - 1) / BK;
  p.k_steps = p.taps * p.c_chunks; p.splits = 1;
  p.Wo = Wo; p.HoWo = Ho * Wo; p.out_hw = out_hw; p.pxb = pxb;
  while ((1 << p.px_shift) < pxb) ++p.px_shift;
  p.bias = bias;
  p.b_tile_bytes = (best * BK * 2 + 1023) & ~1023;
  const int stage_bytes = 2 * BM * BK * 2 + 2 * p.b_tile_bytes;
  const int fixed = 1024 + 2 * OUT_BUF_BYTES + MAX_BN * 4 + 1024;
  p.stages = std::min(MAX_STAGES, (SMEM_LIMIT - fixed) / stage_bytes);
  const int smem_bytes = fixed + p.stages * stage_bytes;
  CUtensorMap m_ah, m_al, m_bh, m_bl, m_out;
  int rc;
  for (int i = 0; i < 2; ++i) {
    cuuint64_t dims[4] = {(cuuint64_t)Cp, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)Cp * 2, (cuuint64_t)Wp * Cp * 2, (cuuint64_t)Hp * Wp * Cp * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_b};
    if ((rc = encode_map(i ? &m_al : &m_ah, i ? a_lo : a_hi, 4, dims, strides, box, "conv activations", CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  for (int i = 0; i < 2; ++i) {
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)Kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)best};
    if ((rc = encode_map(i ? &m_bl : &m_bh, i ? w_lo : w_hi, 2, dims, strides, box, "conv weights", CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)out_hw, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)out_hw * 4, (cuuint64_t)N * out_hw * 4};
    cuuint32_t box[3] = {(cuuint32_t)pxb, (cuuint32_t)CHUNK, (cuuint32_t)(BM / pxb)};
    if ((rc = encode_map(&m_out, out, 3, dims, strides, box, "conv output", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  }
  static int attr_dev[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_dev[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return fail(EDADM_ERR_CUDA, "conv_bf16x3: cannot opt in to %d B shared memory: %s", SMEM_LIMIT, cudaGetErrorString(e));
    attr_dev[dev] = 1;
  }
  const int units = p.m_tiles * p.n_tiles;
  gemm_bf16x3_kernel<<<std::min(units, sm_count()), THREADS, smem_bytes, (cudaStream_t)stream>>>(m_ah, m_al, m_bh, m_bl, m_out, p);
  return check_launch("conv_bf16x3");
}

namespace edadm {
namespace g3 {
// x fp32 [rows][W] -> hi / lo bf16 [S][rows][W], copy s shifted along the row: out[s][r][w] = x[r][w + s - pad] (0 outside).
// A thread owns 8 consecutive pixels of a row: its window x[w0 - pad .. w0 + 7 + S - 1 - pad] is loaded once (two 16-byte loads + the
// edge values), every shifted copy leaves as one 16-byte store of hi and one of lo.  Generic (scalar) form for other shapes.
template <int S_MAX>
__global__ void __launch_bounds__(256)
split_shift_bf16_vec_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long rows,
                            int W, int S, int pad) {
  const long long groups = rows * (W >> 3), total = rows * W;
  for (long long gidx = (long long)blockIdx.x * blockDim.x + threadIdx.x; gidx < groups; gidx += (long long)gridDim.x * blockDim.x) {
    const long long r = gidx / (W >> 3);
    const int w0 = (int)(gidx - r * (W >> 3)) << 3;
    const float* xr = x + r * W;
    float win[8 + S_MAX - 1];
    const float4 a = __ldg(reinterpret_cast<const float4*>(xr + w0)), b = __ldg(reinterpret_cast<const float4*>(xr + w0 + 4));
    const float core[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 8 + S_MAX - 1; ++k) {
      const int w = w0 + k - pad;
      win[k] = (k >= pad && k < pad + 8) ? core[k - pad] : ((w >= 0 && w < W && k < 8 + S - 1) ? __ldg(xr + w) : 0.f);
    }
#pragma unroll
    for (int sft = 0; sft < S_MAX; ++sft) {
      if (sft < S) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float v0 = win[2 * k + sft], v1 = win[2 * k + 1 + sft];
          const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
          const __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
          h[k] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          l[k] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        const long long o = (long long)sft * total + r * W + w0;
        *reinterpret_cast<uint4*>(hi + o) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(lo + o) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
split_shift_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long rows, int W,
                        int S, int pad) {
  const long long total = rows * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / W;
    const int w = (int)(i - r * W);
    for (int sft = 0; sft < S; ++sft) {
      const int ws = w + sft - pad;
      const float v = (ws >= 0 && ws < W) ? __ldg(x + r * W + ws) : 0.f;
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      hi[(long long)sft * total + i] = h;
      lo[(long long)sft * total + i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}
}  // namespace g3
}  // namespace edadm

extern "C" int edadm_split_shift_bf16(const float* x, void* hi, void* lo, int64_t rows, int W, int S, int pad, void* stream) {
  if (!x || !hi || !lo || rows < 1 || W < 1 || S < 1 || pad < 0) return fail(EDADM_ERR_ARG, "split_shift_bf16: bad arguments");
  const long long total = rows * W;
  const bool vec = (W % 8) == 0 && S <= 3 && pad <= 1 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0;
  if (vec) {
    const int blocks = (int)std::min<long long>((total / 8 + 255) / 256, (long long)sm_count() * 16);
    g3::split_shift_bf16_vec_kernel<3><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, rows, W, S, pad);
    return check_launch("split_shift_bf16");
  }
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 16);
  g3::split_shift_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, rows, W, S, pad);
  return check_launch("split_shift_bf16");
}

// Convolution weight gradient on the same kernel: dW[tap][n][c] = sum_{b, pixel} dY[b][n][pixel] * X[b][c][pixel + tap offset].
// dy_* bf16 [B][N][H*W] (hi / lo of the NCHW tensor as it is, edadm_split_bf16 over [B*N][H*W] rows); x_* bf16 [S][B][C][H][W]: S copies
// of the input pre-shifted along W by kw - pad with zero fill (edadm_split_shift_bf16),
// stride 1, padding `pad`, filter R x S (output size == input size: 2 * pad == R - 1).  out fp32 [R*S][N][C]; splits > 1 adds
// partial sums into `out` (zeroed by the caller) by TMA reduce.
extern "C" int edadm_conv_wgrad_bf16x3(const void* dy_hi, const void* dy_lo, const void* x_hi, const void* x_lo, int B, int N, int C, int H,
                                       int W, int R, int S, int pad, float* out, int splits, void* stream) {
  using namespace g3;
  if (!dy_hi || !dy_lo || !x_hi || !x_lo || !out) return fail(EDADM_ERR_ARG, "conv_wgrad_bf16x3: null pointer");
  const int HW = H * W;
  if (B < 1 || N < 1 || C < 1 || R < 1 || S < 1 || 2 * pad != R - 1 || R != S || (HW % BK) || (W % 8) || (C % 4))
    return fail(EDADM_ERR_UNSUPPORTED, "conv_wgrad_bf16x3: geometry B=%d N=%d C=%d H=%d W=%d R=%d pad=%d", B, N, C, H, W, R, pad);
  const int cpi = HW / BK;
  const int k_total = B * cpi;
  if (splits < 1 || (k_total % splits)) return fail(EDADM_ERR_ARG, "conv_wgrad_bf16x3: splits must divide %d", k_total);
  Params p;
  memset(&p, 0, sizeof(p));
  p.M = N; p.N = C; p.K = B * HW;
  p.m_tiles = (N + BM - 1) / BM;
  p.block_n = C >= 128 ? 128 : ((C + 31) / 32) * 32;
  if (C > 128 && C % 128 && C % 96 == 0) p.block_n = 96;
  if (const char* e = getenv("EDADM_WGRAD_BN")) { const int v = atoi(e); if (v >= 32 && v <= 128 && v % 32 == 0) p.block_n = v; }   // experiments
  p.n_tiles = (C + p.block_n - 1) / p.block_n;
  p.splits = splits; p.k_steps = k_total / splits; p.reduce_add = splits > 1;
  p.conv = 2; p.taps = R * S; p.S = S; p.wg_cpi = cpi; p.wg_W = W; p.wg_pad = pad; p.wg_B = B;
  p.b_tile_bytes = (p.block_n * BK * 2 + 1023) & ~1023;
  const int stage_bytes = 2 * BM * BK * 2 + 2 * p.b_tile_bytes;
  const int fixed = 1024 + 2 * OUT_BUF_BYTES + MAX_BN * 4 + 1024;
  p.stages = std::min(MAX_STAGES, (SMEM_LIMIT - fixed) / stage_bytes);
  const int smem_bytes = fixed + p.stages * stage_bytes;
  CUtensorMap m_ah, m_al, m_bh, m_bl, m_out;
  int rc;
  for (int i = 0; i < 2; ++i) {
    cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)HW * 2, (cuuint64_t)N * HW * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BM, 1u};
    if ((rc = encode_map(i ? &m_al : &m_ah, i ? dy_lo : dy_hi, 3, dims, strides, box, "wgrad dY", CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  for (int i = 0; i < 2; ++i) {
    cuuint64_t dims[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)B * S};
    cuuint64_t strides[2] = {(cuuint64_t)HW * 2, (cuuint64_t)C * HW * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)p.block_n, 1u};
    if ((rc = encode_map(i ? &m_bl : &m_bh, i ? x_lo : x_hi, 3, dims, strides, box, "wgrad X", CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)(R * S)};
    cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)N * C * 4};
    cuuint32_t box[3] = {(cuuint32_t)CHUNK, (cuuint32_t)BM, 1u};
    if ((rc = encode_map(&m_out, out, 3, dims, strides, box, "wgrad output", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  static int attr_dev[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_dev[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return fail(EDADM_ERR_CUDA, "conv_wgrad_bf16x3: cannot opt in to %d B shared memory: %s", SMEM_LIMIT, cudaGetErrorString(e));
    attr_dev[dev] = 1;
  }
  const int units = p.m_tiles * p.n_tiles * p.splits * p.taps;
  gemm_bf16x3_kernel<<<std::min(units, sm_count()), THREADS, smem_bytes, (cudaStream_t)stream>>>(m_ah, m_al, m_bh, m_bl, m_out, p);
  return check_launch("conv_wgrad_bf16x3");
}

