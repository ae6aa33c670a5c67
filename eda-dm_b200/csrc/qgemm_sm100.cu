// W4A8 / W8A8 QuantModule GEMM for sm_100a (replaces qdiff/quant_layer.py:434
// F.conv2d / F.conv1d / F.linear on fake-quant operands).
//
//   D[m][n] = sum_k  Aq[m][k] * Wq[n][k]          (u8 x s8 -> s32, exact)
//   out     = dA * dW[n] * (D + cw[n]*rowsum[m] - zA * wsum_eff[n]) + bias[n]   (+ SiLU) (+ prev out)
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer   : A tile (128 pixels x 128 B of channels, one filter tap per K step) through a
//                                rank-4 tiled tensor map over the halo-padded NHWC code tensor -- the implicit
//                                GEMM needs no im2col buffer, out-of-range channels are zero-filled by TMA;
//                                B tile (BLOCK_N out-channels x 128 B) through a rank-3 map over [Np][taps][Cp].
//   warp 1      MMA issuer     : tcgen05.mma.cta_group::1.kind::i8, M=128, N=BLOCK_N (multiple of 16, <=256),
//                                K=32 per instruction, accumulators in TMEM (2 stages x 256 columns).
//   warps 2..9  epilogue       : tcgen05.ld 32x32b (two warps per lane quarter, alternate 16-column chunks, next chunk's
//                                load in flight) -> int32 zero-point fold, per-channel dequant, bias, SiLU ->
//                                coalesced NCHW stores (TMEM lane == output pixel == consecutive address) or float4 rows.
// Shared memory: up to 8 pipeline stages x (16 KB A + BLOCK_N x 128 B), 128B-swizzled K-major operands.
#include "tc05.cuh"
#include <cstdlib>

namespace edadm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 128;        // bytes == int8 elements per K step (one 128B swizzle atom)
constexpr int MAX_BLOCK_N = 256;
constexpr int MAX_STAGES = 8;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = 512;
constexpr int EPI_WARPS = 8;
constexpr int UNPACK_WARPS = 4;     // W4 path: nibble-packed weight tiles -> s8 UMMA operand tiles in shared memory
constexpr int U_STAGES = 3;         // ring of unpacked B tiles
constexpr int GEMM_THREADS_S8 = 64 + EPI_WARPS * 32;                         // TMA warp, MMA warp, 8 epilogue warps
constexpr int GEMM_THREADS_W4 = GEMM_THREADS_S8 + UNPACK_WARPS * 32;         // + 4 unpack warps (W4 storage)
constexpr int SMEM_LIMIT = 227 * 1024;

struct GemmParams {
  // problem
  int M, N, taps, S;        // M output rows, N out channels, taps = R*S filter taps, S = filter width
  int kbytes;               // bytes of K per pipeline step: 128 (128B-swizzled tiles) or 64 (64B-swizzled; channel counts = 64 mod 128)
  int k_chunks;             // ceil(Cp_range / kbytes) K steps per tap
  int k_last_mmas;          // MMAs (of 32 B) in the last chunk of each tap
  int a_c_offset;           // first channel of this K range inside the activation tensor (split shortcut)
  // tile -> coordinate mapping of the activation tensor map (dims: C, W, H, B)
  int Wo, HoWo;             // output row length and pixels per image (2-D GEMM: Wo = HoWo = 2^30)
  int block_n, n_tiles, m_tiles;
  int stages, b_stage_bytes;
  int w4;                   // 1: B arrives as 4-bit codes (two per byte, [Np][taps][Cp/2]) and is unpacked to s8 in shared memory
  int u_stage_bytes;        // bytes of one unpacked B tile (block_n x 128, rounded to 1024)
  const int32_t* zoff;      // [N] per-row offset subtracted from the 4-bit codes while unpacking (the weight zero-point)
  int row_staging;          // 1: row-major output staged through shared memory for 128-byte coalesced row stores
  // epilogue
  int out_hw;               // pixels per image of the OUTPUT layout (1 => row-major [M][N])
  int accumulate, silu;
  const float* residual;   // optional fp32 tensor laid out like `out`, added after bias / accumulate / SiLU
  const float* bias_img;   // optional fp32 [images][N]: per-(image, channel) term added like a residual (timestep embedding)
  const float* post;       // row-major outputs only: optional fp32 [M / post_rows][N] added AFTER the residual, one row per group of
  int post_rows;           //   post_rows consecutive output rows (the one-key cross-attention term of a transformer block)
  int post_shift;          //   log2(post_rows) when it is a power of two, else -1
  const float* delta_a;     // device scalars (nn.Parameter storage): no host sync on the path
  const float* zp_a;
  const float* delta_w;     // [N]
  const int32_t* wsum_eff;  // [N]  sum_k Wq + Kreal*cw
  const int32_t* cw;        // [N] or null
  const int32_t* rowsum;    // [M] or null (required when cw != null)
  const float* bias;        // [N] or null
  float* out;
  int debug;                // debug bits (EDADM_GEMM_DEBUG): 1 = epilogue skips global stores, 2 = skips TMEM loads too
  long long* trace;         // debug: per-CTA, per-tile role timestamps (edadm_debug_set_gemm_trace); null in production
};

struct __align__(8) PipeBarriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tmem_full[ACC_STAGES];
  uint64_t tmem_empty[ACC_STAGES];
  uint64_t ufull[U_STAGES];    // unpacked B tile ready (UNPACK_WARPS arrivals)
  uint64_t uempty[U_STAGES];   // unpacked B tile consumed by the MMAs (tcgen05.commit)
  uint32_t tmem_base;
};

constexpr int EPI_VEC_BYTES = MAX_BLOCK_N * 4 * 4;  // scale, zterm, cw, bias per column
constexpr int SMEM_FIXED = 1024 /*align slack*/ + EPI_VEC_BYTES + 1024 /*barriers*/;
constexpr int ROW_STAGE_LD = 36;                                   // floats per staged row (32 + pad, 16-byte aligned)
constexpr int ROW_STAGE_BYTES = EPI_WARPS * 32 * ROW_STAGE_LD * 4;  // one 32x32 fp32 tile per epilogue warp

// one 16-column chunk of the epilogue for one output row: int32 zero-point fold, fp32 scale + bias, store
template <bool GENERIC, bool RES>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&r)[16], int c0, int n_valid, int rs, const float* epi_scale,
                                               const int* epi_zterm, const int* epi_cw, const float* epi_bias, float* dst,
                                               const float (&t)[16], long long col_stride, bool accumulate, bool silu) {
  if (!GENERIC) {
    // all 16 columns valid, no rowsum term, plain store
    const int4* zt = reinterpret_cast<const int4*>(epi_zterm + c0);
    const float4* sc = reinterpret_cast<const float4*>(epi_scale + c0);
    const float4* bi = reinterpret_cast<const float4*>(epi_bias + c0);
    if (col_stride == 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int4 z = zt[q]; const float4 s = sc[q]; const float4 b = bi[q];
        float4 v;
        v.x = fmaf((float)((int)r[4 * q + 0] + z.x), s.x, b.x);
        v.y = fmaf((float)((int)r[4 * q + 1] + z.y), s.y, b.y);
        v.z = fmaf((float)((int)r[4 * q + 2] + z.z), s.z, b.z);
        v.w = fmaf((float)((int)r[4 * q + 3] + z.w), s.w, b.w);
        if (RES) { v.x += t[4 * q + 0]; v.y += t[4 * q + 1]; v.z += t[4 * q + 2]; v.w += t[4 * q + 3]; }
        *reinterpret_cast<float4*>(dst + 4 * q) = v;
      }
    } else if (RES) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int4 z = zt[q]; const float4 s = sc[q]; const float4 b = bi[q];
        dst[0] = fmaf((float)((int)r[4 * q + 0] + z.x), s.x, b.x) + t[4 * q + 0]; dst += col_stride;
        dst[0] = fmaf((float)((int)r[4 * q + 1] + z.y), s.y, b.y) + t[4 * q + 1]; dst += col_stride;
        dst[0] = fmaf((float)((int)r[4 * q + 2] + z.z), s.z, b.z) + t[4 * q + 2]; dst += col_stride;
        dst[0] = fmaf((float)((int)r[4 * q + 3] + z.w), s.w, b.w) + t[4 * q + 3]; dst += col_stride;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int4 z = zt[q]; const float4 s = sc[q]; const float4 b = bi[q];
        dst[0] = fmaf((float)((int)r[4 * q + 0] + z.x), s.x, b.x); dst += col_stride;
        dst[0] = fmaf((float)((int)r[4 * q + 1] + z.y), s.y, b.y); dst += col_stride;
        dst[0] = fmaf((float)((int)r[4 * q + 2] + z.z), s.z, b.z); dst += col_stride;
        dst[0] = fmaf((float)((int)r[4 * q + 3] + z.w), s.w, b.w); dst += col_stride;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < n_valid) {
        const float iv = (float)((int)r[j] + epi_cw[c0 + j] * rs + epi_zterm[c0 + j]);
        float v = fmaf(iv, epi_scale[c0 + j], epi_bias[c0 + j]);
        float* d = dst + (long long)j * col_stride;
        if (accumulate) v += *d;
        if (silu) v = v / (1.f + __expf(-v));
        if (RES) v += t[j];
        *d = v;
      }
    }
  }
}

// residual values of one 16-column chunk of one output row (issued well ahead of their use: the loads overlap the
// MMA wait and the previous chunk's stores instead of sitting on the epilogue's critical path)
__device__ __forceinline__ void load_residual(float (&t)[16], const float* res, long long col_stride, int n_valid) {
#pragma unroll
  for (int j = 0; j < 16; ++j) t[j] = (j < n_valid) ? __ldg(res + (long long)j * col_stride) : 0.f;
}

// W4 = false: s8 weight tiles straight from TMA, 320 threads (the epilogue keeps its registers); W4 = true: + unpack warps.
template <bool W4>
__global__ void __launch_bounds__(W4 ? GEMM_THREADS_W4 : GEMM_THREADS_S8, 1)
qgemm_i8_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer (LDS/STS, not generic LD/ST)
  uint8_t* smem_a = smem;
  const int a_stage_bytes = BLOCK_M * p.kbytes;
  uint8_t* smem_b = smem + p.stages * a_stage_bytes;
  float* epi_scale = reinterpret_cast<float*>(smem_b + p.stages * p.b_stage_bytes);
  int* epi_zterm = reinterpret_cast<int*>(epi_scale + MAX_BLOCK_N);
  int* epi_cw = epi_zterm + MAX_BLOCK_N;
  float* epi_bias = reinterpret_cast<float*>(epi_cw + MAX_BLOCK_N);
  PipeBarriers* bars = reinterpret_cast<PipeBarriers*>(epi_bias + MAX_BLOCK_N);
  float* row_stage = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 1024);   // only when p.row_staging
  uint8_t* smem_u = reinterpret_cast<uint8_t*>(row_stage) + (p.row_staging ? ROW_STAGE_BYTES : 0);   // only when W4 (1024-aligned by construction)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles;
  const int k_iters = p.taps * p.k_chunks;
  const int stages = p.stages;
  const uint32_t stage_tx = (uint32_t)a_stage_bytes + (uint32_t)p.block_n * (W4 ? BLOCK_K / 2 : p.kbytes);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int i = 0; i < stages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&bars->tmem_full[i], 1); mbar_init(&bars->tmem_empty[i], EPI_WARPS); }
    for (int i = 0; i < U_STAGES; ++i) { mbar_init(&bars->ufull[i], UNPACK_WARPS); mbar_init(&bars->uempty[i], 1); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  // The producer and the MMA issuer run their loops with the WHOLE warp (uniform control flow, loop state and operand
  // descriptors in uniform registers) and issue through one elected lane.  Running them inside `if (lane == 0)` makes every
  // operand "divergent" for the compiler, which then wraps each UTMALDG / UTCIMMA in an ELECT + R2UR + BRA.U.ANY loop: the
  // single issuing thread needed ~380 cycles per MMA where the tensor pipe needs 134 (N=192, scratch/mma_peak.py).
  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    int ti = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
      if (p.trace && lane == 0) p.trace[((size_t)blockIdx.x * 16 + (ti & 15)) * 8 + 5] = clock64();
      const int n_blk = tile % p.n_tiles, m_blk = tile / p.n_tiles;
      const int m0 = m_blk * BLOCK_M;
      const int b0 = m0 / p.HoWo;
      const int rem = m0 - b0 * p.HoWo;
      const int oh0 = rem / p.Wo, ow0 = rem - oh0 * p.Wo;
      const int n0 = n_blk * p.block_n;
      for (int tap = 0; tap < p.taps; ++tap) {
        const int kh = tap / p.S, kw = tap - kh * p.S;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&bars->full[stage], stage_tx);
            tma_load_4d(smem_a + stage * a_stage_bytes, &map_a, &bars->full[stage], p.a_c_offset + kc * p.kbytes, ow0 + kw, oh0 + kh, b0);
            tma_load_3d(smem_b + stage * p.b_stage_bytes, &map_b, &bars->full[stage], W4 ? kc * (BLOCK_K / 2) : kc * p.kbytes, tap, n0);
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
      if (p.trace && lane == 0) p.trace[((size_t)blockIdx.x * 16 + (ti & 15)) * 8 + 6] = clock64();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_i8(p.block_n, /*A u8*/ 0, /*B s8*/ 1);
    const uint32_t a_base = smem_u32(smem_a), b_base = smem_u32(smem_b), u_base = smem_u32(smem_u);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int us = 0;
    uint32_t uphase = 0;
    int ti = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
      mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      if (p.trace && lane == 0) p.trace[((size_t)blockIdx.x * 16 + (ti & 15)) * 8 + 0] = clock64();
      const uint32_t tmem_d = tmem_base + (uint32_t)acc * MAX_BLOCK_N;
      int kc = 0;
      for (int it = 0; it < k_iters; ++it) {
        const int nmma = (kc == p.k_chunks - 1) ? p.k_last_mmas : p.kbytes / UMMA_K;
        if (++kc == p.k_chunks) kc = 0;
        mbar_wait(&bars->full[stage], phase);
        if (W4) mbar_wait(&bars->ufull[us], uphase);
        tc_fence_after();
        if (p.trace && lane == 0 && it == 0) p.trace[((size_t)blockIdx.x * 16 + (ti & 15)) * 8 + 1] = clock64();
        const bool sw64 = !W4 && p.kbytes == 64;
        const uint64_t adesc = sw64 ? make_smem_desc_sw64(a_base + stage * a_stage_bytes) : make_smem_desc(a_base + stage * a_stage_bytes);
        const uint64_t bdesc = sw64 ? make_smem_desc_sw64(b_base + stage * p.b_stage_bytes)
                                    : make_smem_desc(W4 ? u_base + us * p.u_stage_bytes : b_base + stage * p.b_stage_bytes);
        if (elect_one()) {
          umma_i8(tmem_d, adesc, bdesc, idesc, it ? 1u : 0u);
          if (nmma > 1) umma_i8(tmem_d, adesc + 2, bdesc + 2, idesc, 1u);
          if (nmma > 2) umma_i8(tmem_d, adesc + 4, bdesc + 4, idesc, 1u);
          if (nmma > 3) umma_i8(tmem_d, adesc + 6, bdesc + 6, idesc, 1u);
          umma_commit(&bars->empty[stage]);     // A tile and (packed) B tile of this stage are free again
          if (W4) umma_commit(&bars->uempty[us]);
        }
        __syncwarp();
        if (W4 && ++us == U_STAGES) { us = 0; uphase ^= 1; }
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&bars->tmem_full[acc]);
      __syncwarp();
      if (p.trace && lane == 0) p.trace[((size_t)blockIdx.x * 16 + (ti & 15)) * 8 + 2] = clock64();
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 2 + EPI_WARPS) {
    // ===================== W4 unpack (warps 10..13) =====================
    // The packed tile of a stage holds block_n rows of 64 bytes = 128 four-bit codes; inside every 32-bit word the low
    // nibbles are codes 0..3 and the high nibbles codes 4..7 of that word's 8 codes (edadm_pack_weight_w4), so one AND
    // and one shift+AND split a word into two words of byte codes.  code - zoff[n] per byte without borrows between
    // bytes: (code + (0x80 - zoff)) ^ 0x80.  Rows are written in the 128B-swizzled K-major layout the UMMA descriptor
    // expects (16-byte chunk index XOR (row & 7)), exactly what TMA would have produced for s8 weights.
    if (W4) {
      const int ut = threadIdx.x - (64 + EPI_WARPS * 32);     // 0..127
      int stage = 0;
      uint32_t phase = 0;
      int us = 0;
      uint32_t uphase = 0;
      const int chunks = p.block_n * 4;                       // 16-byte packed chunks per tile
      constexpr int UT = UNPACK_WARPS * 32;
      constexpr int MAXC = MAX_BLOCK_N * 4 / UT;              // chunks per thread and K step (8)
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n0 = (tile % p.n_tiles) * p.block_n;
        // this thread always handles the same rows of a tile: row = ut/4 + 32*j -> per-row constants live in registers
        uint32_t kz[MAXC];
#pragma unroll
        for (int j = 0; j < MAXC; ++j) {
          const int n = n0 + (ut >> 2) + (UT / 4) * j;
          const int z = (j * UT + ut < chunks && n < p.N) ? __ldg(p.zoff + n) : 0;
          kz[j] = 0x80808080u - 0x01010101u * (uint32_t)z;
        }
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait_spin(&bars->full[stage], phase);
          const uint8_t* src = smem_b + stage * p.b_stage_bytes + ut * 16;
          uint4 pk[MAXC];
#pragma unroll
          for (int j = 0; j < MAXC; ++j)
            if (j * UT + ut < chunks) pk[j] = *reinterpret_cast<const uint4*>(src + j * UT * 16);
          mbar_wait_spin(&bars->uempty[us], uphase ^ 1);
          uint8_t* dst = smem_u + us * p.u_stage_bytes + (ut >> 2) * 128;
          const int c = ut & 3, sw = (ut >> 2) & 7;           // (row & 7) is the same for all of this thread's rows
          const int off0 = ((2 * c) ^ sw) << 4, off1 = ((2 * c + 1) ^ sw) << 4;
#pragma unroll
          for (int j = 0; j < MAXC; ++j) {
            if (j * UT + ut < chunks) {
              const uint32_t k = kz[j];
              uint4 o0, o1;
              o0.x = ((pk[j].x & 0x0F0F0F0Fu) + k) ^ 0x80808080u; o0.y = (((pk[j].x >> 4) & 0x0F0F0F0Fu) + k) ^ 0x80808080u;
              o0.z = ((pk[j].y & 0x0F0F0F0Fu) + k) ^ 0x80808080u; o0.w = (((pk[j].y >> 4) & 0x0F0F0F0Fu) + k) ^ 0x80808080u;
              o1.x = ((pk[j].z & 0x0F0F0F0Fu) + k) ^ 0x80808080u; o1.y = (((pk[j].z >> 4) & 0x0F0F0F0Fu) + k) ^ 0x80808080u;
              o1.z = ((pk[j].w & 0x0F0F0F0Fu) + k) ^ 0x80808080u; o1.w = (((pk[j].w >> 4) & 0x0F0F0F0Fu) + k) ^ 0x80808080u;
              uint8_t* row = dst + j * (UT / 4) * 128;
              *reinterpret_cast<uint4*>(row + off0) = o0;
              *reinterpret_cast<uint4*>(row + off1) = o1;
            }
          }
          fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->ufull[us]);
          if (++us == U_STAGES) { us = 0; uphase ^= 1; }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9): two warps per TMEM lane quarter, alternate 16-column chunks =====================
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;
    const int row_in_tile = quarter * 32 + lane;
    const int et = threadIdx.x - 64;              // 0..255
    const float da = __ldg(p.delta_a);
    const int za = (int)__ldg(p.zp_a);
    // float4 row stores need 16-byte aligned rows
    const bool generic_all = p.cw != nullptr || p.accumulate || p.silu ||
                             (p.out_hw == 1 && ((p.N & 3) || (reinterpret_cast<uintptr_t>(p.out) & 15) || (reinterpret_cast<uintptr_t>(p.residual) & 15)));
    const long long col_stride = p.out_hw;
    int acc = 0;
    uint32_t acc_phase = 0;
    int ti = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
      const int n_blk = tile % p.n_tiles, m_blk = tile / p.n_tiles;
      const int n0 = n_blk * p.block_n;
      for (int j = et; j < p.block_n; j += EPI_WARPS * 32) {
        const int n = n0 + j;
        const bool ok = n < p.N;
        epi_scale[j] = ok ? da * __ldg(p.delta_w + n) : 0.f;
        epi_zterm[j] = ok ? -za * __ldg(p.wsum_eff + n) : 0;
        epi_cw[j] = (ok && p.cw) ? __ldg(p.cw + n) : 0;
        epi_bias[j] = (ok && p.bias) ? __ldg(p.bias + n) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int m = m_blk * BLOCK_M + row_in_tile;
      const bool row_ok = m < p.M;
      const int rs = (row_ok && p.rowsum) ? __ldg(p.rowsum + m) : 0;
      const long long img = row_ok ? m / p.out_hw : 0;
      const long long pix = row_ok ? m - img * p.out_hw : 0;
      float* out_row = p.out + img * (long long)p.N * p.out_hw + pix + (long long)n0 * p.out_hw;
      // the per-image bias rides on the residual machinery: same prefetch, element (img, n) instead of (m, n)
      const float* res_row = p.residual ? p.residual + (out_row - p.out) : (p.bias_img ? p.bias_img + img * (long long)p.N + n0 : nullptr);
      const long long res_stride = p.residual ? col_stride : 1;

      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * MAX_BLOCK_N;
      if (p.row_staging) {
        // row-major output: 32-column chunks; values go through a per-warp 32x32 smem tile so that every store
        // instruction writes four complete 128-byte row segments instead of 32 scattered 16-byte pieces.  The residual
        // (if any) of a chunk is requested before its TMEM load and conversion, the first chunk's before the accumulator
        // is even ready.
        float* tile_s = row_stage + (warp - 2) * 32 * ROW_STAGE_LD;
        const int m_base = m_blk * BLOCK_M + quarter * 32;
        const int cq = (lane & 7) * 4;                 // 8 lanes x float4 = one 128-byte row segment
        float4 tres[8];
        auto load_res = [&](float4 (&t)[8], int c0) {
          const int ncol = n0 + c0 + cq;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int mm = m_base + i * 4 + (lane >> 3);
            t[i] = (mm < p.M && ncol < p.N) ? __ldg(reinterpret_cast<const float4*>(p.residual + (long long)mm * p.N + ncol))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        };
        const bool has_res = p.residual != nullptr;
        auto do_chunk = [&](int c0, bool res_loaded) {
          uint32_t r[32];
          if (has_res && !res_loaded) load_res(tres, c0);        // requested ahead of the TMEM load + conversion
          tmem_ld32(taddr + c0, r);
          tmem_ld_wait();
          const int4* zt = reinterpret_cast<const int4*>(epi_zterm + c0);
          const float4* sc = reinterpret_cast<const float4*>(epi_scale + c0);
          const float4* bi = reinterpret_cast<const float4*>(epi_bias + c0);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int4 z = zt[q]; const float4 s4 = sc[q]; const float4 b4 = bi[q];
            float4 v;
            v.x = fmaf((float)((int)r[4 * q + 0] + z.x), s4.x, b4.x);
            v.y = fmaf((float)((int)r[4 * q + 1] + z.y), s4.y, b4.y);
            v.z = fmaf((float)((int)r[4 * q + 2] + z.z), s4.z, b4.z);
            v.w = fmaf((float)((int)r[4 * q + 3] + z.w), s4.w, b4.w);
            *reinterpret_cast<float4*>(tile_s + lane * ROW_STAGE_LD + 4 * q) = v;
          }
          __syncwarp();
          const int ncol = n0 + c0 + cq;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + (lane >> 3);
            const int mm = m_base + rr;
            if (mm < p.M && ncol < p.N) {
              float4 v = *reinterpret_cast<const float4*>(tile_s + rr * ROW_STAGE_LD + cq);
              if (has_res) { v.x += tres[i].x; v.y += tres[i].y; v.z += tres[i].z; v.w += tres[i].w; }
              if (p.post) {
                const int grp = p.post_shift >= 0 ? (mm >> p.post_shift) : mm / p.post_rows;      // (token counts are powers of two)
                const float4 pv = __ldg(reinterpret_cast<const float4*>(p.post + (long long)grp * p.N + ncol));
                v.x += pv.x; v.y += pv.y; v.z += pv.z; v.w += pv.w;
              }
              *reinterpret_cast<float4*>(p.out + (long long)mm * p.N + ncol) = v;
            }
          }
          __syncwarp();
        };
        if (has_res && half * 32 < p.block_n) load_res(tres, half * 32);      // first chunk: before the accumulator is ready
        mbar_wait(&bars->tmem_full[acc], acc_phase);
        tc_fence_after();
        for (int c0 = half * 32; c0 < p.block_n; c0 += 64) do_chunk(c0, c0 == half * 32);
      } else {
      uint32_t r0[16], r1[16];
      float t0[16], t1[16];
      int c0 = half * 16;
      const bool has_res = res_row != nullptr && row_ok;
      if (has_res && c0 < p.block_n) load_residual(t0, res_row + (long long)c0 * res_stride, res_stride, p.N - (n0 + c0));
      mbar_wait(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();
      if (p.trace && et == 0) p.trace[((size_t)blockIdx.x * 16 + (ti & 15)) * 8 + 3] = clock64();
      if (p.debug & 2) c0 = p.block_n;
      const bool row_ok_dbg = row_ok && !(p.debug & 1);
      if (c0 < p.block_n) { tmem_ld16(taddr + c0, r0); }
      tmem_ld_wait();
      // software pipeline: the TMEM load (and residual load) of the next chunk is in flight while this one is converted and stored
      for (; c0 < p.block_n; c0 += 64) {
        const int c1 = c0 + 32;
        if (c1 < p.block_n) {
          tmem_ld16(taddr + c1, r1);
          if (has_res) load_residual(t1, res_row + (long long)c1 * res_stride, res_stride, p.N - (n0 + c1));
        }
        if (p.debug & 4) {     // experiment: same bytes as 4 fully coalesced STG.128 per thread (layout is garbage)
          float* cb = p.out + (size_t)tile * 128 * p.block_n + (size_t)(c0 / 16) * 2048 + quarter * 512;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 v;
            v.x = fmaf((float)((int)r0[4 * j + 0] + epi_zterm[c0 + 4 * j + 0]), epi_scale[c0 + 4 * j + 0], epi_bias[c0 + 4 * j + 0]);
            v.y = fmaf((float)((int)r0[4 * j + 1] + epi_zterm[c0 + 4 * j + 1]), epi_scale[c0 + 4 * j + 1], epi_bias[c0 + 4 * j + 1]);
            v.z = fmaf((float)((int)r0[4 * j + 2] + epi_zterm[c0 + 4 * j + 2]), epi_scale[c0 + 4 * j + 2], epi_bias[c0 + 4 * j + 2]);
            v.w = fmaf((float)((int)r0[4 * j + 3] + epi_zterm[c0 + 4 * j + 3]), epi_scale[c0 + 4 * j + 3], epi_bias[c0 + 4 * j + 3]);
            *reinterpret_cast<float4*>(cb + (j * 32 + lane) * 4) = v;
          }
        } else
        if (row_ok_dbg) {
          const int nv = p.N - (n0 + c0);
          float* dptr = out_row + (long long)c0 * col_stride;
          if (has_res) {
            if (!generic_all && nv >= 16) epilogue_chunk<false, true>(r0, c0, 16, rs, epi_scale, epi_zterm, epi_cw, epi_bias, dptr, t0, col_stride, false, false);
            else epilogue_chunk<true, true>(r0, c0, nv, rs, epi_scale, epi_zterm, epi_cw, epi_bias, dptr, t0, col_stride, p.accumulate, p.silu);
          } else {
            if (!generic_all && nv >= 16) epilogue_chunk<false, false>(r0, c0, 16, rs, epi_scale, epi_zterm, epi_cw, epi_bias, dptr, t0, col_stride, false, false);
            else epilogue_chunk<true, false>(r0, c0, nv, rs, epi_scale, epi_zterm, epi_cw, epi_bias, dptr, t0, col_stride, p.accumulate, p.silu);
          }
        }
        tmem_ld_wait();
        const int c2 = c0 + 64;
        if (c2 < p.block_n) {
          tmem_ld16(taddr + c2, r0);
          if (has_res) load_residual(t0, res_row + (long long)c2 * res_stride, res_stride, p.N - (n0 + c2));
        }
        if (p.debug & 4) {
          if (c1 < p.block_n) {
          float* cb = p.out + (size_t)tile * 128 * p.block_n + (size_t)(c1 / 16) * 2048 + quarter * 512;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 v;
            v.x = fmaf((float)((int)r1[4 * j + 0] + epi_zterm[c1 + 4 * j + 0]), epi_scale[c1 + 4 * j + 0], epi_bias[c1 + 4 * j + 0]);
            v.y = fmaf((float)((int)r1[4 * j + 1] + epi_zterm[c1 + 4 * j + 1]), epi_scale[c1 + 4 * j + 1], epi_bias[c1 + 4 * j + 1]);
            v.z = fmaf((float)((int)r1[4 * j + 2] + epi_zterm[c1 + 4 * j + 2]), epi_scale[c1 + 4 * j + 2], epi_bias[c1 + 4 * j + 2]);
            v.w = fmaf((float)((int)r1[4 * j + 3] + epi_zterm[c1 + 4 * j + 3]), epi_scale[c1 + 4 * j + 3], epi_bias[c1 + 4 * j + 3]);
            *reinterpret_cast<float4*>(cb + (j * 32 + lane) * 4) = v;
          }
          }
        } else
        if (c1 < p.block_n && row_ok_dbg) {
          const int nv = p.N - (n0 + c1);
          float* dptr = out_row + (long long)c1 * col_stride;
          if (has_res) {
            if (!generic_all && nv >= 16) epilogue_chunk<false, true>(r1, c1, 16, rs, epi_scale, epi_zterm, epi_cw, epi_bias, dptr, t1, col_stride, false, false);
            else epilogue_chunk<true, true>(r1, c1, nv, rs, epi_scale, epi_zterm, epi_cw, epi_bias, dptr, t1, col_stride, p.accumulate, p.silu);
          } else {
            if (!generic_all && nv >= 16) epilogue_chunk<false, false>(r1, c1, 16, rs, epi_scale, epi_zterm, epi_cw, epi_bias, dptr, t1, col_stride, false, false);
            else epilogue_chunk<true, false>(r1, c1, nv, rs, epi_scale, epi_zterm, epi_cw, epi_bias, dptr, t1, col_stride, p.accumulate, p.silu);
          }
        }
        tmem_ld_wait();
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
      if (p.trace && et == 0) p.trace[((size_t)blockIdx.x * 16 + (ti & 15)) * 8 + 4] = clock64();
      asm volatile("bar.sync 1, 256;" ::: "memory");  // epi_* vectors are rewritten next tile
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// choose the N tile (multiple of `step`, <= 256).  A persistent CTA needs  waves x k_iters x t(bn)  with
// waves = ceil(m_tiles * ceil(N / bn) / SMs) and a per-k-iteration time that grows with the bytes staged per iteration
// (128 activation rows + bn weight rows, plus a fixed issue / latency part worth ~64 rows).  Wide layers with many M tiles
// end up with the fewest, widest tiles; the small-M layers deep in the UNet and the embedding linears get narrow tiles so
// that all SMs take part instead of a dozen.
static int pick_block_n(int N, int m_tiles, int sms, int step) {
  int best = 0;
  long long best_cost = 0;
  for (int bn = step; bn <= MAX_BLOCK_N; bn += step) {
    const long long tiles = (long long)m_tiles * ((N + bn - 1) / bn);
    const long long waves = (tiles + sms - 1) / sms;
    const long long cost = waves * (bn + 192);
    if (best == 0 || cost <= best_cost) { best = bn; best_cost = cost; }
    if (bn >= N) break;           // wider tiles only add padding
  }
  return best;
}

}  // namespace edadm

using namespace edadm;

static long long* g_gemm_trace = nullptr;
// debug hook (scratch/ timeline experiments only): per-CTA [16 tiles][8] clock64 stamps of the warp roles
extern "C" int edadm_debug_set_gemm_trace(long long* buf) { g_gemm_trace = buf; return 0; }
namespace edadm { long long* gemm_trace_buffer() { return g_gemm_trace; } }

// Activation codes q: [B][Hp][Wp][Cp_act] u8 (halo included), filter R x S, stride 1:  Ho = Hp-R+1, Wo = Wp-S+1.
// A 2-D GEMM ([M][Kp] rows) is the special case B=1, Hp=1, Wp=M, R=S=1.
// Weights wq: [Np][R*S][Cp_w] s8 with Np >= N rows allocated (rows past Np are zero-filled by TMA).
static int launch_qgemm(const uint8_t* q, int B, int Hp, int Wp, int Cp_act, int a_c_offset, const void* wq, int w4,
                        const int32_t* zoff, int N, int Np, int R, int S, int Cp_w, const float* delta_a, const float* zp_a,
                        const float* delta_w, const int32_t* wsum_eff, const int32_t* cw, const int32_t* rowsum,
                        const float* bias, const float* bias_img, const float* residual, float* out, int out_hw, int accumulate,
                        int silu, void* stream, const float* post = nullptr, int post_rows = 0) {
  if (bias_img && (residual || out_hw == 1))
    return fail(EDADM_ERR_UNSUPPORTED, "qgemm_i8: bias_img needs an NCHW output (out_hw > 1) and cannot be combined with residual");
  if (!q || !wq || !delta_a || !zp_a || !delta_w || !wsum_eff || !out) return fail(EDADM_ERR_ARG, "qgemm_i8: null pointer");
  if (cw && !rowsum) return fail(EDADM_ERR_ARG, "qgemm_i8: cw given without rowsum");
  if (B < 1 || Hp < R || Wp < S || R < 1 || S < 1 || N < 1 || (Cp_act & 15) || (Cp_w & 15) || Cp_w < 16 || a_c_offset < 0 ||
      a_c_offset + 16 > Cp_act + 15)
    return fail(EDADM_ERR_ARG, "qgemm_i8: bad geometry B=%d Hp=%d Wp=%d Cp_act=%d Cp_w=%d R=%d S=%d N=%d", B, Hp, Wp, Cp_act, Cp_w, R, S, N);
  if ((((uintptr_t)q) & 15) || (((uintptr_t)wq) & 15)) return fail(EDADM_ERR_ARG, "qgemm_i8: operands must be 16-byte aligned");
  if (w4 && (!zoff || (Cp_w & 31))) return fail(EDADM_ERR_ARG, "qgemm_w4a8: packed weights need zoff and a channel pitch that is a multiple of 32 (got %d)", Cp_w);
  const int Ho = Hp - R + 1, Wo = Wp - S + 1;
  const long long M = (long long)B * Ho * Wo;
  if (M > 0x7fffffffLL) return fail(EDADM_ERR_UNSUPPORTED, "qgemm_i8: M too large");
  if (out_hw < 1 || (M % out_hw) != 0) return fail(EDADM_ERR_ARG, "qgemm_i8: out_hw=%d does not divide M=%lld", out_hw, M);

  // tile box over (W, H, B): 128 consecutive output pixels
  int box_w, box_h, box_b;
  const bool flat = (Ho == 1 && B == 1);
  if (flat) {
    box_w = BLOCK_M; box_h = 1; box_b = 1;
  } else if (Wo >= BLOCK_M) {
    if (Wo % BLOCK_M) return fail(EDADM_ERR_UNSUPPORTED, "qgemm_i8: output width %d not a multiple of 128", Wo);
    box_w = BLOCK_M; box_h = 1; box_b = 1;
  } else {
    if (BLOCK_M % Wo) return fail(EDADM_ERR_UNSUPPORTED, "qgemm_i8: output width %d does not divide 128 (use the im2col route)", Wo);
    box_w = Wo;
    const int rows = BLOCK_M / Wo;
    if (rows <= Ho) {
      if (Ho % rows) return fail(EDADM_ERR_UNSUPPORTED, "qgemm_i8: output height %d not a multiple of %d rows per tile", Ho, rows);
      box_h = rows; box_b = 1;
    } else {
      if (rows % Ho) return fail(EDADM_ERR_UNSUPPORTED, "qgemm_i8: %d rows per tile not a multiple of output height %d", rows, Ho);
      box_h = Ho; box_b = rows / Ho;
    }
  }

  if (Np < N) return fail(EDADM_ERR_ARG, "qgemm_i8: weight rows Np=%d < N=%d", Np, N);
  const int m_tiles_all = (int)((M + BLOCK_M - 1) / BLOCK_M);
  // row-major outputs are staged in 32-column pieces; weight rows past Np are zero-filled by TMA
  const int block_n = pick_block_n(N, m_tiles_all, sm_count(), out_hw == 1 ? 32 : 16);
  const int n_tiles = (N + block_n - 1) / block_n;

  // K bytes per pipeline step: channel counts of the form 128 j + 64 (192, 576, 960 ...) would leave every tap's last 128-byte
  // chunk half out of bounds -- such TMA boxes are slow (measured: 711 cycles for a half step vs 465 for a full one) -- so
  // those layers run on 64-byte steps (64B-swizzled tiles, two MMAs per step) with every box in bounds.
  int kbytes = BLOCK_K;
  if (!w4 && (Cp_w % 128) == 64 && Cp_w <= 192) kbytes = 64;
  if (const char* e = getenv("EDADM_GEMM_KBYTES")) { const int v = atoi(e); if (!w4 && (v == 64 || v == 128)) kbytes = v; }   // debug / tuning
  CUtensorMap map_a, map_b;
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cp_act, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)Cp_act, (cuuint64_t)Wp * Cp_act, (cuuint64_t)Hp * Wp * Cp_act};
    cuuint32_t box[4] = {(cuuint32_t)kbytes, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_b};
    int rc = encode_map(&map_a, q, 4, dims, strides, box, "activations", CU_TENSOR_MAP_DATA_TYPE_UINT8,
                        kbytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  if (w4) {   // two codes per byte, plain (unswizzled) 64-byte rows; the unpack warps produce the swizzled s8 tile
    cuuint64_t dims[3] = {(cuuint64_t)(Cp_w / 2), (cuuint64_t)(R * S), (cuuint64_t)Np};
    cuuint64_t strides[2] = {(cuuint64_t)(Cp_w / 2), (cuuint64_t)R * S * (Cp_w / 2)};
    cuuint32_t box[3] = {(cuuint32_t)(BLOCK_K / 2), 1u, (cuuint32_t)block_n};
    int rc = encode_map(&map_b, wq, 3, dims, strides, box, "packed weights", CU_TENSOR_MAP_DATA_TYPE_UINT8, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  } else {
    cuuint64_t dims[3] = {(cuuint64_t)Cp_w, (cuuint64_t)(R * S), (cuuint64_t)Np};
    cuuint64_t strides[2] = {(cuuint64_t)Cp_w, (cuuint64_t)R * S * Cp_w};
    cuuint32_t box[3] = {(cuuint32_t)kbytes, 1u, (cuuint32_t)block_n};
    int rc = encode_map(&map_b, wq, 3, dims, strides, box, "weights", CU_TENSOR_MAP_DATA_TYPE_UINT8,
                        kbytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }

  GemmParams p;
  p.M = (int)M; p.N = N; p.taps = R * S; p.S = S;
  p.kbytes = kbytes;
  p.k_chunks = (Cp_w + kbytes - 1) / kbytes;
  const int last_bytes = Cp_w - (p.k_chunks - 1) * kbytes;
  p.k_last_mmas = (last_bytes + UMMA_K - 1) / UMMA_K;
  p.a_c_offset = a_c_offset;
  p.Wo = flat ? (1 << 30) : Wo;
  p.HoWo = flat ? (1 << 30) : Ho * Wo;
  p.block_n = block_n; p.n_tiles = n_tiles; p.m_tiles = (int)((M + BLOCK_M - 1) / BLOCK_M);
  p.w4 = w4; p.zoff = zoff;
  p.b_stage_bytes = (block_n * (w4 ? BLOCK_K / 2 : kbytes) + 1023) & ~1023;
  p.u_stage_bytes = (block_n * BLOCK_K + 1023) & ~1023;
  // row-major outputs (linear layers) are staged through smem; needs whole float4s per row and no rowsum / accumulate / SiLU
  p.row_staging = (out_hw == 1 && !cw && !accumulate && !silu && (N % 4) == 0 && (block_n % 32) == 0 && ((reinterpret_cast<uintptr_t>(residual) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) ? 1 : 0;
  const int fixed = SMEM_FIXED + (p.row_staging ? ROW_STAGE_BYTES : 0) + (w4 ? U_STAGES * p.u_stage_bytes : 0);
  p.stages = (SMEM_LIMIT - fixed) / (BLOCK_M * kbytes + p.b_stage_bytes);
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  if (const char* e = getenv("EDADM_GEMM_STAGES")) { const int v = atoi(e); if (v >= 2 && v < p.stages) p.stages = v; }   // debug
  const int smem_bytes = fixed + p.stages * (BLOCK_M * kbytes + p.b_stage_bytes);
  p.out_hw = out_hw; p.accumulate = accumulate; p.silu = silu; p.residual = residual; p.bias_img = bias_img;
  p.post = post; p.post_rows = post_rows; p.post_shift = -1;
  if (post_rows > 0 && (post_rows & (post_rows - 1)) == 0) { p.post_shift = 0; while ((1 << p.post_shift) < post_rows) ++p.post_shift; }
  if (post && (!p.row_staging || post_rows < 1 || (M % post_rows) || (reinterpret_cast<uintptr_t>(post) & 15)))
    return fail(EDADM_ERR_UNSUPPORTED, "qgemm_i8: the row-group term needs a staged row-major output (N %% 4 == 0, no rowsum / accumulate) and post_rows dividing M");
  p.delta_a = delta_a; p.zp_a = zp_a; p.delta_w = delta_w; p.wsum_eff = wsum_eff; p.cw = cw; p.rowsum = rowsum;
  p.bias = bias; p.out = out;
  p.trace = g_gemm_trace;
  p.debug = 0;
  if (const char* e = getenv("EDADM_GEMM_DEBUG")) p.debug = atoi(e);

  static bool attr_set_dev[64] = {false};       // the opt-in is per device
  int attr_dev = 0;
  cudaGetDevice(&attr_dev);
  attr_dev = attr_dev < 64 ? attr_dev : 63;
  if (!attr_set_dev[attr_dev]) {
    cudaError_t e = cudaFuncSetAttribute(qgemm_i8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(qgemm_i8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return fail(EDADM_ERR_CUDA, "qgemm_i8: cannot opt in to %d B shared memory: %s", SMEM_LIMIT, cudaGetErrorString(e));
    attr_set_dev[attr_dev] = true;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  if (w4) qgemm_i8_kernel<true><<<grid, GEMM_THREADS_W4, smem_bytes, (cudaStream_t)stream>>>(map_a, map_b, p);
  else qgemm_i8_kernel<false><<<grid, GEMM_THREADS_S8, smem_bytes, (cudaStream_t)stream>>>(map_a, map_b, p);
  return check_launch("qgemm_i8");
}

namespace edadm {
struct Gemm2Args {
  const uint8_t* q; int B, Hp, Wp, Cp_act, a_c_offset;
  const void* wq; int N, Np, R, S, Cp_w;
  const float* delta_a; const float* zp_a; const float* delta_w; const int32_t* wsum_eff; const int32_t* cw; const int32_t* rowsum;
  const float* bias; const float* bias_img; const float* residual;
  void* out; int out_hw; int accumulate;
  int out_mode;
  const float* q_delta; const float* q_zp; int q_levels; int32_t* q_rowsum; int out_pitch;
  const void* wq1 = nullptr; int Cp_w1 = 0; int a_c_offset1 = 0;
  const float* delta_a1 = nullptr; const float* zp_a1 = nullptr; const float* delta_w1 = nullptr; const int32_t* wsum_eff1 = nullptr;
};
int launch_qgemm2(const Gemm2Args& a, void* stream);     // qgemm2_sm100.cu; +1 = case not covered
}

extern "C" int edadm_qgemm_i8(const uint8_t* q, int B, int Hp, int Wp, int Cp_act, int a_c_offset, const int8_t* wq,
                              int N, int Np, int R, int S, int Cp_w, const float* delta_a, const float* zp_a,
                              const float* delta_w, const int32_t* wsum_eff, const int32_t* cw, const int32_t* rowsum,
                              const float* bias, const float* bias_img, const float* residual, float* out, int out_hw,
                              int accumulate, int silu, void* stream) {
  const bool force_v1 = getenv("EDADM_GEMM_V1") != nullptr;
  if (!force_v1 && !silu && q && wq && delta_a && zp_a && delta_w && wsum_eff && out && (!cw || rowsum) && B >= 1 && Hp >= R && Wp >= S &&
      R >= 1 && S >= 1 && N >= 1 && !(Cp_act & 15) && !(Cp_w & 15) && Cp_w >= 16 && a_c_offset >= 0 && a_c_offset + 16 <= Cp_act + 15 &&
      !(((uintptr_t)q) & 15) && !(((uintptr_t)wq) & 15) && out_hw >= 1 && !(bias_img && out_hw == 1)) {
    Gemm2Args a{q, B, Hp, Wp, Cp_act, a_c_offset, wq, N, Np, R, S, Cp_w, delta_a, zp_a, delta_w, wsum_eff, cw, rowsum, bias, bias_img, residual,
                out, out_hw, accumulate, /*out_mode (decided from out_hw)*/ 0, nullptr, nullptr, 0, nullptr, 0};
    // Which generation (profiles/gemm_shapes_r02_*.txt): the second one wins where operand traffic or the residual read
    // dominates (>= 8 K steps per tile, or an NCHW residual / accumulate operand); short-K layers are bound by the fp32 output
    // write, where the first generation's eight independently storing epilogue warps keep more bytes in flight.
    const int k_steps = R * S * ((Cp_w + 127) / 128);
    const bool force_v2 = getenv("EDADM_GEMM_V2") != nullptr;
    if (force_v2 || (out_hw > 1 && (k_steps >= 8 || residual || accumulate))) {
      const int rc = launch_qgemm2(a, stream);
      if (rc <= 0) return rc;
    }
  }
  return launch_qgemm(q, B, Hp, Wp, Cp_act, a_c_offset, wq, 0, nullptr, N, Np, R, S, Cp_w, delta_a, zp_a, delta_w, wsum_eff, cw,
                      rowsum, bias, bias_img, residual, out, out_hw, accumulate, silu, stream);
}

// Split shortcut (quant_layer.py:415-432: the input channels [0, split) and [split, C) have their own activation AND weight
// quantizers): out = conv(x[:, :split], w0) + conv(x[:, split:], w1) (+ bias) (+ residual) as ONE launch -- both K ranges of the
// same NHWC code tensor, two TMEM accumulators combined in the epilogue exactly as the two-launch form (second launch accumulating
// onto the first one's fp32 output) rounds: fma(acc0, s0, bias) + round(acc1 * s1).  Falls back to those two launches when the
// second-generation kernel does not cover the geometry.
extern "C" int edadm_qgemm_i8_split(const uint8_t* q, int B, int Hp, int Wp, int Cp_act, const int8_t* wq0, const int8_t* wq1, int N,
                                    int Np, int R, int S, int Cp_w0, int Cp_w1, int a_c_offset1, const float* delta_a0,
                                    const float* zp_a0, const float* delta_a1, const float* zp_a1, const float* delta_w0,
                                    const float* delta_w1, const int32_t* wsum_eff0, const int32_t* wsum_eff1, const float* bias,
                                    const float* bias_img, const float* residual, float* out, int out_hw, void* stream) {
  if (!q || !wq0 || !wq1 || !delta_a0 || !zp_a0 || !delta_a1 || !zp_a1 || !delta_w0 || !delta_w1 || !wsum_eff0 || !wsum_eff1 || !out)
    return fail(EDADM_ERR_ARG, "qgemm_i8_split: null pointer");
  if (getenv("EDADM_GEMM_V1") == nullptr && B >= 1 && Hp >= R && Wp >= S && R >= 1 && S >= 1 && N >= 1 && !(Cp_act & 15) && !(Cp_w0 & 15) &&
      Cp_w0 >= 16 && !(((uintptr_t)q) & 15) && !(((uintptr_t)wq0) & 15) && out_hw >= 1 && !(bias_img && out_hw == 1)) {
    Gemm2Args a{q, B, Hp, Wp, Cp_act, 0, wq0, N, Np, R, S, Cp_w0, delta_a0, zp_a0, delta_w0, wsum_eff0, nullptr, nullptr, bias, bias_img,
                residual, out, out_hw, 0, 0, nullptr, nullptr, 0, nullptr, 0};
    a.wq1 = wq1; a.Cp_w1 = Cp_w1; a.a_c_offset1 = a_c_offset1;
    a.delta_a1 = delta_a1; a.zp_a1 = zp_a1; a.delta_w1 = delta_w1; a.wsum_eff1 = wsum_eff1;
    const int rc = launch_qgemm2(a, stream);
    if (rc <= 0) return rc;
  }
  int rc = edadm_qgemm_i8(q, B, Hp, Wp, Cp_act, 0, wq0, N, Np, R, S, Cp_w0, delta_a0, zp_a0, delta_w0, wsum_eff0, nullptr, nullptr, bias,
                          nullptr, nullptr, out, out_hw, 0, 0, stream);
  if (rc) return rc;
  return edadm_qgemm_i8(q, B, Hp, Wp, Cp_act, a_c_offset1, wq1, N, Np, R, S, Cp_w1, delta_a1, zp_a1, delta_w1, wsum_eff1, nullptr, nullptr,
                        nullptr, bias_img, residual, out, out_hw, 1, 0, stream);
}

// Linear layer (row-major output) whose epilogue adds, after the residual, one fp32 row per group of `post_rows` output rows:
//   out[m][n] = ((acc * scale + bias[n]) + residual[m][n]) + post[m / post_rows][n]      (each + one fp32 rounding, in this order)
extern "C" int edadm_qgemm_i8_rows_post(const uint8_t* q, int64_t M, int Kp_act, const int8_t* wq, int N, int Np, int Cp_w,
                                        const float* delta_a, const float* zp_a, const float* delta_w, const int32_t* wsum_eff,
                                        const float* bias, const float* residual, const float* post, int post_rows, float* out,
                                        void* stream) {
  if (M < 1 || M > 0x7fffffffLL || !post) return fail(EDADM_ERR_ARG, "qgemm_i8_rows_post: bad arguments");
  return launch_qgemm(q, 1, 1, (int)M, Kp_act, 0, wq, 0, nullptr, N, Np, 1, 1, Cp_w, delta_a, zp_a, delta_w, wsum_eff, nullptr, nullptr,
                      bias, nullptr, residual, out, 1, 0, 0, stream, post, post_rows);
}

// Same GEMM with the weights stored as 4-bit codes, two per byte: wq4 [Np][R*S][Cp_w/2] (Cp_w % 32 == 0; inside each 32-bit
// word the low nibbles hold codes 0..3 and the high nibbles codes 4..7 of the word's 8 channels, see edadm_pack_weight_w4).
// The kernel unpacks every tile to s8 (code - zoff[n]) in shared memory before the MMA; wsum_eff must be sum_k (code - zoff).
extern "C" int edadm_qgemm_w4a8(const uint8_t* q, int B, int Hp, int Wp, int Cp_act, int a_c_offset, const uint8_t* wq4,
                                const int32_t* zoff, int N, int Np, int R, int S, int Cp_w, const float* delta_a,
                                const float* zp_a, const float* delta_w, const int32_t* wsum_eff, const float* bias,
                                const float* bias_img, const float* residual, float* out, int out_hw, int accumulate, int silu,
                                void* stream) {
  if (!zoff) return fail(EDADM_ERR_ARG, "qgemm_w4a8: null zoff");
  return launch_qgemm(q, B, Hp, Wp, Cp_act, a_c_offset, wq4, 1, zoff, N, Np, R, S, Cp_w, delta_a, zp_a, delta_w, wsum_eff, nullptr,
                      nullptr, bias, bias_img, residual, out, out_hw, accumulate, silu, stream);
}
