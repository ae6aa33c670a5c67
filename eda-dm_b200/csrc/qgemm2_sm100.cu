// QuantModule GEMM, second generation (replaces qdiff/quant_layer.py:434 on the integer path).
//
// Same arithmetic as qgemm_sm100.cu (exact u8 x s8 -> s32 implicit GEMM on tcgen05, zero-point fold + dequant in the
// epilogue); what changed is how bytes move, following the round-2 timeline measurements (profiles/gemm_timeline_r02.txt):
//   * the SM <-> L2 port, not the tensor pipe, bounds these layers: ~0.67 TMA rows (<= 128 B) per clock inbound, and every
//     32-byte sector stored by the epilogue's STG stalls the operand stream for about a clock.  So
//   * CTA PAIRS (cta_group::2, CTAS = 2): one MMA covers 256 output rows, each CTA of the pair stages its own 128 activation
//     rows but only HALF of the weight tile -- a third fewer operand rows per K step;
//   * the epilogue goes TMEM -> registers -> shared memory -> TMA STORE (cp.async.bulk.tensor, full 128-byte lines through
//     the async proxy) in 32-column chunks, double buffered; the residual / accumulate operand comes in the same way
//     (TMA load into a small ring, prefetched two chunks ahead) instead of strided LDGs on the critical path;
//   * optional fused consumers: the epilogue can emit the NEXT quantizer's u8 codes instead of fp32 (plain, or GEGLU-gated:
//     a * gelu(g) with the a / g columns of one tile taken from the two halves of the projection), so 4-byte outputs that
//     would be re-read once and thrown away never reach L2.
//
// One persistent CTA (pair) per SM (pair): warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warps 2..9 epilogue.
#include "tc05.cuh"
#include <cstdlib>

namespace edadm {
namespace g2 {

constexpr int BM = 128;                       // rows per CTA
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + EPI_WARPS * 32;  // 320
constexpr int CHUNK = 32;                     // output columns per epilogue step
constexpr int OUT_BUF_BYTES = BM * CHUNK * 4; // 16 KB (fp32) -- u8 code chunks use the first 4 KB
constexpr int MAX_BN = 256;
constexpr int MAX_STAGES = 8;
constexpr int ACC_STAGES = 2;
constexpr int MAX_RES_BUFS = 3;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int EPI_VEC_BYTES = MAX_BN * 4 * 4;

enum OutMode { OUT_F32_NCHW = 0, OUT_F32_ROWS = 1, OUT_U8_ROWS = 2, OUT_U8_GEGLU = 3 };

struct Params {
  int M, N, taps, S;
  int kbytes, k_chunks, k_last_mmas;
  int a_c_offset;
  // split shortcut (two quantizer pairs along K, quant_layer.py:415-432): a second K range with its own weights / scales, summed
  // into a second TMEM accumulator (columns + block_n) and combined in the epilogue -- one launch, no fp32 round trip between them
  int dual, k_chunks1, k_last_mmas1, a_c_offset1;
  // short-K linears (K <= 4 chunks, single CTAs): the weight tile of the CTA's N block stays in smem for the whole launch (b_res) and
  // the stage ring carries activation tiles only -- with 3 K steps per tile an A+B ring is one tile deep and every tile pays the
  // full TMA latency (measured: 370 us of the 475 us GEGLU projection remain with the whole epilogue switched off)
  int b_res;
  long long* trace;                         // debug (edadm_debug_set_gemm_trace): per CTA, per tile (first 16) role timestamps; null in production
  const float* delta_a1; const float* zp_a1; const float* delta_w1; const int32_t* wsum_eff1;
  int Wo, HoWo;
  int block_n, n_tiles, m_units;            // m_units = ceil(m_tiles / CTAS): scheduling units along M
  int stages, a_stage_bytes, b_stage_bytes;
  int out_mode;
  int out_hw, pxb, px_shift;                // NCHW: pixels per image, pixels of one image inside a tile (power of two) and its log2
  int res_mode;                             // 0 none, 1 fp32 tensor shaped like the output (TMA-loaded), 2 per-(image, channel) bias
  int res_bufs;
  int geglu_half;                           // OUT_U8_GEGLU: N / 2 (gate columns start here)
  const float* bias_img;
  const float* delta_a; const float* zp_a; const float* delta_w;
  const int32_t* wsum_eff; const int32_t* cw; const int32_t* rowsum; const float* bias;
  // code-emitting epilogues: the consumer's activation quantizer
  const float* q_delta; const float* q_zp; float q_max;
  int32_t* q_rowsum;                        // optional: += sum of the emitted codes per row (atomic; pre-zeroed by the caller)
};

struct __align__(8) Barriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tmem_full[ACC_STAGES];
  uint64_t tmem_empty[ACC_STAGES];
  uint64_t res_full[MAX_RES_BUFS];
  uint64_t b_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ float gelu_erf(float x) {
  // ATen's CUDA GELU (approximate='none'): x * 0.5 * (1 + erf(x * M_SQRT1_2))
  return x * 0.5f * (1.0f + erff_two_poly(x * 0.70710678118654752440f));
}

// same bits as clamp(round(x / d) + z, 0, qmax) (see pack.cu quant_code_fast)
__device__ __forceinline__ uint32_t q_code(float x, float d, float inv_d, float z, float qmax) {
  const float q0 = x * inv_d;
  const float q1 = __fmaf_rn(__fmaf_rn(-d, q0, x), inv_d, q0);
  const float t = fminf(fmaxf(q1, -z), qmax - z) + 12582912.0f;       // round-to-nearest-even in the low mantissa bits, no conversion pipe
  return (uint32_t)(__float_as_int(t) - 0x4B400000 + (int)z);
}

// v[j] = (acc[j] + zterm[j] (+ cw[j] * rowsum)) * scale[j] + bias[j] for 16 consecutive columns (pointers 16-byte aligned)
__device__ __forceinline__ void dequant16(const uint32_t (&acc)[16], float (&v)[16], const int* __restrict__ zterm,
                                          const float* __restrict__ scale, const float* __restrict__ bias, const int* __restrict__ cw, int rs) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int4 z = reinterpret_cast<const int4*>(zterm)[k];
    const float4 s4 = reinterpret_cast<const float4*>(scale)[k];
    const float4 b4 = reinterpret_cast<const float4*>(bias)[k];
    if (cw) {                                     // 8-bit weight codes with a zero-point off 128: warp-uniform, rare
      const int4 c4 = reinterpret_cast<const int4*>(cw)[k];
      z.x += c4.x * rs; z.y += c4.y * rs; z.z += c4.z * rs; z.w += c4.w * rs;
    }
    v[4 * k + 0] = fmaf((float)((int)acc[4 * k + 0] + z.x), s4.x, b4.x);
    v[4 * k + 1] = fmaf((float)((int)acc[4 * k + 1] + z.y), s4.y, b4.y);
    v[4 * k + 2] = fmaf((float)((int)acc[4 * k + 2] + z.z), s4.z, b4.z);
    v[4 * k + 3] = fmaf((float)((int)acc[4 * k + 3] + z.w), s4.w, b4.w);
  }
}

template <int CTAS>
__global__ void __launch_bounds__(THREADS, 1)
qgemm2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
              const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res,
              const __grid_constant__ CUtensorMap map_b1, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + p.stages * p.a_stage_bytes;
  uint8_t* out_buf = smem_b + (p.b_res ? p.k_chunks : p.stages) * p.b_stage_bytes;
  uint8_t* res_buf = out_buf + 2 * OUT_BUF_BYTES;
  float* epi_scale = reinterpret_cast<float*>(res_buf + p.res_bufs * OUT_BUF_BYTES);
  int* epi_zterm = reinterpret_cast<int*>(epi_scale + MAX_BN);
  int* epi_cw = epi_zterm + MAX_BN;
  float* epi_bias = reinterpret_cast<float*>(epi_cw + MAX_BN);
  Barriers* bars = reinterpret_cast<Barriers*>(epi_bias + MAX_BN);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;
  // b_res: CTA c owns N block c % n_tiles and every (gridDim / n_tiles)-th M tile; otherwise units are dealt round robin
  const int unit0 = CTAS == 2 ? (int)cluster_id_x() : (p.b_res ? ((int)blockIdx.x / p.n_tiles) * p.n_tiles + (int)blockIdx.x % p.n_tiles : (int)blockIdx.x);
  const int unit_step = CTAS == 2 ? (int)num_clusters_x() : (p.b_res ? ((int)gridDim.x / p.n_tiles) * p.n_tiles : (int)gridDim.x);
  const int num_units = p.m_units * p.n_tiles;
  const int k_iters0 = p.taps * p.k_chunks;
  const int k_iters = k_iters0 + (p.dual ? p.taps * p.k_chunks1 : 0);
  const int stages = p.stages;
  const int bn_cta = p.block_n / CTAS;                        // weight rows staged by this CTA
  const uint32_t stage_tx = (uint32_t)p.a_stage_bytes + (p.b_res ? 0u : (uint32_t)bn_cta * p.kbytes);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (p.dual) tma_prefetch_desc(&map_b1);
    tma_prefetch_desc(&map_out);
    if (p.res_mode == 1) tma_prefetch_desc(&map_res);
    for (int i = 0; i < stages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&bars->tmem_full[i], 1); mbar_init(&bars->tmem_empty[i], EPI_WARPS * CTAS); }
    mbar_init(&bars->b_full, 1);
    for (int i = 0; i < MAX_RES_BUFS; ++i) mbar_init(&bars->res_full[i], 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    if (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs of a pair: own A rows, own half of B) =====================
    int stage = 0;
    uint32_t phase = 0;
    if (CTAS == 1 && p.b_res && unit0 < num_units) {      // the CTA's weight rows, all K chunks, once
      const bool geglu = p.out_mode == OUT_U8_GEGLU;
      const int n_blk = unit0 % p.n_tiles;
      const int nrow0 = geglu ? n_blk * (p.block_n / 2) : n_blk * p.block_n;
      if (elect_one()) {
        mbar_expect_tx(&bars->b_full, (uint32_t)p.k_chunks * (uint32_t)p.block_n * p.kbytes);
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          tma_load_3d(smem_b + kc * p.b_stage_bytes, &map_b, &bars->b_full, kc * p.kbytes, 0, nrow0);
          if (geglu) tma_load_3d(smem_b + kc * p.b_stage_bytes + (p.block_n / 2) * p.kbytes, &map_b, &bars->b_full, kc * p.kbytes, 0, nrow0 + p.geglu_half);
        }
      }
      __syncwarp();
    }
    int ti_p = 0;
    for (int unit = unit0; unit < num_units; unit += unit_step, ++ti_p) {
      const int n_blk = unit % p.n_tiles, m_unit = unit / p.n_tiles;
      const int m0 = (m_unit * CTAS + (int)rank) * BM;
      const int b0 = m0 / p.HoWo;
      if (p.trace && lane == 0 && ti_p < 16) p.trace[((size_t)blockIdx.x * 16 + ti_p) * 8 + 5] = clock64();
      const int rem = m0 - b0 * p.HoWo;
      const int oh0 = rem / p.Wo, ow0 = rem - oh0 * p.Wo;
      // weight rows of this CTA: a plain tile takes block_n consecutive rows (each CTA of a pair its half); a GEGLU tile is
      // [block_n/2 value rows | the matching block_n/2 gate rows], which are N/2 apart in the projection
      const bool geglu = p.out_mode == OUT_U8_GEGLU;
      const int nrow0 = geglu ? n_blk * (p.block_n / 2) + ((CTAS == 2 && rank) ? p.geglu_half : 0) : n_blk * p.block_n + (int)rank * bn_cta;
      const int nrow1 = nrow0 + p.geglu_half;                          // second box (GEGLU on a single CTA)
      const int b_half_bytes = (p.block_n / 2) * p.kbytes;
      for (int range = 0; range <= p.dual; ++range) {
        const CUtensorMap* mb = range ? &map_b1 : &map_b;
        const int c_off = range ? p.a_c_offset1 : p.a_c_offset;
        const int kcs = range ? p.k_chunks1 : p.k_chunks;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int kh = tap / p.S, kw = tap - kh * p.S;
          for (int kc = 0; kc < kcs; ++kc) {
            mbar_wait(&bars->empty[stage], phase ^ 1);
            if (elect_one()) {
              if (CTAS == 2) {
                const uint32_t fb = mapa_shared(smem_u32(&bars->full[stage]), 0);
                if (rank == 0) mbar_expect_tx(&bars->full[stage], 2u * stage_tx);
                tma_load_4d_pair(smem_a + stage * p.a_stage_bytes, &map_a, fb, c_off + kc * p.kbytes, ow0 + kw, oh0 + kh, b0);
                tma_load_3d_pair(smem_b + stage * p.b_stage_bytes, mb, fb, kc * p.kbytes, tap, nrow0);
              } else {
                mbar_expect_tx(&bars->full[stage], stage_tx);
                tma_load_4d(smem_a + stage * p.a_stage_bytes, &map_a, &bars->full[stage], c_off + kc * p.kbytes, ow0 + kw, oh0 + kh, b0);
                if (!p.b_res) tma_load_3d(smem_b + stage * p.b_stage_bytes, mb, &bars->full[stage], kc * p.kbytes, tap, nrow0);
                if (geglu && !p.b_res) tma_load_3d(smem_b + stage * p.b_stage_bytes + b_half_bytes, mb, &bars->full[stage], kc * p.kbytes, tap, nrow1);
              }
            }
            __syncwarp();
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (p.trace && lane == 0 && ti_p < 16) p.trace[((size_t)blockIdx.x * 16 + ti_p) * 8 + 6] = clock64();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (rank == 0) {
      const uint32_t idesc = make_idesc_i8_m(BM * CTAS, p.block_n, /*A u8*/ 0, /*B s8*/ 1);
      const uint32_t a_base = smem_u32(smem_a), b_base = smem_u32(smem_b);
      const bool sw64 = p.kbytes == 64;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (CTAS == 1 && p.b_res && unit0 < num_units) { mbar_wait(&bars->b_full, 0); tc_fence_after(); }
      int ti_m = 0;
      for (int unit = unit0; unit < num_units; unit += unit_step, ++ti_m) {
        mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        if (p.trace && lane == 0 && ti_m < 16) p.trace[((size_t)blockIdx.x * 16 + ti_m) * 8 + 0] = clock64();
        uint32_t tmem_d = tmem_base + (uint32_t)acc * MAX_BN;
        int kc = 0, kcs = p.k_chunks, klast = p.k_last_mmas, it_r = 0;
        for (int it = 0; it < k_iters; ++it, ++it_r) {
          if (it == k_iters0 && p.dual) {         // second K range: its own accumulator, block_n columns further
            tmem_d += (uint32_t)p.block_n; kc = 0; kcs = p.k_chunks1; klast = p.k_last_mmas1; it_r = 0;
          }
          const int nmma = (kc == kcs - 1) ? klast : p.kbytes / UMMA_K;
          if (++kc == kcs) kc = 0;
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          if (p.trace && lane == 0 && it == 0 && ti_m < 16) p.trace[((size_t)blockIdx.x * 16 + ti_m) * 8 + 1] = clock64();
          const uint32_t aaddr = a_base + stage * p.a_stage_bytes, baddr = b_base + (p.b_res ? (kc == 0 ? kcs - 1 : kc - 1) : stage) * p.b_stage_bytes;
          const uint64_t adesc = sw64 ? make_smem_desc_sw64(aaddr) : make_smem_desc(aaddr);
          const uint64_t bdesc = sw64 ? make_smem_desc_sw64(baddr) : make_smem_desc(baddr);
          if (elect_one()) {
            if (CTAS == 2) {
              umma_i8_pair(tmem_d, adesc, bdesc, idesc, it_r ? 1u : 0u);
              if (nmma > 1) umma_i8_pair(tmem_d, adesc + 2, bdesc + 2, idesc, 1u);
              if (nmma > 2) umma_i8_pair(tmem_d, adesc + 4, bdesc + 4, idesc, 1u);
              if (nmma > 3) umma_i8_pair(tmem_d, adesc + 6, bdesc + 6, idesc, 1u);
              umma_commit_pair(&bars->empty[stage], 3);
            } else {
              umma_i8(tmem_d, adesc, bdesc, idesc, it_r ? 1u : 0u);
              if (nmma > 1) umma_i8(tmem_d, adesc + 2, bdesc + 2, idesc, 1u);
              if (nmma > 2) umma_i8(tmem_d, adesc + 4, bdesc + 4, idesc, 1u);
              if (nmma > 3) umma_i8(tmem_d, adesc + 6, bdesc + 6, idesc, 1u);
              umma_commit(&bars->empty[stage]);
            }
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) {
          if (CTAS == 2) umma_commit_pair(&bars->tmem_full[acc], 3);
          else umma_commit(&bars->tmem_full[acc]);
        }
        __syncwarp();
        if (p.trace && lane == 0 && ti_m < 16) p.trace[((size_t)blockIdx.x * 16 + ti_m) * 8 + 2] = clock64();
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;             // which 16 columns of a 32-column chunk
    const int r = quarter * 32 + lane;            // row inside the tile
    const int et = threadIdx.x - 64;              // 0..255
    const bool issuer = et == 0;
    const float da = __ldg(p.delta_a);
    const int za = (int)__ldg(p.zp_a);
    const float da1 = p.dual ? __ldg(p.delta_a1) : 0.f;
    const int za1 = p.dual ? (int)__ldg(p.zp_a1) : 0;
    const int mode = p.out_mode;
    const bool has_cw = p.cw != nullptr;
    float qd = 1.f, qinv = 1.f, qz = 0.f;
    if (mode >= OUT_U8_ROWS) { qd = __ldg(p.q_delta); qinv = 1.0f / qd; qz = __ldg(p.q_zp); }
    // where this thread's 16 values of a chunk live in the staging buffer (same geometry for the residual buffer)
    uint32_t st_off[4];                           // byte offsets of the four float4 (row-major) ...
    uint32_t nchw_off = 0, nchw_cstride = 0;      // ... or base + column stride (NCHW)
    if (mode == OUT_F32_ROWS) {
#pragma unroll
      for (int k = 0; k < 4; ++k) st_off[k] = (uint32_t)r * 128u + ((uint32_t)((4 * half + k) ^ (r & 7)) << 4);
    } else if (mode == OUT_F32_NCHW) {
      const int img = r >> p.px_shift, px = r & (p.pxb - 1);
      nchw_cstride = (uint32_t)p.pxb * 4u;
      nchw_off = ((uint32_t)(img * CHUNK + 16 * half) * p.pxb + px) * 4u;
    }
    const int n_chunks = (mode == OUT_U8_GEGLU ? p.block_n / 2 : p.block_n) / CHUNK;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t gchunk = 0;                          // running chunk counter: staging buffer = gchunk & 1
    uint32_t tile_par = 0;                        // code outputs: staging buffer of the current tile
    uint32_t rchunk = 0;                          // running residual chunk counter: buffer = rchunk % res_bufs
    uint32_t rphase_bits = 0;                     // phase bit per residual buffer
    int last_n_blk = -1;
    int ti_e = 0;
    for (int unit = unit0; unit < num_units; unit += unit_step) {
      const int n_blk = unit % p.n_tiles, m_unit = unit / p.n_tiles;
      const int n0 = n_blk * p.block_n;
      const int m0 = (m_unit * CTAS + (int)rank) * BM;
      // per-column epilogue constants of this tile (the previous tile's readers are past their last barrier); a CTA that keeps
      // its N block for the whole launch (b_res) loads them once instead of paying their global-load latency on every tile
      if (n_blk != last_n_blk)
      for (int j = et; j < p.block_n; j += EPI_WARPS * 32) {
        int n = n0 + j;
        if (mode == OUT_U8_GEGLU) n = (j < p.block_n / 2) ? n_blk * (p.block_n / 2) + j : p.geglu_half + n_blk * (p.block_n / 2) + (j - p.block_n / 2);
        const bool ok = (mode == OUT_U8_GEGLU) ? (n < p.N) : (n < p.N);
        epi_scale[j] = ok ? da * __ldg(p.delta_w + n) : 0.f;
        epi_zterm[j] = ok ? -za * __ldg(p.wsum_eff + n) : 0;
        epi_cw[j] = (ok && has_cw) ? __ldg(p.cw + n) : 0;
        epi_bias[j] = (ok && p.bias) ? __ldg(p.bias + n) : 0.f;
        if (p.dual) {                             // second range: constants in the upper half of the vectors (block_n <= 128)
          epi_scale[MAX_BN / 2 + j] = ok ? da1 * __ldg(p.delta_w1 + n) : 0.f;
          epi_zterm[MAX_BN / 2 + j] = ok ? -za1 * __ldg(p.wsum_eff1 + n) : 0;
          epi_bias[MAX_BN / 2 + j] = 0.f;
        }
      }
      last_n_blk = n_blk;
      // coordinates of this tile in the output / residual maps
      int co0, co1, co2;                          // NCHW: (pixel, channel, image); rows: (column, row)
      if (mode == OUT_F32_NCHW) {
        const int img0 = m0 / p.out_hw;
        co0 = m0 - img0 * p.out_hw; co1 = n0; co2 = img0;
      } else {
        co0 = (mode == OUT_U8_GEGLU) ? n_blk * (p.block_n / 2) : n0; co1 = m0; co2 = 0;
      }
      auto issue_res = [&](int ci) {
        const int rb = (int)((rchunk + ci) % (uint32_t)p.res_bufs);
        mbar_expect_tx(&bars->res_full[rb], OUT_BUF_BYTES);
        if (mode == OUT_F32_NCHW) tma_load_3d(res_buf + rb * OUT_BUF_BYTES, &map_res, &bars->res_full[rb], co0, co1 + ci * CHUNK, co2);
        else tma_load_2d(res_buf + rb * OUT_BUF_BYTES, &map_res, &bars->res_full[rb], co0 + ci * CHUNK, co1);
      };
      if (p.res_mode == 1 && issuer) {
        for (int ci = 0; ci < p.res_bufs && ci < n_chunks; ++ci) issue_res(ci);
      }
      // code outputs: the whole tile is staged in one buffer (alternating per tile); the store that used it two tiles ago has drained
      const int wout = n_chunks * CHUNK;
      uint8_t* tile_buf = out_buf + (tile_par & 1u) * OUT_BUF_BYTES;
      if (mode >= OUT_U8_ROWS && issuer) bulk_wait_read<1>();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int m = m0 + r;
      const int rs = (has_cw && m < p.M) ? __ldg(p.rowsum + m) : 0;
      const float* bimg = nullptr;
      if (p.res_mode == 2) { const int img = (m < p.M ? m : p.M - 1) / p.out_hw; bimg = p.bias_img + (size_t)img * p.N + n0; }
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * MAX_BN;
      int rowsum_codes = 0;

      mbar_wait(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();
      if (p.trace && et == 0 && ti_e < 16) p.trace[((size_t)blockIdx.x * 16 + ti_e) * 8 + 3] = clock64();
      // (software-pipelining these TMEM loads against the conversion of the previous chunk was measured and is slower: the epilogue
      // is bound by its instruction count, not by the TMEM latency -- profiles/gemm2_timeline_r02.txt)
      for (int ci = 0; ci < n_chunks; ++ci, ++gchunk) {
        const int c0 = ci * CHUNK + 16 * half;    // first of this thread's 16 columns inside the tile
        uint32_t a[16];
        float v[16];
        tmem_ld16(taddr + c0, a);
        uint32_t g[16];
        if (mode == OUT_U8_GEGLU) tmem_ld16(taddr + p.block_n / 2 + c0, g);
        else if (p.dual) tmem_ld16(taddr + p.block_n + c0, g);
        tmem_ld_wait();
        if (ci == n_chunks - 1) {                 // accumulator fully read: hand the TMEM stage back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CTAS == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&bars->tmem_empty[acc]), 0));
            else mbar_arrive(&bars->tmem_empty[acc]);
          }
          if (p.trace && et == 0 && ti_e < 16) p.trace[((size_t)blockIdx.x * 16 + ti_e) * 8 + 7] = clock64();
        }
        // dequantise: exact int32 zero-point fold, one fp32 FMA; the per-column constants come as 128-bit broadcast loads
        dequant16(a, v, epi_zterm + c0, epi_scale + c0, epi_bias + c0, has_cw ? epi_cw + c0 : nullptr, rs);
        if (p.dual) {                             // + the second range's rounded product, as the accumulating second launch adds it
          float v1[16];
          dequant16(g, v1, epi_zterm + MAX_BN / 2 + c0, epi_scale + MAX_BN / 2 + c0, epi_bias + MAX_BN / 2 + c0, nullptr, 0);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = v1[j] + v[j];
        }
        uint8_t* ob = out_buf + (gchunk & 1u) * OUT_BUF_BYTES;
        if (p.res_mode == 1) {
          const int rb = (int)((rchunk + ci) % (uint32_t)p.res_bufs);
          mbar_wait(&bars->res_full[rb], (rphase_bits >> rb) & 1u);
          rphase_bits ^= 1u << rb;
          const uint8_t* rbuf = res_buf + rb * OUT_BUF_BYTES;
          if (mode == OUT_F32_ROWS) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float4 t = *reinterpret_cast<const float4*>(rbuf + st_off[k]);
              v[4 * k + 0] += t.x; v[4 * k + 1] += t.y; v[4 * k + 2] += t.z; v[4 * k + 3] += t.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += *reinterpret_cast<const float*>(rbuf + nchw_off + j * nchw_cstride);
          }
        } else if (p.res_mode == 2) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += (n0 + c0 + j < p.N) ? __ldg(bimg + c0 + j) : 0.f;
        }
        uint32_t codes[4];
        if (mode >= OUT_U8_ROWS) {
          if (mode == OUT_U8_GEGLU) {
            const int cg = p.block_n / 2 + c0;
            float gate[16];
            dequant16(g, gate, epi_zterm + cg, epi_scale + cg, epi_bias + cg, has_cw ? epi_cw + cg : nullptr, rs);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = v[j] * gelu_erf(gate[j]);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            codes[k] = q_code(v[4 * k + 0], qd, qinv, qz, p.q_max) | (q_code(v[4 * k + 1], qd, qinv, qz, p.q_max) << 8) |
                       (q_code(v[4 * k + 2], qd, qinv, qz, p.q_max) << 16) | (q_code(v[4 * k + 3], qd, qinv, qz, p.q_max) << 24);
          if (p.q_rowsum) {
            // columns past N quantize a 0 to the zero-point code: leave them out of the sum
            const int nout = (mode == OUT_U8_GEGLU) ? p.geglu_half : p.N;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int cbase = co0 + ci * CHUNK + 16 * half + 4 * k;
              if (cbase + 4 <= nout) rowsum_codes += __dp4a(codes[k], 0x01010101u, 0u);
              else for (int b = 0; b < 4; ++b) if (cbase + b < nout) rowsum_codes += (codes[k] >> (8 * b)) & 0xff;
            }
          }
        }
        if (mode >= OUT_U8_ROWS) {                // tile-wide staging: no per-chunk synchronisation, one store after the loop
          *reinterpret_cast<uint4*>(tile_buf + r * wout + ci * CHUNK + 16 * half) = make_uint4(codes[0], codes[1], codes[2], codes[3]);
          continue;
        }
        if (issuer) bulk_wait_read<1>();          // the store issued two chunks ago has drained its staging buffer
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (mode == OUT_F32_ROWS) {
#pragma unroll
          for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(ob + st_off[k]) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        } else if (mode == OUT_F32_NCHW) {
#pragma unroll
          for (int j = 0; j < 16; ++j) *reinterpret_cast<float*>(ob + nchw_off + j * nchw_cstride) = v[j];
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (issuer) {
          if (mode == OUT_F32_NCHW) tma_store_3d(&map_out, ob, co0, co1 + ci * CHUNK, co2);
          else tma_store_2d(&map_out, ob, co0 + ci * CHUNK, co1);
          bulk_commit();
          if (p.res_mode == 1 && ci + p.res_bufs < n_chunks) issue_res(ci + p.res_bufs);
        }
      }
      if (mode >= OUT_U8_ROWS) {
        fence_proxy_async();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (issuer) { tma_store_2d(&map_out, tile_buf, co0, co1); bulk_commit(); }
        ++tile_par;
      }
      if (p.res_mode == 1) rchunk += (uint32_t)n_chunks;
      if (p.q_rowsum && mode >= OUT_U8_ROWS) {
        // the two warps of a quarter hold the two 16-column halves of every chunk of this row
        if (m < p.M && rowsum_codes) atomicAdd(p.q_rowsum + m, rowsum_codes);
      }
      if (p.trace && et == 0 && ti_e < 16) p.trace[((size_t)blockIdx.x * 16 + ti_e) * 8 + 4] = clock64();
      ++ti_e;
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    if (issuer) bulk_wait<0>();                   // all stores complete before the CTA (and its shared memory) goes away
  }

  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CTAS == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// N tile (multiple of 32, <= 256): fewest waves first, then the widest tile (per-MMA cost is nearly flat below N = 192,
// profiles/int8_peak_r02.txt), i.e. the same wave/bytes model as the first-generation kernel
static int pick_block_n(int N, int m_units, int workers, int step, int max_bn = MAX_BN) {
  int best = 0;
  long long best_cost = 0;
  for (int bn = step; bn <= max_bn; bn += step) {
    const long long tiles = (long long)m_units * ((N + bn - 1) / bn);
    const long long waves = (tiles + workers - 1) / workers;
    const long long cost = waves * (bn + 192);
    if (best == 0 || cost <= best_cost) { best = bn; best_cost = cost; }
    if (bn >= N) break;
  }
  return best;
}

}  // namespace g2

struct Gemm2Args {
  const uint8_t* q; int B, Hp, Wp, Cp_act, a_c_offset;
  const void* wq; int N, Np, R, S, Cp_w;
  const float* delta_a; const float* zp_a; const float* delta_w; const int32_t* wsum_eff; const int32_t* cw; const int32_t* rowsum;
  const float* bias; const float* bias_img; const float* residual;
  void* out; int out_hw; int accumulate;
  int out_mode;                 // g2::OutMode
  const float* q_delta; const float* q_zp; int q_levels; int32_t* q_rowsum; int out_pitch;   // code-emitting modes: consumer quantizer, u8 row pitch
  // optional second K range (split shortcut): channels [a_c_offset1, ...) of q against a second weight pack
  const void* wq1 = nullptr; int Cp_w1 = 0; int a_c_offset1 = 0;
  const float* delta_a1 = nullptr; const float* zp_a1 = nullptr; const float* delta_w1 = nullptr; const int32_t* wsum_eff1 = nullptr;
};

long long* gemm_trace_buffer();                   // qgemm_sm100.cu (edadm_debug_set_gemm_trace)

// returns EDADM_OK, an error, or +1 when this kernel does not cover the case (the caller falls back to the first-generation kernel)
int launch_qgemm2(const Gemm2Args& a, void* stream) {
  using namespace g2;
  const int Ho = a.Hp - a.R + 1, Wo = a.Wp - a.S + 1;
  const long long M = (long long)a.B * Ho * Wo;
  if (M > 0x7fffffffLL || M < 1) return 1;
  const bool codes_out = a.out_mode >= OUT_U8_ROWS;
  const bool dual = a.wq1 != nullptr;
  if (dual && (codes_out || a.cw || a.accumulate || a.residual || a.bias_img || !a.delta_a1 || !a.zp_a1 || !a.delta_w1 || !a.wsum_eff1 || (a.Cp_w1 & 15) || a.Cp_w1 < 16 ||
               (((uintptr_t)a.wq1) & 15))) return 1;
  if (a.accumulate && (a.residual || a.bias_img || codes_out)) return 1;
  if (a.residual && a.bias_img) return 1;
  if ((((uintptr_t)a.out) & 15) || (a.residual && (((uintptr_t)a.residual) & 15))) return 1;
  // tile box over (W, H, B): 128 consecutive output pixels
  int box_w, box_h, box_b;
  const bool flat = (Ho == 1 && a.B == 1);
  if (flat) { box_w = BM; box_h = 1; box_b = 1; }
  else if (Wo >= BM) { if (Wo % BM) return 1; box_w = BM; box_h = 1; box_b = 1; }
  else {
    if (BM % Wo) return 1;
    box_w = Wo;
    const int rows = BM / Wo;
    if (rows <= Ho) { if (Ho % rows) return 1; box_h = rows; box_b = 1; }
    else { if (rows % Ho) return 1; box_h = Ho; box_b = rows / Ho; }
  }
  int mode = a.out_mode;
  int pxb = 0, px_shift = 0;
  if (!codes_out) {
    if (a.out_hw == 1) {
      mode = OUT_F32_ROWS;
      if (a.N % 4) return 1;
    } else {
      mode = OUT_F32_NCHW;
      if (a.out_hw % 4 || M % a.out_hw) return 1;
      pxb = a.out_hw >= BM ? BM : a.out_hw;
      if (a.out_hw >= BM ? (a.out_hw % BM) : (BM % a.out_hw)) return 1;
      if (pxb & (pxb - 1)) return 1;
      while ((1 << px_shift) < pxb) ++px_shift;
      if (BM / pxb > 256) return 1;
    }
  } else {
    if (a.out_hw != 1 || a.out_pitch % 16 || a.q_levels > 256 || a.q_levels < 2 || !a.q_delta || !a.q_zp) return 1;
    if (mode == OUT_U8_GEGLU && (a.N % 2 || (a.N / 2) % CHUNK)) return 1;
  }

  const int m_tiles = (int)((M + BM - 1) / BM);
  // CTA pairs when there is enough work to keep 74 pairs busy for at least two rounds AND the main loop is long enough for the
  // halved weight traffic to matter: with a handful of K steps per tile (the transformer linears, K = 384 .. 960) the tile time is
  // its epilogue and the pair's cluster handshakes only cost (measured: 131072 x 3072 x 384 GEGLU 588 -> 533 us with single CTAs)
  const int k_steps_128 = a.R * a.S * ((a.Cp_w + 127) / 128 + (dual ? (a.Cp_w1 + 127) / 128 : 0));
  int ctas = (m_tiles >= 2 * 148 && k_steps_128 >= 8) ? 2 : 1;
  if (const char* e = getenv("EDADM_GEMM_CTAS")) { const int v = atoi(e); if (v == 1 || v == 2) ctas = v; }
  const int sms = sm_count();
  const int workers = ctas == 2 ? sms / 2 : sms;
  const int m_units = (m_tiles + ctas - 1) / ctas;
  int block_n;
  if (mode == OUT_U8_GEGLU) {
    const int h = a.N / 2;      // output columns; a tile covers bh of them (value rows + gate rows = 2 bh weight rows)
    const int bh = (h % 128 == 0) ? 128 : (h % 96 == 0) ? 96 : (h % 64 == 0) ? 64 : 32;
    block_n = 2 * bh;
  } else {
    // code outputs stage a whole tile (128 rows x block_n bytes <= one 16 KB buffer) and leave with ONE TMA store per tile
    block_n = pick_block_n(a.N, m_units, workers, 32, (mode == OUT_U8_ROWS || dual) ? 128 : MAX_BN);      // dual: two accumulators per tile
  }
  const int n_tiles = mode == OUT_U8_GEGLU ? (a.N / 2) / (block_n / 2) : (a.N + block_n - 1) / block_n;
  if (a.Np < a.N) return fail(EDADM_ERR_ARG, "qgemm2: weight rows Np=%d < N=%d", a.Np, a.N);

  int kbytes = 128;
  if (const char* e = getenv("EDADM_GEMM_KBYTES")) { const int v = atoi(e); if (v == 64 || v == 128) kbytes = v; }
  const CUtensorMapSwizzle ksw = kbytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUtensorMap map_a, map_b, map_out, map_res;
  {
    cuuint64_t dims[4] = {(cuuint64_t)a.Cp_act, (cuuint64_t)a.Wp, (cuuint64_t)a.Hp, (cuuint64_t)a.B};
    cuuint64_t strides[3] = {(cuuint64_t)a.Cp_act, (cuuint64_t)a.Wp * a.Cp_act, (cuuint64_t)a.Hp * a.Wp * a.Cp_act};
    cuuint32_t box[4] = {(cuuint32_t)kbytes, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_b};
    int rc = encode_map(&map_a, a.q, 4, dims, strides, box, "activations", CU_TENSOR_MAP_DATA_TYPE_UINT8, ksw);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)a.Cp_w, (cuuint64_t)(a.R * a.S), (cuuint64_t)a.Np};
    cuuint64_t strides[2] = {(cuuint64_t)a.Cp_w, (cuuint64_t)a.R * a.S * a.Cp_w};
    cuuint32_t box[3] = {(cuuint32_t)kbytes, 1u, (cuuint32_t)(mode == OUT_U8_GEGLU ? block_n / 2 : block_n / ctas)};
    int rc = encode_map(&map_b, a.wq, 3, dims, strides, box, "weights", CU_TENSOR_MAP_DATA_TYPE_UINT8, ksw);
    if (rc) return rc;
  }
  CUtensorMap map_b1 = map_b;
  if (dual) {
    cuuint64_t dims[3] = {(cuuint64_t)a.Cp_w1, (cuuint64_t)(a.R * a.S), (cuuint64_t)a.Np};
    cuuint64_t strides[2] = {(cuuint64_t)a.Cp_w1, (cuuint64_t)a.R * a.S * a.Cp_w1};
    cuuint32_t box[3] = {(cuuint32_t)kbytes, 1u, (cuuint32_t)(block_n / ctas)};
    int rc = encode_map(&map_b1, a.wq1, 3, dims, strides, box, "weights (second range)", CU_TENSOR_MAP_DATA_TYPE_UINT8, ksw);
    if (rc) return rc;
  }
  const float* res_src = a.accumulate ? (const float*)a.out : a.residual;
  if (mode == OUT_F32_NCHW) {
    const long long imgs = M / a.out_hw;
    cuuint64_t dims[3] = {(cuuint64_t)a.out_hw, (cuuint64_t)a.N, (cuuint64_t)imgs};
    cuuint64_t strides[2] = {(cuuint64_t)a.out_hw * 4, (cuuint64_t)a.N * a.out_hw * 4};
    cuuint32_t box[3] = {(cuuint32_t)pxb, (cuuint32_t)CHUNK, (cuuint32_t)(BM / pxb)};
    int rc = encode_map(&map_out, a.out, 3, dims, strides, box, "output", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    map_res = map_out;
    if (res_src) { rc = encode_map(&map_res, res_src, 3, dims, strides, box, "residual", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_NONE); if (rc) return rc; }
  } else if (mode == OUT_F32_ROWS) {
    cuuint64_t dims[2] = {(cuuint64_t)a.N, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)a.N * 4};
    cuuint32_t box[2] = {(cuuint32_t)CHUNK, (cuuint32_t)BM};
    int rc = encode_map(&map_out, a.out, 2, dims, strides, box, "output", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    map_res = map_out;
    if (res_src) { rc = encode_map(&map_res, res_src, 2, dims, strides, box, "residual", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B); if (rc) return rc; }
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)a.out_pitch, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)a.out_pitch};
    cuuint32_t box[2] = {(cuuint32_t)(mode == OUT_U8_GEGLU ? block_n / 2 : block_n), (cuuint32_t)BM};
    int rc = encode_map(&map_out, a.out, 2, dims, strides, box, "codes", CU_TENSOR_MAP_DATA_TYPE_UINT8, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    map_res = map_out;
  }

  Params p;
  memset(&p, 0, sizeof(p));
  p.M = (int)M; p.N = a.N; p.taps = a.R * a.S; p.S = a.S;
  p.kbytes = kbytes;
  p.k_chunks = (a.Cp_w + kbytes - 1) / kbytes;
  const int last_bytes = a.Cp_w - (p.k_chunks - 1) * kbytes;
  p.k_last_mmas = (last_bytes + UMMA_K - 1) / UMMA_K;
  p.a_c_offset = a.a_c_offset;
  if (dual) {
    p.dual = 1;
    p.k_chunks1 = (a.Cp_w1 + kbytes - 1) / kbytes;
    p.k_last_mmas1 = (a.Cp_w1 - (p.k_chunks1 - 1) * kbytes + UMMA_K - 1) / UMMA_K;
    p.a_c_offset1 = a.a_c_offset1;
    p.delta_a1 = a.delta_a1; p.zp_a1 = a.zp_a1; p.delta_w1 = a.delta_w1; p.wsum_eff1 = a.wsum_eff1;
  }
  p.Wo = flat ? (1 << 30) : Wo;
  p.HoWo = flat ? (1 << 30) : Ho * Wo;
  p.block_n = block_n; p.n_tiles = n_tiles; p.m_units = m_units;
  p.a_stage_bytes = BM * kbytes;
  p.b_stage_bytes = ((block_n / ctas) * kbytes + 1023) & ~1023;
  p.out_mode = mode; p.out_hw = a.out_hw; p.pxb = pxb; p.px_shift = px_shift;
  p.res_mode = res_src ? 1 : (a.bias_img ? 2 : 0);
  p.geglu_half = a.N / 2;
  p.bias_img = a.bias_img;
  p.delta_a = a.delta_a; p.zp_a = a.zp_a; p.delta_w = a.delta_w; p.wsum_eff = a.wsum_eff; p.cw = a.cw; p.rowsum = a.rowsum; p.bias = a.bias;
  p.q_delta = a.q_delta; p.q_zp = a.q_zp; p.q_max = (float)(a.q_levels - 1); p.q_rowsum = a.q_rowsum;
  p.trace = gemm_trace_buffer();
  const int stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
  const int fixed_no_res = 1024 + 2 * OUT_BUF_BYTES + EPI_VEC_BYTES + 1024;
  p.res_bufs = 0;
  if (p.res_mode == 1) {
    p.res_bufs = (SMEM_LIMIT - fixed_no_res - 3 * OUT_BUF_BYTES) / stage_bytes >= 3 ? 3 : 2;
  }
  const int fixed = fixed_no_res + p.res_bufs * OUT_BUF_BYTES;
  p.stages = (SMEM_LIMIT - fixed) / stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  if (p.stages < 2) return 1;
  int smem_bytes = fixed + p.stages * stage_bytes;
  // weight-resident schedule for the short-K linears: every CTA keeps the K chunks of ONE N block and streams activation tiles
  int grid_res = 0;
  if (ctas == 1 && !dual && a.R * a.S == 1 && p.k_chunks <= 4 && n_tiles <= sms && m_units >= 4 * (sms / n_tiles) && getenv("EDADM_GEMM_NO_BRES") == nullptr) {
    const int a_stages = std::min(MAX_STAGES, (SMEM_LIMIT - fixed - p.k_chunks * p.b_stage_bytes) / p.a_stage_bytes);
    if (a_stages >= p.k_chunks + 2) {               // more than one tile of activations in flight
      p.b_res = 1;
      p.stages = a_stages;
      smem_bytes = fixed + p.k_chunks * p.b_stage_bytes + p.stages * p.a_stage_bytes;
      grid_res = n_tiles * (sms / n_tiles);
    }
  }

  static int attr_dev_mask[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_dev_mask[dev]) {
    cudaError_t e = cudaFuncSetAttribute(qgemm2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(qgemm2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return fail(EDADM_ERR_CUDA, "qgemm2: cannot opt in to %d B shared memory: %s", SMEM_LIMIT, cudaGetErrorString(e));
    attr_dev_mask[dev] = 1;
  }
  const int units = m_units * n_tiles;
  if (ctas == 2) {
    const int pairs = units < workers ? units : workers;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, qgemm2_kernel<2>, map_a, map_b, map_out, map_res, map_b1, p);
    if (e != cudaSuccess) return fail(EDADM_ERR_CUDA, "qgemm2 (pairs): %s", cudaGetErrorString(e));
  } else {
    const int grid = grid_res ? grid_res : (units < sms ? units : sms);
    qgemm2_kernel<1><<<grid, THREADS, smem_bytes, (cudaStream_t)stream>>>(map_a, map_b, map_out, map_res, map_b1, p);
  }
  return check_launch("qgemm2");
}

}  // namespace edadm

// QuantModule linear whose only consumer is the activation quantizer of the NEXT QuantModule (qdiff/quant_layer.py:414-422 of the
// consumer): instead of fp32 outputs the epilogue emits that quantizer's u8 codes, out_codes [M][out_pitch].
//   geglu = 0: codes of  y[m][n]                      (N columns)            -- e.g. to_q / to_k feeding the attention quantizers
//   geglu = 1: codes of  y[m][n] * gelu(y[m][N/2+n])  (N/2 columns)          -- GEGLU.forward (ldm/modules/attention.py:37-44)
// with y = delta_a*delta_w[n]*(acc + cw[n]*rowsum[m] - zp_a*wsum_eff[n]) + bias[n] exactly as edadm_qgemm_i8 computes it.
// q_rowsum (nullable, int32 [M], must be zeroed by the caller): += sum of the emitted codes of each row.
extern "C" int edadm_qgemm_i8_codes(const uint8_t* q, int64_t M, int Kp_act, const int8_t* wq, int N, int Np, int Cp_w,
                                    const float* delta_a, const float* zp_a, const float* delta_w, const int32_t* wsum_eff,
                                    const int32_t* cw, const int32_t* rowsum, const float* bias, int geglu, const float* q_delta,
                                    const float* q_zp, int q_levels, uint8_t* out_codes, int out_pitch, int32_t* q_rowsum,
                                    void* stream) {
  using namespace edadm;
  if (!q || !wq || !delta_a || !zp_a || !delta_w || !wsum_eff || !q_delta || !q_zp || !out_codes) return fail(EDADM_ERR_ARG, "qgemm_i8_codes: null pointer");
  if (cw && !rowsum) return fail(EDADM_ERR_ARG, "qgemm_i8_codes: cw given without rowsum");
  if (M < 1 || M > 0x7fffffffLL || N < 1 || (Kp_act & 15) || (Cp_w & 15) || Cp_w < 16 || (out_pitch & 15) || q_levels < 2 || q_levels > 256)
    return fail(EDADM_ERR_ARG, "qgemm_i8_codes: bad sizes M=%lld N=%d Kp=%d Cp_w=%d pitch=%d levels=%d", (long long)M, N, Kp_act, Cp_w, out_pitch, q_levels);
  if ((((uintptr_t)q) | ((uintptr_t)wq) | ((uintptr_t)out_codes)) & 15) return fail(EDADM_ERR_ARG, "qgemm_i8_codes: operands must be 16-byte aligned");
  const int n_out = geglu ? N / 2 : N;
  if (geglu && ((N & 1) || (n_out % 32))) return fail(EDADM_ERR_UNSUPPORTED, "qgemm_i8_codes: GEGLU needs N/2 to be a multiple of 32 (N=%d)", N);
  if (out_pitch < n_out) return fail(EDADM_ERR_ARG, "qgemm_i8_codes: out_pitch %d < %d output columns", out_pitch, n_out);
  Gemm2Args a{q, 1, 1, (int)M, Kp_act, 0, wq, N, Np, 1, 1, Cp_w, delta_a, zp_a, delta_w, wsum_eff, cw, rowsum, bias, nullptr, nullptr,
              out_codes, 1, 0, geglu ? g2::OUT_U8_GEGLU : g2::OUT_U8_ROWS, q_delta, q_zp, q_levels, q_rowsum, out_pitch};
  const int rc = launch_qgemm2(a, stream);
  if (rc > 0) return fail(EDADM_ERR_UNSUPPORTED, "qgemm_i8_codes: shape not covered");
  return rc;
}
