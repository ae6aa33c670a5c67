// Fused quantized attention for sm_100a (replaces the bmm / softmax / fake-quant chain of
// QuantAttnBlock.forward quant_block.py:431-445, cross_attn_forward :214-233 of the reference).
//
// Inputs are the u8 CODES of q, k (token-major [BH][T][dp]) and v (channel-major [BH][d][Tkp]) plus their per-token /
// per-channel code sums; the kernel never sees fp32 q/k/v and the T x T probability matrix never leaves the SM:
//
//   S[t][s]  = dq*dk*scale * sum_c (q[t][c]-zq)(k[s][c]-zk)            int8 tcgen05 MMA into TMEM, exact
//   P        = softmax_s(S)                                             fp32, two passes over the keys (row max / sum
//                                                                       first, then the normalised probabilities)
//   Pq[t][s] = clamp(rint(P/dP) + zP, 0, L-1)                           quantized in registers -> 128B-swizzled smem
//   O[t][c]  = dP*dv * sum_s (Pq[t][s]-zP)(v[c][s]-zv)                  second tcgen05 MMA, accumulator in TMEM
//
// Zero-points are folded with the code sums (rq, rk, rv, and the running row sum of Pq), so operands stay raw u8.
// One CTA = 128 queries x one chunk (<=256) of the head dim of one (batch, head).  Warp roles: 0 TMA producer,
// 1 MMA issuer, 2..17 softmax + epilogue (four threads per query row == TMEM lane, 32 key columns each).
#include "tc05.cuh"

namespace edadm {

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// d = (c << 16) | (sat_u8(a) << 8) | sat_u8(b)
constexpr int MAGIC_I = 0x4B400000;          // bit pattern of 12582912.0f = 1.5 * 2^23
constexpr float MAGIC_F = 12582912.0f;

// exact (float)v for a biased integer b = v + BIAS (BIAS = MAGIC_I when SMALL, else 0)
template <bool SMALL>
__device__ __forceinline__ float biased_to_float(int b) {
  return SMALL ? (__int_as_float(b) - MAGIC_F) : (float)b;
}

// rint(t) for 0 <= t < 2^22 without the XU pipe: the add rounds to nearest-even, the low mantissa bits are the integer
__device__ __forceinline__ int rint_magic(float t) { return __float_as_int(t + MAGIC_F) - MAGIC_I; }

__device__ __forceinline__ uint32_t pack_sat_u8(int a, int b, uint32_t c) {
  uint32_t d;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

constexpr int ATT_M = 128;
constexpr int ATT_S = 128;
constexpr int ATT_KB = 128;
constexpr int QK_STAGES = 8;                    // upper bound of the K (or Q+K) smem ring; the launch picks what fits
constexpr int V_STAGES = 2;                     // upper bound; 1 when a V tile is larger than 32 KB
constexpr int ATT_TILE_BYTES = 128 * 128;       // 16 KB: Q chunk, K chunk, P tile
constexpr int V_TILE_BYTES = 256 * 128;         // 32 KB
constexpr int SM_PARTS = 4;                 // softmax threads per query row
constexpr int SM_COLS = ATT_S / SM_PARTS;   // key columns of a tile per softmax thread
constexpr int SM_WARPS = 4 * SM_PARTS;
constexpr int ATT_THREADS = 64 + 32 * SM_WARPS;   // TMA warp + MMA warp + 16 softmax warps
constexpr int ATT_TMEM_COLS = 512;              // S0 [0,128) S1 [128,256) O [256,512)

struct AttnParams {
  int Tq, Tk, d;
  int k_chunks, k_last_mmas, s_tiles;
  int d_chunk, heads;
  int s_bufs;                // S accumulators in TMEM: 2 (d_chunk <= 256) or 1 (d_chunk <= 384, O takes columns [128, 512))
  int qk_stages;             // K (q_resident) or Q+K smem ring depth
  int q_resident;            // 1: the CTA's Q tile (k_chunks x 16 KB) is loaded once and stays in smem; 0: Q chunks ride with the K chunks
  int v_stages;              // V smem ring depth (2, or 1 for tiles above 32 KB)
  int q_bytes;               // smem reserved for Q: k_chunks x 16 KB (resident) or qk_stages x 16 KB (streamed)
  int v_halves;              // 1, or 2 when d_chunk > 256: V tile loaded and multiplied as two halves (TMA box / UMMA N <= 256)
  int v_stage_bytes;
  long long o_sb, o_sh, o_st, o_sc;
  float sm_scale;
  const float *dq, *zq, *dk, *zk, *dv, *zv, *dpq, *zpq;
  int p_levels;
  const int32_t *rq, *rk, *rv;
  float* out;
};

struct __align__(8) AttnBarriers {
  uint64_t qk_full[QK_STAGES], qk_empty[QK_STAGES];
  uint64_t v_full[V_STAGES], v_empty[V_STAGES];
  uint64_t s_full[2], s_empty[2];
  uint64_t p_full[2], p_empty[2];
  uint64_t o_full, q_full;
  uint32_t tmem_base;
};

constexpr int ATT_MAX_KEYS = 4096;               // per-key zero-point terms of a whole (batch, head) stay in shared memory
constexpr int ATT_SMEM_BYTES = 220 * 1024;   // dynamic opt-in ceiling (the kernel also has 6 KB of static shared memory); each launch asks for what its tiles need

// SMALL_V: |raw - zq*rk| < 2^22 (head dim <= 64), so int -> float goes through the exact magic-number add instead of I2F.
// The softmax is bound by the 16-op/clk XU pipe (ex2, I2F, F2I); with the conversions moved to the ALU/FMA pipes only the
// two ex2 per score remain there.
template <bool SMALL_V>
__global__ void __launch_bounds__(ATT_THREADS, 1)
qattn_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
             const __grid_constant__ CUtensorMap map_v, AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer (LDS/STS)
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem_q + p.q_bytes;
  uint8_t* smem_v = smem_k + p.qk_stages * ATT_TILE_BYTES;
  uint8_t* smem_p = smem_v + p.v_stages * p.v_stage_bytes;
  AttnBarriers* bars = reinterpret_cast<AttnBarriers*>(smem_p + 2 * ATT_TILE_BYTES);
  int* colint = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [s_tiles*128]: bias - zq*rk[s]

  __shared__ float stat_m[SM_PARTS][ATT_M], stat_l[SM_PARTS][ATT_M];
  __shared__ int stat_rp[SM_PARTS][ATT_M];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_M;
  const int dc = blockIdx.y;
  const int bh = blockIdx.z;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    for (int i = 0; i < QK_STAGES; ++i) { mbar_init(&bars->qk_full[i], 1); mbar_init(&bars->qk_empty[i], 1); }
    for (int i = 0; i < V_STAGES; ++i) { mbar_init(&bars->v_full[i], 1); mbar_init(&bars->v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->s_full[i], 1); mbar_init(&bars->s_empty[i], SM_WARPS);
      mbar_init(&bars->p_full[i], SM_WARPS); mbar_init(&bars->p_empty[i], 1);
    }
    mbar_init(&bars->o_full, 1);
    mbar_init(&bars->q_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(ATT_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  // TMA producer and MMA issuer: the whole warp runs the loop (uniform control flow, descriptors in uniform registers), one
  // elected lane issues -- see the note in qgemm_sm100.cu.
  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0; uint32_t phase = 0;
    int vs = 0; uint32_t vphase = 0;
    if (p.q_resident) {            // the CTA's queries: loaded once, read by every S = Q.K^T of both passes
      if (elect_one()) {
        mbar_expect_tx(&bars->q_full, (uint32_t)p.k_chunks * ATT_TILE_BYTES);
        for (int kc = 0; kc < p.k_chunks; ++kc) tma_load_3d(smem_q + kc * ATT_TILE_BYTES, &map_q, &bars->q_full, kc * ATT_KB, q0, bh);
      }
      __syncwarp();
    }
    for (int pass = 0; pass < 2; ++pass) {
      for (int j = 0; j < p.s_tiles; ++j) {
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(&bars->qk_empty[stage], phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&bars->qk_full[stage], (p.q_resident ? 1 : 2) * ATT_TILE_BYTES);
            if (!p.q_resident) tma_load_3d(smem_q + stage * ATT_TILE_BYTES, &map_q, &bars->qk_full[stage], kc * ATT_KB, q0, bh);
            tma_load_3d(smem_k + stage * ATT_TILE_BYTES, &map_k, &bars->qk_full[stage], kc * ATT_KB, j * ATT_S, bh);
          }
          __syncwarp();
          if (++stage == p.qk_stages) { stage = 0; phase ^= 1; }
        }
        if (pass == 1) {
          mbar_wait(&bars->v_empty[vs], vphase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&bars->v_full[vs], (uint32_t)p.d_chunk * ATT_KB);
            const int vrows = p.d_chunk / p.v_halves;
            tma_load_3d(smem_v + vs * p.v_stage_bytes, &map_v, &bars->v_full[vs], j * ATT_S, dc * p.d_chunk, bh);
            if (p.v_halves == 2)
              tma_load_3d(smem_v + vs * p.v_stage_bytes + vrows * ATT_KB, &map_v, &bars->v_full[vs], j * ATT_S, dc * p.d_chunk + vrows, bh);
          }
          __syncwarp();
          if (++vs == p.v_stages) { vs = 0; vphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_s = make_idesc_i8(ATT_S, 0, 0);
    const int vrows = p.d_chunk / p.v_halves;
    const uint32_t idesc_o = make_idesc_i8(vrows, 0, 0);
    const uint32_t tmem_o = tmem_base + 128u * p.s_bufs;
    const uint32_t q_base = smem_u32(smem_q), k_base = smem_u32(smem_k), v_base = smem_u32(smem_v), p_base = smem_u32(smem_p);
    int stage = 0; uint32_t phase = 0;
    int vs = 0; uint32_t vphase = 0;
    auto issue_pv = [&](int jj) {
      const int pb = jj & 1;
      mbar_wait(&bars->p_full[pb], (uint32_t)((jj >> 1) & 1));
      mbar_wait(&bars->v_full[vs], vphase);
      tc_fence_after();
      const uint64_t adesc = make_smem_desc(p_base + pb * ATT_TILE_BYTES);
      const uint64_t bdesc = make_smem_desc(v_base + vs * p.v_stage_bytes);
      if (elect_one()) {
        umma_i8(tmem_o, adesc, bdesc, idesc_o, jj ? 1u : 0u);
        umma_i8(tmem_o, adesc + 2, bdesc + 2, idesc_o, 1u);
        umma_i8(tmem_o, adesc + 4, bdesc + 4, idesc_o, 1u);
        umma_i8(tmem_o, adesc + 6, bdesc + 6, idesc_o, 1u);
        if (p.v_halves == 2) {      // second half of the head-dim chunk: next `vrows` rows of the V tile -> next `vrows` O columns
          const uint64_t bdesc2 = make_smem_desc(v_base + vs * p.v_stage_bytes + vrows * ATT_KB);
          umma_i8(tmem_o + vrows, adesc, bdesc2, idesc_o, jj ? 1u : 0u);
          umma_i8(tmem_o + vrows, adesc + 2, bdesc2 + 2, idesc_o, 1u);
          umma_i8(tmem_o + vrows, adesc + 4, bdesc2 + 4, idesc_o, 1u);
          umma_i8(tmem_o + vrows, adesc + 6, bdesc2 + 6, idesc_o, 1u);
        }
        umma_commit(&bars->v_empty[vs]);
        umma_commit(&bars->p_empty[pb]);
      }
      __syncwarp();
      if (++vs == p.v_stages) { vs = 0; vphase ^= 1; }
    };
    if (p.q_resident) { mbar_wait(&bars->q_full, 0); tc_fence_after(); }
    int g = 0;
    for (int pass = 0; pass < 2; ++pass) {
      for (int j = 0; j < p.s_tiles; ++j, ++g) {
        const int sb = p.s_bufs == 2 ? (g & 1) : 0;
        mbar_wait(&bars->s_empty[sb], (uint32_t)(((p.s_bufs == 2 ? (g >> 1) : g) & 1) ^ 1));
        tc_fence_after();
        const uint32_t tmem_s = tmem_base + (uint32_t)sb * ATT_S;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          const int nmma = (kc == p.k_chunks - 1) ? p.k_last_mmas : ATT_KB / UMMA_K;
          mbar_wait(&bars->qk_full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(q_base + (p.q_resident ? kc : stage) * ATT_TILE_BYTES);
          const uint64_t bdesc = make_smem_desc(k_base + stage * ATT_TILE_BYTES);
          if (elect_one()) {
            umma_i8(tmem_s, adesc, bdesc, idesc_s, kc ? 1u : 0u);
            if (nmma > 1) umma_i8(tmem_s, adesc + 2, bdesc + 2, idesc_s, 1u);
            if (nmma > 2) umma_i8(tmem_s, adesc + 4, bdesc + 4, idesc_s, 1u);
            if (nmma > 3) umma_i8(tmem_s, adesc + 6, bdesc + 6, idesc_s, 1u);
            umma_commit(&bars->qk_empty[stage]);
          }
          __syncwarp();
          if (++stage == p.qk_stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(&bars->s_full[sb]);
        __syncwarp();
        if (pass == 1 && j > 0) issue_pv(j - 1);
      }
    }
    issue_pv(p.s_tiles - 1);
    if (elect_one()) umma_commit(&bars->o_full);
    __syncwarp();
  } else {
    // ===================== softmax + epilogue (warps 2..17) =====================
    // Four threads per query row: warps w, w+4, w+8, w+12 share a TMEM lane quarter and split every 128-key tile into
    // four parts of 32 columns (two 16-column TMEM loads each).  16 warps = 4 per scheduler, which is what hides the
    // TMEM / MUFU / dependent-issue latencies of the softmax math.  Row statistics and the code row-sum are merged
    // through shared memory.
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;           // 0..3
    const int r = quarter * 32 + lane;
    const int t = q0 + r;
    const bool row_ok = t < p.Tq;
    const int st = threadIdx.x - 64;            // 0..511 among the softmax threads
    const int zq = (int)__ldg(p.zq), zk = (int)__ldg(p.zk);
    const float alpha = __ldg(p.dq) * __ldg(p.dk) * p.sm_scale;
    const float alpha2 = alpha * 1.4426950408889634f;   // to the base-2 exponent domain
    const int32_t* rk = p.rk + (size_t)bh * p.Tk;
    const int row_const = p.d * zq * zk - (row_ok ? zk * __ldg(p.rq + (size_t)bh * p.Tq + t) : 0);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int col_base = part * SM_COLS;
    const float rc_a = (float)row_const * alpha, rc_a2 = (float)row_const * alpha2;   // row constant folded into the FMA

    for (int s = st; s < p.s_tiles * ATT_S; s += 32 * SM_WARPS)
      colint[s] = (SMALL_V ? MAGIC_I : 0) - (s < p.Tk ? zq * __ldg(rk + s) : 0);
    asm volatile("bar.sync 1, 512;" ::: "memory");

    float m = -INFINITY, l = 0.f;     // m in units of x = S*alpha (natural domain), l = sum 2^((x-m)*log2e)
    int g = 0;
    // ---- pass 1: row max and sum ----
    for (int j = 0; j < p.s_tiles; ++j, ++g) {
      const int sb = p.s_bufs == 2 ? (g & 1) : 0;
      mbar_wait(&bars->s_full[sb], (uint32_t)((p.s_bufs == 2 ? (g >> 1) : g) & 1));
      tc_fence_after();
      uint32_t raws[SM_COLS / 16][16];
#pragma unroll
      for (int cc = 0; cc < SM_COLS / 16; ++cc)      // both TMEM loads in flight before the first is consumed
        if (j * ATT_S + col_base + cc * 16 < p.Tk) tmem_ld16(lane_addr + (uint32_t)sb * ATT_S + col_base + cc * 16, raws[cc]);
      tmem_ld_wait();
      // the scores are in registers: hand the S accumulator back right away so the next Q.K^T runs under this tile's softmax math
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_empty[sb]);
      if ((j + 1) * ATT_S <= p.Tk) {
        // full tile (the common case): all 32 columns of this thread at once, no per-chunk bounds logic, one rescale
        const int4* cv = reinterpret_cast<const int4*>(&colint[j * ATT_S + col_base]);
        int v[SM_COLS];
        int vmax = INT_MIN;
#pragma unroll
        for (int q4 = 0; q4 < SM_COLS / 4; ++q4) {
          const int4 c4 = cv[q4];
          const uint32_t* raw = &raws[q4 >> 2][(q4 & 3) * 4];
          v[4 * q4 + 0] = (int)raw[0] + c4.x; v[4 * q4 + 1] = (int)raw[1] + c4.y;
          v[4 * q4 + 2] = (int)raw[2] + c4.z; v[4 * q4 + 3] = (int)raw[3] + c4.w;
          vmax = max(max(max(vmax, v[4 * q4 + 0]), max(v[4 * q4 + 1], v[4 * q4 + 2])), v[4 * q4 + 3]);
        }
        const float m_new = fmaxf(m, fmaf(biased_to_float<SMALL_V>(vmax), alpha, rc_a));
        const float cexp = rc_a2 - m_new * 1.4426950408889634f;
        float add0 = 0.f, add1 = 0.f;
#pragma unroll
        for (int i = 0; i < SM_COLS; i += 2) {
          add0 += ex2_approx(fmaf(biased_to_float<SMALL_V>(v[i]), alpha2, cexp));
          add1 += ex2_approx(fmaf(biased_to_float<SMALL_V>(v[i + 1]), alpha2, cexp));
        }
        l = l * ex2_approx((m - m_new) * 1.4426950408889634f) + (add0 + add1);
        m = m_new;
      } else {
#pragma unroll
      for (int cc = 0; cc < SM_COLS / 16; ++cc) {
        const int c0 = col_base + cc * 16;
        const int s0 = j * ATT_S + c0;
        if (s0 < p.Tk) {
          uint32_t (&raw)[16] = raws[cc];
          const int4* cv = reinterpret_cast<const int4*>(&colint[s0]);
          int v[16];
          if (s0 + 16 <= p.Tk) {
            // full chunk: integer max first (2 instructions / score), then one FMA + ex2 + add per score
            int vmax = INT_MIN;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const int4 c4 = cv[q4];
              v[4 * q4 + 0] = (int)raw[4 * q4 + 0] + c4.x; v[4 * q4 + 1] = (int)raw[4 * q4 + 1] + c4.y;
              v[4 * q4 + 2] = (int)raw[4 * q4 + 2] + c4.z; v[4 * q4 + 3] = (int)raw[4 * q4 + 3] + c4.w;
              vmax = max(max(max(vmax, v[4 * q4 + 0]), max(v[4 * q4 + 1], v[4 * q4 + 2])), v[4 * q4 + 3]);
            }
            const float m_new = fmaxf(m, fmaf(biased_to_float<SMALL_V>(vmax), alpha, rc_a));
            const float cexp = rc_a2 - m_new * 1.4426950408889634f;
            float add = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) add += ex2_approx(fmaf(biased_to_float<SMALL_V>(v[i]), alpha2, cexp));
            l = l * ex2_approx((m - m_new) * 1.4426950408889634f) + add;
            m = m_new;
          } else {
            float cmax = -INFINITY;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              v[i] = (int)raw[i] + colint[s0 + i];
              if (s0 + i < p.Tk) cmax = fmaxf(cmax, fmaf(biased_to_float<SMALL_V>(v[i]), alpha, rc_a));
            }
            const float m_new = fmaxf(m, cmax);
            const float cexp = rc_a2 - m_new * 1.4426950408889634f;
            float add = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) add += (s0 + i < p.Tk) ? ex2_approx(fmaf(biased_to_float<SMALL_V>(v[i]), alpha2, cexp)) : 0.f;
            l = l * ex2_approx((m - m_new) * 1.4426950408889634f) + add;
            m = m_new;
          }
        }
      }
      }
    }
    // merge the column parts of every row
    stat_m[part][r] = m; stat_l[part][r] = l;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    {
      float mm = m;
#pragma unroll
      for (int o = 0; o < SM_PARTS; ++o) mm = fmaxf(mm, stat_m[o][r]);
      float ll = 0.f;
#pragma unroll
      for (int o = 0; o < SM_PARTS; ++o) {
        const float mo = stat_m[o][r];
        ll += (mo == -INFINITY) ? 0.f : stat_l[o][r] * ex2_approx((mo - mm) * 1.4426950408889634f);
      }
      m = mm; l = ll;
    }
    // ---- pass 2: normalised probabilities -> codes -> smem (A operand of the P.V MMA) ----
    const float dpq = __ldg(p.dpq), zpq = __ldg(p.zpq);
    const float qmax = (float)(p.p_levels - 1);
    const float kq = (1.0f / l) / dpq;            // code = rint(2^((x-m)log2e) * kq) + zP
    const float cexp2 = rc_a2 - m * 1.4426950408889634f;
    const int zp_i = (int)zpq;
    const bool fast_codes = p.p_levels == 256;
    int rp = 0;
    for (int j = 0; j < p.s_tiles; ++j, ++g) {
      const int sb = p.s_bufs == 2 ? (g & 1) : 0, pb = j & 1;
      mbar_wait(&bars->s_full[sb], (uint32_t)((p.s_bufs == 2 ? (g >> 1) : g) & 1));
      tc_fence_after();
      uint32_t raws[SM_COLS / 16][16];
#pragma unroll
      for (int cc = 0; cc < SM_COLS / 16; ++cc)
        if (j * ATT_S + col_base + cc * 16 < p.Tk) tmem_ld16(lane_addr + (uint32_t)sb * ATT_S + col_base + cc * 16, raws[cc]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_empty[sb]);
      mbar_wait(&bars->p_empty[pb], (uint32_t)(((j >> 1) & 1) ^ 1));
      uint8_t* prow = smem_p + pb * ATT_TILE_BYTES + r * 128;
      if (fast_codes && (j + 1) * ATT_S <= p.Tk) {
        // full tile, 8-bit codes: FMA + ex2 + mul + magic-round per score, saturating pack, no bounds logic
        const int4* cv = reinterpret_cast<const int4*>(&colint[j * ATT_S + col_base]);
#pragma unroll
        for (int cc = 0; cc < SM_COLS / 16; ++cc) {
          uint32_t packed[4];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int4 c4 = cv[cc * 4 + q4];
            const uint32_t* raw = &raws[cc][q4 * 4];
            const int k0 = rint_magic(ex2_approx(fmaf(biased_to_float<SMALL_V>((int)raw[0] + c4.x), alpha2, cexp2)) * kq) + zp_i;
            const int k1 = rint_magic(ex2_approx(fmaf(biased_to_float<SMALL_V>((int)raw[1] + c4.y), alpha2, cexp2)) * kq) + zp_i;
            const int k2 = rint_magic(ex2_approx(fmaf(biased_to_float<SMALL_V>((int)raw[2] + c4.z), alpha2, cexp2)) * kq) + zp_i;
            const int k3 = rint_magic(ex2_approx(fmaf(biased_to_float<SMALL_V>((int)raw[3] + c4.w), alpha2, cexp2)) * kq) + zp_i;
            const uint32_t w = pack_sat_u8(k1, k0, pack_sat_u8(k3, k2, 0u));
            packed[q4] = w;
            rp = (int)__dp4a(w, 0x01010101u, (unsigned)rp);
          }
          const int chunk = (col_base >> 4) + cc;
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
      } else {
#pragma unroll
      for (int cc = 0; cc < SM_COLS / 16; ++cc) {
        const int c0 = col_base + cc * 16;
        const int s0 = j * ATT_S + c0;
        uint32_t packed[4] = {0, 0, 0, 0};
        if (s0 < p.Tk) {
          uint32_t (&raw)[16] = raws[cc];
          const int4* cv = reinterpret_cast<const int4*>(&colint[s0]);
          if (fast_codes && s0 + 16 <= p.Tk) {
            // full chunk, 8-bit codes: FMA + ex2 + mul + cvt.rni per score, saturating pack 2 codes / instruction
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const int4 c4 = cv[q4];
              const int k0 = rint_magic(ex2_approx(fmaf(biased_to_float<SMALL_V>((int)raw[4 * q4 + 0] + c4.x), alpha2, cexp2)) * kq) + zp_i;
              const int k1 = rint_magic(ex2_approx(fmaf(biased_to_float<SMALL_V>((int)raw[4 * q4 + 1] + c4.y), alpha2, cexp2)) * kq) + zp_i;
              const int k2 = rint_magic(ex2_approx(fmaf(biased_to_float<SMALL_V>((int)raw[4 * q4 + 2] + c4.z), alpha2, cexp2)) * kq) + zp_i;
              const int k3 = rint_magic(ex2_approx(fmaf(biased_to_float<SMALL_V>((int)raw[4 * q4 + 3] + c4.w), alpha2, cexp2)) * kq) + zp_i;
              const uint32_t w = pack_sat_u8(k1, k0, pack_sat_u8(k3, k2, 0u));
              packed[q4] = w;
              rp = (int)__dp4a(w, 0x01010101u, (unsigned)rp);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float x = biased_to_float<SMALL_V>((int)raw[i] + colint[s0 + i]);
              const float e = ex2_approx(fmaf(x, alpha2, cexp2));
              uint32_t code = (uint32_t)fminf(fmaxf(rintf(e * kq) + zpq, 0.f), qmax);
              code = (s0 + i < p.Tk) ? code : 0u;
              rp += (int)code;
              packed[i >> 2] |= code << (8 * (i & 3));
            }
          }
        }
        const int chunk = c0 >> 4;  // one 16-byte chunk per 16 columns, XOR-swizzled by the row (SWIZZLE_128B)
        *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
      }
      }
      fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[pb]);
    }
    // ---- epilogue: O -> fp32 output (the four threads of a row take alternate 16-column groups) ----
    stat_rp[part][r] = rp;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    rp = 0;
#pragma unroll
    for (int o = 0; o < SM_PARTS; ++o) rp += stat_rp[o][r];
    mbar_wait(&bars->o_full, 0);
    tc_fence_after();
    const int zv = (int)__ldg(p.zv);
    const float oscale = dpq * __ldg(p.dv);
    const int32_t* rv = p.rv + (size_t)bh * p.d;
    const int b = bh / p.heads, h = bh - b * p.heads;
    float* obase = p.out + b * p.o_sb + h * p.o_sh + (long long)t * p.o_st;
    const int row_o = p.Tk * zp_i * zv - zv * rp;
    // channel-contiguous outputs ([B, T, heads*d]): a thread owns 16 consecutive floats of its row per step -> four 16-byte stores
    // (the scalar form costs one 32-byte sector transaction per element: 49 152 per CTA at d = 384)
    const bool vec_out = p.o_sc == 1 && ((p.o_sb | p.o_sh | p.o_st) & 3) == 0 && (p.d & 3) == 0 && (p.d_chunk & 3) == 0 &&
                         (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
    for (int c0 = part * 16; c0 < p.d_chunk; c0 += 16 * SM_PARTS) {
      uint32_t raw[16];
      tmem_ld16(lane_addr + 128u * p.s_bufs + c0, raw);
      tmem_ld_wait();
      if (row_ok) {
        const int cb = dc * p.d_chunk + c0;
        if (vec_out) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int c = cb + 4 * k;
            if (c < p.d) {                                              // d % 4 == 0: the four channels are all inside
              const int4 r4 = __ldg(reinterpret_cast<const int4*>(rv + c));
              float4 o;
              o.x = (float)((int)raw[4 * k + 0] + row_o - zp_i * r4.x) * oscale;
              o.y = (float)((int)raw[4 * k + 1] + row_o - zp_i * r4.y) * oscale;
              o.z = (float)((int)raw[4 * k + 2] + row_o - zp_i * r4.z) * oscale;
              o.w = (float)((int)raw[4 * k + 3] + row_o - zp_i * r4.w) * oscale;
              *reinterpret_cast<float4*>(obase + c) = o;
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = cb + i;
            if (c < p.d) obase[(long long)c * p.o_sc] = (float)((int)raw[i] + row_o - zp_i * __ldg(rv + c)) * oscale;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(ATT_TMEM_COLS));
  }
}

}  // namespace edadm

using namespace edadm;

// qc, kc: u8 codes [BH][Tq|Tk][dp] (dp % 16 == 0, padded bytes 0); vc: u8 codes [BH][d][Tkp] (Tkp % 16 == 0, padding 0);
// rq[BH][Tq], rk[BH][Tk]: per-token code sums; rv[BH][d]: per-channel code sums.  Scalars are device pointers.
// out fp32: element (b, h, t, c) at b*o_sb + h*o_sh + t*o_st + c*o_sc with bh = b*heads + h.
extern "C" int edadm_qattn_fwd(const uint8_t* qc, const uint8_t* kc, const uint8_t* vc, const int32_t* rq, const int32_t* rk,
                               const int32_t* rv, int BH, int heads, int Tq, int Tk, int d, int dp, int Tkp,
                               const float* dq, const float* zq, const float* dk, const float* zk, const float* dv,
                               const float* zv, const float* dpq, const float* zpq, int p_levels, float sm_scale, float* out,
                               int64_t o_sb, int64_t o_sh, int64_t o_st, int64_t o_sc, void* stream) {
  if (!qc || !kc || !vc || !rq || !rk || !rv || !dq || !zq || !dk || !zk || !dv || !zv || !dpq || !zpq || !out)
    return fail(EDADM_ERR_ARG, "qattn_fwd: null pointer");
  if (BH < 1 || heads < 1 || (BH % heads) || Tq < 1 || Tk < 1 || d < 1 || dp < d || (dp & 15) || Tkp < Tk || (Tkp & 15) ||
      p_levels < 2 || p_levels > 256)
    return fail(EDADM_ERR_ARG, "qattn_fwd: bad sizes BH=%d heads=%d Tq=%d Tk=%d d=%d dp=%d Tkp=%d", BH, heads, Tq, Tk, d, dp, Tkp);
  if (BH > 65535) return fail(EDADM_ERR_UNSUPPORTED, "qattn_fwd: more than 65535 (batch x heads) per launch");
  if (Tk > ATT_MAX_KEYS) return fail(EDADM_ERR_UNSUPPORTED, "qattn_fwd: more than %d keys per (batch, head)", ATT_MAX_KEYS);
  if ((((uintptr_t)qc | (uintptr_t)kc | (uintptr_t)vc) & 15)) return fail(EDADM_ERR_ARG, "qattn_fwd: operands must be 16-byte aligned");

  AttnParams p;
  p.Tq = Tq; p.Tk = Tk; p.d = d;
  p.k_chunks = (dp + ATT_KB - 1) / ATT_KB;
  p.k_last_mmas = (dp - (p.k_chunks - 1) * ATT_KB + UMMA_K - 1) / UMMA_K;
  p.s_tiles = (Tk + ATT_S - 1) / ATT_S;
  // head-dim chunk of one CTA: two S accumulators + O need 256 + d_chunk <= 512 TMEM columns; a single S accumulator (handed
  // back as soon as the softmax warps hold the scores in registers) leaves 384 columns for O, so d = 384 (LDM-4 ImageNet,
  // 32x32 level) needs ONE pass over the scores instead of two
  p.s_bufs = d <= 256 ? 2 : 1;
  const int d_cap = 512 - 128 * p.s_bufs;
  const int d_chunks = (d + d_cap - 1) / d_cap;
  p.d_chunk = (((d + d_chunks - 1) / d_chunks) + 15) & ~15;
  p.v_halves = p.d_chunk > 256 ? 2 : 1;
  if (p.v_halves == 2) p.d_chunk = (p.d_chunk + 31) & ~31;
  p.v_stage_bytes = (p.d_chunk * ATT_KB + 1023) & ~1023;
  // shared memory: [Q][K ring][V ring][2 P tiles][barriers][per-key terms].  The queries of the CTA stay resident when their
  // k_chunks x 16 KB leave room for a K ring of at least 3 tiles (every S tile of both passes re-reads them: streaming them with
  // the K chunks doubled the operand traffic and halved the useful bytes in flight); V is double-buffered up to 32 KB tiles.
  p.v_stages = p.v_stage_bytes > 32 * 1024 ? 1 : V_STAGES;
  const int colint_bytes = p.s_tiles * ATT_S * 4;
  const int fixed_smem = 1024 + p.v_stages * p.v_stage_bytes + 2 * ATT_TILE_BYTES + 256 + colint_bytes;
  p.q_resident = (ATT_SMEM_BYTES - fixed_smem - p.k_chunks * ATT_TILE_BYTES) / ATT_TILE_BYTES >= 3 ? 1 : 0;
  if (const char* e = getenv("EDADM_ATTN_QRES")) p.q_resident = atoi(e) && p.q_resident;
  if (p.q_resident) {
    p.qk_stages = std::min(QK_STAGES, (ATT_SMEM_BYTES - fixed_smem - p.k_chunks * ATT_TILE_BYTES) / ATT_TILE_BYTES);
    p.q_bytes = p.k_chunks * ATT_TILE_BYTES;
  } else {
    p.qk_stages = std::min(QK_STAGES, (ATT_SMEM_BYTES - fixed_smem) / (2 * ATT_TILE_BYTES));
    p.q_bytes = p.qk_stages * ATT_TILE_BYTES;
  }
  const int att_smem = fixed_smem + p.q_bytes + p.qk_stages * ATT_TILE_BYTES;
  if (p.qk_stages < 2) return fail(EDADM_ERR_UNSUPPORTED, "qattn_fwd: shared memory budget exceeded (d_chunk %d)", p.d_chunk);
  if (att_smem > ATT_SMEM_BYTES) return fail(EDADM_ERR_UNSUPPORTED, "qattn_fwd: shared memory budget exceeded (d_chunk %d)", p.d_chunk);
  p.heads = heads;
  p.o_sb = o_sb; p.o_sh = o_sh; p.o_st = o_st; p.o_sc = o_sc;
  p.sm_scale = sm_scale;
  p.dq = dq; p.zq = zq; p.dk = dk; p.zk = zk; p.dv = dv; p.zv = zv; p.dpq = dpq; p.zpq = zpq;
  p.p_levels = p_levels;
  p.rq = rq; p.rk = rk; p.rv = rv; p.out = out;

  CUtensorMap map_q, map_k, map_v;
  {
    cuuint64_t dims[3] = {(cuuint64_t)dp, (cuuint64_t)Tq, (cuuint64_t)BH};
    cuuint64_t strides[2] = {(cuuint64_t)dp, (cuuint64_t)Tq * dp};
    cuuint32_t box[3] = {(cuuint32_t)ATT_KB, (cuuint32_t)ATT_M, 1u};
    int rc = encode_map(&map_q, qc, 3, dims, strides, box, "attention q codes");
    if (rc) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)dp, (cuuint64_t)Tk, (cuuint64_t)BH};
    cuuint64_t strides[2] = {(cuuint64_t)dp, (cuuint64_t)Tk * dp};
    cuuint32_t box[3] = {(cuuint32_t)ATT_KB, (cuuint32_t)ATT_S, 1u};
    int rc = encode_map(&map_k, kc, 3, dims, strides, box, "attention k codes");
    if (rc) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)Tkp, (cuuint64_t)d, (cuuint64_t)BH};
    cuuint64_t strides[2] = {(cuuint64_t)Tkp, (cuuint64_t)d * Tkp};
    cuuint32_t box[3] = {(cuuint32_t)ATT_S, (cuuint32_t)(p.d_chunk / p.v_halves), 1u};
    int rc = encode_map(&map_v, vc, 3, dims, strides, box, "attention v codes");
    if (rc) return rc;
  }
  static bool attr_set_dev[64] = {false};       // the opt-in is per device
  int attr_dev = 0;
  cudaGetDevice(&attr_dev);
  attr_dev = attr_dev < 64 ? attr_dev : 63;
  if (!attr_set_dev[attr_dev]) {
    cudaError_t e = cudaFuncSetAttribute(qattn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(qattn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES);
    if (e != cudaSuccess) return fail(EDADM_ERR_CUDA, "qattn_fwd: cannot opt in to %d B shared memory: %s", ATT_SMEM_BYTES, cudaGetErrorString(e));
    attr_set_dev[attr_dev] = true;
  }
  dim3 grid((Tq + ATT_M - 1) / ATT_M, d_chunks, BH);
  // |raw - zq*rk| <= 255*255*max(d, dp): the magic-number int->float conversion is exact below 2^22
  if ((long long)dp * 65025LL < (1LL << 22))
    qattn_kernel<true><<<grid, ATT_THREADS, att_smem, (cudaStream_t)stream>>>(map_q, map_k, map_v, p);
  else
    qattn_kernel<false><<<grid, ATT_THREADS, att_smem, (cudaStream_t)stream>>>(map_q, map_k, map_v, p);
  return check_launch("qattn_fwd");
}
