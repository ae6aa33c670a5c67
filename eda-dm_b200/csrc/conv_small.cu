// fp32 3x3 convolution with a handful of output channels (the UNet's last layer: C -> 3 or 4 channels).
// The reference keeps the network output layer's INPUT un-quantized (quant_model.py `disable_network_output_quantization`),
// so this layer is an fp32 activation x fake-quantized 8-bit weight convolution, not an integer GEMM.  With N <= 4 it is a
// memory-bound stencil that library implicit-GEMM kernels handle poorly (one 0.49 ms launch in the LSUN-church step);
// here a block owns an 8 x 32 pixel tile, four thread groups split the input channels, and every thread keeps the 27..36
// weights of a channel in registers for four horizontally adjacent pixels (7 + 18 shared-memory loads per 108 FMAs).
#include <type_traits>

#include "common.cuh"

namespace edadm {

constexpr int CS_TH = 8, CS_TW = 32;          // output tile
constexpr int CS_PW = CS_TW + 2 + 2;          // patch row pitch (34 used, padded to 36 floats)
constexpr int CS_PH = CS_TH + 2;
constexpr int CS_CCH = 4;                     // channels staged per step and group
constexpr int CS_GROUPS = 4;                  // channel groups (64 threads each)

template <int NOUT>
__global__ void __launch_bounds__(256, 2)
conv3x3_small_n_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                       const float* __restrict__ aff_a, const float* __restrict__ aff_s, int silu,
                       float* __restrict__ out, int C, int H, int W, int tiles_w) {
  extern __shared__ float cs_smem[];
  float* wsm = cs_smem;                                              // [C][NOUT][12] (9 used, 16-byte rows)
  float* patch = cs_smem + (size_t)C * NOUT * 12;                    // [CS_GROUPS][CS_CCH][CS_PH][CS_PW]
  const int b = blockIdx.z;
  const int ty0 = blockIdx.y * CS_TH, tx0 = (blockIdx.x % tiles_w) * CS_TW;
  const int grp = threadIdx.x >> 6, tg = threadIdx.x & 63;
  const int py = tg >> 3, px = (tg & 7) * 4;                         // this thread's 4 pixels: row py, cols px..px+3
  for (int i = threadIdx.x; i < C * NOUT * 9; i += 256) {
    const int k = i % 9, r = i / 9, n = r % NOUT, c = r / NOUT;
    wsm[((size_t)c * NOUT + n) * 12 + k] = __ldg(w + ((size_t)n * C + c) * 9 + k);
  }
  float acc[NOUT][4];
#pragma unroll
  for (int n = 0; n < NOUT; ++n)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[n][j] = 0.f;
  const float* xb = x + (size_t)b * C * H * W;
  float* pg = patch + (size_t)grp * CS_CCH * CS_PH * CS_PW;
  const int c_per_group = (C + CS_GROUPS - 1) / CS_GROUPS;
  const int cbeg = grp * c_per_group, cend = min(C, cbeg + c_per_group);
  const int steps = (c_per_group + CS_CCH - 1) / CS_CCH;             // same trip count for every group (block barriers)
  // patch elements of one step handled by this thread (CS_CCH * CS_PH * 34 = 1360 values over 64 threads -> 22 each); the
  // values of step s+1 are requested into registers before step s is computed, so global latency overlaps the FMAs.
  // Where each patch element lives does not depend on the step: a table built once per block holds its offset inside the
  // step's 4-channel slab (packed with the channel index, -1 outside the image), so a fetch is one table read + one load.
  // Optional per-(image, channel) affine + SiLU on load: the GroupNorm + SiLU in front of the output conv (edadm_gn_fold),
  // in the same branch-free arithmetic as the activation producers (pack.cu norm_act_m).
  constexpr int PER_T = (CS_CCH * CS_PH * 34 + 63) / 64;
  __shared__ int tbl[CS_CCH * CS_PH * 34];
  for (int i = threadIdx.x; i < CS_CCH * CS_PH * 34; i += 256) {
    const int col = i % 34, r = i / 34, row = r % CS_PH, cc = r / CS_PH;
    const int iy = ty0 + row - 1, ix = tx0 + col - 1;
    tbl[i] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? (((cc * H + iy) * W + ix) << 2 | cc) : -1;
  }
  __syncthreads();
  float stage[PER_T];
  // fetch: raw loads only (they stay in flight while the previous step's FMAs run); the activation is applied when the values
  // are committed to shared memory one step later
  auto fetch = [&](int s) {
    const int c0 = cbeg + s * CS_CCH;
    const float* xs = xb + (size_t)c0 * H * W;
    const int live = cend - c0;                                       // channels of this step that exist (<= 0 .. CS_CCH)
#pragma unroll
    for (int u = 0; u < PER_T; ++u) {
      const int i = tg + u * 64;
      float v = 0.f;
      if (i < CS_CCH * CS_PH * 34) {
        const int e = tbl[i];
        if (e >= 0 && (e & 3) < live) v = __ldg(xs + (e >> 2));
      }
      stage[u] = v;
    }
  };
  auto commit_mode = [&](int s, auto mode_tag) {
    constexpr int MODE = decltype(mode_tag)::value;
    const int c0 = cbeg + s * CS_CCH;
    const int live = cend - c0;
    float a4[CS_CCH], s4[CS_CCH];
#pragma unroll
    for (int cc = 0; cc < CS_CCH; ++cc) {
      const bool ok = MODE >= 0 && cc < live;
      a4[cc] = ok ? __ldg(aff_a + (size_t)b * C + c0 + cc) : 1.f;
      s4[cc] = ok ? __ldg(aff_s + (size_t)b * C + c0 + cc) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < PER_T; ++u) {
      const int i = tg + u * 64;
      if (i < CS_CCH * CS_PH * 34) {
        float v = stage[u];
        const int e = tbl[i];
        const int cc = e & 3;
        if (MODE >= 0 && e >= 0 && cc < live) {                      // zero padding applies AFTER the activation
          v = fmaf(v, a4[cc], s4[cc]);
          if (MODE == 17) v = __fdividef(v, 1.0f + __expf(-v));
          else if (MODE == 1) {            // v / (1 + expf(-v)), IEEE quotient (see pack.cu)
            const float d = fminf(__fadd_rn(1.0f, expf(-v)), 8.507059e37f);
            float r;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
            r = fmaf(r, fmaf(-d, r, 1.0f), r);
            const float q0 = __fmul_rn(v, r);
            v = fmaf(fmaf(-d, q0, v), r, q0);
          } else if (MODE == 2) {          // v * sigmoid(v) with the IEEE reciprocal
            const float d = fminf(__fadd_rn(1.0f, expf(-v)), 8.507059e37f);
            float r;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
            r = fmaf(r, fmaf(-d, r, 1.0f), r);
            v = __fmul_rn(v, r);
          }
        }
        const int col = i % 34, rr = i / 34;                          // rr = cc * CS_PH + row
        pg[rr * CS_PW + col] = v;
      }
    }
  };
  const int mode = !aff_a ? -1 : ((silu & 16) ? ((silu & 3) ? 17 : 0) : silu);
  auto commit = [&](int s) {
    switch (mode) {
      case -1: commit_mode(s, std::integral_constant<int, -1>{}); break;
      case 0: commit_mode(s, std::integral_constant<int, 0>{}); break;
      case 1: commit_mode(s, std::integral_constant<int, 1>{}); break;
      case 2: commit_mode(s, std::integral_constant<int, 2>{}); break;
      default: commit_mode(s, std::integral_constant<int, 17>{}); break;
    }
  };
  fetch(0);
  for (int s = 0; s < steps; ++s) {
    const int c0 = cbeg + s * CS_CCH;
    __syncthreads();                 // previous step's reads of the patch are done
    commit(s);
    __syncthreads();
    if (s + 1 < steps) fetch(s + 1);
#pragma unroll
    for (int cc = 0; cc < CS_CCH; ++cc) {
      const int c = c0 + cc;
      if (c < cend) {
        float in[3][6];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float* rowp = pg + (cc * CS_PH + py + r) * CS_PW + px;      // 16-byte aligned: px % 4 == 0, CS_PW % 4 == 0
          const float4 lo = *reinterpret_cast<const float4*>(rowp);
          const float2 hi = *reinterpret_cast<const float2*>(rowp + 4);
          in[r][0] = lo.x; in[r][1] = lo.y; in[r][2] = lo.z; in[r][3] = lo.w; in[r][4] = hi.x; in[r][5] = hi.y;
        }
#pragma unroll
        for (int n = 0; n < NOUT; ++n) {
          const float4* wp = reinterpret_cast<const float4*>(wsm + ((size_t)c * NOUT + n) * 12);
          const float4 w0 = wp[0], w1 = wp[1];
          const float w8 = wsm[((size_t)c * NOUT + n) * 12 + 8];
          const float wk[9] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w8};
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int q = 0; q < 3; ++q) acc[n][j] = fmaf(in[r][j + q], wk[r * 3 + q], acc[n][j]);
        }
      }
    }
  }
  // combine the channel groups through shared memory (reusing the patch buffer)
  __syncthreads();
  float* red = patch;                                                // [CS_GROUPS][NOUT][256 pixels]
#pragma unroll
  for (int n = 0; n < NOUT; ++n)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[((size_t)grp * NOUT + n) * 256 + py * CS_TW + px + j] = acc[n][j];
  __syncthreads();
  for (int i = threadIdx.x; i < NOUT * 256; i += 256) {
    const int pix = i & 255, n = i >> 8;
    const int oy = ty0 + (pix >> 5), ox = tx0 + (pix & 31);
    if (oy < H && ox < W) {
      float v = bias ? __ldg(bias + n) : 0.f;
#pragma unroll
      for (int g2 = 0; g2 < CS_GROUPS; ++g2) v += red[((size_t)g2 * NOUT + n) * 256 + pix];
      out[(((size_t)b * NOUT + n) * H + oy) * W + ox] = v;
    }
  }
}

}  // namespace edadm

using namespace edadm;

// x fp32 [B][C][H][W], w fp32 [N][C][3][3] (already fake-quantized by the caller), bias fp32 [N] or null,
// out fp32 [B][N][H][W]; stride 1, zero padding 1, N in {1..4}.  aff_a / aff_s (nullable, [B][C]) + silu: the input is
// silu(a*x+s), i.e. the GroupNorm + SiLU of the output head folded into the load (zero padding applies AFTER the activation).
extern "C" int edadm_conv3x3_small_n(const float* x, const float* w, const float* bias, const float* aff_a, const float* aff_s,
                                     int silu, float* out, int B, int C, int H, int W, int N, void* stream) {
  if (!x || !w || !out || ((aff_a == nullptr) != (aff_s == nullptr))) return fail(EDADM_ERR_ARG, "conv3x3_small_n: null pointer");
  if (B < 0 || C < 1 || H < 1 || W < 1 || N < 1) return fail(EDADM_ERR_ARG, "conv3x3_small_n: bad sizes");
  if (N > 4) return fail(EDADM_ERR_UNSUPPORTED, "conv3x3_small_n: at most 4 output channels (got %d)", N);
  if (B == 0) return EDADM_OK;
  if (B > 65535) return fail(EDADM_ERR_UNSUPPORTED, "conv3x3_small_n: batch above 65535");
  const size_t patch_floats = (size_t)CS_GROUPS * CS_CCH * CS_PH * CS_PW;
  const size_t red_floats = (size_t)CS_GROUPS * N * 256;
  const size_t smem = ((size_t)C * N * 12 + (patch_floats > red_floats ? patch_floats : red_floats)) * sizeof(float);
  if (smem > 200 * 1024) return fail(EDADM_ERR_UNSUPPORTED, "conv3x3_small_n: %d input channels do not fit shared memory", C);
  const int tiles_w = (W + CS_TW - 1) / CS_TW, tiles_h = (H + CS_TH - 1) / CS_TH;
  dim3 grid(tiles_w, tiles_h, B);
  cudaStream_t s = (cudaStream_t)stream;
#define EDADM_LAUNCH_CS(NO)                                                                                            \
  {                                                                                                                    \
    if (smem > 48 * 1024) cudaFuncSetAttribute(conv3x3_small_n_kernel<NO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    conv3x3_small_n_kernel<NO><<<grid, 256, smem, s>>>(x, w, bias, aff_a, aff_s, silu, out, C, H, W, tiles_w);                             \
  }
  switch (N) {
    case 1: EDADM_LAUNCH_CS(1); break;
    case 2: EDADM_LAUNCH_CS(2); break;
    case 3: EDADM_LAUNCH_CS(3); break;
    default: EDADM_LAUNCH_CS(4); break;
  }
#undef EDADM_LAUNCH_CS
  return check_launch("conv3x3_small_n");
}
