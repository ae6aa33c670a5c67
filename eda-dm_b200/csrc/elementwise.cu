// HBM-bound kernels of the quantized-UNet path: fused fake-quant (UniformAffineQuantizer),
// AdaRound soft/hard rounding, the rounding regulariser and the L_p reconstruction loss, each
// as ONE streaming pass with its backward.  All run at one read + one write of the tensor
// (vs ~8 ATen passes in the reference) and are sized in whole waves of 148 SMs.
//
// Arithmetic contract (bit-exact codes): true IEEE fp32 division x/delta, rintf (half-to-even,
// == torch.round), zero-point added AFTER rounding, clamp to [0, n_levels-1]
// (reference qdiff/quant_layer.py:267-269).  Compiled WITHOUT --use_fast_math.
#include <algorithm>
#include "common.cuh"

namespace edadm {

struct QDrop {
  const uint8_t* mask;  // explicit keep-mask (1 = take quantized value), may be null
  const float* rnd;     // explicit uniform draws (torch.rand_like): keep iff rnd[i] < prob -- the reference's own stream
  float prob;           // used when mask == null and prob < 1: keep iff u < prob
  uint2 key;
  uint64_t offset;
};

__device__ __forceinline__ bool qdrop_keep(const QDrop& q, int64_t i, uint4& rnd, int64_t& rnd_quad) {
  if (q.mask) return q.mask[i] != 0;
  if (q.prob >= 1.0f) return true;
  if (q.rnd) return __ldcs(q.rnd + i) < q.prob;
  const int64_t quad = (i + (int64_t)q.offset) >> 2;
  if (quad != rnd_quad) {
    rnd = philox4x32_10(make_uint4((uint32_t)quad, (uint32_t)(quad >> 32), 0u, 0u), q.key);
    rnd_quad = quad;
  }
  const int lane = (int)((i + (int64_t)q.offset) & 3);
  const uint32_t r = lane == 0 ? rnd.x : lane == 1 ? rnd.y : lane == 2 ? rnd.z : rnd.w;
  return u01(r) < q.prob;
}

// ---------------------------------------------------------------------------------------------
// K2 forward.  y = where(keep, (clamp(rint(x/d)+zp, 0, L-1) - zp) * d, x)
// channel index of element i is (i / inner) % channels; channels == 1 -> per-tensor.
// ---------------------------------------------------------------------------------------------
template <bool PER_TENSOR>
__global__ void __launch_bounds__(kThreads)
uaq_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ codes,
               const float* __restrict__ delta, const float* __restrict__ zp, int64_t n,
               int64_t channels, int64_t inner, float qmax, QDrop qd) {
  const int64_t nvec = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float d0 = 0.f, z0 = 0.f;
  if (PER_TENSOR) { d0 = __ldg(delta); z0 = __ldg(zp); }
  uint4 rnd = make_uint4(0, 0, 0, 0);
  int64_t rq = -1;
  const bool vec_ok = ((((uintptr_t)x | (uintptr_t)y) & 15) == 0) && (!codes || (((uintptr_t)codes & 3) == 0));
  int64_t done = 0;
  if (vec_ok) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
      const float4 xv = __ldcs(reinterpret_cast<const float4*>(x) + v);
      float xi[4] = {xv.x, xv.y, xv.z, xv.w};
      float yo[4];
      uint32_t cpack = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t i = (v << 2) + j;
        float d = d0, z = z0;
        if (!PER_TENSOR) {
          const int64_t c = (i / inner) % channels;
          d = __ldg(delta + c);
          z = __ldg(zp + c);
        }
        const float q = fminf(fmaxf(rintf(xi[j] / d) + z, 0.f), qmax);
        const float dq = (q - z) * d;
        yo[j] = qdrop_keep(qd, i, rnd, rq) ? dq : xi[j];
        cpack |= ((uint32_t)q & 0xffu) << (8 * j);
      }
      __stcs(reinterpret_cast<float4*>(y) + v, make_float4(yo[0], yo[1], yo[2], yo[3]));
      if (codes) reinterpret_cast<uint32_t*>(codes)[v] = cpack;
    }
    done = nvec << 2;
  }
  for (int64_t i = done + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float d = d0, z = z0;
    if (!PER_TENSOR) {
      const int64_t c = (i / inner) % channels;
      d = __ldg(delta + c);
      z = __ldg(zp + c);
    }
    const float xv = x[i];
    const float q = fminf(fmaxf(rintf(xv / d) + z, 0.f), qmax);
    y[i] = qdrop_keep(qd, i, rnd, rq) ? (q - z) * d : xv;
    if (codes) codes[i] = (uint8_t)q;
  }
}

// ---------------------------------------------------------------------------------------------
// K2 backward (straight-through estimator of round_ste, quant_layer.py:19-23, through clamp and
// the dequant multiply; LSQ-style step-size gradient exactly as autograd derives it):
//   in   = 0 <= rint(x/d)+zp <= L-1
//   gx   = keep ? (in ? (gy*d)/d : 0) : gy        ((gy*d)/d: the two roundings autograd performs)
//   gd  += keep ? gy * ((q - zp) - (in ? x/d : 0)) : 0          (per-tensor delta only)
// ---------------------------------------------------------------------------------------------
template <bool PER_TENSOR, bool WANT_GD>
__global__ void __launch_bounds__(kThreads)
uaq_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x, float* __restrict__ gx,
               const float* __restrict__ delta, const float* __restrict__ zp, int64_t n,
               int64_t channels, int64_t inner, float qmax, QDrop qd, double* __restrict__ partials) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float d0 = 0.f, z0 = 0.f;
  if (PER_TENSOR) { d0 = __ldg(delta); z0 = __ldg(zp); }
  uint4 rnd = make_uint4(0, 0, 0, 0);
  int64_t rq = -1;
  double acc = 0.0;
  const int64_t nvec = n >> 2;
  const bool vec_ok = ((((uintptr_t)x | (uintptr_t)gy | (uintptr_t)gx) & 15) == 0);
  int64_t done = 0;
  if (vec_ok) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
      const float4 xv = __ldcs(reinterpret_cast<const float4*>(x) + v);
      const float4 gv = __ldcs(reinterpret_cast<const float4*>(gy) + v);
      float xi[4] = {xv.x, xv.y, xv.z, xv.w}, gi[4] = {gv.x, gv.y, gv.z, gv.w}, go[4];
      float local = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t i = (v << 2) + j;
        float d = d0, z = z0;
        if (!PER_TENSOR) {
          const int64_t c = (i / inner) % channels;
          d = __ldg(delta + c);
          z = __ldg(zp + c);
        }
        const float r = xi[j] / d;
        const float xint = rintf(r) + z;
        const bool in = (xint >= 0.f) && (xint <= qmax);
        const float q = fminf(fmaxf(xint, 0.f), qmax);
        const bool keep = qdrop_keep(qd, i, rnd, rq);
        go[j] = keep ? (in ? (gi[j] * d) / d : 0.f) : gi[j];
        if (WANT_GD && keep) local += gi[j] * ((q - z) - (in ? r : 0.f));
      }
      __stcs(reinterpret_cast<float4*>(gx) + v, make_float4(go[0], go[1], go[2], go[3]));
      acc += (double)local;
    }
    done = nvec << 2;
  }
  for (int64_t i = done + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float d = d0, z = z0;
    if (!PER_TENSOR) {
      const int64_t c = (i / inner) % channels;
      d = __ldg(delta + c);
      z = __ldg(zp + c);
    }
    const float r = x[i] / d, g = gy[i];
    const float xint = rintf(r) + z;
    const bool in = (xint >= 0.f) && (xint <= qmax);
    const float q = fminf(fmaxf(xint, 0.f), qmax);
    const bool keep = qdrop_keep(qd, i, rnd, rq);
    gx[i] = keep ? (in ? (g * d) / d : 0.f) : g;
    if (WANT_GD && keep) acc += (double)(g * ((q - z) - (in ? r : 0.f)));
  }
  if (WANT_GD) {
    acc = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
  }
}

// final stage of the two-stage reductions: out (+)= scale * sum(partials[0..m))
__global__ void finish_sum_kernel(const double* __restrict__ partials, int m, float* __restrict__ out,
                                  double scale, int accumulate) {
  double v = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) v += partials[i];
  v = block_sum(v);
  if (threadIdx.x == 0) {
    const float r = (float)(v * scale);
    out[0] = accumulate ? out[0] + r : r;
  }
}

// ---------------------------------------------------------------------------------------------
// K3 AdaRound (adaptive_rounding.py:49-64).  h(a) = clamp(sigmoid(a)*(zeta-gamma)+gamma, 0, 1),
// zeta = 1.1, gamma = -0.1.  soft: floor(w/d)+h(a); hard: floor(w/d)+(a>=0).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_exact(float a) { return 1.0f / (1.0f + expf(-a)); }

__device__ __forceinline__ float soft_target(float a) {
  return fminf(fmaxf(sigmoidf_exact(a) * 1.2f + (-0.1f), 0.f), 1.f);
}

__global__ void __launch_bounds__(kThreads)
adaround_fwd_kernel(const float* __restrict__ w, const float* __restrict__ alpha, float* __restrict__ out,
                    uint8_t* __restrict__ codes, const float* __restrict__ delta,
                    const float* __restrict__ zp, int64_t n, int64_t channels, int64_t inner,
                    float qmax, int soft) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t c = (channels == 1) ? 0 : (i / inner) % channels;
    const float d = __ldg(delta + c), z = __ldg(zp + c);
    const float a = alpha[i];
    const float fl = floorf(w[i] / d);
    const float h = soft ? soft_target(a) : (a >= 0.f ? 1.f : 0.f);
    const float q = fminf(fmaxf(fl + h + z, 0.f), qmax);
    out[i] = (q - z) * d;
    if (codes) codes[i] = (uint8_t)q;
  }
}

// galpha = gout * d * 1[0 <= floor+h+zp <= L-1] * h'(a),  h' = 1.2 s (1-s) * 1[0 <= 1.2 s - 0.1 <= 1]
__global__ void __launch_bounds__(kThreads)
adaround_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ w,
                    const float* __restrict__ alpha, float* __restrict__ galpha,
                    const float* __restrict__ delta, const float* __restrict__ zp, int64_t n,
                    int64_t channels, int64_t inner, float qmax, int accumulate) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t c = (channels == 1) ? 0 : (i / inner) % channels;
    const float d = __ldg(delta + c), z = __ldg(zp + c);
    const float a = alpha[i];
    const float s = sigmoidf_exact(a);
    const float hr = s * 1.2f + (-0.1f);
    const float h = fminf(fmaxf(hr, 0.f), 1.f);
    const float xi = floorf(w[i] / d) + h + z;
    const bool in = (xi >= 0.f) && (xi <= qmax);
    const bool hin = (hr >= 0.f) && (hr <= 1.f);
    const float g = (in && hin) ? gout[i] * d * (1.2f * (s * (1.f - s))) : 0.f;
    galpha[i] = accumulate ? galpha[i] + g : g;
  }
}

// alpha init: rest = w/d - floor(w/d); alpha = -log((zeta-gamma)/(rest-gamma) - 1)
// (adaptive_rounding.py:66-72)
__global__ void __launch_bounds__(kThreads)
adaround_init_kernel(const float* __restrict__ w, float* __restrict__ alpha,
                     const float* __restrict__ delta, int64_t n, int64_t channels, int64_t inner) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t c = (channels == 1) ? 0 : (i / inner) % channels;
    const float r = w[i] / __ldg(delta + c);
    const float rest = r - floorf(r);
    alpha[i] = -logf(1.2f / (rest - (-0.1f)) - 1.f);
  }
}

// rounding regulariser (block_recon.py:286-291): weight * sum(1 - |2h-1|^b); optional grad.
__global__ void __launch_bounds__(kThreads)
round_reg_kernel(const float* __restrict__ alpha, int64_t n, float b, float weight,
                 double* __restrict__ partials, float* __restrict__ galpha) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float a = alpha[i];
    const float s = sigmoidf_exact(a);
    const float hr = s * 1.2f + (-0.1f);
    const float h = fminf(fmaxf(hr, 0.f), 1.f);
    const float u = (h - 0.5f);
    const float t = fabsf(u) * 2.f;
    acc += (double)(1.f - powf(t, b));
    if (galpha) {
      const bool hin = (hr >= 0.f) && (hr <= 1.f);
      // d/dh (1 - (2|h-.5|)^b) = -b (2|u|)^(b-1) * 2 sign(u)
      float g = 0.f;
      if (hin && u != 0.f) g = -b * powf(t, b - 1.f) * 2.f * (u > 0.f ? 1.f : -1.f) * (1.2f * s * (1.f - s));
      galpha[i] += weight * g;
    }
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// ---------------------------------------------------------------------------------------------
// K4 L_p loss (quant_layer.py:26-33): sum(|pred-tgt|^p) * inv_rest, inv_rest = dim1 / numel
// ---------------------------------------------------------------------------------------------
template <bool P2>
__global__ void __launch_bounds__(kThreads)
lp_loss_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, int64_t n, float p,
                   double* __restrict__ partials) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  const int64_t nvec = n >> 2;
  int64_t done = 0;
  if ((((uintptr_t)pred | (uintptr_t)tgt) & 15) == 0) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
      const float4 a = __ldcs(reinterpret_cast<const float4*>(pred) + v);
      const float4 b = __ldcs(reinterpret_cast<const float4*>(tgt) + v);
      const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
      float s;
      if (P2) s = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
      else s = powf(fabsf(d0), p) + powf(fabsf(d1), p) + powf(fabsf(d2), p) + powf(fabsf(d3), p);
      acc += (double)s;
    }
    done = nvec << 2;
  }
  for (int64_t i = done + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = pred[i] - tgt[i];
    acc += (double)(P2 ? d * d : powf(fabsf(d), p));
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// gpred = gloss * inv_rest * p * |d|^(p-1) * sign(d)
template <bool P2>
__global__ void __launch_bounds__(kThreads)
lp_loss_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, int64_t n, float p,
                   float inv_rest, const float* __restrict__ gloss, float* __restrict__ gpred) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float k = __ldg(gloss) * inv_rest * p;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = pred[i] - tgt[i];
    float g;
    if (P2) g = k * d;
    else g = (d == 0.f) ? 0.f : k * powf(fabsf(d), p - 1.f) * (d > 0.f ? 1.f : -1.f);
    gpred[i] = g;
  }
}

}  // namespace edadm

using namespace edadm;

static QDrop make_qdrop(const uint8_t* mask, const float* rnd, float prob, uint64_t seed, uint64_t offset) {
  QDrop q;
  q.mask = mask;
  q.rnd = rnd;
  q.prob = prob;
  q.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  q.offset = offset;
  return q;
}

extern "C" int edadm_reduce_slots(void) { return sm_count() * 8; }

extern "C" int edadm_uaq_fwd(const float* x, float* y, uint8_t* codes, const float* delta,
                             const float* zero_point, int64_t n, int64_t channels, int64_t inner,
                             int n_levels, const uint8_t* keep_mask, const float* keep_rand, float qdrop_prob,
                             uint64_t seed, uint64_t offset, void* stream) {
  if (n == 0) return EDADM_OK;   // empty tensor (torch hands out a null data pointer for it): nothing to do
  if (!x || !y || !delta || !zero_point) return fail(EDADM_ERR_ARG, "uaq_fwd: null pointer");
  if (n < 0 || channels < 1 || inner < 1 || n_levels < 2 || n_levels > 256)
    return fail(EDADM_ERR_ARG, "uaq_fwd: bad sizes n=%lld channels=%lld inner=%lld levels=%d",
                (long long)n, (long long)channels, (long long)inner, n_levels);
  if (n == 0) return EDADM_OK;
  const QDrop qd = make_qdrop(keep_mask, keep_rand, qdrop_prob, seed, offset);
  const int grid = stream_grid((n + 3) / 4);
  const float qmax = (float)(n_levels - 1);
  cudaStream_t s = (cudaStream_t)stream;
  if (channels == 1)
    uaq_fwd_kernel<true><<<grid, kThreads, 0, s>>>(x, y, codes, delta, zero_point, n, 1, 1, qmax, qd);
  else
    uaq_fwd_kernel<false><<<grid, kThreads, 0, s>>>(x, y, codes, delta, zero_point, n, channels, inner, qmax, qd);
  return check_launch("uaq_fwd");
}

extern "C" int edadm_uaq_bwd(const float* gy, const float* x, const float* delta, const float* zero_point,
                             int64_t n, int64_t channels, int64_t inner, int n_levels,
                             const uint8_t* keep_mask, const float* keep_rand, float qdrop_prob, uint64_t seed,
                             uint64_t offset, float* gx, float* gdelta, int accumulate_gdelta, double* partials,
                             void* stream) {
  if (!gy || !x || !gx || !delta || !zero_point) return fail(EDADM_ERR_ARG, "uaq_bwd: null pointer");
  if (gdelta && channels != 1)
    return fail(EDADM_ERR_UNSUPPORTED, "uaq_bwd: step-size gradient only for per-tensor delta");
  if (gdelta && !partials) return fail(EDADM_ERR_ARG, "uaq_bwd: partials workspace required for gdelta");
  if (n < 0 || channels < 1 || inner < 1 || n_levels < 2 || n_levels > 256)
    return fail(EDADM_ERR_ARG, "uaq_bwd: bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  const QDrop qd = make_qdrop(keep_mask, keep_rand, qdrop_prob, seed, offset);
  const int grid = stream_grid((n + 3) / 4);
  const float qmax = (float)(n_levels - 1);
  if (n > 0) {
    if (channels == 1) {
      if (gdelta)
        uaq_bwd_kernel<true, true><<<grid, kThreads, 0, s>>>(gy, x, gx, delta, zero_point, n, 1, 1, qmax, qd, partials);
      else
        uaq_bwd_kernel<true, false><<<grid, kThreads, 0, s>>>(gy, x, gx, delta, zero_point, n, 1, 1, qmax, qd, nullptr);
    } else {
      uaq_bwd_kernel<false, false><<<grid, kThreads, 0, s>>>(gy, x, gx, delta, zero_point, n, channels, inner, qmax, qd, nullptr);
    }
  }
  if (gdelta) finish_sum_kernel<<<1, 256, 0, s>>>(partials, n > 0 ? grid : 0, gdelta, 1.0, accumulate_gdelta);
  return check_launch("uaq_bwd");
}

extern "C" int edadm_adaround_fwd(const float* w, const float* alpha, const float* delta,
                                  const float* zero_point, int64_t n, int64_t channels, int64_t inner,
                                  int n_levels, int soft, float* out, uint8_t* codes, void* stream) {
  if (!w || !alpha || !delta || !zero_point || !out) return fail(EDADM_ERR_ARG, "adaround_fwd: null pointer");
  if (n < 0 || channels < 1 || inner < 1 || n_levels < 2 || n_levels > 256)
    return fail(EDADM_ERR_ARG, "adaround_fwd: bad sizes");
  if (n == 0) return EDADM_OK;
  adaround_fwd_kernel<<<stream_grid(n), kThreads, 0, (cudaStream_t)stream>>>(
      w, alpha, out, codes, delta, zero_point, n, channels, inner, (float)(n_levels - 1), soft);
  return check_launch("adaround_fwd");
}

extern "C" int edadm_adaround_bwd(const float* gout, const float* w, const float* alpha, const float* delta,
                                  const float* zero_point, int64_t n, int64_t channels, int64_t inner,
                                  int n_levels, float* galpha, int accumulate, void* stream) {
  if (!gout || !w || !alpha || !delta || !zero_point || !galpha)
    return fail(EDADM_ERR_ARG, "adaround_bwd: null pointer");
  if (n < 0 || channels < 1 || inner < 1) return fail(EDADM_ERR_ARG, "adaround_bwd: bad sizes");
  if (n == 0) return EDADM_OK;
  adaround_bwd_kernel<<<stream_grid(n), kThreads, 0, (cudaStream_t)stream>>>(
      gout, w, alpha, galpha, delta, zero_point, n, channels, inner, (float)(n_levels - 1), accumulate);
  return check_launch("adaround_bwd");
}

extern "C" int edadm_adaround_init_alpha(const float* w, const float* delta, int64_t n, int64_t channels,
                                         int64_t inner, float* alpha, void* stream) {
  if (!w || !delta || !alpha) return fail(EDADM_ERR_ARG, "adaround_init_alpha: null pointer");
  if (n < 0 || channels < 1 || inner < 1) return fail(EDADM_ERR_ARG, "adaround_init_alpha: bad sizes");
  if (n == 0) return EDADM_OK;
  adaround_init_kernel<<<stream_grid(n), kThreads, 0, (cudaStream_t)stream>>>(w, alpha, delta, n, channels, inner);
  return check_launch("adaround_init_alpha");
}

extern "C" int edadm_round_reg(const float* alpha, int64_t n, float b, float weight, double* partials,
                               float* loss, int accumulate_loss, float* galpha, void* stream) {
  if (!alpha || !partials || !loss) return fail(EDADM_ERR_ARG, "round_reg: null pointer");
  if (n < 0) return fail(EDADM_ERR_ARG, "round_reg: bad size");
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = stream_grid(n);
  if (n > 0) round_reg_kernel<<<grid, kThreads, 0, s>>>(alpha, n, b, weight, partials, galpha);
  finish_sum_kernel<<<1, 256, 0, s>>>(partials, n > 0 ? grid : 0, loss, (double)weight, accumulate_loss);
  return check_launch("round_reg");
}

extern "C" int edadm_lp_loss_fwd(const float* pred, const float* tgt, int64_t n, float p, float inv_rest,
                                 double* partials, float* loss, void* stream) {
  if (!pred || !tgt || !partials || !loss) return fail(EDADM_ERR_ARG, "lp_loss_fwd: null pointer");
  if (n < 0) return fail(EDADM_ERR_ARG, "lp_loss_fwd: bad size");
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = stream_grid((n + 3) / 4);
  if (n > 0) {
    if (p == 2.0f) lp_loss_fwd_kernel<true><<<grid, kThreads, 0, s>>>(pred, tgt, n, p, partials);
    else lp_loss_fwd_kernel<false><<<grid, kThreads, 0, s>>>(pred, tgt, n, p, partials);
  }
  finish_sum_kernel<<<1, 256, 0, s>>>(partials, n > 0 ? grid : 0, loss, (double)inv_rest, 0);
  return check_launch("lp_loss_fwd");
}

extern "C" int edadm_lp_loss_bwd(const float* pred, const float* tgt, int64_t n, float p, float inv_rest,
                                 const float* gloss, float* gpred, void* stream) {
  if (!pred || !tgt || !gloss || !gpred) return fail(EDADM_ERR_ARG, "lp_loss_bwd: null pointer");
  if (n < 0) return fail(EDADM_ERR_ARG, "lp_loss_bwd: bad size");
  if (n == 0) return EDADM_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = stream_grid(n);
  if (p == 2.0f) lp_loss_bwd_kernel<true><<<grid, kThreads, 0, s>>>(pred, tgt, n, p, inv_rest, gloss, gpred);
  else lp_loss_bwd_kernel<false><<<grid, kThreads, 0, s>>>(pred, tgt, n, p, inv_rest, gloss, gpred);
  return check_launch("lp_loss_bwd");
}

// ------------------------------------------------------------------------------------------------
// K7  scale search: scores of ALL clipping candidates in one pass over the tensor
// (replaces the 100 (x chunked) fake-quant + reduce passes of UniformAffineQuantizer.perform_1D_search,
// qdiff/quant_layer.py:150-213, scored by lp_loss(x, Q(x), p=2.4, 'all') / the per-channel mean :63-70)
// ------------------------------------------------------------------------------------------------
namespace edadm {

constexpr int kSearchElems = 16;        // elements a thread keeps in registers per sweep over the candidates
constexpr int kSearchMaxCand = 128;

// x: [segments][inner]; candidate k of segment s quantizes with (delta[s*K+k], zp[s*K+k]).
// partial[(s * gridDim.x + block) * K + k] = this block's share of sum_i |Q(x_i) - x_i|^p (fp64; the caller divides by `inner`).
// grid = (blocks per segment, segments).  Every sum has a fixed order (per-warp slots, then warps, then blocks in
// mse_search_finish_kernel): the scores -- and with them the arg-min over candidates -- are reproducible run to run.
__global__ void __launch_bounds__(256)
mse_search_kernel(const float* __restrict__ x, long long inner, const float* __restrict__ delta, const float* __restrict__ zp,
                  int K, float qmax, float p, double* __restrict__ partial) {
  __shared__ double acc[8][kSearchMaxCand];
  __shared__ float sd[kSearchMaxCand], sz[kSearchMaxCand];
  const int seg = blockIdx.y;
  const float* xs = x + (long long)seg * inner;
  for (int k = threadIdx.x; k < K; k += blockDim.x) { sd[k] = __ldg(delta + (long long)seg * K + k); sz[k] = __ldg(zp + (long long)seg * K + k); }
  for (int i = threadIdx.x; i < 8 * kSearchMaxCand; i += blockDim.x) (&acc[0][0])[i] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long stride = (long long)gridDim.x * blockDim.x * kSearchElems;
  // the trip count is warp-uniform (the shuffles below use the full mask); lanes past the end carry zeros
  for (long long wb = ((long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * kSearchElems; wb < inner; wb += stride) {
    const long long base = wb + (long long)lane * kSearchElems;
    float v[kSearchElems];
#pragma unroll
    for (int j = 0; j < kSearchElems; ++j) v[j] = (base + j < inner) ? __ldg(xs + base + j) : 0.f;
    const int nvalid = (int)max(0LL, min((long long)kSearchElems, inner - base));
    for (int k = 0; k < K; ++k) {
      const float d = sd[k], z = sz[k];
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < kSearchElems; ++j) {
        const float q = fminf(fmaxf(rintf(v[j] / d) + z, 0.f), qmax);
        const float e = fabsf((q - z) * d - v[j]);
        s += (j < nvalid) ? (p == 2.0f ? e * e : powf(e, p)) : 0.f;
      }
      double sdbl = (double)s;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sdbl += __shfl_xor_sync(0xffffffffu, sdbl, o);
      if (lane == 0) acc[warp][k] += sdbl;                 // this warp's own slot
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += acc[w][k];
    partial[((long long)seg * gridDim.x + blockIdx.x) * K + k] = t;
  }
}

__global__ void __launch_bounds__(128)
mse_search_finish_kernel(const double* __restrict__ partial, int blocks, int K, double* __restrict__ scores) {
  const int seg = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    double t = 0.0;
    for (int b = 0; b < blocks; ++b) t += partial[((long long)seg * blocks + b) * K + k];
    scores[(long long)seg * K + k] += t;
  }
}

}  // namespace edadm

// scores[s*K + k] (fp64, zeroed by the caller) += sum over segment s of |Q_k(x) - x|^p with Q_k = clamp(round(x/delta)+zp, 0, L-1)
// dequantised, all K <= 128 candidates of a segment in one pass over x ([segments][inner] fp32).
extern "C" int edadm_mse_search_scores(const float* x, int64_t segments, int64_t inner, const float* delta, const float* zp, int K,
                                       int n_levels, float p, double* scores, void* stream) {
  using namespace edadm;
  if (!x || !delta || !zp || !scores) return fail(EDADM_ERR_ARG, "mse_search_scores: null pointer");
  if (segments < 1 || segments > 65535 || inner < 1 || K < 1 || K > kSearchMaxCand || n_levels < 2)
    return fail(EDADM_ERR_ARG, "mse_search_scores: bad sizes segments=%lld inner=%lld K=%d", (long long)segments, (long long)inner, K);
  long long per_seg = (inner + 256LL * kSearchElems - 1) / (256LL * kSearchElems);
  const long long cap = std::max<long long>(1, (long long)sm_count() * 8 / segments);
  if (per_seg > cap) per_seg = cap;
  dim3 grid((unsigned)per_seg, (unsigned)segments);
  cudaStream_t st = (cudaStream_t)stream;
  double* partial = nullptr;                     // stream-ordered scratch: [segments][blocks][K] block partials
  cudaError_t e = cudaMallocAsync((void**)&partial, sizeof(double) * (size_t)segments * per_seg * K, st);
  if (e != cudaSuccess) return fail(EDADM_ERR_CUDA, "mse_search_scores: scratch allocation failed: %s", cudaGetErrorString(e));
  mse_search_kernel<<<grid, 256, 0, st>>>(x, inner, delta, zp, K, (float)(n_levels - 1), p, partial);
  mse_search_finish_kernel<<<(unsigned)segments, 128, 0, st>>>(partial, (int)per_seg, K, scores);
  cudaFreeAsync(partial, st);
  return check_launch("mse_search_scores");
}
