// K6: the two Adam updates of a reconstruction iteration as ONE streaming pass.
//
// The reference loop (qdiff/block_recon.py:113-117, 199-206; layer_recon.py:82-86; attn_layer_recon.py:68-72) keeps two
// torch.optim.Adam instances -- AdaRound alphas at lr_w, activation step sizes at lr_a, both cosine-annealed -- and calls
// .step() on each after the backward (and, data parallel, after the gradient all-reduce).  On the GPU that is ~14
// multi-tensor launches moving every alpha-sized tensor about ten times.  Here the gradients already sit in one flat bucket
// (qdiff/dist.py GradBucket: every .grad is a view into it), the moments are flat buffers with the same offsets, and a
// segment table maps flat ranges back to the parameter tensors, so one launch reads g, p, m, v and writes p, m, v (and the
// zeroed gradient for the next backward): 32 B per parameter element, HBM bound.
//
// Arithmetic = torch.optim.Adam's single-tensor form (torch/optim/adam.py _single_tensor_adam, amsgrad=False,
// weight_decay=0, maximize=False), one rounding per tensor op (this file is compiled with -fmad=false):
//   m     = m + (1-b1) * (g - m)                      exp_avg.lerp_(grad, 1 - beta1)
//   v     = v * b2 + ((1-b2) * g) * g                 exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
//   denom = sqrt(v) / sqrt(1 - b2^t) + eps
//   p     = p + (-(lr / (1 - b1^t))) * (m / denom)    param.addcdiv_(exp_avg, denom, value=-step_size)
// with the bias corrections and lr / (1 - b1^t) evaluated in double like the Python scalars they are in torch.
// lr and t live on the device (the captured CUDA graph of the iteration sees the schedule and the step count).
#include "common.cuh"

namespace edadm {

struct AdamSegment {
  float* param;      // first element of this segment inside its parameter tensor
  int64_t offset;    // the same element's index in the flat gradient / moment buffers
  int32_t count;     // elements in the segment (<= kAdamSegment)
  int32_t group;     // 0: lr[0] (AdaRound alphas), 1: lr[1] (activation step sizes)
};
static_assert(sizeof(AdamSegment) == 24, "host table layout (qdiff/_fused_adam.py) is 3 x int64");

__global__ void __launch_bounds__(kThreads)
fused_adam_kernel(const AdamSegment* __restrict__ segments, float* __restrict__ grad, float* __restrict__ exp_avg,
                  float* __restrict__ exp_avg_sq, const float* __restrict__ lr0, const float* __restrict__ lr1,
                  const int64_t* __restrict__ step, double beta1, double beta2, float eps, int zero_grad) {
  const AdamSegment seg = segments[blockIdx.x];
  __shared__ float s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    const double t = (double)step[0];
    const double lr = (double)(seg.group ? lr1[0] : lr0[0]);
    s_step_size = (float)(lr / (1.0 - pow(beta1, t)));
    s_bc2_sqrt = (float)sqrt(1.0 - pow(beta2, t));
  }
  __syncthreads();
  const float step_size_neg = -s_step_size, bc2_sqrt = s_bc2_sqrt;
  const float w1 = (float)(1.0 - beta1), b2 = (float)beta2, w2 = (float)(1.0 - beta2);
  float* __restrict__ p = seg.param;
  float* __restrict__ g = grad + seg.offset;
  float* __restrict__ m = exp_avg + seg.offset;
  float* __restrict__ v = exp_avg_sq + seg.offset;

  auto update = [&](float& pi, float gi, float& mi, float& vi) {
    mi = mi + w1 * (gi - mi);
    vi = vi * b2 + (w2 * gi) * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi = pi + step_size_neg * (mi / denom);
  };

  const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
  int done = 0;
  if (vec) {
    const int nvec = seg.count >> 2;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      float4 pv = reinterpret_cast<float4*>(p)[i];
      const float4 gv = reinterpret_cast<const float4*>(g)[i];
      float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
      update(pv.x, gv.x, mv.x, vv.x);
      update(pv.y, gv.y, mv.y, vv.y);
      update(pv.z, gv.z, mv.z, vv.z);
      update(pv.w, gv.w, mv.w, vv.w);
      reinterpret_cast<float4*>(p)[i] = pv;
      reinterpret_cast<float4*>(m)[i] = mv;
      reinterpret_cast<float4*>(v)[i] = vv;
      if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    done = nvec << 2;
  }
  for (int i = done + threadIdx.x; i < seg.count; i += blockDim.x) {
    float pi = p[i], mi = m[i], vi = v[i];
    update(pi, g[i], mi, vi);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
    if (zero_grad) g[i] = 0.f;
  }
}

}  // namespace edadm

// One Adam step over every parameter listed in `segments` (device table of n_segments {param pointer, flat offset, count,
// group}; a parameter tensor is cut into segments of at most 2^31-1 elements, the caller uses 8192).  grad / exp_avg /
// exp_avg_sq are the flat buffers; lr points to two device floats (group 0, group 1); step to the device int64 step count t >= 1.
// zero_grad != 0: the consumed gradient is overwritten with 0 (the next backward accumulates into it).
extern "C" int edadm_fused_adam(const void* segments, int n_segments, float* grad, float* exp_avg, float* exp_avg_sq,
                                const float* lr, const int64_t* step, double beta1, double beta2, float eps, int zero_grad,
                                void* stream) {
  using namespace edadm;
  if (n_segments == 0) return EDADM_OK;
  if (!segments || !grad || !exp_avg || !exp_avg_sq || !lr || !step) return fail(EDADM_ERR_ARG, "fused_adam: null pointer");
  if (n_segments < 0 || !(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0) || !(eps >= 0.f))
    return fail(EDADM_ERR_ARG, "fused_adam: bad arguments n_segments=%d beta1=%g beta2=%g eps=%g", n_segments, beta1, beta2, (double)eps);
  fused_adam_kernel<<<(unsigned)n_segments, kThreads, 0, (cudaStream_t)stream>>>(
      (const AdamSegment*)segments, grad, exp_avg, exp_avg_sq, lr, lr + 1, step, beta1, beta2, eps, zero_grad);
  return check_launch("fused_adam");
}
