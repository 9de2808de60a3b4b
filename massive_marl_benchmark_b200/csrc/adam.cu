// adam.cu - one optimiser step for MANY networks in two launches.
//
// Replaces, for the 10 agents x (actor, critic) of the MARL runs, the 20 separate `torch.optim.Adam` instances the
// reference builds (agents/algorithms/marl/mappo_policy.py:32-37, ippo_policy.py:39-45, happo_policy.py) and steps one after
// the other inside the per-agent loop (runner.py:266-317 -> mappo_trainer.py:143-170), each step preceded by
// `nn.utils.clip_grad_norm_` over that network (mappo_trainer.py:147-150,164-167): per network ~6 clip kernels + the ~9
// foreach kernels of Adam = ~300 launches per update for a TenAnt team.  Here every parameter, gradient and moment of
// every network lives in ONE flat fp32 buffer each (grouped_adam.GroupedAdam re-points the modules' parameters at views of
// it), a group = one network = one contiguous slice, and an update is:
//   mmb_grad_sumsq_group   per-group sum of squares of the gradients (fp64 accumulation, one atomic per CTA)
//   mmb_adam_group         clip coefficient from that sum, weight decay, both moments, bias corrections, parameter update
// The flat gradient buffer is also the NCCL all-reduce buffer of the env-sharded runs (dist.GradBuckets): no flatten /
// unflatten copies.
//
// Arithmetic = torch.optim.Adam (foreach path, amsgrad off, maximize off) after clip_grad_norm_(max_norm, 2):
//   total_norm = ||g_group||_2;  coef = min(max_norm / (total_norm + 1e-6), 1);  g = g * coef        (clip; skipped if max_norm <= 0)
//   g = g + weight_decay * p                                                              (if weight_decay != 0)
//   m = m + (g - m) * (1 - beta1);   v = v * beta2 + (1 - beta2) * g * g
//   p = p - (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps),   bc1 = 1 - beta1^step, bc2 = 1 - beta2^step (host doubles)
#include "../../include/mmb.h"
#include "mmb_common.cuh"
#include "mmb_math.cuh"

namespace mmb {
namespace {

__device__ __forceinline__ int group_of(const mmb_adam_params& p, int64_t i) {
  int g = 0;                                  // <= MMB_ADAM_MAX_GROUPS boundaries: a short scan per thread, done once
  while (g + 1 < p.num_groups && i >= p.group_start[g + 1]) ++g;
  return g;
}

// grid = (blocks per group, groups): every CTA reduces a strided part of ONE group's slice
__global__ void __launch_bounds__(256) grad_sumsq_kernel(const __grid_constant__ mmb_adam_params p) {
  __shared__ double sh[8];
  const int g = blockIdx.y;
  const int64_t lo = p.group_start[g], hi = (g + 1 < p.num_groups) ? p.group_start[g + 1] : p.total;
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // group slices start at multiples of 4 elements (GroupedAdam pads): 128-bit loads
  const int64_t n4 = (hi - lo) >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(p.grads + lo);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(g4 + i);
    acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  for (int64_t i = lo + (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
    const float v = __ldg(p.grads + i);
    acc += (double)v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    atomicAdd(p.sumsq + g, t);
  }
}

__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ mmb_adam_params p) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < (p.total >> 2); q += stride) {
    const int64_t i = q << 2;
    const int g = group_of(p, i);             // a 4-element quad never straddles groups (slices are padded to 4)
    float coef = 1.0f;
    if (p.max_grad_norm[g] > 0.0f) {          // clip_grad_norm_: clamp(max_norm / (total_norm + 1e-6), max = 1)
      const float total_norm = (float)sqrt(p.sumsq[g]);
      coef = fminf(fdiv(p.max_grad_norm[g], fadd(total_norm, 1e-6f)), 1.0f);
    }
    const float lr_bc1 = p.step_size[g], bc2_sqrt = p.bc2_sqrt[g], eps = p.eps[g], wd = p.weight_decay[g];
    const float b1w = p.one_minus_beta1, b2 = p.beta2, b2w = p.one_minus_beta2;
    float4 pv = *reinterpret_cast<const float4*>(p.params + i);
    const float4 gv = __ldg(reinterpret_cast<const float4*>(p.grads + i));
    float4 mv = *reinterpret_cast<const float4*>(p.exp_avg + i);
    float4 vv = *reinterpret_cast<const float4*>(p.exp_avg_sq + i);
    float* pp = reinterpret_cast<float*>(&pv);
    const float* gp = reinterpret_cast<const float*>(&gv);
    float* mp = reinterpret_cast<float*>(&mv);
    float* vp = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // the multiply-adds are fused where torch's own elementwise kernels contract them (nvcc -fmad on `a + s * b`)
      float gr = fmul(gp[k], coef);
      if (wd != 0.0f) gr = __fmaf_rn(wd, pp[k], gr);                           // grad.add(param, alpha = weight_decay)
      mp[k] = __fmaf_rn(b1w, fsub(gr, mp[k]), mp[k]);                          // lerp(m, g, 1 - beta1), |weight| < 0.5 branch
      vp[k] = __fmaf_rn(fmul(b2w, gr), gr, fmul(vp[k], b2));                   // mul_(beta2).addcmul_(g, g, value = 1 - beta2)
      const float denom = fadd(fdiv(fsqrt(vp[k]), bc2_sqrt), eps);             // sqrt().div_(sqrt(bc2)).add_(eps)
      pp[k] = __fmaf_rn(-lr_bc1, fdiv(mp[k], denom), pp[k]);                   // addcdiv_(m, denom, value = -step_size)
    }
    *reinterpret_cast<float4*>(p.params + i) = pv;
    *reinterpret_cast<float4*>(p.exp_avg + i) = mv;
    *reinterpret_cast<float4*>(p.exp_avg_sq + i) = vv;
  }
}

inline bool adam_valid(const mmb_adam_params& p) {
  if (p.num_groups < 1 || p.num_groups > MMB_ADAM_MAX_GROUPS || p.total <= 0 || (p.total & 3)) return false;
  if (!p.params || !p.grads || !p.exp_avg || !p.exp_avg_sq || !p.sumsq) return false;
  if ((reinterpret_cast<uintptr_t>(p.params) | reinterpret_cast<uintptr_t>(p.grads) | reinterpret_cast<uintptr_t>(p.exp_avg) |
       reinterpret_cast<uintptr_t>(p.exp_avg_sq)) & 15u)
    return false;
  if (p.group_start[0] != 0) return false;
  for (int g = 0; g < p.num_groups; ++g) {
    const int64_t hi = (g + 1 < p.num_groups) ? p.group_start[g + 1] : p.total;
    if ((p.group_start[g] & 3) || hi <= p.group_start[g]) return false;
  }
  return true;
}

}  // namespace
}  // namespace mmb

using namespace mmb;

extern "C" int32_t mmb_grad_sumsq_group(const mmb_adam_params* pp, void* stream) {
  if (!pp) return MMB_EINVAL;
  mmb_adam_params p = *pp;
  if (!adam_valid(p)) return MMB_EINVAL;
  int64_t longest = 0;
  for (int g = 0; g < p.num_groups; ++g) {
    const int64_t hi = (g + 1 < p.num_groups) ? p.group_start[g + 1] : p.total;
    if (hi - p.group_start[g] > longest) longest = hi - p.group_start[g];
  }
  int64_t bx = (longest / 4 + 256 * 4 - 1) / (256 * 4);
  const int64_t cap = (sm_count() * 8 + p.num_groups - 1) / p.num_groups;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  {
    LaunchScope ls(K_ADAM_NORM, (cudaStream_t)stream);
    grad_sumsq_kernel<<<dim3((unsigned)bx, p.num_groups), 256, 0, (cudaStream_t)stream>>>(p);
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

extern "C" int32_t mmb_adam_group(const mmb_adam_params* pp, void* stream) {
  if (!pp) return MMB_EINVAL;
  mmb_adam_params p = *pp;
  if (!adam_valid(p)) return MMB_EINVAL;
  int64_t bx = (p.total / 4 + 255) / 256;
  if (bx > sm_count() * 8) bx = sm_count() * 8;
  {
    LaunchScope ls(K_ADAM, (cudaStream_t)stream);
    adam_kernel<<<(unsigned)bx, 256, 0, (cudaStream_t)stream>>>(p);
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}
