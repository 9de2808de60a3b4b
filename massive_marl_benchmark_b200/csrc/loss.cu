// loss.cu - the PPO minibatch loss (forward value + the gradients autograd would produce) in one launch.
//
// Replaces, after the two MLPs have produced the action mean and the value of a minibatch
// (agents/algorithms/rl/ppo/ppo.py:266-302, module.py:92-107):
//   MultivariateNormal(mean, scale_tril = diag(exp(log_std)^2)).log_prob / .entropy   (effective std = sigma^2)
//   the adaptive-schedule KL estimate (ppo.py:271-275)
//   ratio / clipped surrogate (ppo.py:286-290), clipped value loss (ppo.py:293-300), total loss (ppo.py:302)
// and their backward pass down to d loss / d mean, d loss / d log_std and d loss / d value - about forty elementwise /
// reduction kernels in torch.  HBM-bound: 5 rows of A floats per sample (mean, actions, old mean, old sigma in, mean
// gradient out) + 6 scalars.  Gradient rules follow torch's: `max(a, b)` splits the gradient evenly on exact ties,
// `clamp` passes it on the closed interval.
#include "../../include/mmb.h"
#include "mmb_common.cuh"
#include "mmb_math.cuh"

namespace mmb {
namespace {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_MAX_K = 8;           // columns per lane: act_dim <= 32 * 8
constexpr float LOG_2PI = 1.8378770664093453f;

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// G lanes per minibatch row (8 / 16 / 32 by action width); lane `sub` owns the columns sub, sub + G, ...
template <int G, int KMAX>
__global__ void __launch_bounds__(LOSS_THREADS) ppo_loss_kernel(const __grid_constant__ mmb_ppo_loss_params p) {
  __shared__ double s_gls[32 * LOSS_MAX_K];
  __shared__ double s_sum[3];
  const int A = p.act_dim, B = p.num_rows;
  const int tid = threadIdx.x, sub = tid % G;
  for (int j = tid; j < A; j += LOSS_THREADS) s_gls[j] = 0.0;
  if (tid < 3) s_sum[tid] = 0.0;
  __syncthreads();

  // per-column constants of the new policy
  float ls[KMAX], sd[KMAX], two_var[KMAX], gls[KMAX];
  float hld_part = 0.0f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int j = sub + k * G;
    gls[k] = 0.0f;
    if (j < A) {
      ls[k] = __ldg(p.log_std + j);
      const float e = expf(ls[k]);
      sd[k] = e * e;                     // module.py:95: diag(exp(log_std) * exp(log_std)) used as scale_tril
      two_var[k] = 2.0f * (e * e);       // ppo.py:273: 2.0 * square(sigma.exp())
      hld_part += logf(sd[k]);
    } else {
      ls[k] = 0.0f; sd[k] = 1.0f; two_var[k] = 1.0f;
    }
  }
  const float half_log_det = group_sum<G>(hld_part);
  const float inv_b = 1.0f / (float)B;

  double acc_s = 0.0, acc_v = 0.0, acc_kl = 0.0;
  const int rows_per_block = LOSS_THREADS / G;
  // warp-uniform trip count (the group reductions below shuffle with the full mask); rows past the end are predicated off
  for (int base = blockIdx.x * rows_per_block; base < B; base += gridDim.x * rows_per_block) {
    const int row = base + tid / G;
    const bool live = row < B;
    const int A_live = live ? A : 0;
    const float* mu = p.mu + (int64_t)row * p.mu_stride;
    const float* act = p.actions + (int64_t)row * A;
    float z[KMAX];
    float maha = 0.0f, kl = 0.0f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int j = sub + k * G;
      z[k] = 0.0f;
      if (j < A_live) {
        const float m = __ldg(mu + j);
        z[k] = (__ldg(act + j) - m) / sd[k];
        maha = __fadd_rn(maha, __fmul_rn(z[k], z[k]));       // pow(2) then sum, as torch rounds them (no FMA contraction)
        if (p.old_mu) {
          const float om = __ldcs(p.old_mu + (int64_t)row * A + j), os = __ldcs(p.old_sigma + (int64_t)row * A + j);
          const float eo = expf(os), d = om - m;
          // every term is a difference of O(1) numbers: keep the reference's rounding sequence (products rounded before the add)
          const float t = __fadd_rn(__fadd_rn(ls[k], -os), __fdiv_rn(__fadd_rn(__fmul_rn(eo, eo), __fmul_rn(d, d)), two_var[k]));
          kl = __fadd_rn(kl, __fadd_rn(t, -0.5f));
        }
      }
    }
    maha = group_sum<G>(maha);
    if (p.old_mu) kl = group_sum<G>(kl);
    const float logp = __fadd_rn(__fmul_rn(-0.5f, __fadd_rn(p.k_log_2pi, maha)), -half_log_det);

    // surrogate (ppo.py:286-290)
    const float adv = live ? __ldg(p.advantages + row) : 0.0f;
    const float ratio = expf(logp - (live ? __ldg(p.old_logp + row) : 0.0f));
    const float clamped = fminf(fmaxf(ratio, p.ratio_lo), p.ratio_hi);
    const float surr = -adv * ratio, surr_c = -adv * clamped;
    const bool inside = ratio >= p.ratio_lo && ratio <= p.ratio_hi;
    const float d_ratio = -adv * ratio;                 // d surr / d logp; also d surr_c / d logp where the clamp passes
    float g;
    if (surr > surr_c) g = d_ratio;
    else if (surr == surr_c) g = 0.5f * d_ratio + (inside ? 0.5f * d_ratio : 0.0f);
    else g = inside ? d_ratio : 0.0f;
    g = live ? g * inv_b : 0.0f;          // (a dead row's ratio may overflow: keep 0 * inf out of the column sums)

    if (p.grad_mu) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int j = sub + k * G;
        if (j < A_live) __stcs(p.grad_mu + (int64_t)row * A + j, g * (z[k] / sd[k]));
      }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) gls[k] += g * (2.0f * z[k] * z[k] - 2.0f);   // d logp / d log_std = 2 z^2 - 2

    if (sub == 0 && live) {
      // value loss (ppo.py:293-300)
      const float v = __ldg(p.value + row), ret = __ldg(p.returns + row);
      float lv, gv;
      if (p.use_clipped_value_loss) {
        const float tv = __ldg(p.target_values + row);
        const float d = v - tv;
        const float dc = fminf(fmaxf(d, -p.clip_param), p.clip_param);
        const float vc = tv + dc;
        const float e1 = v - ret, e2 = vc - ret;
        const float l1 = e1 * e1, l2 = e2 * e2;
        const bool in_v = d >= -p.clip_param && d <= p.clip_param;
        const float g1 = 2.0f * e1, g2 = in_v ? 2.0f * e2 : 0.0f;
        lv = fmaxf(l1, l2);
        gv = l1 > l2 ? g1 : (l1 == l2 ? 0.5f * g1 + 0.5f * g2 : g2);
      } else {
        const float e = ret - v;
        lv = e * e;
        gv = -2.0f * e;
      }
      if (p.grad_value) p.grad_value[row] = p.value_loss_coef * gv * inv_b;
      if (p.logp) p.logp[row] = logp;
      acc_s += (double)fmaxf(surr, surr_c);
      acc_v += (double)lv;
      acc_kl += (double)kl;
    }
  }

  // block totals, then one double atomic per slot per block
  if (sub == 0) {
    atomicAdd(&s_sum[0], acc_s); atomicAdd(&s_sum[1], acc_v); atomicAdd(&s_sum[2], acc_kl);
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int j = sub + k * G;
    if (j < A) atomicAdd(&s_gls[j], (double)gls[k]);
  }
  __syncthreads();
  if (tid < 3) atomicAdd(p.sums + tid, s_sum[tid]);
  for (int j = tid; j < A; j += LOSS_THREADS) {
    double v = s_gls[j];
    if (blockIdx.x == 0) v -= 2.0 * (double)p.entropy_coef;     // entropy = const + sum_j 2 log_std_j, equal on every row
    atomicAdd(p.sums + 4 + j, v);
  }
  const double entropy = 0.5 * (double)A * (1.0 + (double)LOG_2PI) + (double)half_log_det;
  if (p.out == nullptr) {
    if (blockIdx.x == 0 && tid == 0) p.sums[3] = entropy;
    return;
  }
  // Finalisation by the last block to finish (ticket): the fp32 terms torch would compute from the sums
  // (ppo.py:285-302: means over the batch, total loss) - and the scratch is handed back zeroed.
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  volatile double* sums = p.sums;
  if (tid == 0) {
    const float s0 = (float)sums[0], s1 = (float)sums[1], s2 = (float)sums[2], ent = (float)entropy;
    const float fb = (float)p.num_rows;
    p.out[0] = s0 * (1.0f / fb) + s1 * (p.value_loss_coef / fb) - p.entropy_coef * ent;
    p.out[1] = s0 / fb; p.out[2] = s1 / fb; p.out[3] = s2 / fb; p.out[4] = ent;
  }
  for (int j = tid; j < A; j += LOSS_THREADS) p.out[5 + j] = (float)sums[4 + j];
  __syncthreads();
  for (int j = tid; j < 4 + A; j += LOSS_THREADS) p.sums[j] = 0.0;
  if (tid == 0) *p.ticket = 0u;
}

template <int G>
cudaError_t launch_loss(const mmb_ppo_loss_params& p, int kmax, int grid, cudaStream_t st) {
  if (G < 32) {                          // narrow groups cover the whole row with one column per lane
    ppo_loss_kernel<G, 1><<<grid, LOSS_THREADS, 0, st>>>(p);
    return cudaGetLastError();
  }
  switch (kmax) {
    case 1: ppo_loss_kernel<32, 1><<<grid, LOSS_THREADS, 0, st>>>(p); break;
    case 2: ppo_loss_kernel<32, 2><<<grid, LOSS_THREADS, 0, st>>>(p); break;
    case 3: ppo_loss_kernel<32, 3><<<grid, LOSS_THREADS, 0, st>>>(p); break;
    case 4: ppo_loss_kernel<32, 4><<<grid, LOSS_THREADS, 0, st>>>(p); break;
    default: ppo_loss_kernel<32, LOSS_MAX_K><<<grid, LOSS_THREADS, 0, st>>>(p); break;
  }
  return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------------------------------
// MAPPO minibatch losses (agents/algorithms/marl/mappo_trainer.py:62-103,127-168) for one agent: per-dimension Normal
// log-probs (distributions.py:32-35), importance weight exp(sum_j (logp_j - old_logp_j)), clipped surrogate, huber / mse
// value loss on (PopArt-)normalised returns, optional active masks - and the gradients of the policy loss with respect to
// mean / std and of the value loss with respect to the values.  Same thread layout as ppo_loss_kernel.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void value_term(float e, float delta, bool huber, float& loss, float& dloss_de) {
  if (huber) {
    // agents/utils/util.py:23-26: a = |e| <= d, b = e > d (no branch for e < -d: zero loss, zero gradient - kept)
    const float a = fabsf(e) <= delta ? 1.0f : 0.0f, b = e > delta ? 1.0f : 0.0f;
    loss = a * (e * e) / 2.0f + (b * delta) * (fabsf(e) - delta / 2.0f);
    dloss_de = a * e + b * delta;
  } else {
    loss = (e * e) / 2.0f;                                      // util.py:28-29
    dloss_de = e;
  }
}

template <int G, int KMAX>
__global__ void __launch_bounds__(LOSS_THREADS) mappo_loss_kernel(const __grid_constant__ mmb_mappo_loss_params p) {
  __shared__ double s_gstd[32 * LOSS_MAX_K];
  __shared__ double s_sum[2];
  const int A = p.act_dim, B = p.num_rows;
  const int tid = threadIdx.x, sub = tid % G;
  for (int j = tid; j < A; j += LOSS_THREADS) s_gstd[j] = 0.0;
  if (tid < 2) s_sum[tid] = 0.0;
  __syncthreads();

  float sd[KMAX], two_var[KMAX], log_sd[KMAX], gstd[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int j = sub + k * G;
    gstd[k] = 0.0f;
    sd[k] = j < A ? __ldg(p.std + j) : 1.0f;
    two_var[k] = 2.0f * (sd[k] * sd[k]);                         // Normal.log_prob: var = scale ** 2; ... / (2 * var)
    log_sd[k] = logf(sd[k]);
  }
  const float inv_b = 1.0f / (float)B;
  const float inv_mask_sum = p.mask_sum ? 1.0f / __ldg(p.mask_sum) : 0.0f;
  // the reference calls its PopArt normaliser once per error term and each call first updates the running moments
  // (mappo_trainer.py:80-81, popart.py:38-60): the clipped and the original error may see different (mean, var)
  float ret_mean = 0.0f, ret_sd = 1.0f, ret_mean_o = 0.0f, ret_sd_o = 1.0f;
  const bool norm_ret = p.ret_mean != nullptr;
  if (norm_ret) {
    ret_mean = __ldg(p.ret_mean); ret_sd = sqrtf(__ldg(p.ret_var));
    ret_mean_o = p.ret_mean_orig ? __ldg(p.ret_mean_orig) : ret_mean;
    ret_sd_o = p.ret_var_orig ? sqrtf(__ldg(p.ret_var_orig)) : ret_sd;
  }

  double acc_p = 0.0, acc_v = 0.0;
  const int rows_per_block = LOSS_THREADS / G;
  for (int base = blockIdx.x * rows_per_block; base < B; base += gridDim.x * rows_per_block) {
    const int row = base + tid / G;
    const bool live = row < B;
    const int A_live = live ? A : 0;
    const float* mean = p.mean + (int64_t)row * p.mean_stride;
    float diff[KMAX];
    float dsum = 0.0f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int j = sub + k * G;
      diff[k] = 0.0f;
      if (j < A_live) {
        diff[k] = __ldg(p.actions + (int64_t)row * A + j) - __ldg(mean + j);
        // -((a - m) ** 2) / (2 var) - log(scale) - log(sqrt(2 pi)), rounded like torch's separate kernels
        const float lp = __fadd_rn(__fadd_rn(__fdiv_rn(-__fmul_rn(diff[k], diff[k]), two_var[k]), -log_sd[k]), -0.9189385332046727f);
        if (p.logp) __stcs(p.logp + (int64_t)row * A + j, lp);
        dsum = __fadd_rn(dsum, __fadd_rn(lp, -__ldcs(p.old_logp + (int64_t)row * A + j)));
      }
    }
    dsum = group_sum<G>(dsum);
    const float imp = expf(dsum);                                // mappo_trainer.py:128
    const float active = (live && p.active_masks) ? __ldg(p.active_masks + row) : 1.0f;
    const float adv = live ? __ldg(p.adv_targ + row) : 0.0f;
    const float clamped = fminf(fmaxf(imp, p.ratio_lo), p.ratio_hi);
    const float surr1 = imp * adv, surr2 = clamped * adv;
    const bool inside = imp >= p.ratio_lo && imp <= p.ratio_hi;
    float dmin;                                                  // d min(surr1, surr2) / d imp
    if (surr1 < surr2) dmin = adv;
    else if (surr1 == surr2) dmin = 0.5f * adv + (inside ? 0.5f * adv : 0.0f);
    else dmin = inside ? adv : 0.0f;
    const float w_p = p.use_policy_active_masks ? active * inv_mask_sum : inv_b;
    const float gl = live ? -dmin * imp * w_p : 0.0f;           // d policy_loss / d logp_j, the same for every j
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int j = sub + k * G;
      const float var = sd[k] * sd[k];
      if (j < A_live && p.grad_mean) __stcs(p.grad_mean + (int64_t)row * A + j, gl * (diff[k] / var));
      gstd[k] += gl * ((diff[k] * diff[k]) / (var * sd[k]) - 1.0f / sd[k]);
    }

    if (sub == 0 && live) {
      if (p.imp_weights) p.imp_weights[row] = imp;
      acc_p += (double)(-fminf(surr1, surr2) * (p.use_policy_active_masks ? active : 1.0f));
      // cal_value_loss, mappo_trainer.py:73-103
      const float v = __ldg(p.values + row), vp = __ldg(p.value_preds + row);
      const float ret_raw = __ldg(p.returns + row);
      float ret_c = ret_raw, ret_o = ret_raw;
      if (norm_ret) { ret_c = (ret_raw - ret_mean) / ret_sd; ret_o = (ret_raw - ret_mean_o) / ret_sd_o; }   // popart.py:59-60
      const float d = v - vp;
      const float vpc = vp + fminf(fmaxf(d, -p.clip_param), p.clip_param);
      const bool in_v = d >= -p.clip_param && d <= p.clip_param;
      float lo, go, lc, gc;
      value_term(ret_o - v, p.huber_delta, p.use_huber_loss != 0, lo, go);
      value_term(ret_c - vpc, p.huber_delta, p.use_huber_loss != 0, lc, gc);
      go = -go;                                                 // d error / d values = -1 (through the clamp: where it passes)
      gc = in_v ? -gc : 0.0f;
      float lv = lo, gv = go;
      if (p.use_clipped_value_loss) {
        lv = fmaxf(lo, lc);
        gv = lo > lc ? go : (lo == lc ? 0.5f * go + 0.5f * gc : gc);
      }
      const float w_v = p.use_value_active_masks ? active * inv_mask_sum : inv_b;
      if (p.grad_values) p.grad_values[row] = gv * w_v;
      acc_v += (double)(lv * (p.use_value_active_masks ? active : 1.0f));
    }
  }

  if (sub == 0) { atomicAdd(&s_sum[0], acc_p); atomicAdd(&s_sum[1], acc_v); }
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int j = sub + k * G;
    if (j < A) atomicAdd(&s_gstd[j], (double)gstd[k]);
  }
  __syncthreads();
  if (tid < 2) atomicAdd(p.sums + tid, s_sum[tid]);
  for (int j = tid; j < A; j += LOSS_THREADS) atomicAdd(p.sums + 2 + j, s_gstd[j]);
  if (p.out == nullptr) return;
  // Finalisation by the last block to finish (ticket): the two means over the reference's denominators
  // (mappo_trainer.py:139-144, 98-101: .sum() / active_masks.sum() or .mean()) - and the scratch is handed back zeroed.
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  volatile double* sums = p.sums;
  if (tid == 0) {
    const double msum = p.mask_sum ? (double)__ldg(p.mask_sum) : 0.0;
    p.out[0] = (float)(sums[0] / (p.use_policy_active_masks ? msum : (double)B));
    p.out[1] = (float)(sums[1] / (p.use_value_active_masks ? msum : (double)B));
  }
  for (int j = tid; j < A; j += LOSS_THREADS) p.out[2 + j] = (float)sums[2 + j];
  __syncthreads();
  for (int j = tid; j < 2 + A; j += LOSS_THREADS) p.sums[j] = 0.0;
  if (tid == 0) *p.ticket = 0u;
}

template <int G>
cudaError_t launch_mappo_loss(const mmb_mappo_loss_params& p, int kmax, int grid, cudaStream_t st) {
  if (G < 32) {
    mappo_loss_kernel<G, 1><<<grid, LOSS_THREADS, 0, st>>>(p);
    return cudaGetLastError();
  }
  switch (kmax) {
    case 1: mappo_loss_kernel<32, 1><<<grid, LOSS_THREADS, 0, st>>>(p); break;
    case 2: mappo_loss_kernel<32, 2><<<grid, LOSS_THREADS, 0, st>>>(p); break;
    case 3: mappo_loss_kernel<32, 3><<<grid, LOSS_THREADS, 0, st>>>(p); break;
    case 4: mappo_loss_kernel<32, 4><<<grid, LOSS_THREADS, 0, st>>>(p); break;
    default: mappo_loss_kernel<32, LOSS_MAX_K><<<grid, LOSS_THREADS, 0, st>>>(p); break;
  }
  return cudaGetLastError();
}

}  // namespace
}  // namespace mmb

using namespace mmb;

extern "C" int32_t mmb_ppo_loss(const mmb_ppo_loss_params* pp, void* stream) {
  if (!pp) return MMB_EINVAL;
  mmb_ppo_loss_params p = *pp;
  if (p.num_rows <= 0 || p.act_dim <= 0 || p.act_dim > 32 * LOSS_MAX_K || p.mu_stride < p.act_dim) return MMB_EINVAL;
  if (!p.mu || !p.log_std || !p.actions || !p.old_logp || !p.advantages || !p.value || !p.returns || !p.sums) return MMB_EINVAL;
  if (p.use_clipped_value_loss && !p.target_values) return MMB_EINVAL;
  if ((p.old_mu == nullptr) != (p.old_sigma == nullptr)) return MMB_EINVAL;
  if (p.out && !p.ticket) return MMB_EINVAL;
  p.k_log_2pi = (float)((double)p.act_dim * 1.8378770664093453);   // torch: the Python double k * log(2 pi), then cast
  cudaStream_t st = (cudaStream_t)stream;
  const int G = p.act_dim <= 8 ? 8 : (p.act_dim <= 16 ? 16 : 32);
  const int kmax = (p.act_dim + G - 1) / G;
  const int rows_per_block = LOSS_THREADS / G;
  int grid = (p.num_rows + rows_per_block - 1) / rows_per_block;
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  cudaError_t e;
  {
    LaunchScope ls(K_PPO_LOSS, st);
    e = G == 8 ? launch_loss<8>(p, kmax, grid, st) : (G == 16 ? launch_loss<16>(p, kmax, grid, st) : launch_loss<32>(p, kmax, grid, st));
  }
  return e == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

extern "C" int32_t mmb_mappo_loss(const mmb_mappo_loss_params* pp, void* stream) {
  if (!pp) return MMB_EINVAL;
  const mmb_mappo_loss_params& p = *pp;
  if (p.num_rows <= 0 || p.act_dim <= 0 || p.act_dim > 32 * LOSS_MAX_K || p.mean_stride < p.act_dim) return MMB_EINVAL;
  if (!p.mean || !p.std || !p.actions || !p.old_logp || !p.adv_targ || !p.values || !p.value_preds || !p.returns || !p.sums)
    return MMB_EINVAL;
  if ((p.use_value_active_masks || p.use_policy_active_masks) && (!p.active_masks || !p.mask_sum)) return MMB_EINVAL;
  if ((p.ret_mean == nullptr) != (p.ret_var == nullptr)) return MMB_EINVAL;
  if ((p.ret_mean_orig == nullptr) != (p.ret_var_orig == nullptr) || (p.ret_mean_orig && !p.ret_mean)) return MMB_EINVAL;
  if (p.out && !p.ticket) return MMB_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const int G = p.act_dim <= 8 ? 8 : (p.act_dim <= 16 ? 16 : 32);
  const int kmax = (p.act_dim + G - 1) / G;
  const int rows_per_block = LOSS_THREADS / G;
  int grid = (p.num_rows + rows_per_block - 1) / rows_per_block;
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  cudaError_t e;
  {
    LaunchScope ls(K_MAPPO_LOSS, st);
    e = G == 8 ? launch_mappo_loss<8>(p, kmax, grid, st)
               : (G == 16 ? launch_mappo_loss<16>(p, kmax, grid, st) : launch_mappo_loss<32>(p, kmax, grid, st));
  }
  return e == cudaSuccess ? MMB_OK : MMB_ECUDA;
}
