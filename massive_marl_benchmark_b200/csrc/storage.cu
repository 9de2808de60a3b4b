// storage.cu - rollout-storage kernels: fused add_transitions, reverse-time GAE scans (PPO and MARL
// with PopArt/ValueNorm denormalisation), advantage statistics + normalisation, device-side
// get_statistics and the MARL mask logic.
//
// Replaces agents/algorithms/rl/ppo/storage.py:32-73, agents/algorithms/marl/utils/separated_buffer.py:
// 124-168, popart.py:64-75 (denormalize), mappo_trainer.py:189-199, runner.py:229-255.
//
// GAE is sequential in t and embarrassingly parallel over envs; the recurrence is evaluated in the reference's exact op
// order (bit-identical `returns`).  Small rollouts: one thread per env, all T loads issued before the dependent chain.
// From 8192 envs: four consecutive envs per thread with 128-bit accesses, two [t] rows in flight - few, wide, page-sized
// row streams instead of 80 narrow ones (0.42 -> 0.96 of HBM peak at 64 M transitions).  Sum and sum of squares of the raw
// advantages are accumulated in fp64 (thread -> warp shuffle -> one atomic per CTA of a capped grid).  Env-sharded
// multi-GPU: mmb_adv_normalize_xchg exchanges the three doubles over NVLink peer memory inside the normalise kernel.
#include "../../include/mmb.h"
#include "mmb_common.cuh"
#include "mmb_math.cuh"
#include <cstdlib>
#include <cstring>

namespace mmb {
namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of two doubles, result valid in thread 0
__device__ __forceinline__ void block_sum2(double& a, double& b) {
  __shared__ double sa[32], sb[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) { sa[wid] = a; sb[wid] = b; }
  __syncthreads();
  if (wid == 0) {
    a = lane < nw ? sa[lane] : 0.0;
    b = lane < nw ? sb[lane] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
  }
}

// ------------------------------------------------------------------------------------------------
// RolloutStorage.compute_returns (storage.py:51-65)
// ------------------------------------------------------------------------------------------------
// ---- statistics exchange over peer memory (mmb.h "Env-sharded multi-GPU") ----------------------
struct XchgDev {  // kernel-parameter copy of mmb_xchg
  int world, rank, slots, enabled;
  long long spin_ns;   // bound of the flag wait in ns (%globaltimer); < 0: wait for ever
  unsigned long long* state;
  unsigned long long* mailbox[MMB_MAX_RANKS];
};

inline XchgDev make_xchg(const mmb_xchg* x) {
  XchgDev d{};
  if (!x) return d;
  d.world = x->world; d.rank = x->rank; d.slots = x->slots; d.enabled = 1;
  d.spin_ns = x->timeout_ms < 0 ? -1ll : (x->timeout_ms == 0 ? 10000ll : (long long)x->timeout_ms) * 1000000ll;
  d.state = reinterpret_cast<unsigned long long*>(x->state);
  for (int r = 0; r < MMB_MAX_RANKS; ++r) d.mailbox[r] = reinterpret_cast<unsigned long long*>(x->mailbox[r]);
  return d;
}

inline bool xchg_valid(const mmb_xchg* x) {
  if (!x) return true;
  if (x->world < 1 || x->world > MMB_MAX_RANKS || x->rank < 0 || x->rank >= x->world || x->slots < 2 || !x->state) return false;
  for (int r = 0; r < x->world; ++r)
    if (!x->mailbox[r]) return false;
  return true;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// {count, sum, sumsq} of an accumulator: the three base words, plus (MMB_NORM_SLOTS) the slot area the fused GAE of
// mmb_ten_ant_step fills.  Block-wide: lane i of warp 0 reads slot i (one 16-byte load), a fixed shuffle tree adds them
// (the same order in every block and on every rank), the result goes to all threads through shared memory.  Must be
// called by every thread of the block.
__device__ __forceinline__ void read_stats(const double* stats, bool slots, double& cnt, double& s1, double& s2) {
  __shared__ double sh[3];
  if (threadIdx.x < 32) {
    double a = 0.0, b = 0.0;
    if (slots) {
      static_assert(MMB_STAT_SLOTS == 32, "one slot per lane");
      const double2 v = __ldcg(reinterpret_cast<const double2*>(stats + 4 + threadIdx.x * MMB_STAT_SLOT_STRIDE));
      a = v.x; b = v.y;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
    }
    if (threadIdx.x == 0) { sh[0] = __ldcg(stats); sh[1] = __ldcg(stats + 1) + a; sh[2] = __ldcg(stats + 2) + b; }
  }
  __syncthreads();
  cnt = sh[0]; s1 = sh[1]; s2 = sh[2];
}
__device__ __forceinline__ void clear_stats_words(double* stats, bool slots) {
  stats[0] = 0.0; stats[1] = 0.0; stats[2] = 0.0;
  if (slots)
    for (int i = 0; i < MMB_STAT_SLOTS; ++i) { stats[4 + i * MMB_STAT_SLOT_STRIDE] = 0.0; stats[4 + i * MMB_STAT_SLOT_STRIDE + 1] = 0.0; }
}

constexpr int GAE_CHUNK = 16;  // horizon 16 (the reference's default is 8): all loads of a rollout in flight at once

// Grid-stride over env tiles: one fp64 atomic pair per CTA, and the grid is capped - same-address fp64 atomics retire at
// ~13 ns each on B200 (measured), so one pair per 256 envs was the whole run time of the scan at >= 1 M envs.
__global__ void __launch_bounds__(256) gae_ppo_kernel(const __grid_constant__ mmb_gae_ppo_params p) {
  const int N = p.num_envs, T = p.num_steps;
  const float gamma = (float)p.gamma, lam = (float)p.lam;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t e64 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e64 < N; e64 += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)e64;
    float adv = 0.0f;
    float next_v = __ldg(p.last_values + e);
    for (int t_hi = T; t_hi > 0; t_hi -= GAE_CHUNK) {
      const int cnt = t_hi < GAE_CHUNK ? t_hi : GAE_CHUNK;
      float r[GAE_CHUNK], v[GAE_CHUNK], m[GAE_CHUNK];
#pragma unroll
      for (int j = 0; j < GAE_CHUNK; ++j) {
        if (j < cnt) {
          const int64_t idx = (int64_t)(t_hi - 1 - j) * N + e;
          r[j] = __ldg(p.rewards + idx);
          v[j] = __ldg(p.values + idx);
          m[j] = fsub(1.0f, (float)__ldg(p.dones + idx));  // next_is_not_terminal = 1.0 - dones.float()
        }
      }
#pragma unroll
      for (int j = 0; j < GAE_CHUNK; ++j) {
        if (j < cnt) {
          const int64_t idx = (int64_t)(t_hi - 1 - j) * N + e;
          const float mg = fmul(m[j], gamma);
          const float delta = fsub(fadd(r[j], fmul(mg, next_v)), v[j]);
          adv = fadd(delta, fmul(fmul(mg, lam), adv));
          const float ret = fadd(adv, v[j]);
          p.returns[idx] = ret;
          const float a = fsub(ret, v[j]);  // self.advantages = self.returns - self.values
          p.advantages[idx] = a;
          s1 += (double)a;
          s2 += (double)a * (double)a;
          next_v = v[j];
        }
      }
    }
  }
  if (p.stats) {
    block_sum2(s1, s2);
    if (threadIdx.x == 0) {
      atomicAdd(p.stats + 1, s1);
      atomicAdd(p.stats + 2, s2);
      if (blockIdx.x == 0) atomicAdd(p.stats + 0, (double)N * (double)T);
    }
  }
}

// (x - mean) / denom over a grid-stride span: 128-bit streaming accesses (read once, written once, no reuse)
__device__ __forceinline__ void normalize_span(float* __restrict__ adv, int64_t n, int64_t n4, int64_t i, int64_t stride, float mean,
                                               float denom) {
  float4* a4 = reinterpret_cast<float4*>(adv);
  for (int64_t j = i; j < n4; j += stride) {
    float4 v = __ldcs(a4 + j);
    v.x = fdiv(fsub(v.x, mean), denom); v.y = fdiv(fsub(v.y, mean), denom);
    v.z = fdiv(fsub(v.z, mean), denom); v.w = fdiv(fsub(v.w, mean), denom);
    __stcs(a4 + j, v);
  }
  for (int64_t k = (n4 << 2) + i; k < n; k += stride) adv[k] = fdiv(fsub(adv[k], mean), denom);
}

// Statistics exchange + normalisation in one kernel.  Block 0 publishes this shard's {count,sum,sumsq}: thread r
// stores the three words and then the sequence flag (release, system scope) into rank r's mailbox - `world` NVLink
// peer stores in flight at once - and clears the accumulator.  Every block then waits for the `world` flags of
// exchange q = done + 1 in its OWN mailbox (local memory; the peers wrote them over NVLink), sums the shards in rank
// order (bit-identical on every rank) and normalises its slice.  The last block to finish advances `done`.
// The exchange proper, ONE warp: lane r publishes this shard's {count,sum,sumsq} + sequence flag into rank r's mailbox
// (`world` NVLink peer stores in flight at once), then waits for rank r's flag of exchange q = done + 1 in its OWN mailbox
// (local memory; the peers wrote it over NVLink); the shards are summed in rank order (bit-identical on every rank) and
// the global moments go to state[4..6] for the normalise launch that follows on the same stream.  The first version waited
// inside the normalise kernel itself - 64 CTAs of 256 threads spinning on every SM they landed on while the next rollout's
// step kernel wanted those slots (2 GPUs: +1.9 us per rollout); one spinning warp costs nothing.
__global__ void __launch_bounds__(32) xchg_exchange_kernel(double* stats, const __grid_constant__ XchgDev x, int slots_on) {
  const int lane = threadIdx.x;
  const unsigned long long q = ld_relaxed_sys(x.state + 1) + 1ull;
  const size_t slot = (size_t)(q % (unsigned)x.slots) * x.world;
  // this shard's moments: base words + (fused GAE) the slot area, lane i reads slot i
  double a = 0.0, b = 0.0;
  if (slots_on) {
    const double2 v = __ldcg(reinterpret_cast<const double2*>(stats + 4 + lane * MMB_STAT_SLOT_STRIDE));
    a = v.x; b = v.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
  }
  const double cd = __ldcg(stats), ad = __ldcg(stats + 1) + a, bd = __ldcg(stats + 2) + b;
  if (lane < x.world) {
    unsigned long long* box = x.mailbox[lane] + (slot + x.rank) * 4;
    st_relaxed_sys(box + 0, (unsigned long long)__double_as_longlong(cd));
    st_relaxed_sys(box + 1, (unsigned long long)__double_as_longlong(ad));
    st_relaxed_sys(box + 2, (unsigned long long)__double_as_longlong(bd));
    st_release_sys(box + 3, q);
  }
  __syncwarp();
  if (lane == 0) clear_stats_words(stats, slots_on != 0);
  double c = 0.0, s1 = 0.0, s2 = 0.0;
  bool bad = false;
  if (lane < x.world) {
    const unsigned long long* box = x.mailbox[x.rank] + (slot + lane) * 4;
    const unsigned long long t0 = globaltimer_ns();
    unsigned long long f = ld_acquire_sys(box + 3);
    while (f < q && (x.spin_ns < 0 || (long long)(globaltimer_ns() - t0) < x.spin_ns)) {
      __nanosleep(32);
      f = ld_acquire_sys(box + 3);
    }
    bad = f != q;                          // timed out, or the slot was overrun: no partial moments are ever used
    c = __longlong_as_double((long long)ld_relaxed_sys(box + 0));
    s1 = __longlong_as_double((long long)ld_relaxed_sys(box + 1));
    s2 = __longlong_as_double((long long)ld_relaxed_sys(box + 2));
  }
  const bool any_bad = __any_sync(0xffffffffu, bad);
  // sum over ranks in rank order: lane 0 collects lane r's values one after the other
  double tc = 0.0, t1 = 0.0, t2 = 0.0;
  for (int r = 0; r < x.world; ++r) {
    tc += __shfl_sync(0xffffffffu, c, r);
    t1 += __shfl_sync(0xffffffffu, s1, r);
    t2 += __shfl_sync(0xffffffffu, s2, r);
  }
  if (lane == 0) {
    double* g = reinterpret_cast<double*>(x.state + 4);
    if (any_bad) {                         // loud: the normalise launch turns the whole plane into NaN
      atomicAdd(x.state + 3, 1ull);
      const double nan = __longlong_as_double(0x7ff8000000000000ll);
      g[0] = nan; g[1] = nan; g[2] = nan;
    } else {
      g[0] = tc; g[1] = t1; g[2] = t2;
    }
    __threadfence();
    st_relaxed_sys(x.state + 1, q);
  }
}

// Four consecutive envs per thread, 128-bit accesses: a CTA's load of one [t] row is 4 KB contiguous per plane (whole
// DRAM pages instead of 1 KB runs spread over 80 concurrent row streams), a quarter of the memory instructions, and four
// independent recurrences per thread.  Same operations in the same order per env as the scalar kernel.
template <int CH>
__global__ void __launch_bounds__(256) gae_ppo_vec4_kernel(const __grid_constant__ mmb_gae_ppo_params p) {
  const int N = p.num_envs, T = p.num_steps;
  const int N4 = N >> 2;
  const float gamma = (float)p.gamma, lam = (float)p.lam;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < N4; q += (int64_t)gridDim.x * blockDim.x) {
    float adv[4] = {0.f, 0.f, 0.f, 0.f};
    const float4 lv = __ldg(reinterpret_cast<const float4*>(p.last_values) + q);
    float nv[4] = {lv.x, lv.y, lv.z, lv.w};
    for (int t_hi = T; t_hi > 0; t_hi -= CH) {
      const int cnt = t_hi < CH ? t_hi : CH;
      float4 r[CH], v[CH];
      uchar4 d[CH];
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        if (j < cnt) {
          const int64_t row = (int64_t)(t_hi - 1 - j) * N4 + q;
          r[j] = __ldg(reinterpret_cast<const float4*>(p.rewards) + row);
          v[j] = __ldg(reinterpret_cast<const float4*>(p.values) + row);
          d[j] = __ldg(reinterpret_cast<const uchar4*>(p.dones) + row);
        }
      }
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        if (j < cnt) {
          const int64_t row = (int64_t)(t_hi - 1 - j) * N4 + q;
          const float rr[4] = {r[j].x, r[j].y, r[j].z, r[j].w}, vv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
          const float dd[4] = {(float)d[j].x, (float)d[j].y, (float)d[j].z, (float)d[j].w};
          float ret[4], a[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float mg = fmul(fsub(1.0f, dd[h]), gamma);
            const float delta = fsub(fadd(rr[h], fmul(mg, nv[h])), vv[h]);
            adv[h] = fadd(delta, fmul(fmul(mg, lam), adv[h]));
            ret[h] = fadd(adv[h], vv[h]);
            a[h] = fsub(ret[h], vv[h]);
            s1 += (double)a[h];
            s2 += (double)a[h] * (double)a[h];
            nv[h] = vv[h];
          }
          reinterpret_cast<float4*>(p.returns)[row] = make_float4(ret[0], ret[1], ret[2], ret[3]);
          reinterpret_cast<float4*>(p.advantages)[row] = make_float4(a[0], a[1], a[2], a[3]);
        }
      }
    }
  }
  if (p.stats) {
    block_sum2(s1, s2);
    if (threadIdx.x == 0) {
      atomicAdd(p.stats + 1, s1);
      atomicAdd(p.stats + 2, s2);
      if (blockIdx.x == 0) atomicAdd(p.stats + 0, (double)N * (double)T);
    }
  }
}

// (adv - mean) / (std_unbiased + eps).  With `clear_stats` the last block to have read the statistics clears them
// (ticket counter in stats[3]), so the accumulator is ready for the next rollout without a memset launch and
// the launch sequence can be replayed from a CUDA graph.
__global__ void __launch_bounds__(256) adv_normalize_kernel(float* __restrict__ adv, int64_t n, double* stats,
                                                            float eps, int flags) {
  double cnt, s1, s2;
  read_stats(stats, (flags & MMB_NORM_SLOTS) != 0, cnt, s1, s2);
  if (flags & MMB_NORM_CLEAR) {
    __syncthreads();  // every thread of this block holds its copy
    if (threadIdx.x == 0) {
      unsigned long long* ticket = reinterpret_cast<unsigned long long*>(stats + 3);
      if (atomicAdd(ticket, 1ull) == (unsigned long long)gridDim.x - 1ull) {
        clear_stats_words(stats, (flags & MMB_NORM_SLOTS) != 0);
        *ticket = 0ull;
      }
    }
  }
  const double mean_d = s1 / cnt;
  double var_d = (s2 - s1 * mean_d) / (cnt - 1.0);  // torch.std(): unbiased
  if (var_d < 0.0) var_d = 0.0;
  const float mean = (float)mean_d;
  const float denom = fadd((float)sqrt(var_d), eps);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n4 = aligned16(adv) ? (n >> 2) : 0;
  normalize_span(adv, n, n4, i, stride, mean, denom);
}

// RolloutStorage.get_statistics (storage.py:67-73).  With the last row forced done and the env-major
// flatten, the trajectory lengths sum to T*N, so the mean length is T*N / (#dones in rows < T-1  +  N).
__global__ void __launch_bounds__(1024) rollout_statistics_kernel(const uint8_t* __restrict__ dones,
                                                                   const float* __restrict__ rewards, int T, int N,
                                                                   float* __restrict__ out2) {
  double cnt = 0.0, rs = 0.0;
  const int64_t total = (int64_t)T * N;
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
    rs += (double)rewards[i];
    if (i < (int64_t)(T - 1) * N) cnt += dones[i] ? 1.0 : 0.0;
  }
  block_sum2(cnt, rs);
  if (threadIdx.x == 0) {
    out2[0] = (float)total / (float)(cnt + (double)N);
    out2[1] = (float)(rs / (double)total);
  }
}

// RolloutStorage.add_transitions (storage.py:32-46): nine copies in one launch; blockIdx.y = field
__global__ void __launch_bounds__(256) rollout_add_kernel(const __grid_constant__ mmb_rollout_add_params p) {
  const int field = blockIdx.y;
  const int64_t N = p.num_envs;
  const float* src = nullptr;
  float* dst = nullptr;
  int64_t n = 0;
  switch (field) {
    case 0: src = p.observations; dst = p.dst_observations; n = N * p.obs_dim; break;
    case 1: src = p.states; dst = p.dst_states; n = N * p.states_dim; break;
    case 2: src = p.actions; dst = p.dst_actions; n = N * p.act_dim; break;
    case 3: src = p.rewards; dst = p.dst_rewards; n = N; break;
    case 4: src = p.values; dst = p.dst_values; n = N; break;
    case 5: src = p.actions_log_prob; dst = p.dst_actions_log_prob; n = N; break;
    case 6: src = p.mu; dst = p.dst_mu; n = N * p.act_dim; break;
    case 7: src = p.sigma; dst = p.dst_sigma; n = N * p.act_dim; break;
    default: break;
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (field == 8) {  // dones int64 -> uint8 (copy_ cast)
    if (p.dones && p.dst_dones)
      for (int64_t i = i0; i < N; i += stride) p.dst_dones[i] = (uint8_t)p.dones[i];
    return;
  }
  if (!src || !dst || n == 0) return;
  if (field == 4 && p.values_stride > 1) {   // one column of a wider matrix
    for (int64_t i = i0; i < N; i += stride) dst[i] = src[i * p.values_stride];
    return;
  }
  if (aligned16(src) && aligned16(dst)) {
    const int64_t n4 = n >> 2;
    for (int64_t i = i0; i < n4; i += stride) stg4(dst + 4 * i, ldg4(src + 4 * i));
    for (int64_t i = (n4 << 2) + i0; i < n; i += stride) dst[i] = src[i];
  } else {
    for (int64_t i = i0; i < n; i += stride) dst[i] = src[i];
  }
}

// ------------------------------------------------------------------------------------------------
// SeparatedReplayBuffer.compute_returns (separated_buffer.py:124-168) + advantage prologue
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gae_marl_kernel(const __grid_constant__ mmb_gae_marl_params p) {
  const int N = p.num_envs, T = p.num_steps;
  const int a = blockIdx.y;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t e64 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e64 < N; e64 += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)e64;
    const float gamma = (float)p.gamma;
    const float gl = (float)(p.gamma * p.gae_lambda);  // Python float product, then cast by torch
    float mean = 0.0f, sd = 1.0f;
    if (p.use_denorm) {
      mean = __ldg(p.denorm_mean + a);
      sd = fsqrt(__ldg(p.denorm_var + a));
    }
    auto D = [&](float x) { return p.use_denorm ? fadd(fmul(x, sd), mean) : x; };  // popart.py:71
    const int64_t ro = (int64_t)e * p.rew_e + (int64_t)a * p.rew_a;
    const int64_t vo = (int64_t)e * p.val_e + (int64_t)a * p.val_a;
    const int64_t mo = (int64_t)e * p.msk_e + (int64_t)a * p.msk_a;
    const int64_t bo = (int64_t)e * p.bad_e + (int64_t)a * p.bad_a;
    const int64_t to = (int64_t)e * p.ret_e + (int64_t)a * p.ret_a;
    const int64_t ao = (int64_t)e * p.adv_e + (int64_t)a * p.adv_a;
    const float nv = __ldg(p.next_value + (int64_t)e * p.nv_e + (int64_t)a * p.nv_a);
    if (p.use_gae) {
      p.value_preds[(int64_t)T * p.val_t + vo] = nv;  // self.value_preds[-1] = next_value
      float gae = 0.0f;
      float v1 = nv;
      for (int t = T - 1; t >= 0; --t) {
        const float r = __ldg(p.rewards + (int64_t)t * p.rew_t + ro);
        const float v0 = p.value_preds[(int64_t)t * p.val_t + vo];
        const float m1 = __ldg(p.masks + (int64_t)(t + 1) * p.msk_t + mo);
        const float d0 = D(v0);
        const float delta = fsub(fadd(r, fmul(fmul(gamma, D(v1)), m1)), d0);
        gae = fadd(delta, fmul(fmul(gl, m1), gae));
        if (p.use_proper_time_limits) gae = fmul(gae, __ldg(p.bad_masks + (int64_t)(t + 1) * p.bad_t + bo));
        const float ret = fadd(gae, d0);
        p.returns[(int64_t)t * p.ret_t + to] = ret;
        if (p.advantages) {
          const float adv = fsub(ret, d0);  // mappo_trainer.py:190-192
          p.advantages[(int64_t)t * p.adv_t + ao] = adv;
          s1 += (double)adv;
          s2 += (double)adv * (double)adv;
        }
        v1 = v0;
      }
    } else {
      p.returns[(int64_t)T * p.ret_t + to] = nv;  // self.returns[-1] = next_value
      float r1 = nv;
      for (int t = T - 1; t >= 0; --t) {
        const float r = __ldg(p.rewards + (int64_t)t * p.rew_t + ro);
        const float m1 = __ldg(p.masks + (int64_t)(t + 1) * p.msk_t + mo);
        const float v0 = p.value_preds[(int64_t)t * p.val_t + vo];
        float ret;
        if (p.use_proper_time_limits) {
          const float b1 = __ldg(p.bad_masks + (int64_t)(t + 1) * p.bad_t + bo);
          const float vp = (p.use_popart && p.use_denorm) ? D(v0) : v0;
          ret = fadd(fmul(fadd(fmul(fmul(r1, gamma), m1), r), b1), fmul(fsub(1.0f, b1), vp));
        } else {
          ret = fadd(fmul(fmul(r1, gamma), m1), r);
        }
        p.returns[(int64_t)t * p.ret_t + to] = ret;
        if (p.advantages) {
          const float adv = fsub(ret, D(v0));
          p.advantages[(int64_t)t * p.adv_t + ao] = adv;
          s1 += (double)adv;
          s2 += (double)adv * (double)adv;
        }
        r1 = ret;
      }
    }
  }
  if (p.stats && p.advantages) {
    block_sum2(s1, s2);
    if (threadIdx.x == 0) {
      atomicAdd(p.stats + 4 * a + 1, s1);
      atomicAdd(p.stats + 4 * a + 2, s2);
      if (blockIdx.x == 0) atomicAdd(p.stats + 4 * a + 0, (double)N * (double)T);
    }
  }
}

// Runner.insert (runner.py:229-255)
__global__ void marl_masks_kernel(const int64_t* __restrict__ dones, int N, int A, float* masks, int64_t m_e, int64_t m_a,
                                  float* active, int64_t am_e, int64_t am_a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  bool all_done = true;
  for (int a = 0; a < A; ++a) all_done = all_done && (dones[(int64_t)e * A + a] != 0);
  for (int a = 0; a < A; ++a) {
    const bool d = dones[(int64_t)e * A + a] != 0;
    if (masks) masks[(int64_t)e * m_e + (int64_t)a * m_a] = all_done ? 0.0f : 1.0f;
    if (active) active[(int64_t)e * am_e + (int64_t)a * am_a] = (d && !all_done) ? 0.0f : 1.0f;
  }
}

}  // namespace
}  // namespace mmb

using namespace mmb;

extern "C" int32_t mmb_gae_ppo(const mmb_gae_ppo_params* pp, void* stream) {
  if (!pp) return MMB_EINVAL;
  mmb_gae_ppo_params p = *pp;
  if (p.num_envs <= 0 || p.num_steps <= 0) return MMB_EINVAL;
  if (!p.rewards || !p.values || !p.dones || !p.last_values || !p.returns || !p.advantages) return MMB_EINVAL;
  {
    LaunchScope ls(K_GAE_PPO, (cudaStream_t)stream);
    const uintptr_t al = reinterpret_cast<uintptr_t>(p.rewards) | reinterpret_cast<uintptr_t>(p.values) |
                         reinterpret_cast<uintptr_t>(p.last_values) | reinterpret_cast<uintptr_t>(p.returns) |
                         reinterpret_cast<uintptr_t>(p.advantages);
    static const int vec_chunk = [] { const char* v = getenv("MMB_GAE_CHUNK"); return v ? atoi(v) : 2; }();
    if (p.num_envs >= 8192 && (p.num_envs & 3) == 0 && (al & 15u) == 0 && (reinterpret_cast<uintptr_t>(p.dones) & 3u) == 0 &&
        vec_chunk > 0) {
      int blocks = (p.num_envs / 4 + 255) / 256;
      if (blocks > sm_count() * 8) blocks = sm_count() * 8;
      if (vec_chunk == 8) gae_ppo_vec4_kernel<8><<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
      else if (vec_chunk == 2) gae_ppo_vec4_kernel<2><<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
      else if (vec_chunk == 16) gae_ppo_vec4_kernel<16><<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
      else gae_ppo_vec4_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
    } else {
      // small rollouts: 64-thread CTAs, so that 4096 envs are 64 CTAs on 64 SMs instead of 16 on 16 (the scan is a
      // latency chain per thread; spreading it costs nothing and finds free SM slots next to a running step kernel)
      const int bs = p.num_envs >= 256 * sm_count() ? 256 : 64;
      int blocks = (p.num_envs + bs - 1) / bs;
      if (blocks > sm_count() * 8) blocks = sm_count() * 8;
      gae_ppo_kernel<<<blocks, bs, 0, (cudaStream_t)stream>>>(p);
    }
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

extern "C" int32_t mmb_adv_normalize(float* advantages, int64_t n, double* stats, float eps, int32_t flags,
                                     void* stream) {
  if (!advantages || !stats || n <= 0 || (flags & ~(MMB_NORM_CLEAR | MMB_NORM_SLOTS))) return MMB_EINVAL;
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  {
    LaunchScope ls(K_ADV_NORM, (cudaStream_t)stream);
    adv_normalize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(advantages, n, stats, eps, flags);
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

extern "C" int32_t mmb_adv_normalize_xchg(float* advantages, int64_t n, double* stats, const mmb_xchg* xchg, float eps,
                                          int32_t flags, void* stream) {
  if (!advantages || !stats || !xchg || n <= 0 || !xchg_valid(xchg) || (flags & ~(MMB_NORM_CLEAR | MMB_NORM_SLOTS)))
    return MMB_EINVAL;
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  const XchgDev xd = make_xchg(xchg);
  {
    LaunchScope ls(K_ADV_NORM_XCHG, (cudaStream_t)stream);
    xchg_exchange_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(stats, xd, (flags & MMB_NORM_SLOTS) ? 1 : 0);
  }
  if (cudaGetLastError() != cudaSuccess) return MMB_ECUDA;
  {
    LaunchScope ls(K_ADV_NORM, (cudaStream_t)stream);
    // the global moments sit in state[4..6]; nothing to clear there (the next exchange overwrites them)
    adv_normalize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(advantages, n, reinterpret_cast<double*>(xchg->state + 4), eps, 0);
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

extern "C" int64_t mmb_xchg_mailbox_bytes(int32_t world, int32_t slots) {
  if (world < 1 || world > MMB_MAX_RANKS || slots < 2) return MMB_EINVAL;
  return (int64_t)slots * world * 4 * 8;
}

extern "C" int32_t mmb_xchg_alloc(int64_t bytes, void** dev_ptr, uint8_t* handle64) {
  if (bytes <= 0 || !dev_ptr || !handle64) return MMB_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) return MMB_ECUDA;
  if (cudaMemset(p, 0, (size_t)bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { cudaFree(p); return MMB_ECUDA; }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaFree(p); (void)cudaGetLastError(); return MMB_ECUDA; }
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return MMB_OK;
}

extern "C" int32_t mmb_xchg_open(const uint8_t* handle64, void** dev_ptr) {
  if (!handle64 || !dev_ptr) return MMB_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  if (cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); return MMB_ECUDA; }
  return MMB_OK;
}

extern "C" int32_t mmb_xchg_close(void* dev_ptr) {
  if (!dev_ptr) return MMB_EINVAL;
  return cudaIpcCloseMemHandle(dev_ptr) == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

extern "C" int32_t mmb_xchg_free(void* dev_ptr) {
  if (!dev_ptr) return MMB_EINVAL;
  return cudaFree(dev_ptr) == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

extern "C" int32_t mmb_rollout_statistics(const uint8_t* dones, const float* rewards, int32_t num_steps,
                                          int32_t num_envs, float* out2, void* stream) {
  if (!dones || !rewards || !out2 || num_steps <= 0 || num_envs <= 0) return MMB_EINVAL;
  {
    LaunchScope ls(K_STATS, (cudaStream_t)stream);
    rollout_statistics_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(dones, rewards, num_steps, num_envs, out2);
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

extern "C" int32_t mmb_rollout_add(const mmb_rollout_add_params* pp, void* stream) {
  if (!pp) return MMB_EINVAL;
  mmb_rollout_add_params p = *pp;
  if (p.num_envs <= 0 || p.obs_dim < 0 || p.states_dim < 0 || p.act_dim < 0) return MMB_EINVAL;
  // the grid follows the widest plane that is actually copied (a NULL source = the producer wrote the slot itself)
  int64_t maxn = p.num_envs;
  if (p.observations && p.dst_observations && (int64_t)p.num_envs * p.obs_dim > maxn) maxn = (int64_t)p.num_envs * p.obs_dim;
  if (p.states && p.dst_states && (int64_t)p.num_envs * p.states_dim > maxn) maxn = (int64_t)p.num_envs * p.states_dim;
  if (((p.actions && p.dst_actions) || (p.mu && p.dst_mu) || (p.sigma && p.dst_sigma)) && (int64_t)p.num_envs * p.act_dim > maxn)
    maxn = (int64_t)p.num_envs * p.act_dim;
  int64_t bx = (maxn / 4 + 255) / 256;
  if (bx < 1) bx = 1;
  if (bx > sm_count() * 8) bx = sm_count() * 8;
  {
    LaunchScope ls(K_ROLLOUT_ADD, (cudaStream_t)stream);
    rollout_add_kernel<<<dim3((unsigned)bx, 9), 256, 0, (cudaStream_t)stream>>>(p);
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

namespace mmb {
namespace {
// blockIdx.y = segment; rows (i0, i1) of n2 contiguous elements each, one warp per row chunk: coalesced on both sides
__global__ void __launch_bounds__(256) copy_group_kernel(const __grid_constant__ mmb_copy_group_params p) {
  const mmb_copy_seg& g = p.seg[blockIdx.y];
  const int64_t rows = (int64_t)g.n0 * g.n1;
  const int lane = threadIdx.x & 31;
  const bool vec = (g.n2 & 3) == 0 && ((g.dst_s0 | g.dst_s1 | g.src_s0 | g.src_s1) & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(g.dst) | reinterpret_cast<uintptr_t>(g.src)) & 15u) == 0;
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int64_t i0 = r / g.n1, i1 = r - i0 * g.n1;
    float* d = g.dst + i0 * g.dst_s0 + i1 * g.dst_s1;
    const float* s = g.src + i0 * g.src_s0 + i1 * g.src_s1;
    if (vec) {
      for (int j = lane * 4; j < g.n2; j += 128) *reinterpret_cast<float4*>(d + j) = __ldg(reinterpret_cast<const float4*>(s + j));
    } else {
      for (int j = lane; j < g.n2; j += 32) d[j] = __ldg(s + j);
    }
  }
}
}  // namespace
}  // namespace mmb

extern "C" int32_t mmb_copy_group(const mmb_copy_group_params* pp, void* stream) {
  if (!pp || pp->count <= 0 || pp->count > MMB_MAX_COPY_SEGS) return MMB_EINVAL;
  int64_t max_rows = 1;
  for (int i = 0; i < pp->count; ++i) {
    const mmb_copy_seg& g = pp->seg[i];
    if (!g.dst || !g.src || g.n0 <= 0 || g.n1 <= 0 || g.n2 <= 0) return MMB_EINVAL;
    if ((int64_t)g.n0 * g.n1 > max_rows) max_rows = (int64_t)g.n0 * g.n1;
  }
  int64_t bx = (max_rows + 7) / 8;
  if (bx > mmb::sm_count() * 8) bx = mmb::sm_count() * 8;
  {
    mmb::LaunchScope ls(mmb::K_ROLLOUT_ADD, (cudaStream_t)stream);
    mmb::copy_group_kernel<<<dim3((unsigned)bx, (unsigned)pp->count), 256, 0, (cudaStream_t)stream>>>(*pp);
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

extern "C" int32_t mmb_gae_marl(const mmb_gae_marl_params* pp, void* stream) {
  if (!pp) return MMB_EINVAL;
  mmb_gae_marl_params p = *pp;
  if (p.num_envs <= 0 || p.num_steps <= 0 || p.num_agents <= 0 || p.num_agents > 65535) return MMB_EINVAL;
  if (!p.rewards || !p.value_preds || !p.masks || !p.next_value || !p.returns) return MMB_EINVAL;
  if (p.use_proper_time_limits && !p.bad_masks) return MMB_EINVAL;
  if (p.use_denorm && (!p.denorm_mean || !p.denorm_var)) return MMB_EINVAL;
  {
    LaunchScope ls(K_GAE_MARL, (cudaStream_t)stream);
    int bx = (p.num_envs + 255) / 256;
    const int cap = (sm_count() * 8 + p.num_agents - 1) / p.num_agents;
    if (bx > cap) bx = cap;
    gae_marl_kernel<<<dim3(bx, p.num_agents), 256, 0, (cudaStream_t)stream>>>(p);
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

extern "C" int32_t mmb_marl_masks(const int64_t* dones, int32_t num_envs, int32_t num_agents, float* masks, int64_t m_e,
                                  int64_t m_a, float* active_masks, int64_t am_e, int64_t am_a, void* stream) {
  if (!dones || num_envs <= 0 || num_agents <= 0) return MMB_EINVAL;
  {
    LaunchScope ls(K_MASKS, (cudaStream_t)stream);
    marl_masks_kernel<<<(num_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dones, num_envs, num_agents, masks, m_e,
                                                                               m_a, active_masks, am_e, am_a);
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}
