// reset.cu - reset_idx for TenAnt / OneAnt / MultiIngenuity as one kernel.
//
// Replaces ten_ant.py:810-884, one_ant.py:363-391, multi_ingenuity.py:231-266:
//   env_ids = reset_buf.nonzero().flatten()                  -> ordered stream compaction
//   torch.unique(cat(actor_indices[env_ids] ...)).int32      -> ascending {apn*e + j}: no sort needed,
//                                                               the lists are emitted in order
//   dof_pos_k[env_ids] = clamp(initial + U(-.2,.2)), dof_vel_k[env_ids] = U(-.1,.1), same noise for all ants
//   Ingenuity: rotor speeds for ALL envs when any env resets; forces rows of reset envs zeroed.
// The reference pays a host sync (`len(env_ids)`) and ~45 small kernels per step for this; here the
// count stays on the device (counts[f]) and nothing is read back.
//
// One CTA of 1024 threads per row f (a row = the flags of one frame), four consecutive envs per thread: shuffle scan of
// the per-thread counts inside a warp, warp totals scanned by warp 0, running offset across 4096-env passes.  O(N) flag
// bytes per row; fine up to ~1M envs per row (one SM streams the flags); a multi-CTA decoupled look-back scan is the
// planned upgrade for larger N.
#include "../../include/mmb.h"
#include "mmb_common.cuh"
#include "mmb_math.cuh"

namespace mmb {
namespace {

struct TaskShape { int apn, na, nb, dofs, ants; };
__device__ __forceinline__ TaskShape shape_of(int task) {
  if (task == MMB_TASK_TEN_ANT) return {11, 11, 10, 80, 10};
  if (task == MMB_TASK_ONE_ANT) return {2, 2, 1, 8, 1};
  return {4, 4, 4, 16, 0};
}

__global__ void __launch_bounds__(1024) reset_kernel(const __grid_constant__ mmb_reset_params p) {
  __shared__ int warp_tot[32];
  __shared__ int s_running, s_chunk_total;
  const int f = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int N = p.num_envs;
  const TaskShape sh = shape_of(p.task);
  const int64_t* f64 = p.flags_i64 ? p.flags_i64 + (int64_t)f * p.flags_i64_row_stride : nullptr;
  const uint8_t* f8 = p.flags_u8 ? p.flags_u8 + (int64_t)f * p.flags_u8_row_stride : nullptr;
  int64_t* env_ids = p.env_ids + (int64_t)f * p.env_ids_row_stride;
  int32_t* ia = p.index_a ? p.index_a + (int64_t)f * p.index_a_row_stride : nullptr;
  int32_t* ib = p.index_b ? p.index_b + (int64_t)f * p.index_b_row_stride : nullptr;

  if (tid == 0) s_running = 0;
  __syncthreads();
  // Four consecutive envs per thread (one 32-bit load of the uint8 flags), 4096 envs per pass: the ordered rank of a
  // flagged env = running total + exclusive scan of the per-thread counts (shuffle scan inside a warp, the 32 warp totals
  // scanned by warp 0) + its rank among the thread's four.  One pass and two barriers for N <= 4096.
  for (int base = 0; base < N; base += 4096) {
    const int e4 = base + 4 * tid;
    unsigned bits = 0;
    if (f8 && e4 + 3 < N && (reinterpret_cast<uintptr_t>(f8 + e4) & 3u) == 0) {
      const uchar4 v = *reinterpret_cast<const uchar4*>(f8 + e4);
      bits = (v.x ? 1u : 0u) | (v.y ? 2u : 0u) | (v.z ? 4u : 0u) | (v.w ? 8u : 0u);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = e4 + j;
        if (e < N && (f64 ? (f64[e] != 0) : (f8[e] != 0))) bits |= 1u << j;
      }
    }
    const int cnt = __popc(bits);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const int wt = warp_tot[lane];
      int wincl = wt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, wincl, o);
        if (lane >= o) wincl += v;
      }
      warp_tot[lane] = wincl - wt;          // exclusive prefix of the warp totals
      if (lane == 31) s_chunk_total = wincl;
    }
    __syncthreads();
    int i = s_running + warp_tot[wid] + incl - cnt;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (bits & (1u << j)) {
        const int e = e4 + j;
        env_ids[i] = e;
        if (ia) for (int q = 0; q < sh.na; ++q) ia[i * sh.na + q] = sh.apn * e + q;
        if (ib) for (int q = 0; q < sh.nb; ++q) ib[i * sh.nb + q] = sh.apn * e + q;
        ++i;
      }
    }
    __syncthreads();
    if (tid == 0) s_running += s_chunk_total;
    __syncthreads();
  }
  const int count = s_running;
  if (tid == 0 && p.counts) p.counts[f] = count;
  if (count == 0) return;

  float* dof = p.dof_state ? p.dof_state + (int64_t)f * p.dof_state_row_stride : nullptr;
  if (sh.ants > 0 && dof) {
    // ten_ant.py:822-857 / one_ant.py:371-376: one (pos, vel) pair per (reset env, dof)
    const float* npos = p.noise_pos ? p.noise_pos + (int64_t)f * p.noise_row_stride : nullptr;
    const float* nvel = p.noise_vel ? p.noise_vel + (int64_t)f * p.noise_row_stride : nullptr;
    const int items = count * sh.dofs;
    for (int it = tid; it < items; it += 1024) {
      const int i = it / sh.dofs, d = it - i * sh.dofs, j = d & 7;
      const int e = (int)env_ids[i];
      float np_, nv_;
      if (p.noise_mode == 0) {
        np_ = npos[(int64_t)i * 8 + j];
        nv_ = nvel[(int64_t)i * 8 + j];
      } else {  // torch_rand_float(lo, hi) = (hi - lo) * U[0,1) + lo, keyed by env so every ant shares the draw
        uint4 r = philox4x32_10(make_uint4((uint32_t)e, (uint32_t)(p.step + f), (uint32_t)(j >> 1), 0u),
                                make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
        float up = u01((j & 1) ? r.z : r.x), uv = u01((j & 1) ? r.w : r.y);
        np_ = fadd(fmul(0.4f, up), -0.2f);
        nv_ = fadd(fmul(0.2f, uv), -0.1f);
      }
      float pos = fadd(p.c.initial_dof_pos[j], np_);
      pos = fmaxf(fminf(pos, p.c.dof_upper[j]), p.c.dof_lower[j]);  // tensor_clamp = max(min(t, hi), lo)
      *reinterpret_cast<float2*>(dof + ((int64_t)e * sh.dofs + d) * 2) = make_float2(pos, nv_);
    }
  }
  if (p.task == MMB_TASK_INGENUITY) {
    if (dof) {  // multi_ingenuity.py:234-241: every env, whenever reset_idx runs
      for (int it = tid; it < N * 4; it += 1024) {
        float* d = dof + (int64_t)it * 8;  // 4 dofs x (pos, vel) per helicopter
        d[3] = -50.0f;
        d[7] = 50.0f;
      }
    }
    if (p.forces_state) {  // multi_ingenuity.py:243-244
      const int items = count * 72;
      for (int it = tid; it < items; it += 1024) {
        const int i = it / 72, r = it - i * 72;
        p.forces_state[(int64_t)env_ids[i] * 72 + r] = 0.0f;
      }
    }
  }
}

// One CTA per 4096-env chunk of a flag row (grid = chunks x rows), decoupled look-back: every CTA publishes its chunk total
// to the caller's scratch, sums the totals of the chunks before it (earlier CTAs in launch order: forward progress as in
// any single-pass scan), and writes its own ordered segment.  The CTA that completes the row's look-backs last clears
// the scratch for the next launch / graph replay.
__global__ void __launch_bounds__(1024) reset_scan_kernel(const __grid_constant__ mmb_reset_params p) {
  __shared__ int warp_tot[32];
  __shared__ int s_running, s_chunk_total;
  const int f = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  unsigned long long* status = reinterpret_cast<unsigned long long*>(p.scan_scratch) + (size_t)f * (chunks + 1);   // [chunks] totals, [chunks] done counter
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int N = p.num_envs;
  const TaskShape sh = shape_of(p.task);
  const int64_t* f64 = p.flags_i64 ? p.flags_i64 + (int64_t)f * p.flags_i64_row_stride : nullptr;
  const uint8_t* f8 = p.flags_u8 ? p.flags_u8 + (int64_t)f * p.flags_u8_row_stride : nullptr;
  int64_t* env_ids = p.env_ids + (int64_t)f * p.env_ids_row_stride;
  int32_t* ia = p.index_a ? p.index_a + (int64_t)f * p.index_a_row_stride : nullptr;
  int32_t* ib = p.index_b ? p.index_b + (int64_t)f * p.index_b_row_stride : nullptr;

  if (tid == 0) s_running = 0;
  __syncthreads();
  // Four consecutive envs per thread (one 32-bit load of the uint8 flags), 4096 envs per pass: the ordered rank of a
  // flagged env = running total + exclusive scan of the per-thread counts (shuffle scan inside a warp, the 32 warp totals
  // scanned by warp 0) + its rank among the thread's four.  One pass and two barriers for N <= 4096.
  {
    const int base = chunk * 4096;
    const int e4 = base + 4 * tid;
    unsigned bits = 0;
    if (f8 && e4 + 3 < N && (reinterpret_cast<uintptr_t>(f8 + e4) & 3u) == 0) {
      const uchar4 v = *reinterpret_cast<const uchar4*>(f8 + e4);
      bits = (v.x ? 1u : 0u) | (v.y ? 2u : 0u) | (v.z ? 4u : 0u) | (v.w ? 8u : 0u);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = e4 + j;
        if (e < N && (f64 ? (f64[e] != 0) : (f8[e] != 0))) bits |= 1u << j;
      }
    }
    const int cnt = __popc(bits);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const int wt = warp_tot[lane];
      int wincl = wt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, wincl, o);
        if (lane >= o) wincl += v;
      }
      warp_tot[lane] = wincl - wt;          // exclusive prefix of the warp totals
      if (lane == 31) s_chunk_total = wincl;
    }
    __syncthreads();
    if (tid == 0) {                       // publish this chunk's total: bit 63 = valid
      *reinterpret_cast<volatile unsigned long long*>(status + chunk) = (1ull << 63) | (unsigned long long)s_chunk_total;
      s_running = 0;
    }
    __syncthreads();
    {                                     // look back: totals of chunks 0..chunk-1
      int part = 0;
      for (int j = tid; j < chunk; j += 1024) {
        unsigned long long v;
        do { v = *reinterpret_cast<volatile unsigned long long*>(status + j); } while (!(v >> 63));
        part += (int)(v & 0xffffffffu);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0 && part) atomicAdd(&s_running, part);
    }
    __syncthreads();
    int i = s_running + warp_tot[wid] + incl - cnt;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (bits & (1u << j)) {
        const int e = e4 + j;
        env_ids[i] = e;
        if (ia) for (int q = 0; q < sh.na; ++q) ia[i * sh.na + q] = sh.apn * e + q;
        if (ib) for (int q = 0; q < sh.nb; ++q) ib[i * sh.nb + q] = sh.apn * e + q;
        ++i;
      }
    }
    __syncthreads();
  }
  const int seg0 = s_running, seg_n = s_chunk_total;   // this CTA's ordered segment of the row's lists
  if (tid == 0) {
    if (chunk == chunks - 1 && p.counts) p.counts[f] = seg0 + seg_n;
    __threadfence();
    if (atomicAdd(status + chunks, 1ull) == (unsigned long long)chunks - 1ull) {   // every CTA of the row has looked back
      for (int j = 0; j <= chunks; ++j) status[j] = 0ull;
    }
  }
  const int count = seg_n;
  if (count == 0) return;

  float* dof = p.dof_state ? p.dof_state + (int64_t)f * p.dof_state_row_stride : nullptr;
  if (sh.ants > 0 && dof) {
    // ten_ant.py:822-857 / one_ant.py:371-376: one (pos, vel) pair per (reset env, dof)
    const float* npos = p.noise_pos ? p.noise_pos + (int64_t)f * p.noise_row_stride : nullptr;
    const float* nvel = p.noise_vel ? p.noise_vel + (int64_t)f * p.noise_row_stride : nullptr;
    const int items = count * sh.dofs;
    for (int it = tid; it < items; it += 1024) {
      const int i = seg0 + it / sh.dofs, d = it - (it / sh.dofs) * sh.dofs, j = d & 7;   // i: ordinal in the row
      const int e = (int)env_ids[i];
      float np_, nv_;
      if (p.noise_mode == 0) {
        np_ = npos[(int64_t)i * 8 + j];
        nv_ = nvel[(int64_t)i * 8 + j];
      } else {  // torch_rand_float(lo, hi) = (hi - lo) * U[0,1) + lo, keyed by env so every ant shares the draw
        uint4 r = philox4x32_10(make_uint4((uint32_t)e, (uint32_t)(p.step + f), (uint32_t)(j >> 1), 0u),
                                make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
        float up = u01((j & 1) ? r.z : r.x), uv = u01((j & 1) ? r.w : r.y);
        np_ = fadd(fmul(0.4f, up), -0.2f);
        nv_ = fadd(fmul(0.2f, uv), -0.1f);
      }
      float pos = fadd(p.c.initial_dof_pos[j], np_);
      pos = fmaxf(fminf(pos, p.c.dof_upper[j]), p.c.dof_lower[j]);  // tensor_clamp = max(min(t, hi), lo)
      *reinterpret_cast<float2*>(dof + ((int64_t)e * sh.dofs + d) * 2) = make_float2(pos, nv_);
    }
  }
}

}  // namespace
}  // namespace mmb

extern "C" int32_t mmb_reset_compact(const mmb_reset_params* pp, void* stream) {
  using namespace mmb;
  if (!pp) return MMB_EINVAL;
  mmb_reset_params p = *pp;
  if (p.num_envs <= 0 || p.num_rows <= 0) return MMB_EINVAL;
  if (p.task < MMB_TASK_TEN_ANT || p.task > MMB_TASK_INGENUITY) return MMB_EINVAL;
  if ((!p.flags_i64 && !p.flags_u8) || !p.env_ids) return MMB_EINVAL;
  if (p.task != MMB_TASK_INGENUITY && p.dof_state && p.noise_mode == 0 && (!p.noise_pos || !p.noise_vel)) return MMB_EINVAL;
  if (p.noise_mode != 0 && p.noise_mode != 1) return MMB_EINVAL;
  if (p.dof_state && (reinterpret_cast<uintptr_t>(p.dof_state) & 7u)) return MMB_EALIGN;
  {
    LaunchScope ls(K_RESET, (cudaStream_t)stream);
    const int chunks = (p.num_envs + 4095) / 4096;
    if (p.scan_scratch && chunks > 1 && p.task != MMB_TASK_INGENUITY)
      reset_scan_kernel<<<dim3(chunks, p.num_rows), 1024, 0, (cudaStream_t)stream>>>(p);
    else
      reset_kernel<<<p.num_rows, 1024, 0, (cudaStream_t)stream>>>(p);
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}
