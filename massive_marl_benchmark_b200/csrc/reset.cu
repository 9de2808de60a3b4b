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
// One CTA of 256 threads per 4096 envs of a flag row, SIXTEEN consecutive envs per thread (one 128-bit load of the uint8
// flags): shuffle scan of the per-thread counts inside a warp, the 8 warp totals scanned by warp 0.  The first version used
// 1024-thread CTAs with 4 flags per thread: next to a running step kernel (4 x 352 threads resident per SM) such a CTA
// waits until two step CTAs of one SM have retired before it can start at all; a 256-thread CTA fits into the slack of
// any SM.  The DOF re-randomisation runs one item per (reset env, DOF pair): one Philox call feeds the two DOFs of the
// pair for all ten ants (the reference draws ONE (len, 8) noise tensor for all ants, ten_ant.py:822-826), instead of one
// call per (env, ant, dof) as before.
#include "../../include/mmb.h"
#include "mmb_common.cuh"
#include "mmb_math.cuh"

namespace mmb {
namespace {

constexpr int RNT = 256;       // threads per CTA
constexpr int RFPT = 16;       // flags per thread
constexpr int RCHUNK = RNT * RFPT;   // 4096 envs per pass / per CTA of the multi-CTA scan

struct TaskShape { int apn, na, nb, dofs, ants; };
__device__ __forceinline__ TaskShape shape_of(int task) {
  if (task == MMB_TASK_TEN_ANT) return {11, 11, 10, 80, 10};
  if (task == MMB_TASK_ONE_ANT) return {2, 2, 1, 8, 1};
  return {4, 4, 4, 16, 0};
}

// the 16 flags of envs [e0, e0 + 16) as a bit mask
__device__ __forceinline__ unsigned load_flags16(const int64_t* f64, const uint8_t* f8, int e0, int N) {
  unsigned bits = 0;
  if (e0 >= N) return 0;
  if (f8 && e0 + RFPT <= N && (reinterpret_cast<uintptr_t>(f8 + e0) & 15u) == 0) {
    const uint4 v = *reinterpret_cast<const uint4*>(f8 + e0);
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if ((w[q] >> (8 * b)) & 0xffu) bits |= 1u << (4 * q + b);
  } else {
#pragma unroll
    for (int j = 0; j < RFPT; ++j) {
      const int e = e0 + j;
      if (e < N && (f64 ? (f64[e] != 0) : (f8[e] != 0))) bits |= 1u << j;
    }
  }
  return bits;
}

// exclusive rank of this thread's first flagged env inside the CTA's 4096-env pass; *total = flagged envs of the pass
__device__ __forceinline__ int block_exclusive_scan(int cnt, int* warp_tot, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const int wt = lane < RNT / 32 ? warp_tot[lane] : 0;
    int wincl = wt;
#pragma unroll
    for (int o = 1; o < RNT / 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, wincl, o);
      if (lane >= o) wincl += v;
    }
    if (lane < RNT / 32) warp_tot[lane] = wincl - wt;   // exclusive prefix of the warp totals
    if (lane == RNT / 32 - 1) *total = wincl;
  }
  __syncthreads();
  return warp_tot[wid] + incl - cnt;
}

__device__ __forceinline__ void write_lists(unsigned bits, int e0, int i, const TaskShape& sh, int64_t* env_ids, int32_t* ia,
                                            int32_t* ib) {
  while (bits) {
    const int j = __ffs(bits) - 1;
    bits &= bits - 1;
    const int e = e0 + j;
    env_ids[i] = e;
    if (ia) for (int q = 0; q < sh.na; ++q) ia[i * sh.na + q] = sh.apn * e + q;
    if (ib) for (int q = 0; q < sh.nb; ++q) ib[i * sh.nb + q] = sh.apn * e + q;
    ++i;
  }
}

// ten_ant.py:822-857 / one_ant.py:371-376 for the reset envs with row ordinals [seg0, seg0 + count): one item per
// (reset env, DOF pair); the pair's (pos, vel, pos, vel) goes out as one 128-bit store per ant
__device__ __forceinline__ void rerandomise_dofs(const mmb_reset_params& p, const TaskShape& sh, int f, int seg0, int count,
                                                 const int64_t* env_ids, float* dof, uint64_t step) {
  const float* npos = p.noise_pos ? p.noise_pos + (int64_t)f * p.noise_row_stride : nullptr;
  const float* nvel = p.noise_vel ? p.noise_vel + (int64_t)f * p.noise_row_stride : nullptr;
  const bool al16 = (reinterpret_cast<uintptr_t>(dof) & 15u) == 0;
  for (int it = threadIdx.x; it < count * 4; it += blockDim.x) {
    const int i = seg0 + (it >> 2), jp = it & 3, j0 = 2 * jp;
    const int e = (int)env_ids[i];
    float np0, np1, nv0, nv1;
    if (p.noise_mode == 0) {
      np0 = npos[(int64_t)i * 8 + j0]; np1 = npos[(int64_t)i * 8 + j0 + 1];
      nv0 = nvel[(int64_t)i * 8 + j0]; nv1 = nvel[(int64_t)i * 8 + j0 + 1];
    } else {  // torch_rand_float(lo, hi) = (hi - lo) * U[0,1) + lo, keyed by env so every ant shares the draw
      const uint4 r = philox4x32_10(make_uint4((uint32_t)e, (uint32_t)(step + f), (uint32_t)jp, 0u),
                                    make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
      np0 = fadd(fmul(0.4f, u01(r.x)), -0.2f); nv0 = fadd(fmul(0.2f, u01(r.y)), -0.1f);
      np1 = fadd(fmul(0.4f, u01(r.z)), -0.2f); nv1 = fadd(fmul(0.2f, u01(r.w)), -0.1f);
    }
    // tensor_clamp = max(min(t, hi), lo)
    const float pos0 = fmaxf(fminf(fadd(p.c.initial_dof_pos[j0], np0), p.c.dof_upper[j0]), p.c.dof_lower[j0]);
    const float pos1 = fmaxf(fminf(fadd(p.c.initial_dof_pos[j0 + 1], np1), p.c.dof_upper[j0 + 1]), p.c.dof_lower[j0 + 1]);
    for (int a = 0; a < sh.ants; ++a) {
      float* d = dof + ((int64_t)e * sh.dofs + a * 8 + j0) * 2;
      if (al16) {
        stg4(d, make_float4(pos0, nv0, pos1, nv1));
      } else {
        *reinterpret_cast<float2*>(d) = make_float2(pos0, nv0);
        *reinterpret_cast<float2*>(d + 2) = make_float2(pos1, nv1);
      }
    }
  }
}

constexpr int INGENUITY_FILL_CTAS = 8;   // extra CTAs (blockIdx.y >= 1) that spread the all-envs rotor-speed write of MultiIngenuity

__global__ void __launch_bounds__(RNT) reset_kernel(const __grid_constant__ mmb_reset_params p) {
  __shared__ int warp_tot[RNT / 32];
  __shared__ int s_chunk_total;
  // per-step path (mmb_ten_ant_env_step): the step kernel behind this launch is a programmatic dependent - it may load and
  // compute its frame while this CTA compacts, and waits for this grid only before it touches the task state
  griddep_launch_dependents();
  const int f = blockIdx.x;
  const int tid = threadIdx.x;
  const int N = p.num_envs;
  const TaskShape sh = shape_of(p.task);
  const int64_t* f64 = p.flags_i64 ? p.flags_i64 + (int64_t)f * p.flags_i64_row_stride : nullptr;
  const uint8_t* f8 = p.flags_u8 ? p.flags_u8 + (int64_t)f * p.flags_u8_row_stride : nullptr;
  if (blockIdx.y > 0) {
    // MultiIngenuity only (multi_ingenuity.py:234-241): whenever reset_idx runs - i.e. any flag of the row is set - the rotor
    // speeds of EVERY env are rewritten.  One CTA doing that for 4 N helicopters (two scattered 4-byte stores each) took ~30 us;
    // here each fill CTA finds out by itself whether the row has a flag (the flags are a few KB, L2-resident) and writes its
    // slice.
    float* dof = p.dof_state ? p.dof_state + (int64_t)f * p.dof_state_row_stride : nullptr;
    if (!dof) return;
    int any = 0;
    for (int e0 = RFPT * tid; e0 < N; e0 += RCHUNK) any |= load_flags16(f64, f8, e0, N) != 0;
    if (!__syncthreads_or(any)) return;
    const int fills = gridDim.y - 1, me = blockIdx.y - 1;
    for (int it = me * RNT + tid; it < N * 4; it += fills * RNT) {
      float* d = dof + (int64_t)it * 8;  // 4 dofs x (pos, vel) per helicopter
      d[3] = -50.0f;
      d[7] = 50.0f;
    }
    return;
  }
  int64_t* env_ids = p.env_ids + (int64_t)f * p.env_ids_row_stride;
  int32_t* ia = p.index_a ? p.index_a + (int64_t)f * p.index_a_row_stride : nullptr;
  int32_t* ib = p.index_b ? p.index_b + (int64_t)f * p.index_b_row_stride : nullptr;
  // Philox counter base: the device-resident word (CUDA-graph replays draw fresh numbers) or the host's value
  const uint64_t step = p.step_counter ? *p.step_counter : p.step;

  // ordered rank of a flagged env = flagged envs of the earlier passes + exclusive scan of the per-thread counts + its
  // rank among the thread's sixteen.  One pass and two barriers for N <= 4096.
  int running = 0;
  for (int base = 0; base < N; base += RCHUNK) {
    const int e0 = base + RFPT * tid;
    const unsigned bits = load_flags16(f64, f8, e0, N);
    const int excl = block_exclusive_scan(__popc(bits), warp_tot, &s_chunk_total);
    write_lists(bits, e0, running + excl, sh, env_ids, ia, ib);
    running += s_chunk_total;
    __syncthreads();                      // warp_tot / s_chunk_total are rewritten by the next pass
  }
  const int count = running;
  if (tid == 0 && p.counts) p.counts[f] = count;
  // a single-CTA launch (the per-step path) advances the device counter itself: every thread has read it before the
  // barriers of the scan above.  Larger grids leave it to the one-thread kernel behind the launch.
  if (tid == 0 && p.step_counter && p.noise_mode == 1 && gridDim.x * gridDim.y == 1) *p.step_counter = step + 1;
  if (count == 0) return;
  __syncthreads();                        // env_ids of this row (written by this CTA) are read back below

  float* dof = p.dof_state ? p.dof_state + (int64_t)f * p.dof_state_row_stride : nullptr;
  if (sh.ants > 0 && dof) rerandomise_dofs(p, sh, f, 0, count, env_ids, dof, step);
  if (p.task == MMB_TASK_INGENUITY) {
    // (the all-envs rotor-speed write of multi_ingenuity.py:234-241 is done by the fill CTAs, blockIdx.y >= 1)
    if (p.forces_state) {  // multi_ingenuity.py:243-244
      const int items = count * 72;
      for (int it = tid; it < items; it += RNT) {
        const int i = it / 72, r = it - i * 72;
        p.forces_state[(int64_t)env_ids[i] * 72 + r] = 0.0f;
      }
    }
  }
}

__global__ void bump_step_counter_kernel(uint64_t* counter, uint64_t by) {
  griddep_launch_dependents();
  *counter += by;
}

// One CTA per 4096-env chunk of a flag row (grid = chunks x rows), decoupled look-back: every CTA publishes its chunk total
// to the caller's scratch, sums the totals of the chunks before it (earlier CTAs in launch order: forward progress as in
// any single-pass scan), and writes its own ordered segment.  The CTA that completes the row's look-backs last clears
// the scratch for the next launch / graph replay.
__global__ void __launch_bounds__(RNT) reset_scan_kernel(const __grid_constant__ mmb_reset_params p) {
  __shared__ int warp_tot[RNT / 32];
  __shared__ int s_running, s_chunk_total;
  griddep_launch_dependents();            // (see reset_kernel)
  const int f = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  unsigned long long* status = reinterpret_cast<unsigned long long*>(p.scan_scratch) + (size_t)f * (chunks + 1);   // [chunks] totals, [chunks] done counter
  const int tid = threadIdx.x, lane = tid & 31;
  const int N = p.num_envs;
  const TaskShape sh = shape_of(p.task);
  const int64_t* f64 = p.flags_i64 ? p.flags_i64 + (int64_t)f * p.flags_i64_row_stride : nullptr;
  const uint8_t* f8 = p.flags_u8 ? p.flags_u8 + (int64_t)f * p.flags_u8_row_stride : nullptr;
  int64_t* env_ids = p.env_ids + (int64_t)f * p.env_ids_row_stride;
  int32_t* ia = p.index_a ? p.index_a + (int64_t)f * p.index_a_row_stride : nullptr;
  int32_t* ib = p.index_b ? p.index_b + (int64_t)f * p.index_b_row_stride : nullptr;

  // Philox counter base, read before this CTA publishes anything: the CTA that sees every look-back of the row complete
  // (below) may then advance the device-resident word
  // (not a conditional expression: with a volatile operand its result is a volatile lvalue, and nvcc 12.9 then reads
  // p.step - a __grid_constant__ member - through a generic address it computes from a null base: a fault)
  uint64_t step = p.step;
  if (p.step_counter) step = *reinterpret_cast<const volatile uint64_t*>(p.step_counter);
  const int e0 = chunk * RCHUNK + RFPT * tid;
  const unsigned bits = load_flags16(f64, f8, e0, N);
  const int excl = block_exclusive_scan(__popc(bits), warp_tot, &s_chunk_total);
  if (tid == 0) {                         // publish this chunk's total: bit 63 = valid
    *reinterpret_cast<volatile unsigned long long*>(status + chunk) = (1ull << 63) | (unsigned long long)s_chunk_total;
    s_running = 0;
  }
  __syncthreads();
  {                                       // look back: totals of chunks 0..chunk-1
    int part = 0;
    for (int j = tid; j < chunk; j += RNT) {
      unsigned long long v;
      do { v = *reinterpret_cast<volatile unsigned long long*>(status + j); } while (!(v >> 63));
      part += (int)(v & 0xffffffffu);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0 && part) atomicAdd(&s_running, part);
  }
  __syncthreads();
  const int seg0 = s_running, seg_n = s_chunk_total;   // this CTA's ordered segment of the row's lists
  write_lists(bits, e0, seg0 + excl, sh, env_ids, ia, ib);
  if (tid == 0) {
    if (chunk == chunks - 1 && p.counts) p.counts[f] = seg0 + seg_n;
    __threadfence();
    if (atomicAdd(status + chunks, 1ull) == (unsigned long long)chunks - 1ull) {   // every CTA of the row has looked back
      for (int j = 0; j <= chunks; ++j) status[j] = 0ull;
      if (p.step_counter && p.noise_mode == 1 && gridDim.y == 1) *p.step_counter = step + 1;   // single row: no kernel behind
    }
  }
  if (seg_n == 0) return;
  __syncthreads();                        // this CTA's env_ids segment is read back below

  float* dof = p.dof_state ? p.dof_state + (int64_t)f * p.dof_state_row_stride : nullptr;
  if (sh.ants > 0 && dof) rerandomise_dofs(p, sh, f, seg0, seg_n, env_ids, dof, step);
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
    const int chunks = (p.num_envs + RCHUNK - 1) / RCHUNK;
    if (p.scan_scratch && chunks > 1 && p.task != MMB_TASK_INGENUITY)
      reset_scan_kernel<<<dim3(chunks, p.num_rows), RNT, 0, (cudaStream_t)stream>>>(p);
    else
      reset_kernel<<<dim3(p.num_rows, (p.task == MMB_TASK_INGENUITY && p.dof_state) ? 1 + INGENUITY_FILL_CTAS : 1), RNT, 0,
                     (cudaStream_t)stream>>>(p);
    // a single-row launch advances the device counter itself (reset_kernel: its only CTA; reset_scan_kernel: the CTA that
    // completes the row); several rows - the horizon-batched reset lists - leave it to a one-thread kernel behind the launch
    const bool in_kernel = p.num_rows == 1 && p.task != MMB_TASK_INGENUITY;
    if (p.step_counter && p.noise_mode == 1 && !in_kernel)
      bump_step_counter_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(p.step_counter, (uint64_t)p.num_rows);
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}
