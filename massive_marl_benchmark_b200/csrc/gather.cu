// gather.cu - minibatch shuffle + fused multi-field gather.
//
// Replaces RolloutStorage.mini_batch_generator + the nine `.view(-1, .)[indices]` gathers of PPO.update
// (agents/algorithms/rl/ppo/storage.py:75-87, ppo.py:252-264) and SeparatedReplayBuffer.
// feed_forward_generator (agents/algorithms/marl/utils/separated_buffer.py:170-228).  The reference
// materialises a Python list of ints per minibatch and performs one indexed gather per field (9-11 launches,
// each with a host->device copy of the index list).  Here one launch gathers every field of a minibatch into
// contiguous buffers: blockIdx.y = field, every thread moves 16-byte units with four independent loads in flight.
//
// Index modes:  0 = indices supplied (parity mode: the host permutation, e.g. torch.randperm, is the
// reference's);  1 = a stateless bijection on [0,total) keyed by `seed` (fast mode: position j of the epoch's
// permutation is computed, not stored).  The bijection is three rounds of (odd multiply + add, xorshift)
// on the next power of two with cycle walking, hence a permutation by construction.
#include "../../include/mmb.h"
#include "mmb_common.cuh"
#include "mmb_math.cuh"

namespace mmb {
namespace {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct Bijection {
  uint64_t mask, mul[3], add[3];
  int shift;
  int64_t total;
};

__host__ __device__ inline Bijection make_bijection(int64_t total, uint64_t seed) {
  Bijection b;
  int bits = 1;
  while ((1ull << bits) < (uint64_t)total) ++bits;
  b.mask = (bits >= 64) ? ~0ull : ((1ull << bits) - 1ull);
  b.shift = bits > 1 ? bits / 2 : 1;
  b.total = total;
  uint64_t s = seed ^ 0xA5A5A5A55A5A5A5Aull;
  for (int i = 0; i < 3; ++i) {
    b.mul[i] = splitmix64(s) | 1ull;
    b.add[i] = splitmix64(s);
  }
  return b;
}

__device__ __forceinline__ int64_t bijection_eval(const Bijection& b, int64_t pos) {
  uint64_t x = (uint64_t)pos;
  do {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      x = (x * b.mul[i] + b.add[i]) & b.mask;
      x ^= x >> b.shift;
    }
  } while (x >= (uint64_t)b.total);  // cycle walking keeps the map a bijection on [0,total)
  return (int64_t)x;
}

// Grouped shuffle (mmb.h, `group`): the epoch's permutation is a bijection over GROUPS of G consecutive rows plus a
// rotation inside each group, so every random access moves G * row_bytes contiguous bytes (a whole number of 32-byte DRAM
// sectors for every field once G * row_bytes >= 32) instead of one row.
__device__ __forceinline__ int group_rot(int64_t src_group, int log_g) {
  return log_g ? (int)(((uint32_t)src_group * 2654435761u) >> (32 - log_g)) : 0;
}

// Rows are mapped to lane groups, not elements to threads: with `upr` units (of 16 / 4 / 1 bytes) per (super-)row, a warp
// holds 32 / upr whole rows side by side (lane -> (sub-row, unit) fixed for the whole kernel: no per-element division), the
// first lane of each group fetches / computes the source row index and shuffles it to its group, four row groups are in
// flight per warp, and a warp's stores are one contiguous run of the output.  Rows wider than a warp are walked 32 units
// at a time.  With G > 1 a "row" here is a group of G consecutive rows (upr = G * units of one row) read with its rows
// rotated by group_rot.
template <typename V>
__device__ __forceinline__ void gather_field(const mmb_gather_params& p, const Bijection& bj, const V* __restrict__ src,
                                             V* __restrict__ dst, int upr1, int G, int log_g) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int upr = upr1 * G;
  const int64_t B = p.batch_size >> log_g, start = p.batch_start >> log_g;
  auto src_index = [&](int64_t row) -> int64_t {
    if (p.index_mode == 0) return __ldg(p.indices + row);
    if (p.index_mode == 1) return bijection_eval(bj, start + row);
    return start + row;
  };
  auto rotated = [&](int u, int64_t si) -> int {   // unit u of the output group comes from unit (u + rot rows) of the source group
    if (G == 1) return u;
    int v = u + group_rot(si, log_g) * upr1;
    return v >= upr ? v - upr : v;
  };
  if (upr <= 32) {
    const int rpw = 32 / upr;                 // rows per warp pass
    const int sub = lane / upr, unit = lane - sub * upr;
    const bool lane_on = sub < rpw;
    const int leader = sub * upr;
    constexpr int U = 4;
    for (int64_t g0 = warp * U; g0 * rpw < B; g0 += nwarps * U) {
      int64_t srow[U];
      V v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t row = (g0 + u) * rpw + sub;
        int64_t si = 0;
        if (lane_on && unit == 0 && row < B) si = src_index(row);
        srow[u] = __shfl_sync(0xffffffffu, si, lane_on ? leader : 0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t row = (g0 + u) * rpw + sub;
        if (lane_on && row < B) v[u] = __ldg(src + srow[u] * upr + rotated(unit, srow[u]));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t row = (g0 + u) * rpw + sub;
        if (lane_on && row < B) dst[row * upr + unit] = v[u];
      }
    }
  } else {
    for (int64_t row = warp; row < B; row += nwarps) {
      int64_t si = 0;
      if (lane == 0) si = src_index(row);
      si = __shfl_sync(0xffffffffu, si, 0);
      const V* s = src + si * upr;
      V* d = dst + row * upr;
      const int rot = (G == 1) ? 0 : group_rot(si, log_g) * upr1;
      auto su = [&](int u) { int v = u + rot; return v >= upr ? v - upr : v; };
      int u0 = lane;
      for (; u0 + 96 < upr; u0 += 128) {      // four independent loads in flight
        V a0 = __ldg(s + su(u0)), a1 = __ldg(s + su(u0 + 32)), a2 = __ldg(s + su(u0 + 64)), a3 = __ldg(s + su(u0 + 96));
        d[u0] = a0; d[u0 + 32] = a1; d[u0 + 64] = a2; d[u0 + 96] = a3;
      }
      for (; u0 < upr; u0 += 32) d[u0] = __ldg(s + su(u0));
    }
  }
}

// blockIdx.y = field.  Every field is moved in the widest unit its row size and alignment allow (16 / 4 / 1 bytes).
__global__ void __launch_bounds__(256) gather_kernel(const __grid_constant__ mmb_gather_params p, const Bijection bj, const int log_g) {
  const int f = blockIdx.y;
  const int rb = p.row_bytes[f];
  const int G = 1 << log_g;
  const char* s = static_cast<const char*>(p.src[f]);
  char* d = static_cast<char*>(p.dst[f]);
  if (f == 0 && p.indices_out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < p.batch_size; row += stride) {
      int64_t sr;
      if (p.index_mode == 0) sr = __ldg(p.indices + row);
      else if (p.index_mode == 2) sr = p.batch_start + row;
      else {
        const int64_t sg = bijection_eval(bj, (p.batch_start + row) >> log_g);
        sr = (sg << log_g) + ((int)(row & (G - 1)) + group_rot(sg, log_g)) % G;
      }
      p.indices_out[row] = sr;
    }
  }
  if ((rb & 15) == 0 && aligned16(s) && aligned16(d))
    gather_field<uint4>(p, bj, reinterpret_cast<const uint4*>(s), reinterpret_cast<uint4*>(d), rb >> 4, G, log_g);
  else if ((rb & 3) == 0 && ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 3u) == 0)
    gather_field<uint32_t>(p, bj, reinterpret_cast<const uint32_t*>(s), reinterpret_cast<uint32_t*>(d), rb >> 2, G, log_g);
  else
    gather_field<uint8_t>(p, bj, reinterpret_cast<const uint8_t*>(s), reinterpret_cast<uint8_t*>(d), rb, G, log_g);
}

__global__ void __launch_bounds__(256) permutation_kernel(const Bijection bj, int64_t n, int64_t* __restrict__ out, const int log_g) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int G = 1 << log_g;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t sg = bijection_eval(bj, i >> log_g);
    out[i] = (sg << log_g) + ((int)(i & (G - 1)) + group_rot(sg, log_g)) % G;
  }
}

}  // namespace
}  // namespace mmb

using namespace mmb;

static int log2_group(int32_t g) {   // 1, 2, 4, 8, 16 -> 0..4; anything else -> -1
  for (int l = 0; l <= 4; ++l)
    if (g == (1 << l)) return l;
  return g == 0 ? 0 : -1;
}

extern "C" int32_t mmb_shuffle_gather(const mmb_gather_params* pp, void* stream) {
  if (!pp) return MMB_EINVAL;
  mmb_gather_params p = *pp;
  if (p.num_fields <= 0 || p.num_fields > MMB_MAX_GATHER_FIELDS || p.batch_size <= 0 || p.total <= 0) return MMB_EINVAL;
  if (p.index_mode == 0 && !p.indices) return MMB_EINVAL;
  if (p.index_mode < 0 || p.index_mode > 2) return MMB_EINVAL;
  if (p.index_mode >= 1 && (p.batch_start < 0 || p.batch_start + p.batch_size > p.total)) return MMB_EINVAL;
  for (int f = 0; f < p.num_fields; ++f)
    if (!p.src[f] || !p.dst[f] || p.row_bytes[f] <= 0) return MMB_EINVAL;
  const int log_g = log2_group(p.group);
  if (log_g < 0) return MMB_EINVAL;
  const int64_t G = 1ll << log_g;
  if (log_g > 0 && (p.index_mode != 1 || (p.total % G) || (p.batch_start % G) || (p.batch_size % G))) return MMB_EINVAL;
  Bijection bj = make_bijection(p.total >> log_g, p.seed);
  int max_rb = 0;
  for (int f = 0; f < p.num_fields; ++f) max_rb = p.row_bytes[f] > max_rb ? p.row_bytes[f] : max_rb;
  // one warp pass moves 32 units x 4 row groups of the widest field; 8 warps per block
  const int64_t units = p.batch_size * ((max_rb + 15) / 16);
  int64_t blocks = (units + 8 * 128 - 1) / (8 * 128);
  if (blocks < 1) blocks = 1;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  {
    LaunchScope ls(K_GATHER, (cudaStream_t)stream);
    gather_kernel<<<dim3((unsigned)blocks, p.num_fields), 256, 0, (cudaStream_t)stream>>>(p, bj, log_g);
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

extern "C" int32_t mmb_permutation(int64_t n, uint64_t seed, int32_t group, int64_t* out, void* stream) {
  const int log_g = log2_group(group);
  if (n <= 0 || !out || log_g < 0 || (n % (1ll << log_g))) return MMB_EINVAL;
  Bijection bj = make_bijection(n >> log_g, seed);
  int64_t blocks = (n + 255) / 256;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  {
    LaunchScope ls(K_PERM, (cudaStream_t)stream);
    permutation_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(bj, n, out, log_g);
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}
