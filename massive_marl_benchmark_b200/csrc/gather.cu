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

// Rows are mapped to lane groups, not elements to threads: with `upr` units (of 16 / 4 / 1 bytes) per row, a warp holds
// 32 / upr whole rows side by side (lane -> (sub-row, unit) fixed for the whole kernel: no per-element division), the
// first lane of each group fetches / computes the source row index and shuffles it to its group, four row groups are in
// flight per warp, and a warp's stores are one contiguous run of the output.  Rows wider than a warp are walked 32 units
// at a time.
template <typename V>
__device__ __forceinline__ void gather_field(const mmb_gather_params& p, const Bijection& bj, const V* __restrict__ src,
                                             V* __restrict__ dst, int upr) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t B = p.batch_size;
  auto src_index = [&](int64_t row) -> int64_t {
    if (p.index_mode == 0) return __ldg(p.indices + row);
    if (p.index_mode == 1) return bijection_eval(bj, p.batch_start + row);
    return p.batch_start + row;
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
        if (lane_on && row < B) v[u] = __ldg(src + srow[u] * upr + unit);
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
      int u0 = lane;
      for (; u0 + 96 < upr; u0 += 128) {      // four independent loads in flight
        V a0 = __ldg(s + u0), a1 = __ldg(s + u0 + 32), a2 = __ldg(s + u0 + 64), a3 = __ldg(s + u0 + 96);
        d[u0] = a0; d[u0 + 32] = a1; d[u0 + 64] = a2; d[u0 + 96] = a3;
      }
      for (; u0 < upr; u0 += 32) d[u0] = __ldg(s + u0);
    }
  }
}

// blockIdx.y = field.  Every field is moved in the widest unit its row size and alignment allow (16 / 4 / 1 bytes).
__global__ void __launch_bounds__(256) gather_kernel(const __grid_constant__ mmb_gather_params p, const Bijection bj) {
  const int f = blockIdx.y;
  const int rb = p.row_bytes[f];
  const char* s = static_cast<const char*>(p.src[f]);
  char* d = static_cast<char*>(p.dst[f]);
  if (f == 0 && p.indices_out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < p.batch_size; row += stride) {
      int64_t sr = p.index_mode == 0 ? __ldg(p.indices + row) : (p.index_mode == 1 ? bijection_eval(bj, p.batch_start + row) : p.batch_start + row);
      p.indices_out[row] = sr;
    }
  }
  if ((rb & 15) == 0 && aligned16(s) && aligned16(d))
    gather_field<uint4>(p, bj, reinterpret_cast<const uint4*>(s), reinterpret_cast<uint4*>(d), rb >> 4);
  else if ((rb & 3) == 0 && ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 3u) == 0)
    gather_field<uint32_t>(p, bj, reinterpret_cast<const uint32_t*>(s), reinterpret_cast<uint32_t*>(d), rb >> 2);
  else
    gather_field<uint8_t>(p, bj, reinterpret_cast<const uint8_t*>(s), reinterpret_cast<uint8_t*>(d), rb);
}

__global__ void __launch_bounds__(256) permutation_kernel(const Bijection bj, int64_t n, int64_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = bijection_eval(bj, i);
}

}  // namespace
}  // namespace mmb

using namespace mmb;

extern "C" int32_t mmb_shuffle_gather(const mmb_gather_params* pp, void* stream) {
  if (!pp) return MMB_EINVAL;
  mmb_gather_params p = *pp;
  if (p.num_fields <= 0 || p.num_fields > MMB_MAX_GATHER_FIELDS || p.batch_size <= 0 || p.total <= 0) return MMB_EINVAL;
  if (p.index_mode == 0 && !p.indices) return MMB_EINVAL;
  if (p.index_mode < 0 || p.index_mode > 2) return MMB_EINVAL;
  if (p.index_mode >= 1 && (p.batch_start < 0 || p.batch_start + p.batch_size > p.total)) return MMB_EINVAL;
  for (int f = 0; f < p.num_fields; ++f)
    if (!p.src[f] || !p.dst[f] || p.row_bytes[f] <= 0) return MMB_EINVAL;
  Bijection bj = make_bijection(p.total, p.seed);
  int max_rb = 0;
  for (int f = 0; f < p.num_fields; ++f) max_rb = p.row_bytes[f] > max_rb ? p.row_bytes[f] : max_rb;
  // one warp pass moves 32 units x 4 row groups of the widest field; 8 warps per block
  const int64_t units = p.batch_size * ((max_rb + 15) / 16);
  int64_t blocks = (units + 8 * 128 - 1) / (8 * 128);
  if (blocks < 1) blocks = 1;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  {
    LaunchScope ls(K_GATHER, (cudaStream_t)stream);
    gather_kernel<<<dim3((unsigned)blocks, p.num_fields), 256, 0, (cudaStream_t)stream>>>(p, bj);
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

extern "C" int32_t mmb_permutation(int64_t n, uint64_t seed, int64_t* out, void* stream) {
  if (n <= 0 || !out) return MMB_EINVAL;
  Bijection bj = make_bijection(n, seed);
  int64_t blocks = (n + 255) / 256;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  {
    LaunchScope ls(K_PERM, (cudaStream_t)stream);
    permutation_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(bj, n, out);
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}
