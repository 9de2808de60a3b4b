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

template <typename V>
__device__ __forceinline__ void gather_field(const mmb_gather_params& p, const Bijection& bj, const V* __restrict__ src,
                                             V* __restrict__ dst, int units_per_row) {
  // element i of the output = (row i / upr, unit i % upr); consecutive threads write consecutive units (coalesced
  // stores), the loads of four independent elements are in flight per thread
  const int64_t total = p.batch_size * units_per_row;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto src_index = [&](int64_t row) -> int64_t {
    if (p.index_mode == 0) return __ldg(p.indices + row);
    if (p.index_mode == 1) return bijection_eval(bj, p.batch_start + row);
    return p.batch_start + row;
  };
  for (; i + 3 * stride < total; i += 4 * stride) {
    V v[4];
    int64_t off[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t e = i + u * stride;
      const int64_t row = e / units_per_row;
      const int unit = (int)(e - row * units_per_row);
      off[u] = src_index(row) * units_per_row + unit;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(src + off[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) dst[i + u * stride] = v[u];
  }
  for (; i < total; i += stride) {
    const int64_t row = i / units_per_row;
    const int unit = (int)(i - row * units_per_row);
    dst[i] = __ldg(src + src_index(row) * units_per_row + unit);
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
  int64_t blocks = (p.batch_size * ((max_rb + 15) / 16) + 1023) / 1024;  // ~4 elements per thread for the widest field
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
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
  if (blocks > 148 * 16) blocks = 148 * 16;
  {
    LaunchScope ls(K_PERM, (cudaStream_t)stream);
    permutation_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(bj, n, out);
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}
