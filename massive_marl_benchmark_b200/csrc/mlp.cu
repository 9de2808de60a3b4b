// mlp.cu - actor-critic MLP forward on the 5th-generation tensor cores (tcgen05 + TMEM), the one dense
// contraction on the rollout path.
//
// Replaces the per-layer `nn.Linear -> ELU [-> LayerNorm]` chains of
//   PPO   ActorCritic.act / evaluate      agents/algorithms/rl/ppo/module.py:25-55,73-107   (Linear, ELU)
//   MARL  Actor / Critic forward          agents/algorithms/marl/actor_critic.py:42-69,149-168,
//                                         agents/algorithms/utils/mlp.py:6-65                (LayerNorm, Linear, ELU, LayerNorm)
// which the reference runs as fp32 cuBLAS SGEMMs plus separate elementwise / LayerNorm kernels.
//
// One launch = one layer:  Y[M,N] = epilogue(X[M,K] . W[N,K]^T + b)
//   operands   bf16 (kind::f16), K-major for both: X rows and nn.Linear weight rows are contiguous in K
//   accumulate fp32 in TMEM (128 lanes x up to 512 columns per CTA)
//   epilogue   fp32: bias, ELU, optional LayerNorm over the full row (one thread owns one row after tcgen05.ld),
//              output bf16 (next layer's operand) or fp32 (last layer)
// CTA = 128 rows x n_tile columns (n_tile <= 512), 128 threads.  K is walked in blocks of 64 (= one 128-byte
// swizzle row of bf16); a 2-stage shared-memory ring lets the tensor core work on block k while the threads
// stage block k+1.  Shared-memory tiles use the canonical K-major SWIZZLE_128B layout (8-row x 128-byte atoms,
// 16-byte chunk index XOR row%8), described to the tensor core by 64-bit matrix descriptors; a single thread
// issues tcgen05.mma and tcgen05.commit arrives on an mbarrier when the stage may be overwritten.
//
// Round-1 state: correct and validated against fp32 torch; tiles are staged by the threads (LDG -> swizzled STS)
// rather than by TMA and there is no warp specialisation yet - see DESIGN.md section 7.
#include <cuda_bf16.h>

#include "../../include/mmb.h"
#include "mmb_common.cuh"
#include "mmb_math.cuh"

namespace mmb {
namespace {

constexpr int BM = 128;       // rows per CTA = TMEM lanes = UMMA_M
constexpr int BK = 64;        // bf16 elements per k-block (128 bytes)
constexpr int MLP_THREADS = 128;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int MAX_STAGES = 4;
constexpr int SMEM_BUDGET = 200 * 1024;

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits
// [0,14), leading byte offset (ignored for swizzled K-major, canonical value 1) in [16,30), stride byte offset =
// 1024 B between 8-row groups in [32,46), descriptor version 1 in [46,48), layout type 2 = SWIZZLE_128B in [61,64).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_byte_addr) {
  return (uint64_t)((smem_byte_addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

// Instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, both K-major, dense.
__device__ __forceinline__ uint32_t umma_idesc_bf16(int umma_n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(umma_n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 consecutive fp32 accumulator columns of this thread's TMEM lane (row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float elu1(float x) { return x > 0.0f ? x : expm1f(x); }  // nn.ELU(alpha=1)

// stage `rows` x 64 bf16 (row-major, leading dimension ld) into the swizzled tile at `tile` with asynchronous 16-byte
// copies (cp.async / LDGSTS): nothing waits here, the caller commits a group per k-block
__device__ __forceinline__ void stage_tile_async(uint8_t* tile, const __nv_bfloat16* __restrict__ g, int64_t ld, int rows, int tid) {
  const int chunks = rows * 8;  // 16-byte chunks
  const uint32_t base = smem_u32(tile);
  for (int i = tid; i < chunks; i += MLP_THREADS) {
    const int r = i >> 3, c = i & 7;
    const uint32_t dst = base + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g + (int64_t)r * ld + c * 8) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(MLP_THREADS, 1) mlp_layer_kernel(const __grid_constant__ mmb_mlp_layer_params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t mma_done[MAX_STAGES];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * p.n_tile;
  const int n_tile = p.n_tile;
  const int stage_bytes = A_STAGE_BYTES + n_tile * BK * 2;   // [A tile | B tile], both multiples of 1024 B
  const int S = p.stages;

  if (tid == 0)
    for (int i = 0; i < S; ++i) mbar_init(&mma_done[i], 1);
  if (warp == 0) {  // one warp allocates all 512 TMEM columns (1 CTA per SM) and publishes the base address
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  const __nv_bfloat16* X = static_cast<const __nv_bfloat16*>(p.x) + (int64_t)m0 * p.Kpad;
  const __nv_bfloat16* W = static_cast<const __nv_bfloat16*>(p.w) + (int64_t)n0 * p.Kpad;
  const int nkb = p.Kpad / BK;
  const int umma_n = n_tile > 256 ? 256 : n_tile;
  const uint32_t idesc = umma_idesc_bf16(umma_n);

  // S-stage ring: k-blocks kb+1 .. kb+S-1 are in flight (cp.async) while the tensor core works on block kb.
  // A stage is refilled only after the tcgen05.commit of the MMAs that read it has arrived on its mbarrier.
  auto issue_loads = [&](int kb) {
    uint8_t* st = smem + (kb % S) * stage_bytes;
    stage_tile_async(st, X + kb * BK, p.Kpad, BM, tid);
    stage_tile_async(st + A_STAGE_BYTES, W + kb * BK, p.Kpad, n_tile, tid);
  };
  for (int kb = 0; kb < S - 1; ++kb) {
    if (kb < nkb) issue_loads(kb);
    cp_async_commit();
  }
  for (int kb = 0; kb < nkb; ++kb) {
    const int s = kb % S;
    cp_async_wait<MAX_STAGES - 2>();   // conservative for S < MAX_STAGES: see the group accounting below
    if (S < MAX_STAGES) {               // exact wait: all but the newest S-2 groups are complete
      if (S == 2) cp_async_wait<0>();
      else if (S == 3) cp_async_wait<1>();
    }
    fence_async_smem();  // LDGSTS / generic-proxy writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + s * stage_bytes), b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
      for (int j = 0; j < BK / 16; ++j) {  // UMMA_K = 16 bf16 = 32 bytes along the swizzled row
        const bool acc = (kb > 0) || (j > 0);
        umma_bf16(tmem, umma_desc_k_sw128(a_addr + j * 32), umma_desc_k_sw128(b_addr + j * 32), idesc, acc);
        if (n_tile > 256)
          umma_bf16(tmem + 256, umma_desc_k_sw128(a_addr + j * 32), umma_desc_k_sw128(b_addr + 256 * 128 + j * 32), idesc, acc);
      }
      umma_commit(&mma_done[s]);  // arrives when every MMA issued so far has completed
    }
    // refill the stage that block kb-1 used with block kb+S-1 (its MMAs were issued one iteration ago)
    const int kn = kb + S - 1;
    if (kn < nkb) {
      if (kb >= 1) mbar_wait(&mma_done[kn % S], (uint32_t)((((kb - 1) / S)) & 1));
      issue_loads(kn);
    }
    cp_async_commit();
  }
  mbar_wait(&mma_done[(nkb - 1) % S], (uint32_t)(((nkb - 1) / S) & 1));
  tc_fence_after();

  // ---- epilogue: thread = one accumulator row (TMEM lane 32*warp + lane) ----
  const int row = warp * 32 + lane;
  const int m = m0 + row;
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  const bool row_ok = m < p.M;
  float v[32];
  if (p.epilogue == 2) {  // bias + ELU + LayerNorm over the whole row (n_tile == N): two passes over TMEM
    float sum = 0.0f, sumsq = 0.0f;
    for (int c0 = 0; c0 < n_tile; c0 += 32) {
      tmem_ld32(taddr + c0, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x = elu1(v[i] + __ldg(p.bias + n0 + c0 + i));
        sum += x;
        sumsq += x * x;
      }
    }
    const float mean = sum / (float)p.N;
    const float var = fmaxf(sumsq / (float)p.N - mean * mean, 0.0f);  // biased variance, as nn.LayerNorm
    const float rstd = rsqrtf(var + p.ln_eps);
    for (int c0 = 0; c0 < n_tile; c0 += 32) {
      tmem_ld32(taddr + c0, v);
      if (row_ok) {
        __nv_bfloat16* y = static_cast<__nv_bfloat16*>(p.y) + (int64_t)m * p.y_stride + n0 + c0;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int n = n0 + c0 + i;
          const float x0 = (elu1(v[i] + __ldg(p.bias + n)) - mean) * rstd * __ldg(p.ln_gamma + n) + __ldg(p.ln_beta + n);
          const float x1 = (elu1(v[i + 1] + __ldg(p.bias + n + 1)) - mean) * rstd * __ldg(p.ln_gamma + n + 1) + __ldg(p.ln_beta + n + 1);
          *reinterpret_cast<__nv_bfloat162*>(y + i) = __floats2bfloat162_rn(x0, x1);
        }
      }
    }
  } else {
    for (int c0 = 0; c0 < n_tile; c0 += 32) {
      tmem_ld32(taddr + c0, v);
      if (!row_ok) continue;
      if (p.epilogue == 1) {  // bias + ELU -> bf16
        __nv_bfloat16* y = static_cast<__nv_bfloat16*>(p.y) + (int64_t)m * p.y_stride + n0 + c0;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int n = n0 + c0 + i;
          if (n < p.N) {  // columns >= N of the (zero-initialised, K-padded) activation buffer stay zero
            const float x0 = elu1(v[i] + __ldg(p.bias + n));
            const float x1 = (n + 1) < p.N ? elu1(v[i + 1] + __ldg(p.bias + n + 1)) : 0.0f;
            *reinterpret_cast<__nv_bfloat162*>(y + i) = __floats2bfloat162_rn(x0, x1);
          }
        }
      } else {  // bias -> fp32 (last layer)
        float* y = static_cast<float*>(p.y) + (int64_t)m * p.y_stride + n0 + c0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int n = n0 + c0 + i;
          if (n < p.N) y[i] = v[i] + __ldg(p.bias + n);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// fp32 [M][K] -> optional LayerNorm (mlp.py:58-59 feature_norm) -> bf16 [Mpad][Kpad], zero padded: the first
// layer's A operand.  One warp per row.
__global__ void __launch_bounds__(256) ln_cast_kernel(const float* __restrict__ x, int M, int Mpad, int K, int Kpad,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      float eps, int use_ln, __nv_bfloat16* __restrict__ y) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= Mpad) return;
  __nv_bfloat16* yr = y + (int64_t)warp * Kpad;
  if (warp >= M) {
    for (int k = lane; k < Kpad; k += 32) yr[k] = __float2bfloat16(0.0f);
    return;
  }
  const float* xr = x + (int64_t)warp * K;
  float mean = 0.0f, rstd = 1.0f;
  if (use_ln) {
    float s = 0.0f;
    for (int k = lane; k < K; k += 32) s += xr[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    mean = s / (float)K;
    float q = 0.0f;
    for (int k = lane; k < K; k += 32) { const float d = xr[k] - mean; q += d * d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    rstd = rsqrtf(q / (float)K + eps);
  }
  for (int k = lane; k < Kpad; k += 32) {
    float v = 0.0f;
    if (k < K) v = use_ln ? (xr[k] - mean) * rstd * gamma[k] + beta[k] : xr[k];
    yr[k] = __float2bfloat16(v);
  }
}

}  // namespace
}  // namespace mmb

using namespace mmb;

extern "C" int32_t mmb_mlp_layer(const mmb_mlp_layer_params* pp, void* stream) {
  if (!pp) return MMB_EINVAL;
  mmb_mlp_layer_params p = *pp;
  if (p.M <= 0 || p.N <= 0 || p.K <= 0 || !p.x || !p.w || !p.bias || !p.y) return MMB_EINVAL;
  if (p.Kpad % BK || p.Kpad < p.K || p.n_tile % 32 || p.n_tile < 32 || p.n_tile > 512 || p.Npad % p.n_tile || p.Npad < p.N)
    return MMB_EINVAL;
  if (p.n_tile > 256 && p.n_tile != 512) return MMB_EINVAL;
  if (p.Mpad % BM || p.Mpad < p.M) return MMB_EINVAL;
  if (p.epilogue < 0 || p.epilogue > 2) return MMB_EINVAL;
  if (p.epilogue == 2 && (p.n_tile != p.N || p.Npad != p.N || !p.ln_gamma || !p.ln_beta)) return MMB_EINVAL;
  if ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.w)) & 15u) return MMB_EALIGN;
  const int stage_bytes = A_STAGE_BYTES + p.n_tile * BK * 2;
  int stages = SMEM_BUDGET / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return MMB_EUNSUPPORTED;
  p.stages = stages;
  const int smem = stages * stage_bytes;
  static bool attr_done[MMB_MAX_DEVICES] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= MMB_MAX_DEVICES) return MMB_EUNSUPPORTED;
  if (!attr_done[dev]) {
    if (cudaFuncSetAttribute(mlp_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET) != cudaSuccess)
      return MMB_ECUDA;
    attr_done[dev] = true;
  }
  {
    LaunchScope ls(K_MLP_LAYER, (cudaStream_t)stream);
    mlp_layer_kernel<<<dim3(p.Mpad / BM, p.Npad / p.n_tile), MLP_THREADS, smem, (cudaStream_t)stream>>>(p);
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

extern "C" int32_t mmb_ln_cast(const float* x, int32_t M, int32_t Mpad, int32_t K, int32_t Kpad, const float* gamma,
                               const float* beta, float eps, int32_t use_ln, void* y_bf16, void* stream) {
  if (!x || !y_bf16 || M <= 0 || Mpad < M || K <= 0 || Kpad < K) return MMB_EINVAL;
  if (use_ln && (!gamma || !beta)) return MMB_EINVAL;
  {
    LaunchScope ls(K_LN_CAST, (cudaStream_t)stream);
    ln_cast_kernel<<<(Mpad * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, M, Mpad, K, Kpad, gamma, beta, eps, use_ln,
                                                                            static_cast<__nv_bfloat16*>(y_bf16));
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}
