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
// CTA = 128 rows x n_tile columns (n_tile <= 512).  K is walked in blocks of 64 (= one 128-byte swizzle row of bf16)
// through a shared-memory ring of up to 8 stages.  Shared-memory tiles use the canonical K-major SWIZZLE_128B layout
// (8-row x 128-byte atoms, 16-byte chunk index XOR row%8), described to the tensor core by 64-bit matrix descriptors; a
// single thread issues tcgen05.mma, and tcgen05.commit arrives on an mbarrier when a stage may be overwritten.
//
// Kernels (they share the descriptors and the epilogue math):
//   mlp_layer_ws_kernel (default)  warp-specialised, 320 threads: warp 0 = TMA producer (2-D tensor maps, SWIZZLE_128B
//       boxes, one elected lane), warp 1 = tcgen05.mma issuer + TMEM owner, warps 2-9 = epilogue (TMEM lane quarter =
//       warp % 4, two warps per quarter split the columns).  full / empty mbarrier ring between producer and MMA;
//       tcgen05.commit frees a stage and finally publishes the accumulator.  Epilogue: tcgen05.ld -> bias / ELU
//       (ex2.approx) / LayerNorm in fp32 -> output tile staged in the drained operand ring in the TMA swizzle -> 2-D TMA
//       stores, each finished 128-byte-wide sub-tile handed over while the next is computed.  With overlap_prev the launch
//       is programmatic-dependent: prologue and the first ring of WEIGHT tiles run while the previous layer drains, only
//       the activation loads wait for it.  Optional modes, measured and off by default:
//       MMB_MLP_CLUSTER = 2 / 4 (weight tile multicast across a cluster), MMB_MLP_PAIR = 1 (cta_group::2).
//   mlp_layer_ws_pair_kernel       the cta_group::2 instantiation (a kernel containing cta_group::2 instructions can only
//       be launched as a cluster, hence its own entry)
//   mlp_layer_ws_group_kernel      grid z = one of up to 16 independent problems of identical geometry (per-agent nets)
//   mlp_layer_kernel (MMB_MLP_VARIANT=legacy)  first version: all threads stage tiles with cp.async, barrier per k-block
//   ln_cast_kernel / ln_cast_group_kernel      fp32 input -> optional LayerNorm -> bf16 zero-padded first operand
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/mmb.h"
#include "mmb_common.cuh"
#include "mmb_math.cuh"

namespace mmb {
namespace {

constexpr int BM = 128;       // rows per CTA = TMEM lanes = UMMA_M
constexpr int BK = 64;        // bf16 elements per k-block (128 bytes)
constexpr int MLP_THREADS = 128;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int MAX_STAGES = 8;   // ring depth = min(8, smem budget / stage bytes): 4 at n_tile 256, 6 at 128, 8 below
constexpr int SMEM_BUDGET = 200 * 1024;

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits
// [0,14), leading byte offset (ignored for swizzled K-major, canonical value 1) in [16,30), stride byte offset =
// 1024 B between 8-row groups in [32,46), descriptor version 1 in [46,48), layout type 2 = SWIZZLE_128B in [61,64).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_byte_addr) {
  return (uint64_t)((smem_byte_addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

// Instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, both K-major, dense.
__device__ __forceinline__ uint32_t umma_idesc_bf16(int umma_n, int umma_m = BM) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(umma_n >> 3) << 17) | ((uint32_t)(umma_m >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// accumulate = 1 with a constant predicate (the per-instruction setp of umma_bf16 is one of the things that made the MMA
// issuer's loop - a single thread whose instruction latencies are the k-loop's clock - 345 ns per k-block whatever the tile)
__device__ __forceinline__ void umma_bf16_acc(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
               "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
}
// kind::tf32: fp32 operands in shared memory (the tensor core reads the upper 19 bits), 8 elements of K per instruction =
// the same 32 bytes of a SWIZZLE_128B row as 16 bf16, at half the rate.
__device__ __forceinline__ uint32_t umma_idesc_tf32(int umma_n, int umma_m = BM) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(umma_n >> 3) << 17) | ((uint32_t)(umma_m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ float rn_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
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

// The same load without the wait: the registers are valid after tmem_ld_wait(r) (which names them, so that the compiler
// cannot move their uses above the wait).  Lets the next 32 columns travel while the current ones are processed.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                 "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                 "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :: "memory");
}

// nn.ELU(alpha=1).  The negative branch is exp(x) - 1 through ex2.approx (2 ulp): its absolute error of ~1e-7 is far below
// the bf16 rounding of the stored activation (2^-9 relative) even where x -> 0 and the subtraction cancels; expm1f costs
// ~35 instructions per element and made the epilogue, not the MMA, the longest phase of a tile.
__device__ __forceinline__ float elu1(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
  return x > 0.0f ? x : e - 1.0f;
}
__device__ __forceinline__ void load_bias32(const float* __restrict__ b, float* out) {  // 32 consecutive floats, 16 B aligned
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(b) + i);
    out[4 * i] = v.x; out[4 * i + 1] = v.y; out[4 * i + 2] = v.z; out[4 * i + 3] = v.w;
  }
}

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

__device__ __forceinline__ void epilogue_row(const mmb_mlp_layer_params& p, uint32_t taddr, int m, int n0, int n_tile, int cb, int ce);

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
    if (S == 2) cp_async_wait<0>();     // all but the newest S-2 groups are complete
    else if (S == 3) cp_async_wait<1>();
    else cp_async_wait<2>();
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
  epilogue_row(p, tmem + ((uint32_t)(warp * 32) << 16), m0 + warp * 32 + lane, n0, n_tile, 0, n_tile);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// ------------------------------------------------------------------------------------------------------
// warp-specialised, TMA-fed variant
// ------------------------------------------------------------------------------------------------------
constexpr int WS_THREADS = 320;  // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter, half the columns each)

__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst_smem)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst_smem, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   smem_u32(dst_smem)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
// multicast variants for thread-block clusters: the tile lands at the same shared-memory offset, and completes on the
// mbarrier at the same offset, in every CTA of `cta_mask`
__device__ __forceinline__ void tma_load_2d_mc(void* dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
          smem_u32(dst_smem)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask) : "memory");
}
// ---- cta_group::2 (a CTA pair = the two SMs of a TPC run ONE 256-row MMA): instruction forms as in CUTLASS
// (cute/arch/copy_sm100_tma.hpp SM100_TMA_2SM_LOAD_2D, mma_sm100_umma.hpp SM100_MMA_F16BF16_2x1SM_SS, cutlass/arch/barrier.h
// umma_arrive_multicast_2x1SM).  Both CTAs load into their own shared memory; the transaction bytes of both land on the
// LEADER's (rank 0) mbarrier: clearing bit 24 of a shared::cluster address selects the same offset in the pair's CTA 0.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_2sm(void* dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst_smem)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate), "r"(z)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask) : "memory");
}
// bounded mbarrier wait for the paired kernel: a protocol error is recorded (first one wins) and the kernel runs on with
// garbage instead of hanging the GPU; mmb_mlp_debug_status() reads the record
__device__ unsigned int g_pair_dbg[4];
__device__ __forceinline__ void mbar_wait_or_flag(uint64_t* bar, uint32_t parity, unsigned code) {
  const long long t0 = clock64();
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!done && clock64() - t0 > 100000000ll) {   // ~50 ms
      if (atomicCAS(&g_pair_dbg[0], 0u, code) == 0u) { g_pair_dbg[1] = blockIdx.x; g_pair_dbg[2] = blockIdx.y; g_pair_dbg[3] = parity; }
      return;
    }
  }
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// Deferred LayerNorm (include/mmb.h): mean / rstd of input row m from the producing layer's partial sums (L2 reads: the
// producer was the previous kernel in the stream), and this launch's own partials of its bf16-rounded outputs.
__device__ __forceinline__ void ln_row_moments(const mmb_mlp_layer_params& p, int m, float& mean, float& rstd) {
  const float2* st = reinterpret_cast<const float2*>(p.ln_in_stats) + (int64_t)m * p.ln_in_parts;
  float s = 0.0f, q = 0.0f;
  int i = 0;
  for (; i + 4 <= p.ln_in_parts; i += 4) {        // four independent L2 loads in flight, the sums in chunk order
    const float2 v0 = __ldcg(st + i), v1 = __ldcg(st + i + 1), v2 = __ldcg(st + i + 2), v3 = __ldcg(st + i + 3);
    s += v0.x; q += v0.y; s += v1.x; q += v1.y; s += v2.x; q += v2.y; s += v3.x; q += v3.y;
  }
  for (; i < p.ln_in_parts; ++i) { const float2 v = __ldcg(st + i); s += v.x; q += v.y; }
  const float inv_n = 1.0f / (float)p.ln_in_n;
  mean = s * inv_n;
  rstd = rsqrtf(fmaxf(q * inv_n - mean * mean, 0.0f) + p.ln_in_eps);      // biased variance, as nn.LayerNorm
}
// one partial per 64 columns (an epilogue thread's share of a 256-column tile is two or one of them), summed in column order:
// 64 bytes = two sectors of partials per 512-wide row for the consumer to read, and the same bits however many warps split a tile
__device__ __forceinline__ void ln_store_partial(const mmb_mlp_layer_params& p, int m, int n, float psum, float psq) {
  reinterpret_cast<float2*>(p.ln_out_stats)[(int64_t)m * (p.Npad >> 6) + (n >> 6)] = make_float2(psum, psq);
}

// epilogue of columns [cb, ce) of one accumulator row held in a TMEM lane (shared by both kernels; the LayerNorm
// epilogue needs the whole row: cb = 0, ce = n_tile)
__device__ __forceinline__ void epilogue_row(const mmb_mlp_layer_params& p, uint32_t taddr, int m, int n0, int n_tile, int cb, int ce) {
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
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int n = n0 + c0 + i;
          const float x0 = (elu1(v[i] + __ldg(p.bias + n)) - mean) * rstd * __ldg(p.ln_gamma + n) + __ldg(p.ln_beta + n);
          const float x1 = (elu1(v[i + 1] + __ldg(p.bias + n + 1)) - mean) * rstd * __ldg(p.ln_gamma + n + 1) + __ldg(p.ln_beta + n + 1);
          __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(y)[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
      }
    }
  } else {
    float mean = 0.0f, rstd = 1.0f;
    const bool corr = p.ln_in_stats != nullptr;
    if (corr) ln_row_moments(p, m, mean, rstd);
    float psum = 0.0f, psq = 0.0f;
    for (int c0 = cb; c0 < ce; c0 += 32) {
      tmem_ld32(taddr + c0, v);
      if (!row_ok) continue;
      if (corr) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = (n0 + c0 + i < p.N) ? rstd * (v[i] - mean * __ldg(p.ln_c + n0 + c0 + i)) : 0.0f;
      }
      if (p.epilogue == 1) {  // bias + ELU -> bf16
        __nv_bfloat16* y = static_cast<__nv_bfloat16*>(p.y) + (int64_t)m * p.y_stride + n0 + c0;
        if (n0 + c0 + 32 <= p.N) {  // full group: 64 contiguous bytes of the row as four 16-byte stores
          uint32_t pk[16];
          float bv[32];
          load_bias32(p.bias + n0 + c0, bv);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            __nv_bfloat162 h = __floats2bfloat162_rn(elu1(v[i] + bv[i]), elu1(v[i + 1] + bv[i + 1]));
            pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
            if (p.ln_out_stats) { const float2 f = __bfloat1622float2(h); psum += f.x; psq += f.x * f.x; psum += f.y; psq += f.y * f.y; }   // (column order, as the staged path)
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(y)[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const int n = n0 + c0 + i;
            if (n < p.N) {  // columns >= N of the (zero-initialised, K-padded) activation buffer stay zero
              const float x0 = elu1(v[i] + __ldg(p.bias + n));
              const float x1 = (n + 1) < p.N ? elu1(v[i + 1] + __ldg(p.bias + n + 1)) : 0.0f;
              const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
              *reinterpret_cast<__nv_bfloat162*>(y + i) = h;
              if (p.ln_out_stats) { const float2 f = __bfloat1622float2(h); psum += f.x; psq += f.x * f.x; psum += f.y; psq += f.y * f.y; }   // (column order, as the staged path)
            }
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
      if (p.ln_out_stats && p.epilogue == 1 && ((c0 + 32) & 63) == 0) { ln_store_partial(p, m, n0 + c0, psum, psq); psum = 0.0f; psq = 0.0f; }
    }
  }
}

// Epilogue of the warp-specialised kernel: columns [cb, ce) of one accumulator row -> the output tile in shared memory
// (the drained operand ring), laid out as 128-byte-wide sub-tiles in the TMA SWIZZLE_128B pattern (16-byte chunk index
// XOR row % 8: conflict-free for the row-per-lane stores), from where ONE thread writes the whole tile with 2-D TMA
// stores.  Row-per-thread global stores were the longest phase of a tile: every 16-byte store instruction of a warp
// touched 32 different sectors.  Out-of-range rows / columns are clipped by the tensor map; columns in [N, y_stride)
// of a bf16 activation buffer are written as zeros (they are the K padding of the next layer).
template <bool F32>
__device__ __forceinline__ void stage_out32(uint8_t* tile, int row, int c0, const float* x, int sub_stride = BM * 128) {
  // bf16: 64 columns per 128-byte sub-tile row, this group = 4 chunks; fp32: 32 columns per sub-tile row = 8 chunks
  if (F32) {
    uint8_t* sub = tile + (c0 >> 5) * sub_stride + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      *reinterpret_cast<float4*>(sub + ((c ^ (row & 7)) << 4)) = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
  } else {
    uint8_t* sub = tile + (c0 >> 6) * sub_stride + (row >> 3) * 1024 + (row & 7) * 128;
    const int cbase = (c0 & 63) >> 3;  // first 16-byte chunk of this 32-column group inside the 128-byte row
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t pk[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(x[8 * c + 2 * i], x[8 * c + 2 * i + 1]);
        pk[i] = *reinterpret_cast<uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(sub + (((cbase + c) ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

// `flush(j)` is called by every thread of the calling warps right after sub-tile j (64 bf16 / 32 fp32 columns) of their
// rows has been staged: the kernel uses it to hand finished sub-tiles to the TMA store while the next ones are computed.
template <class Flush>
__device__ __forceinline__ void epilogue_row_staged(const mmb_mlp_layer_params& p, uint32_t taddr, int row, int n0, int n_tile,
                                                    int cb, int ce, uint8_t* tile, Flush flush, int sub_stride = BM * 128, int m = 0) {
  // m = global row (only read with the deferred LayerNorm fields of p)
  const int sub_cols = (p.epilogue == 0) ? 32 : 64;
  float v[32], bv[32], x[32];
  if (p.epilogue == 2) {  // bias + ELU + LayerNorm over the whole row (n_tile == N): two passes over TMEM
    float sum = 0.0f, sumsq = 0.0f;
    for (int c0 = 0; c0 < n_tile; c0 += 32) {
      tmem_ld32(taddr + c0, v);
      load_bias32(p.bias + n0 + c0, bv);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float e = elu1(v[i] + bv[i]);
        sum += e;
        sumsq += e * e;
      }
    }
    const float mean = sum / (float)p.N;
    const float var = fmaxf(sumsq / (float)p.N - mean * mean, 0.0f);  // biased variance, as nn.LayerNorm
    const float rstd = rsqrtf(var + p.ln_eps);
    for (int c0 = 0; c0 < n_tile; c0 += 32) {
      tmem_ld32(taddr + c0, v);
      load_bias32(p.bias + n0 + c0, bv);
      float g[32], b[32];
      load_bias32(p.ln_gamma + n0 + c0, g);
      load_bias32(p.ln_beta + n0 + c0, b);
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = (elu1(v[i] + bv[i]) - mean) * rstd * g[i] + b[i];
      stage_out32<false>(tile, row, c0, x, sub_stride);
      if ((c0 + 32) % sub_cols == 0 || c0 + 32 >= n_tile) flush(c0 / sub_cols);
    }
    return;
  }
  float mean = 0.0f, rstd = 1.0f, psum = 0.0f, psq = 0.0f;
  const bool corr = p.ln_in_stats != nullptr;
  if (corr) ln_row_moments(p, m, mean, rstd);
  for (int c0 = cb; c0 < ce; c0 += 32) {
    tmem_ld32(taddr + c0, v);
    const int n = n0 + c0;
    if (n + 32 <= p.N) {
      load_bias32(p.bias + n, bv);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) bv[i] = (n + i < p.N) ? __ldg(p.bias + n + i) : 0.0f;
    }
    if (corr) {               // LN(e) . W^T = rstd * (e . (W gamma)^T - mean * c) (+ W beta, folded into the bias)
      if (n + 32 <= p.N) {
        load_bias32(p.ln_c + n, x);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = rstd * (v[i] - mean * x[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = (n + i < p.N) ? rstd * (v[i] - mean * __ldg(p.ln_c + n + i)) : 0.0f;
      }
    }
    if (p.epilogue == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = (n + i < p.N) ? elu1(v[i] + bv[i]) : 0.0f;
      if (p.ln_out_stats) {   // the sums run over the values as the next layer's GEMM will see them
#pragma unroll
        for (int i = 0; i < 32; ++i) { x[i] = __bfloat162float(__float2bfloat16_rn(x[i])); psum += x[i]; psq += x[i] * x[i]; }
        if (((c0 + 32) & 63) == 0) { if (m < p.M) ln_store_partial(p, m, n, psum, psq); psum = 0.0f; psq = 0.0f; }
      }
      stage_out32<false>(tile, row, c0, x, sub_stride);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = v[i] + bv[i];
      stage_out32<true>(tile, row, c0, x, sub_stride);
    }
    if ((c0 + 32) % sub_cols == 0 || c0 + 32 >= ce) flush(c0 / sub_cols);
  }
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src_smem)),
               "r"(c0), "r"(c1) : "memory");
}

template <bool PAIR>   // PAIR: the cta_group::2 variant; a kernel that contains cta_group::2 instructions can only be launched as a cluster
__device__ __forceinline__ void mlp_layer_ws_body(const mmb_mlp_layer_params& p, const CUtensorMap& map_x, const CUtensorMap& map_w,
                                                  const CUtensorMap& map_y) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], accum_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * p.n_tile;
  const int n_tile = p.n_tile;
  // PAIR: cta_group::2 - the CTA pair (cluster of 2 along M) runs one 256 x n_tile MMA; each CTA holds its
  // 128 rows of A and HALF of the weight tile, so 32 KB instead of 48 KB enter each SM per k-block (the measured bound)
  constexpr bool pair = PAIR;
  const int b_bytes = (pair ? n_tile / 2 : n_tile) * BK * 2;
  const int stage_bytes = A_STAGE_BYTES + b_bytes;   // [A tile | B tile], both multiples of 1024 B
  const int S = p.stages;
  const int nkb = p.Kpad / BK;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < n_tile) tmem_cols <<= 1;
  // thread-block cluster along M (launch attribute): the CM CTAs of a cluster share one weight tile per k-block - each
  // loads 1/CM of its rows and multicasts them to all - so the L2 -> SM operand traffic per CTA drops from A + B to
  // A + B/CM; a stage is free again only when the MMAs of ALL CTAs of the cluster have read it
  uint32_t cm, crank;
  asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(cm));
  asm volatile("mov.u32 %0, %%cluster_ctaid.x;" : "=r"(crank));
  const uint16_t cmask = (uint16_t)((1u << cm) - 1u);
  const int w_rows = n_tile / (int)cm;     // rows of the weight tile this CTA fetches (host guarantees divisibility by 8)

  if (tid == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], pair ? 1 : cm); }
    mbar_init(&accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
  }
  if (warp == 1) {  // the MMA warp owns the tensor memory
    if constexpr (pair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (cm > 1) cluster_sync_all();          // every CTA's barriers are initialised before a peer multicasts into them
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (p.overlap_prev) griddep_launch_dependents();  // the next layer may start its prologue and weight loads now

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      auto load_w = [&](int kb, int s) {
        uint8_t* st = smem + s * stage_bytes;
        if constexpr (pair) {  // this CTA's half of the weight tile into its own shared memory, bytes counted on the leader's barrier
          tma_load_2d_2sm(st + A_STAGE_BYTES, &map_w, kb * BK, n0 + (int)crank * (n_tile / 2), &full_bar[s]);
          return;
        }
        if (cm > 1) {
          tma_load_2d_mc(st + A_STAGE_BYTES + (int)crank * w_rows * 128, &map_w, kb * BK, n0 + (int)crank * w_rows, &full_bar[s], cmask);
          return;
        }
        tma_load_2d(st + A_STAGE_BYTES, &map_w, kb * BK, n0, &full_bar[s]);
        if (n_tile > 256) tma_load_2d(st + A_STAGE_BYTES + 256 * 128, &map_w, kb * BK, n0 + 256, &full_bar[s]);
      };
      // the first ring of weight tiles does not depend on the previous kernel: with overlap_prev it is in flight while
      // that kernel (the previous layer) is still running; the activation tiles wait for it
      const int pre = nkb < S ? nkb : S;
      for (int kb = 0; kb < pre; ++kb) {
        if (!pair) mbar_expect_tx(&full_bar[kb], (uint32_t)stage_bytes);
        else if (crank == 0) mbar_expect_tx(&full_bar[kb], 2u * (uint32_t)stage_bytes);   // both CTAs' tiles
        load_w(kb, kb);
      }
      if (p.overlap_prev) griddep_wait();
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S, u = kb / S;
        if (u > 0) {
          if (pair) mbar_wait_or_flag(&empty_bar[s], (uint32_t)((u - 1) & 1), 0x100u + (unsigned)kb);
          else mbar_wait(&empty_bar[s], (uint32_t)((u - 1) & 1));
          if (!pair) mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
          else if (crank == 0) mbar_expect_tx(&full_bar[s], 2u * (uint32_t)stage_bytes);
          load_w(kb, s);
        }
        if constexpr (pair) tma_load_2d_2sm(smem + s * stage_bytes, &map_x, kb * BK, m0, &full_bar[s]);
        else tma_load_2d(smem + s * stage_bytes, &map_x, kb * BK, m0, &full_bar[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if constexpr (pair) {
      if (crank == 0 && elect_one()) {       // only the leader issues; the MMA reads both CTAs' operand tiles
        const uint32_t idesc = umma_idesc_bf16(n_tile, 2 * BM);
        for (int kb = 0; kb < nkb; ++kb) {
          const int s = kb % S, u = kb / S;
          mbar_wait_or_flag(&full_bar[s], (uint32_t)(u & 1), 0x200u + (unsigned)kb);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * stage_bytes), b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
          for (int j = 0; j < BK / 16; ++j)
            umma_bf16_2sm(tmem, umma_desc_k_sw128(a_addr + j * 32), umma_desc_k_sw128(b_addr + j * 32), idesc, (kb > 0) || (j > 0));
          umma_commit_2sm(&empty_bar[s], 3);   // frees the stage in both CTAs
        }
        umma_commit_2sm(&accum_bar, 3);        // accumulators of both CTAs ready
      }
    } else if (elect_one()) {
      const int umma_n = n_tile > 256 ? 256 : n_tile;
      const uint32_t idesc = umma_idesc_bf16(umma_n);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S, u = kb / S;
        mbar_wait(&full_bar[s], (uint32_t)(u & 1));
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * stage_bytes), b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
        for (int j = 0; j < BK / 16; ++j) {  // UMMA_K = 16 bf16 = 32 bytes along the swizzled row
          const bool acc = (kb > 0) || (j > 0);
          umma_bf16(tmem, umma_desc_k_sw128(a_addr + j * 32), umma_desc_k_sw128(b_addr + j * 32), idesc, acc);
          if (n_tile > 256)
            umma_bf16(tmem + 256, umma_desc_k_sw128(a_addr + j * 32), umma_desc_k_sw128(b_addr + 256 * 128 + j * 32), idesc, acc);
        }
        if (cm > 1) umma_commit_mc(&empty_bar[s], cmask);   // ... on every CTA of the cluster: their producers write into this stage too
        else umma_commit(&empty_bar[s]);       // the stage may be refilled once these MMAs have read it
      }
      umma_commit(&accum_bar);                 // every MMA of the tile has completed: accumulator ready
    }
  } else {
    // ===== epilogue warps: TMEM lane quarter = warp % 4 =====
    if (pair) mbar_wait_or_flag(&accum_bar, 0, 0x300u + crank);
    else mbar_wait(&accum_bar, 0);
    tc_fence_after();
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    int cb = 0, ce = n_tile;
    if (p.epilogue != 2 && n_tile >= 64) {      // split the columns between the two warps of a lane quarter
      const int mid = ((n_tile / 32 + 1) / 2) * 32;
      cb = half ? mid : 0;
      ce = half ? n_tile : mid;
    } else if (half) {
      ce = 0;
    }
    // by now every MMA has completed, so every operand stage has been consumed: the ring is free to hold the output tile
    // staged output needs a TMA-addressable destination: 16-byte aligned base and row pitch; bf16 sub-tiles are 64 wide
    const bool staged = (p.epilogue == 0) ? (((p.y_stride & 3) | (reinterpret_cast<uintptr_t>(p.y) & 15u)) == 0)
                                          : (n_tile % 64 == 0);
    const int sub_cols = (p.epilogue == 0) ? 32 : 64;       // columns per 128-byte sub-tile row (fp32 / bf16)
    // pipelined stores: the four warps of a column half hand every finished sub-tile to the TMA store (their own named
    // barrier, one issuing lane) while they compute the next one - the 64 KB tile no longer leaves in one burst after the
    // math.  Needs the halves to own whole sub-tiles; otherwise one store phase after a barrier of all eight warps.
    const bool pipelined = staged && (p.epilogue == 2 || (cb % sub_cols == 0 && (ce - cb) % sub_cols == 0 && n_tile >= 2 * sub_cols));
    if (!staged) {
      if (cb < ce) epilogue_row(p, tmem + ((uint32_t)(q * 32) << 16), m0 + row, n0, n_tile, cb, ce);
    } else if (cb < ce) {
      if (pipelined) {
        epilogue_row_staged(p, tmem + ((uint32_t)(q * 32) << 16), row, n0, n_tile, cb, ce, smem, [&](int j) {
          fence_async_smem();
          if (half) asm volatile("bar.sync 3, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
          if (q == 0 && lane == 0) {
            tma_store_2d(&map_y, smem + j * (BM * 128), n0 + j * sub_cols, m0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }, BM * 128, m0 + row);
        if (q == 0 && lane == 0) tma_store_wait_read();     // the tile must stay intact until the stores have read it
      } else {
        epilogue_row_staged(p, tmem + ((uint32_t)(q * 32) << 16), row, n0, n_tile, cb, ce, smem, [](int) {}, BM * 128, m0 + row);
      }
    }
    if (!pipelined) {
      fence_async_smem();                                   // generic-proxy tile writes -> visible to the TMA store
      asm volatile("bar.sync 1, 256;" ::: "memory");        // the eight epilogue warps
      if (staged && tid == 64) {
        for (int j = 0; j * sub_cols < n_tile; ++j) tma_store_2d(&map_y, smem + j * (BM * 128), n0 + j * sub_cols, m0);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        tma_store_wait_read();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (cm > 1) cluster_sync_all();          // no CTA leaves while a peer's commit may still arrive on its barriers
  if (warp == 1) {
    if constexpr (pair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
  }
}

__global__ void __launch_bounds__(WS_THREADS, 1) mlp_layer_ws_kernel(const __grid_constant__ mmb_mlp_layer_params p,
                                                                     const __grid_constant__ CUtensorMap map_x,
                                                                     const __grid_constant__ CUtensorMap map_w,
                                                                     const __grid_constant__ CUtensorMap map_y) {
  mlp_layer_ws_body<false>(p, map_x, map_w, map_y);
}
__global__ void __launch_bounds__(WS_THREADS, 1) mlp_layer_ws_pair_kernel(const __grid_constant__ mmb_mlp_layer_params p,
                                                                          const __grid_constant__ CUtensorMap map_x,
                                                                          const __grid_constant__ CUtensorMap map_w,
                                                                          const __grid_constant__ CUtensorMap map_y) {
  mlp_layer_ws_body<true>(p, map_x, map_w, map_y);
}

// Grouped launch: blockIdx.z selects one of up to MMB_MAX_GROUP independent problems of identical geometry (the ten
// per-agent networks of the MARL policies, actor_critic.py / runner.py:205-217): one launch per layer for the whole team.
// Parameters and tensor maps travel as one large __grid_constant__ kernel argument (8 KB; CUDA 12.1+ allows 32 KB).
struct WsGroupArgs {
  mmb_mlp_layer_params p[MMB_MAX_GROUP];
  CUtensorMap map_x[MMB_MAX_GROUP], map_w[MMB_MAX_GROUP], map_y[MMB_MAX_GROUP];
};
// Persistent variant for launches with more tiles than SMs (large batches): one CTA per SM walks tiles m-major, the operand
// ring runs on across tiles, and TWO accumulators in tensor memory (2 x n_tile <= 512 columns) let the epilogue of tile i
// (TMEM -> bias / ELU -> staged sub-tiles -> TMA stores) overlap the k-loop of tile i + 1.  Output staging is two 16 KB
// sub-tile buffers (one per column half) outside the ring.  n_tile <= 256, no LayerNorm epilogue, TMA-addressable output.
// `sel` hands out the parameters and tensor maps of problem a of `count` problems of identical geometry (one: the plain
// launch; several: the per-agent networks of a team): a CTA's tile sequence runs problem-major, then m-major, across all of them.
struct PersistOne {
  const mmb_mlp_layer_params& p0; const CUtensorMap &mx, &mw, &my;
  __device__ __forceinline__ const mmb_mlp_layer_params& p(int) const { return p0; }
  __device__ __forceinline__ const CUtensorMap* map_x(int) const { return &mx; }
  __device__ __forceinline__ const CUtensorMap* map_w(int) const { return &mw; }
  __device__ __forceinline__ const CUtensorMap* map_y(int) const { return &my; }
};
struct PersistGroup {
  const WsGroupArgs& g;
  __device__ __forceinline__ const mmb_mlp_layer_params& p(int a) const { return g.p[a]; }
  __device__ __forceinline__ const CUtensorMap* map_x(int a) const { return &g.map_x[a]; }
  __device__ __forceinline__ const CUtensorMap* map_w(int a) const { return &g.map_w[a]; }
  __device__ __forceinline__ const CUtensorMap* map_y(int a) const { return &g.map_y[a]; }
};
// P = warps per TMEM lane quarter in the epilogue (2: 320 threads; 4: 576 threads, <= 113 registers - four warps on every
// scheduler hide the ex2 / tensor-memory / barrier latencies two could not: the epilogue, not the k-loop, bounds these launches)
__device__ __forceinline__ void bar_part(int part) { asm volatile("bar.sync %0, 128;" ::"r"(2 + part) : "memory"); }
template <class Sel, int P>
__device__ __forceinline__ void mlp_layer_ws_persist_body(const Sel& sel, const int count) {
  const mmb_mlp_layer_params& p = sel.p(0);           // the geometry (identical for every problem)
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_s[256], c_s[256];   // hidden-layer fast path: this tile's bias / c rows
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tile = p.n_tile;
  const int stage_bytes = A_STAGE_BYTES + n_tile * BK * 2;
  const int S = p.stages;
  const int nkb = p.Kpad / BK;
  const int tiles_n = p.Npad / n_tile, tiles_m = p.Mpad / BM, per_problem = tiles_m * tiles_n, num_tiles = count * per_problem;
  uint8_t* out_buf = smem + S * stage_bytes;          // 2 x 16 KB, 1024-byte aligned (stage_bytes is a multiple of 1024)

  if (tid == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4 * P); }   // every epilogue warp releases an accumulator
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int a = 0; a < count; ++a) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(sel.map_x(a)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(sel.map_w(a)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(sel.map_y(a)) : "memory");
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (p.overlap_prev) griddep_launch_dependents();

  if (warp == 0) {
    // ===== TMA producer: the ring position runs on across tiles =====
    if (elect_one()) {
      if (p.overlap_prev) griddep_wait();
      int s = 0;
      uint32_t u = 0;                          // (no division in these single-thread loops: tools/probe/tma_mma_rate.cu)
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int a = tile / per_problem, rem = tile - a * per_problem;
        const int m0 = (rem / tiles_n) * BM, n0 = (rem % tiles_n) * n_tile;
        const CUtensorMap* mx = sel.map_x(a);
        const CUtensorMap* mw = sel.map_w(a);
        for (int kb = 0; kb < nkb; ++kb) {
          if (u > 0) mbar_wait(&empty_bar[s], (u - 1) & 1u);
          uint8_t* st = smem + s * stage_bytes;
          mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
          tma_load_2d(st, mx, kb * BK, m0, &full_bar[s]);
          tma_load_2d(st + A_STAGE_BYTES, mw, kb * BK, n0, &full_bar[s]);
          if (++s == S) { s = 0; ++u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: accumulator i & 1 =====
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(n_tile);
      const uint64_t da0 = umma_desc_k_sw128(smem_u32(smem)), db0 = umma_desc_k_sw128(smem_u32(smem) + A_STAGE_BYTES);
      const uint32_t dstep = (uint32_t)stage_bytes >> 4;
      int s = 0, i = 0;
      uint32_t u = 0;
      uint64_t da = da0, db = db0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++i) {
        const int acc = i & 1, use = i >> 1;
        if (use > 0) mbar_wait(&acc_empty[acc], (uint32_t)((use - 1) & 1));   // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem + (uint32_t)(acc * n_tile);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[s], u & 1u);
          if (kb == 0) umma_bf16(tacc, da, db, idesc, false);
          else umma_bf16_acc(tacc, da, db, idesc);
#pragma unroll
          for (int j = 1; j < BK / 16; ++j) umma_bf16_acc(tacc, da + (uint64_t)(2 * j), db + (uint64_t)(2 * j), idesc);
          umma_commit(&empty_bar[s]);
          if (++s == S) { s = 0; ++u; da = da0; db = db0; }
          else { da += dstep; db += dstep; }
        }
        umma_commit(&acc_full[acc]);
      }
    }
  } else {
    // ===== epilogue warps =====
    const int q = warp & 3, part = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int sub_cols = (p.epilogue == 0) ? 32 : 64;
    // the P warps of a lane quarter split the columns into whole sub-tiles; a tile too narrow for that (an 8- or 1-wide head)
    // goes to the first of them
    const bool split = (n_tile / P) % sub_cols == 0 && n_tile >= P * sub_cols;
    const int share = n_tile / P;
    const int cb = split ? part * share : 0, ce = split ? cb + share : (part ? 0 : n_tile);
    uint8_t* buf = out_buf + part * (BM * 128);
    int i = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++i) {
      const int a = tile / per_problem, rem = tile - a * per_problem;
      const int m0 = (rem / tiles_n) * BM, n0 = (rem % tiles_n) * n_tile;
      const mmb_mlp_layer_params& pa = sel.p(a);
      const CUtensorMap* my = sel.map_y(a);
      const int acc = i & 1, use = i >> 1;
      const bool fast = pa.epilogue == 1 && (ce - cb) % 64 == 0 && n0 + n_tile <= pa.N;   // (the same for all eight warps: they meet at barrier 1)
      const int m = m0 + row;
      const bool corr = pa.ln_in_stats != nullptr;
      float mean = 0.0f, rstd = 1.0f;
      if (fast) {
        // everything that does not need the accumulator happens BEFORE it is waited for: this tile's bias / c rows into shared
        // memory (one element per thread), the row's LayerNorm moments from the producing layer's partials (L2 reads)
        asm volatile("bar.sync 1, %0;" ::"n"(P * 128) : "memory");      // the previous tile's readers of bias_s / c_s are done
        { const int e = tid - 64; if (e < n_tile) { bias_s[e] = __ldg(pa.bias + n0 + e); c_s[e] = corr ? __ldg(pa.ln_c + n0 + e) : 0.0f; } }
        if (corr) ln_row_moments(pa, m, mean, rstd);
        asm volatile("bar.sync 1, %0;" ::"n"(P * 128) : "memory");
      }
      mbar_wait(&acc_full[acc], (uint32_t)(use & 1));
      tc_fence_after();
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * n_tile);
      if (fast) {
        // Hidden layers, the hot case (the generic path below took ~12 us per 128 x 256 tile and bounded the whole launch):
        // tensor-memory loads one 32-column chunk ahead, bias / c rows in shared memory, the staging buffer's previous store
        // only waited for when the buffer is written again.
        const bool stats = pa.ln_out_stats != nullptr && m < pa.M;
        uint8_t* rowp = buf + (row >> 3) * 1024 + (row & 7) * 128;
        const bool leader = q == 0 && lane == 0;
        uint32_t ra[32], rb[32], pk[32];
        float psum = 0.0f, psq = 0.0f;
        auto chunk = [&](const uint32_t (&r)[32], int col, uint32_t* out16) {   // 32 columns from tile column col: -> 16 packed bf16 pairs
          const float4* b4 = reinterpret_cast<const float4*>(bias_s + col);
          const float4* c4 = reinterpret_cast<const float4*>(c_s + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b = b4[i];
            float t0 = __uint_as_float(r[4 * i]), t1 = __uint_as_float(r[4 * i + 1]), t2 = __uint_as_float(r[4 * i + 2]), t3 = __uint_as_float(r[4 * i + 3]);
            if (corr) {
              const float4 c = c4[i];
              t0 = rstd * (t0 - mean * c.x); t1 = rstd * (t1 - mean * c.y); t2 = rstd * (t2 - mean * c.z); t3 = rstd * (t3 - mean * c.w);
            }
            const __nv_bfloat162 h0 = __floats2bfloat162_rn(elu1(t0 + b.x), elu1(t1 + b.y)), h1 = __floats2bfloat162_rn(elu1(t2 + b.z), elu1(t3 + b.w));
            out16[2 * i] = *reinterpret_cast<const uint32_t*>(&h0);
            out16[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&h1);
            if (stats) {
              const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
              psum += f0.x; psq += f0.x * f0.x; psum += f0.y; psq += f0.y * f0.y;
              psum += f1.x; psq += f1.x * f1.x; psum += f1.y; psq += f1.y * f1.y;
            }
          }
        };
        tmem_ld32_issue(taddr + cb, ra);
        for (int c0 = cb; c0 < ce; c0 += 64) {
          psum = 0.0f; psq = 0.0f;
          if constexpr (P <= 2) {
            tmem_ld_wait(ra);
            tmem_ld32_issue(taddr + c0 + 32, rb);
            chunk(ra, c0, pk);
            tmem_ld_wait(rb);
            if (c0 + 64 < ce) tmem_ld32_issue(taddr + c0 + 64, ra);
            chunk(rb, c0 + 32, pk + 16);
          } else {                  // four warps per scheduler: no second register buffer (113 registers), the other warps cover
            tmem_ld_wait(ra);
            chunk(ra, c0, pk);
            tmem_ld32_issue(taddr + c0 + 32, ra);
            tmem_ld_wait(ra);
            chunk(ra, c0 + 32, pk + 16);
            if (c0 + 64 < ce) tmem_ld32_issue(taddr + c0 + 64, ra);
          }
          if (stats) ln_store_partial(pa, m, n0 + c0, psum, psq);
          if (leader) tma_store_wait_read();                // the buffer fed the previous sub-tile's (or tile's) store
          bar_part(part);
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(rowp + ((c ^ (row & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          fence_async_smem();
          bar_part(part);
          if (leader) {
            tma_store_2d(my, buf, n0 + c0, m0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      } else
      // every finished sub-tile goes out through this half's staging buffer: staged -> barrier -> one lane stores and waits
      // until the TMA has read the buffer -> barrier -> the buffer is free for the next sub-tile
      epilogue_row_staged(pa, taddr, row, n0, n_tile, cb, ce, buf - (cb / sub_cols) * 0, [&](int j) {
        fence_async_smem();
        bar_part(part);
        if (q == 0 && lane == 0) {
          tma_store_2d(my, buf, n0 + j * sub_cols, m0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          tma_store_wait_read();
        }
        bar_part(part);
      }, 0, m0 + row);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[acc])) : "memory");
    }
    if (q == 0 && lane == 0) tma_store_wait_read();       // the fast path leaves its last store in flight: the buffer must outlive it
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
__global__ void __launch_bounds__(WS_THREADS, 1) mlp_layer_ws_persist_kernel(const __grid_constant__ mmb_mlp_layer_params p,
                                                                             const __grid_constant__ CUtensorMap map_x,
                                                                             const __grid_constant__ CUtensorMap map_w,
                                                                             const __grid_constant__ CUtensorMap map_y) {
  mlp_layer_ws_persist_body<PersistOne, 2>(PersistOne{p, map_x, map_w, map_y}, 1);
}
// the team's networks as ONE persistent launch per layer: 10 agents x 32 row blocks x 2 column tiles = 640 tiles walked by one
// CTA per SM, each tile's epilogue under the next tile's k-loop (the non-persistent grouped launch paid prologue, pipeline fill
// and an exposed epilogue per tile: 4.3 waves of ~10 us for 2.6 us of k-loop each)
// (four parts - sixteen epilogue warps, 96 registers - were measured: the hidden layers 40.9 -> 46.8 us, layer 0 unchanged at
// 27 us, the head 18.3 -> 25.5 us: the epilogue's 6.3 us per tile is not a shortage of warps to switch between)
constexpr int PERSIST_GROUP_PARTS = 2, PERSIST_GROUP_THREADS = 64 + PERSIST_GROUP_PARTS * 128;
__global__ void __launch_bounds__(PERSIST_GROUP_THREADS, 1) mlp_layer_ws_persist_group_kernel(const __grid_constant__ WsGroupArgs g, const int count) {
  mlp_layer_ws_persist_body<PersistGroup, PERSIST_GROUP_PARTS>(PersistGroup{g}, count);
}

__global__ void __launch_bounds__(WS_THREADS, 1) mlp_layer_ws_group_kernel(const __grid_constant__ WsGroupArgs g) {
  const int a = blockIdx.z;
  mlp_layer_ws_body<false>(g.p[a], g.map_x[a], g.map_w[a], g.map_y[a]);
}

struct LnCastGroupArgs {
  const float* x[MMB_MAX_GROUP];
  const float* gamma[MMB_MAX_GROUP];
  const float* beta[MMB_MAX_GROUP];
  __nv_bfloat16* y[MMB_MAX_GROUP];
};


// ------------------------------------------------------------------------------------------------------
// Whole-MLP forward in ONE launch (PPO ActorCritic.actor / .critic: Linear-ELU chains, module.py:25-55).
// A cluster of 4 CTAs owns one 128-row block for ALL layers; CTA r of the cluster computes columns [r * N_l / 4, (r + 1) *
// N_l / 4) of every layer l.  Layer l + 1 of a row block needs exactly the four column slices its own cluster produced, so
// there is no kernel boundary between the layers and no barrier that every thread waits at either: activations travel
// through L2 (TMA store -> cp.async.bulk.wait_group 0), then the storing thread of each epilogue half arrives on two
// single-use mbarriers - `own_bar[l]` of its CTA and `peer_bar[l]` of all four CTAs (release.cluster) - and only the TMA
// producer thread waits: for own_bar before it requests the k-blocks this CTA produced itself (every CTA starts the next
// layer with its own slices), for peer_bar before the peers' k-blocks.  The operand ring and its mbarrier phases run on
// across the layers, the next layer's first weight tiles are in flight while this layer's epilogue runs (the output is
// staged outside the ring), and launch prologue / pipeline drain are paid once per forward instead of once per layer.
// The epilogue of a hidden layer is the serial part (the next layer's MMAs need all of it): bias slice staged in shared
// memory beforehand, tensor-memory loads one 32-column chunk ahead, store waits deferred to the buffer's next use.
// Same warp roles as mlp_layer_ws_kernel (TMA producer, MMA issuer, 8 epilogue warps), one accumulator in tensor memory.
// grid = (4 x row blocks, 1, networks): blockIdx.z selects one of up to two networks of identical geometry (actor and
// critic side by side).  Geometry: hidden widths multiples of 256, <= 1024; TMA-addressable fp32 output.
// Timeline of a launch (tools/probe/chain_trace.py, build with EXTRA=-DMMB_CHAIN_TRACE_BUILD): profiles/r02_chain_trace.json.
// ------------------------------------------------------------------------------------------------------
constexpr int CHAIN_CLUSTER = 4;
constexpr int CHAIN_STAGES = 4;
constexpr int CHAIN_STAGE_BYTES = A_STAGE_BYTES + 256 * BK * 2;   // 48 KB: the widest slice (256 columns)
constexpr int CHAIN_SMEM = CHAIN_STAGES * CHAIN_STAGE_BYTES + 2 * BM * 128;   // ring + two 16 KB output staging buffers = 224 KB
constexpr int CHAIN_MAX_NETS = 2;
constexpr int DUO_MAX_STAGES = 8;   // mlp_chain_duo_kernel: stages of the narrowest slices (20 KB each)
constexpr int DUO_A_WARP = 10;      // its second TMA producer (activation tiles)
constexpr int DUO_THREADS = WS_THREADS + 32;
struct ChainLayer {
  CUtensorMap map_x, map_w, map_y;
  const float* bias;
  int N, nkb, n_tile, epilogue;
  int kgroup;       // mlp_chain_duo_kernel: k-blocks per tensor load (map_x / map_w are then 3-D maps, see make_map_bf16_kgroup)
  int pad_[9];      // keeps the next element's tensor maps 64-byte aligned
};
static_assert(sizeof(ChainLayer) % 64 == 0, "tensor maps need 64-byte alignment");
struct ChainArgs {
  ChainLayer l[CHAIN_MAX_NETS][MMB_MLP_MAX_LAYERS];
  const float* x32[CHAIN_MAX_NETS];   // != NULL: layer 0's A operand is cast from this fp32 [M][K0] matrix inside the kernel
  int num_layers, M, overlap_prev, K0;
  unsigned long long* trace;          // diagnostics (MMB_CHAIN_TRACE=1): %globaltimer stamps, see mmb_mlp_chain_trace; else NULL
};
enum { TR_LAYER_BEGIN = 0, TR_FIRST_OPERANDS = 1, TR_MMA_ISSUED = 2, TR_ACC_COMPLETE = 3, TR_COMPUTED = 4, TR_LANDED = 5, TR_BOUNDARY = 6 };
[[maybe_unused]] constexpr int CHAIN_TRACE_EVENTS = 8 * MMB_MLP_MAX_LAYERS * 8;      // then [8 CTAs][6 layers][16]: operands of the i-th k-block in smem
constexpr int CHAIN_TRACE_WORDS = 8 * MMB_MLP_MAX_LAYERS * (8 + 16);
// Compiled in only with -DMMB_CHAIN_TRACE_BUILD (make EXTRA=-DMMB_CHAIN_TRACE_BUILD; tools/probe/chain_trace.py): even
// predicated-off, the stamps cost the kernel ~4 us of scheduling freedom.
__device__ __forceinline__ void chain_trace(const ChainArgs& g, int l, int ev, int kb = -1) {
#ifdef MMB_CHAIN_TRACE_BUILD
  if (g.trace != nullptr && blockIdx.x < 8 && blockIdx.z == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (kb < 0) g.trace[(blockIdx.x * MMB_MLP_MAX_LAYERS + l) * 8 + ev] = t;
    else if (kb < 16) g.trace[CHAIN_TRACE_EVENTS + (blockIdx.x * MMB_MLP_MAX_LAYERS + l) * 16 + kb] = t;
  }
#endif
}

__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta_rank) {   // same barrier, CTA cta_rank of the cluster
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta_rank));
  // relaxed: what the peer is told about are TMA stores this warp has seen complete (async proxy, already in L2) - a
  // release at cluster scope here costs ~0.8 us per arrival (measured)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// The i-th k-block layer l > 0 consumes in CTA `crank`: its own n_own output sub-tiles of layer l - 1 first, then the peers'.
__device__ __forceinline__ int chain_kb(int i, int n_own, int crank, int nkb) {
  const int kb = i + crank * n_own;
  return kb >= nkb ? kb - nkb : kb;
}

// TF32: operands stay fp32 end to end (the observations, the nn.Linear weights as they are, fp32 hidden activations):
// 32-element k-blocks and 32-column sub-tiles instead of 64, kind::tf32 MMAs, no input cast and no kernel-side weight copies.
template <bool TF32>
__device__ __forceinline__ void mlp_chain_body(const ChainArgs& g) {
  constexpr int KBE = TF32 ? 32 : BK;                // elements of K per k-block (one 128-byte swizzled row)
  constexpr int SUB_SHIFT = TF32 ? 5 : 6;            // log2(columns of a hidden sub-tile = elements of the next layer's k-block)
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[CHAIN_STAGES], empty_bar[CHAIN_STAGES], a_bar[CHAIN_STAGES], acc_bar;
  __shared__ uint64_t own_bar[MMB_MLP_MAX_LAYERS];   // single use: this CTA's own slices of layer l have landed (one arrival per epilogue half)
  __shared__ uint64_t peer_bar[MMB_MLP_MAX_LAYERS];  // single use: ALL slices of layer l have landed (one arrival per epilogue half of each CTA)
  __shared__ __align__(16) float bias_s[256];       // this CTA's bias slice of the current hidden layer
  __shared__ uint32_t tmem_slot;
  constexpr int S = CHAIN_STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctaid.x;" : "=r"(crank));
  const int m0 = (blockIdx.x / CHAIN_CLUSTER) * BM;
  const ChainLayer* Ls = g.l[blockIdx.z];
  const int L = g.num_layers;
  // a programmatically launched successor (mmb_gaussian_act: everything behind its griddepcontrol.wait) may take the SM
  // slots this kernel leaves free and be resident when the last tile retires
  griddep_launch_dependents();
  uint8_t* out_buf = smem + S * CHAIN_STAGE_BYTES;
  // Input cast folded into layer 0: the eight epilogue warps - idle until the first accumulator is complete - convert the
  // fp32 rows of this row block into the bf16 SWIZZLE_128B operand tiles themselves (one arrival per warp on a_bar), instead
  // of a separate cast kernel whose completion the whole forward had to wait for.
  const float* x32 = g.x32[blockIdx.z];

  if (tid == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); mbar_init(&a_bar[i], 8); }
    mbar_init(&acc_bar, 1);
    for (int l = 0; l < L; ++l) {
      const uint32_t halves = (Ls[l].n_tile >> SUB_SHIFT) >= 2 ? 2 : 1;
      mbar_init(&own_bar[l], halves);
      mbar_init(&peer_bar[l], halves * CHAIN_CLUSTER);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int l = 0; l < L; ++l) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&Ls[l].map_x) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&Ls[l].map_w) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&Ls[l].map_y) : "memory");
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // every CTA's barriers exist before a peer arrives on them
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  int it_p = 0, it_m = 0;          // ring positions of the producer / the MMA issuer: they run on across the layers
  for (int l = 0; l < L; ++l) {
    const ChainLayer& C = Ls[l];
    const int nkb = C.nkb, n_tile = C.n_tile, n0 = (int)crank * n_tile;
    if (warp == 0) {
      // ===== TMA producer =====
      if (elect_one()) {
        chain_trace(g, l, TR_LAYER_BEGIN);
        auto load_w = [&](const ChainLayer& D, int kb, int itx, bool a_by_tma = true) {   // claims ring slot itx for k-block kb of layer D: weight slice now
          const int s = itx % S, u = itx / S;
          if (u > 0) mbar_wait(&empty_bar[s], (uint32_t)((u - 1) & 1));
          mbar_expect_tx(&full_bar[s], (uint32_t)((a_by_tma ? A_STAGE_BYTES : 0) + D.n_tile * BK * 2));
          tma_load_2d(smem + s * CHAIN_STAGE_BYTES + A_STAGE_BYTES, &D.map_w, kb * KBE, (int)crank * D.n_tile, &full_bar[s]);
        };
        const int pre = nkb < S ? nkb : S;   // k-blocks whose weight slices were requested ahead (before the data they multiply existed)
        const bool a_tma = !(l == 0 && x32 != nullptr);
        // Layers l > 0 take their k-blocks starting with the ones this CTA produced itself (chain_kb): they are requested as
        // soon as this CTA's own stores have landed (own_bar), the peers' slices once every epilogue half of the cluster has
        // reported (peer_bar: remote arrivals, release.cluster).  No other warp waits at a layer boundary.  Stores and
        // loads are both the async proxy and meet in L2 (a fence.proxy.async here would drain the loads in flight: 0.8 us).
        const int n_own = l == 0 ? 0 : (Ls[l - 1].n_tile >> SUB_SHIFT);
        if (l == 0) {
          for (int kb = 0; kb < pre; ++kb) load_w(C, kb, it_p + kb, a_tma);
          if (g.overlap_prev && a_tma) griddep_wait();        // the input cast is the previous kernel in the stream
        } else {
          mbar_wait(&own_bar[l - 1], 0u);
        }
        for (int i = 0; i < nkb; ++i) {
          const int itx = it_p + i, kb = chain_kb(i, n_own, (int)crank, nkb);
          if (i >= pre) load_w(C, kb, itx, a_tma);
          if (l > 0 && i == n_own) mbar_wait(&peer_bar[l - 1], 0u);
          if (a_tma) tma_load_2d(smem + (itx % S) * CHAIN_STAGE_BYTES, &C.map_x, kb * KBE, m0, &full_bar[itx % S]);
        }
        if (l + 1 < L) {          // the next layer's first weight slices: in flight while this layer computes and stores
          const ChainLayer& D = Ls[l + 1];
          const int pre1 = D.nkb < S ? D.nkb : S;
          for (int i = 0; i < pre1; ++i) load_w(D, chain_kb(i, n_tile >> SUB_SHIFT, (int)crank, D.nkb), it_p + nkb + i);
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ===== MMA issuer =====
      if (elect_one()) {
        // This thread's instruction latencies are the k-loop's clock (tools/probe/tma_mma_rate.cu: the generic form of this
        // loop takes 345 ns per k-block whatever the tile width, a lean one 226-315): stage descriptors advance by adding to
        // their address field, the accumulate predicate is a constant after the first instruction, and the tcgen05 fence is
        // only issued where generic-proxy writers (the input cast) or the epilogue's accumulator loads must be ordered.
        const uint32_t idesc = TF32 ? umma_idesc_tf32(n_tile) : umma_idesc_bf16(n_tile);
        const bool cast_a = l == 0 && x32 != nullptr;
        const uint64_t da0 = umma_desc_k_sw128(smem_u32(smem)), db0 = umma_desc_k_sw128(smem_u32(smem) + A_STAGE_BYTES);
        constexpr uint32_t dstep = (uint32_t)CHAIN_STAGE_BYTES >> 4;
        int s = it_m % S;
        uint32_t u = (uint32_t)(it_m / S);
        uint64_t da = da0 + (uint64_t)((uint32_t)s * dstep), db = db0 + (uint64_t)((uint32_t)s * dstep);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[s], u & 1u);
          if (cast_a) mbar_wait(&a_bar[s], u & 1u);   // the A tile the epilogue warps converted
          if (kb == 0) chain_trace(g, l, TR_FIRST_OPERANDS);
          chain_trace(g, l, 0, kb);
          if (cast_a || kb == 0) tc_fence_after();
          if (kb == 0) {
            if constexpr (TF32) umma_tf32(tmem, da, db, idesc, false);
            else umma_bf16(tmem, da, db, idesc, false);
          } else {
            if constexpr (TF32) umma_tf32(tmem, da, db, idesc, true);
            else umma_bf16_acc(tmem, da, db, idesc);
          }
#pragma unroll
          for (int j = 1; j < BK / 16; ++j)
            if constexpr (TF32) umma_tf32(tmem, da + (uint64_t)(2 * j), db + (uint64_t)(2 * j), idesc, true);
            else umma_bf16_acc(tmem, da + (uint64_t)(2 * j), db + (uint64_t)(2 * j), idesc);
          umma_commit(&empty_bar[s]);
          if (++s == S) { s = 0; ++u; da = da0; db = db0; }
          else { da += dstep; db += dstep; }
        }
        umma_commit(&acc_bar);
        chain_trace(g, l, TR_MMA_ISSUED);
      }
      __syncwarp();
    } else {
      // ===== epilogue warps: TMEM lane quarter = warp % 4, two warps per quarter split the columns =====
      if (l == 0 && x32 != nullptr) {
        // layer 0's A tiles from the fp32 input.  A warp owns 16 rows of the 128 x 64 tile of every k-block; one load
        // instruction covers two whole 256-byte row segments (coalesced), and the loads of k-block kb + 1 are in flight
        // while k-block kb is converted and handed to the tensor core.
        if (g.overlap_prev) griddep_wait();                   // the input may come from the previous kernel in the stream
        const int w8 = (tid - 64) >> 5, sub = lane >> 4, c4 = lane & 15;
        const int K0 = g.K0;
        const float* xbase = x32 + (int64_t)(m0 + w8 * 16 + sub) * K0 + c4 * 4;
        const int rows_left = g.M - (m0 + w8 * 16 + sub);     // row i of this thread exists iff 2 i < rows_left
        auto load_tile = [&](int kb, float4 (&v)[8]) {
          const bool col_ok = kb * BK + c4 * 4 < K0;          // K0 % 4 == 0 (host): the float4 is inside the row or outside it
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col_ok && 2 * i < rows_left) v[i] = __ldg(reinterpret_cast<const float4*>(xbase + (int64_t)(2 * i) * K0 + kb * BK));
          }
        };
        // three k-blocks of this thread's rows in registers: 2 x 32 KB per CTA in flight hide the ~1 us the rows take
        // to arrive (one k-block ahead left the tensor core waiting 0.5 us per k-block)
        float4 t0[8], t1[8], t2[8];
        load_tile(0, t0);
        if (nkb > 1) load_tile(1, t1);
        auto convert = [&](int kb, const float4 (&cur)[8]) {
          const int s = kb % S, u = kb / S;
          if (u > 0) mbar_wait(&empty_bar[s], (uint32_t)((u - 1) & 1));
          uint8_t* tile = smem + s * CHAIN_STAGE_BYTES;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = w8 * 16 + 2 * i + sub;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(cur[i].x, cur[i].y), h1 = __floats2bfloat162_rn(cur[i].z, cur[i].w);
            *reinterpret_cast<uint2*>(tile + (r >> 3) * 1024 + (r & 7) * 128 + ((((c4 >> 1) ^ (r & 7)) << 4) | ((c4 & 1) << 3))) =
                make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
          }
          fence_async_smem();                                  // generic-proxy tile writes -> visible to the tensor core's reads
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&a_bar[s])) : "memory");
        };
        for (int kb = 0; kb < nkb; kb += 3) {                   // the three buffers rotate: no register copies
          if (kb + 2 < nkb) load_tile(kb + 2, t2);
          convert(kb, t0);
          if (kb + 1 >= nkb) break;
          if (kb + 3 < nkb) load_tile(kb + 3, t0);
          convert(kb + 1, t1);
          if (kb + 2 >= nkb) break;
          if (kb + 4 < nkb) load_tile(kb + 4, t1);
          convert(kb + 2, t2);
        }
      }
      const int q = warp & 3, half = (warp - 2) >> 2;
      const int row = q * 32 + lane;
      if (l + 1 < L) {
        // ---- hidden layer: bias + ELU -> bf16, 64-column sub-tiles through this half's staging buffer ----
        // The bias slice waits in shared memory before the accumulator is complete, the next 32 accumulator columns travel
        // from tensor memory while the current ones are processed, and a sub-tile's store is only waited for when its
        // buffer is needed again: the epilogue is the serial part of a layer (the MMAs of layer l + 1 need all of it).
        asm volatile("bar.sync 1, 256;" ::: "memory");        // the other half may still be reading the previous layer's slice
        { const int e = tid - 64; if (e < n_tile) bias_s[e] = __ldg(C.bias + n0 + e); }      // N % 256 == 0 (host): all columns exist
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mbar_wait(&acc_bar, (uint32_t)(l & 1));
        tc_fence_after();
        if (tid == 64) chain_trace(g, l, TR_ACC_COMPLETE);
        const int n_sub = n_tile >> SUB_SHIFT, halves = n_sub >= 2 ? 2 : 1, per_half = n_sub / halves;
        const bool leader = q == 0 && lane == 0;
        if constexpr (TF32) {
          if (half < halves) {     // fp32 sub-tiles of 32 columns = one tensor-memory chunk each
            const int cbeg = half * per_half * 32;
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)cbeg;
            uint8_t* buf = out_buf + half * (BM * 128);
            uint8_t* rowp = buf + (row >> 3) * 1024 + (row & 7) * 128;
            uint32_t ra[32], rb[32];
            auto sub_tile = [&](int js, uint32_t (&r)[32], uint32_t (&nxt)[32]) {
              tmem_ld_wait(r);
              if (js + 1 < per_half) tmem_ld32_issue(taddr + (js + 1) * 32, nxt);
              const float4* b4 = reinterpret_cast<const float4*>(bias_s + cbeg + js * 32);
              float4 x[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 b = b4[i];
                // rounded to tf32 here (round-to-nearest): the tensor core itself truncates its fp32 operands
                x[i] = make_float4(rn_tf32(elu1(__uint_as_float(r[4 * i]) + b.x)), rn_tf32(elu1(__uint_as_float(r[4 * i + 1]) + b.y)),
                                   rn_tf32(elu1(__uint_as_float(r[4 * i + 2]) + b.z)), rn_tf32(elu1(__uint_as_float(r[4 * i + 3]) + b.w)));
              }
              if (js > 0) {                                     // the buffer fed the previous sub-tile's store
                if (leader) tma_store_wait_read();
                if (half) asm volatile("bar.sync 3, 128;" ::: "memory");
                else asm volatile("bar.sync 2, 128;" ::: "memory");
              }
#pragma unroll
              for (int c = 0; c < 8; ++c) *reinterpret_cast<float4*>(rowp + ((c ^ (row & 7)) << 4)) = x[c];
              fence_async_smem();
              if (half) asm volatile("bar.sync 3, 128;" ::: "memory");
              else asm volatile("bar.sync 2, 128;" ::: "memory");
              if (leader) {
                tma_store_2d(&C.map_y, buf, n0 + cbeg + js * 32, m0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
            };
            tmem_ld32_issue(taddr, ra);
            for (int js = 0; js < per_half; js += 2) {
              sub_tile(js, ra, rb);
              if (js + 1 < per_half) sub_tile(js + 1, rb, ra);
            }
            if (q == 0) {
              if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
              __syncwarp();
              if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&own_bar[l])) : "memory");
              else if (lane <= CHAIN_CLUSTER) mbar_arrive_remote(&peer_bar[l], (uint32_t)(lane - 1));
            }
          }
        } else
        if (half < halves) {
          const int cbeg = half * per_half * 64;
          const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)cbeg;
          uint8_t* buf = out_buf + half * (BM * 128);
          uint8_t* rowp = buf + (row >> 3) * 1024 + (row & 7) * 128;
          uint32_t ra[32], rb[32], pk[32];
          auto act32 = [&](const uint32_t (&r)[32], int col, uint32_t* out16) {      // 32 columns: bias + ELU -> 16 packed bf16 pairs
            const float4* b4 = reinterpret_cast<const float4*>(bias_s + col);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = b4[i];
              __nv_bfloat162 h0 = __floats2bfloat162_rn(elu1(__uint_as_float(r[4 * i]) + b.x), elu1(__uint_as_float(r[4 * i + 1]) + b.y));
              __nv_bfloat162 h1 = __floats2bfloat162_rn(elu1(__uint_as_float(r[4 * i + 2]) + b.z), elu1(__uint_as_float(r[4 * i + 3]) + b.w));
              out16[2 * i] = *reinterpret_cast<uint32_t*>(&h0);
              out16[2 * i + 1] = *reinterpret_cast<uint32_t*>(&h1);
            }
          };
          tmem_ld32_issue(taddr, ra);
          for (int js = 0; js < per_half; ++js) {
            tmem_ld_wait(ra);
            tmem_ld32_issue(taddr + js * 64 + 32, rb);
            act32(ra, cbeg + js * 64, pk);
            tmem_ld_wait(rb);
            if (js + 1 < per_half) tmem_ld32_issue(taddr + js * 64 + 64, ra);
            act32(rb, cbeg + js * 64 + 32, pk + 16);
            if (js > 0) {                                       // the buffer fed the previous sub-tile's store
              if (leader) tma_store_wait_read();
              if (half) asm volatile("bar.sync 3, 128;" ::: "memory");
              else asm volatile("bar.sync 2, 128;" ::: "memory");
            }
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<uint4*>(rowp + ((c ^ (row & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            fence_async_smem();
            if (half) asm volatile("bar.sync 3, 128;" ::: "memory");
            else asm volatile("bar.sync 2, 128;" ::: "memory");
            if (leader) {
              tma_store_2d(&C.map_y, buf, n0 + cbeg + js * 64, m0);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
          if (leader && half == 0) chain_trace(g, l, TR_COMPUTED);
          // the slices must have LANDED (not only left shared memory) before the peers are told to read them
          if (q == 0) {           // the storing lane's warp: lane 0 reports to this CTA, lanes 1-4 to the four CTAs in parallel
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&own_bar[l])) : "memory");
            else if (lane <= CHAIN_CLUSTER) mbar_arrive_remote(&peer_bar[l], (uint32_t)(lane - 1));
          }
          if (leader && half == 0) chain_trace(g, l, TR_LANDED);
        }
        tc_fence_before();
        __syncwarp();
      } else {
      mbar_wait(&acc_bar, (uint32_t)(l & 1));
      tc_fence_after();
      if (tid == 64) chain_trace(g, l, TR_ACC_COMPLETE);
      const int sub_cols = (C.epilogue == 0) ? 32 : 64;
      int cb = 0, ce = half ? 0 : n_tile;
      if (n_tile >= 2 * sub_cols) { cb = half ? n_tile / 2 : 0; ce = half ? n_tile : n_tile / 2; }
      mmb_mlp_layer_params p = {};
      p.M = g.M; p.N = C.N; p.epilogue = C.epilogue; p.bias = C.bias;
      uint8_t* buf = out_buf + half * (BM * 128);
      if (cb < ce) {
        epilogue_row_staged(p, tmem + ((uint32_t)(q * 32) << 16), row, n0, n_tile, cb, ce, buf, [&](int j) {
          fence_async_smem();
          if (half) asm volatile("bar.sync 3, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
          if (q == 0 && lane == 0) {
            tma_store_2d(&C.map_y, buf, n0 + j * sub_cols, m0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            tma_store_wait_read();
          }
          if (half) asm volatile("bar.sync 3, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
        }, 0);
        // the slice must have LANDED (not only left shared memory) before the peers are told to read it
        if (q == 0 && lane == 0 && half == 0) chain_trace(g, l, TR_COMPUTED);   // all sub-tiles of this half computed, staged, stores issued
        if (q == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (q == 0 && lane == 0 && half == 0) chain_trace(g, l, TR_LANDED);
      }
      tc_fence_before();
      __syncwarp();
      }
    }
    it_p += nkb;
    it_m += nkb;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // no CTA leaves while a peer's arrival may still be on its way to the barriers here
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

__global__ void __launch_bounds__(WS_THREADS, 1) mlp_chain_kernel(const __grid_constant__ ChainArgs g) { mlp_chain_body<false>(g); }
__global__ void __launch_bounds__(WS_THREADS, 1) mlp_chain_tf32_kernel(const __grid_constant__ ChainArgs g) { mlp_chain_body<true>(g); }


// ------------------------------------------------------------------------------------------------------
// Two networks on the SAME rows in one CTA (PPO act(): actor and critic of module.py:25-55 read the same observations).
// Same cluster-of-4 / column-slice decomposition as mlp_chain_kernel, but a CTA walks BOTH networks, alternating
// between them layer by layer - task (net n, layer l) in the order (0,0) (1,0) (0,1) (1,1) ... - with one 256-column
// accumulator per network in tensor memory (512 columns = all of it).  The two chains are independent, so the serial part
// of one network's layer (accumulator -> bias / ELU -> bf16 -> L2 -> peers, ~4 us of the tensor core idling in
// mlp_chain_kernel) runs UNDER the other network's k-loop: from layer 1 on the tensor core only waits for operands.
// grid z = 1: 4 x row blocks CTAs do the work 8 x row blocks CTAs of the side-by-side launch did in two waves.
// Layer 0: the fp32 rows are cast ONCE per k-block and feed both networks - ring slot 2 kb holds [A(kb) | W0(kb)], slot
// 2 kb + 1 holds [unused | W1(kb)]: 80 KB instead of 96 KB enter the SM per k-block pair, and both accumulators complete
// together.  Requires x_fp32[0] == x_fp32[1] (cast folded in) and bf16 operands; anything else takes mlp_chain_kernel.
// Every mbarrier wait is bounded (50 ms, then the CTA stops waiting altogether and records a code for
// mmb_mlp_debug_status): a protocol error yields garbage, never a hung GPU.
// ------------------------------------------------------------------------------------------------------
// `bar_addr` = shared-window address of the mbarrier.  The common case - the phase has completed or completes within the
// hardware's own try_wait time slice - costs one instruction and a branch; only then the clock is read.
__device__ __noinline__ void duo_wait_slow(uint32_t bar_addr, uint32_t parity, unsigned code, volatile int* abort_flag) {
  const long long t0 = clock64();
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar_addr), "r"(parity) : "memory");
    if (!done) {
      if (*abort_flag) return;
      if (clock64() - t0 > 100000000ll) {   // ~50 ms
        *abort_flag = 1;
        if (atomicCAS(&g_pair_dbg[0], 0u, code) == 0u) { g_pair_dbg[1] = blockIdx.x; g_pair_dbg[2] = threadIdx.x; g_pair_dbg[3] = parity; }
        return;
      }
    }
  }
}
__device__ __forceinline__ void duo_wait_addr(uint32_t bar_addr, uint32_t parity, unsigned code, volatile int* abort_flag) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar_addr), "r"(parity) : "memory");
  if (!done) duo_wait_slow(bar_addr, parity, code, abort_flag);
}
__device__ __forceinline__ void duo_wait(uint64_t* bar, uint32_t parity, unsigned code, volatile int* abort_flag) {
  duo_wait_addr(smem_u32(bar), parity, code, abort_flag);
}
__device__ __forceinline__ void umma_commit_addr(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_addr(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_addr) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(map),
               "r"(c0), "r"(c1), "r"(bar_addr) : "memory");
}
__device__ __forceinline__ void tma_load_3d_addr(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar_addr) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst), "l"(map),
               "r"(c0), "r"(c1), "r"(c2), "r"(bar_addr) : "memory");
}

// diagnostics (-DMMB_CHAIN_TRACE_BUILD + MMB_CHAIN_TRACE=1, tools/probe/duo_trace.py): [4 CTAs][16 tasks][8 events]
enum { DT_PRODUCER = 0, DT_FIRST_OPERANDS = 1, DT_MMA_ISSUED = 2, DT_ACC_COMPLETE = 3, DT_COMPUTED = 4, DT_LANDED = 5, DT_PEERS = 6, DT_EPI_BEGIN = 7 };
__device__ __forceinline__ void duo_trace(const ChainArgs& g, int task, int ev) {
#ifdef MMB_CHAIN_TRACE_BUILD
  if (g.trace != nullptr && blockIdx.x < 4 && task < 16) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g.trace[(blockIdx.x * 16 + task) * 8 + ev] = t;
  }
#endif
}
__device__ __forceinline__ void duo_trace_kb(const ChainArgs& g, int task, int kb) {     // CTA 0: operands of k-block kb in shared memory
#ifdef MMB_CHAIN_TRACE_BUILD
  if (g.trace != nullptr && blockIdx.x == 0 && task < 32 && kb < 16) {   // tasks 16..31: the producer's requests
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g.trace[512 + task * 16 + kb] = t;
  }
#endif
}

__global__ void __launch_bounds__(DUO_THREADS, 1) mlp_chain_duo_kernel(const __grid_constant__ ChainArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int SMAX = DUO_MAX_STAGES;
  __shared__ uint64_t full_bar[SMAX], empty_bar[SMAX], a_bar[SMAX], acc_bar[2];
  __shared__ uint64_t own_bar[2][MMB_MLP_MAX_LAYERS], peer_bar[2][MMB_MLP_MAX_LAYERS];
  __shared__ __align__(16) float bias_s[256];
  __shared__ uint32_t tmem_slot;
  __shared__ int s_abort;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctaid.x;" : "=r"(crank));
  const int m0 = (blockIdx.x / CHAIN_CLUSTER) * BM;
  const int L = g.num_layers;
  uint8_t* out_buf = smem + CHAIN_STAGES * CHAIN_STAGE_BYTES;
  const float* x32 = g.x32[0];
  volatile int* ab = &s_abort;
  griddep_launch_dependents();   // see mlp_chain_body

  if (tid == 0) {
    s_abort = 0;
    for (int i = 0; i < SMAX; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); mbar_init(&a_bar[i], 8); }
    mbar_init(&acc_bar[0], 1); mbar_init(&acc_bar[1], 1);
    for (int n = 0; n < 2; ++n)
      for (int l = 0; l < L; ++l) {
        const uint32_t halves = (g.l[n][l].n_tile >> 6) >= 2 ? 2 : 1;
        mbar_init(&own_bar[n][l], halves);
        mbar_init(&peer_bar[n][l], halves * CHAIN_CLUSTER);
      }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int n = 0; n < 2; ++n)
      for (int l = 0; l < L; ++l) {
        if (l > 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&g.l[n][l].map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&g.l[n][l].map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&g.l[n][l].map_y) : "memory");
      }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // every CTA's barriers exist before a peer arrives on them
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int nkb0 = g.l[0][0].nkb;
  // Ring geometry per layer.  What paces a k-loop here is the TMA unit: every tensor load costs it ~0.16 us whatever its
  // size (0.35 / 0.38 / 0.42 us per k-block = two loads of 160 / 256 / 384 rows: measured with 4, 6 and 8 stages, with one
  // and with two issuing threads, with and without the MMAs - always the same).  So layers >= 1 load k-blocks in GROUPS of
  // kgroup (2) through 3-D tensor maps - one load for the activation tiles, one for the weight slices of the group - and a
  // stage holds a group: [A tile x G | W slice x G].  Where the stage size SHRINKS from one layer to the next, a new stage j
  // only overlaps old stages of index <= j, whose barriers the producers re-claim - i.e. wait for - in order before they
  // reach stage j; where it GROWS the producers first wait for every barrier's last use (one exposed load latency).
  // Barrier s serves stage s of every geometry; its phase parities are tracked in bit masks.
  auto stage_bytes = [&](int l) { return l == 0 ? CHAIN_STAGE_BYTES : g.l[0][l].kgroup * (A_STAGE_BYTES + g.l[0][l].n_tile * BK * 2); };
  auto stages_of = [](int sz) { const int s = (CHAIN_STAGES * CHAIN_STAGE_BYTES) / sz; return s > SMAX ? SMAX : s; };

  if (warp == 0 || warp == DUO_A_WARP) {
    // ===== two TMA producers: warp 0 requests the WEIGHT slices, warp 10 the ACTIVATION tiles =====
    // Both walk the same stage sequence and wait for a stage to be free on their own (non-consuming parity waits); the
    // weight thread arms the full barrier with the bytes of both loads (an activation group that lands first only drives
    // the transaction count negative until then).  The weight thread depends on no activations at all, so every layer's
    // first weight slices are in flight while the previous layer's outputs still travel.
    if (elect_one()) {
      const bool wprod = warp == 0;
      int nxt = 0, sz = stage_bytes(0), S = CHAIN_STAGES;    // nxt: the stage the next claim takes (no division in this loop:
      uint32_t used = 0, par = 0;                            // a single thread's instruction latencies show in the k-loop)
      auto claim = [&](bool wait) {          // the next stage of the sequence; waits until it is free
        const int s = nxt;
        const uint32_t bit = 1u << s;
        if (used & bit) { if (wait) duo_wait(&empty_bar[s], (par >> s) & 1u, wprod ? 101u : 104u, ab); par ^= bit; }
        used |= bit;
        nxt = (s + 1 == S) ? 0 : s + 1;
        return s;
      };
      for (int kb = 0; kb < nkb0; ++kb)
        for (int n = 0; n < 2; ++n) {
          const ChainLayer& C = g.l[n][0];
          const int s = claim(wprod);        // (layer 0's A tiles are converted by the epilogue warps: the A thread only counts)
          if (wprod) {
            mbar_expect_tx(&full_bar[s], (uint32_t)(C.n_tile * BK * 2));
            tma_load_2d(smem + s * sz + A_STAGE_BYTES, &C.map_w, kb * BK, (int)crank * C.n_tile, &full_bar[s]);
          }
        }
      // Layers >= 1.  These loops are kept as short as they can be made - a single thread's instruction latencies ARE the
      // k-loop's clock here (tools/probe/tma_mma_rate.cu: ~270 ns per iteration for a producer that waits, arms and issues
      // two tensor loads, whatever their size): stage address, barrier addresses and the k coordinate advance incrementally.
      const uint32_t ring = smem_u32(smem), full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      for (int l = 1; l < L; ++l) {
        const int nsz = stage_bytes(l);
        if (nsz != sz) {
          if (nsz > sz)                      // stages grow: every old stage must have been consumed
            for (int k = 0; k < S; ++k)
              if (used & (1u << k)) duo_wait(&empty_bar[k], (par >> k) & 1u, 105u, ab);
          sz = nsz; S = stages_of(sz); nxt = 0;
        }
        for (int n = 0; n < 2; ++n) {
          const ChainLayer& C = g.l[n][l];
          const int nkb = C.nkb, n_own = g.l[n][l - 1].n_tile >> 6, G = C.kgroup;
          const uint32_t a_bytes = (uint32_t)(G * A_STAGE_BYTES), tx = a_bytes + (uint32_t)(G * C.n_tile * BK * 2);
          const CUtensorMap* map = wprod ? &C.map_w : &C.map_x;
          const int row = wprod ? (int)crank * C.n_tile : m0;
          const uint32_t off = wprod ? a_bytes : 0u;
          int kb = (int)crank * n_own;       // chain_kb(0): this CTA's own slices first
          if (kb >= nkb) kb -= nkb;
          if (!wprod) {
            duo_wait(&own_bar[n][l - 1], 0u, 102u, ab);
            duo_trace(g, 2 * l + n, DT_PRODUCER);
          }
          int s = nxt;
          for (int part = 0; part < 2; ++part) {           // own k-blocks, then the peers'
            const int cnt = part == 0 ? n_own : nkb - n_own;
            if (part == 1 && !wprod && cnt > 0) { duo_wait(&peer_bar[n][l - 1], 0u, 103u, ab); duo_trace(g, 2 * l + n, DT_PEERS); }
            for (int i = 0; i < cnt; i += G) {
              const uint32_t bit = 1u << s, fb = full0 + 8u * (uint32_t)s;
              if (used & bit) { duo_wait_addr(empty0 + 8u * (uint32_t)s, (par >> s) & 1u, wprod ? 101u : 104u, ab); par ^= bit; }
              used |= bit;
              const uint32_t dst = ring + (uint32_t)(s * sz) + off;
              if (wprod) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(tx) : "memory");
                duo_trace_kb(g, 16 + 2 * l + n, (part ? n_own : 0) + i);
              }
              if (G == 1) tma_load_2d_addr(dst, map, kb * BK, row, fb);
              else tma_load_3d_addr(dst, map, 0, row, kb, fb);
              kb += G;
              if (kb >= nkb) kb -= nkb;
              s = (s + 1 == S) ? 0 : s + 1;
            }
          }
          nxt = s;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      int nxt = 0, sz = stage_bytes(0), S = CHAIN_STAGES;
      uint32_t par = 0;                      // per barrier: parity of the next full phase
      {
        const uint32_t idesc = umma_idesc_bf16(g.l[0][0].n_tile);
        for (int kb = 0; kb < nkb0; ++kb) {
          const int s0 = nxt, s1 = nxt + 1;  // (four stages: slots in pairs)
          duo_wait(&full_bar[s0], (par >> s0) & 1u, 111u, ab);
          duo_wait(&a_bar[s0], (par >> s0) & 1u, 112u, ab);            // the A tile the epilogue warps converted
          if (kb == 0) { duo_trace(g, 0, DT_FIRST_OPERANDS); duo_trace(g, 1, DT_FIRST_OPERANDS); }
          duo_trace_kb(g, 0, kb);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s0 * sz);
          const uint32_t b0_addr = a_addr + A_STAGE_BYTES, b1_addr = smem_u32(smem + s1 * sz) + A_STAGE_BYTES;
#pragma unroll
          for (int j = 0; j < BK / 16; ++j)
            umma_bf16(tmem, umma_desc_k_sw128(a_addr + j * 32), umma_desc_k_sw128(b0_addr + j * 32), idesc, (kb > 0) || (j > 0));
          duo_wait(&full_bar[s1], (par >> s1) & 1u, 113u, ab);
          tc_fence_after();
#pragma unroll
          for (int j = 0; j < BK / 16; ++j)
            umma_bf16(tmem + 256, umma_desc_k_sw128(a_addr + j * 32), umma_desc_k_sw128(b1_addr + j * 32), idesc, (kb > 0) || (j > 0));
          umma_commit(&empty_bar[s0]);       // (a commit covers every MMA issued before it: slot s0's A tile fed both)
          umma_commit(&empty_bar[s1]);
          par ^= (1u << s0) | (1u << s1);
          nxt = (s1 + 1 == S) ? 0 : s1 + 1;
        }
        umma_commit(&acc_bar[0]);
        umma_commit(&acc_bar[1]);
        duo_trace(g, 0, DT_MMA_ISSUED); duo_trace(g, 1, DT_MMA_ISSUED);
      }
      // Layers >= 1: the lean loop (see the producers).  No tcgen05 fence per k-block: the operands were written by the
      // async proxy (TMA) and are read by it (the MMA); what orders them is the mbarrier.  Descriptors advance by adding to
      // their address field (units of 16 bytes; the ring lies below 256 KB, so no carry leaves the field).
      const uint32_t ring = smem_u32(smem), full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      for (int l = 1; l < L; ++l) {
        const int nsz = stage_bytes(l);
        if (nsz != sz) { sz = nsz; S = stages_of(sz); nxt = 0; }
        const uint32_t dstep = (uint32_t)sz >> 4;
        for (int n = 0; n < 2; ++n) {
          const ChainLayer& C = g.l[n][l];
          const uint32_t idesc = umma_idesc_bf16(C.n_tile);
          const uint32_t acc = tmem + (uint32_t)(n * 256);
          const int G = C.kgroup, groups = C.nkb / G;
          const uint32_t wstep = (uint32_t)(C.n_tile * BK * 2) >> 4;
          const uint64_t da0 = umma_desc_k_sw128(ring), db0 = umma_desc_k_sw128(ring + (uint32_t)(G * A_STAGE_BYTES));
          int s = nxt;
          uint64_t da = da0 + (uint64_t)((uint32_t)s * dstep), db = db0 + (uint64_t)((uint32_t)s * dstep);
          bool fresh = true;
          duo_trace(g, 2 * l + n, DT_FIRST_OPERANDS);
          for (int i = 0; i < groups; ++i) {
            duo_wait_addr(full0 + 8u * (uint32_t)s, (par >> s) & 1u, 114u, ab);
            par ^= 1u << s;
            duo_trace_kb(g, 2 * l + n, i * G);
            if (G == 1) {
#pragma unroll
              for (int j = 0; j < BK / 16; ++j) {
                if (fresh) { tc_fence_after(); umma_bf16(acc, da, db, idesc, false); fresh = false; }   // (fence: the epilogue's loads of this accumulator)
                else umma_bf16_acc(acc, da + (uint64_t)(2 * j), db + (uint64_t)(2 * j), idesc);
              }
            } else {
#pragma unroll
              for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int j = 0; j < BK / 16; ++j) {
                  if (fresh) { tc_fence_after(); umma_bf16(acc, da, db, idesc, false); fresh = false; }
                  else umma_bf16_acc(acc, da + (uint64_t)(q * (A_STAGE_BYTES >> 4) + 2 * j), db + (uint64_t)(q * wstep + 2 * j), idesc);
                }
            }
            umma_commit_addr(empty0 + 8u * (uint32_t)s);
            if (++s == S) { s = 0; da = da0; db = db0; }
            else { da += dstep; db += dstep; }
          }
          nxt = s;
          umma_commit(&acc_bar[n]);
          duo_trace(g, 2 * l + n, DT_MMA_ISSUED);
        }
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps: TMEM lane quarter = warp % 4, two warps per quarter split the columns =====
    {
      // layer 0's A tiles from the fp32 input, once for both networks (see mlp_chain_body): slot of k-block kb = 2 kb
      if (g.overlap_prev) griddep_wait();
      const int w8 = (tid - 64) >> 5, sub = lane >> 4, c4 = lane & 15;
      const int K0 = g.K0;
      const float* xbase = x32 + (int64_t)(m0 + w8 * 16 + sub) * K0 + c4 * 4;
      const int rows_left = g.M - (m0 + w8 * 16 + sub);
      auto load_tile = [&](int kb, float4 (&v)[8]) {
        const bool col_ok = kb * BK + c4 * 4 < K0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (col_ok && 2 * i < rows_left) v[i] = __ldg(reinterpret_cast<const float4*>(xbase + (int64_t)(2 * i) * K0 + kb * BK));
        }
      };
      float4 t0[8], t1[8], t2[8];
      load_tile(0, t0);
      if (nkb0 > 1) load_tile(1, t1);
      constexpr int sz0 = CHAIN_STAGE_BYTES, S0 = CHAIN_STAGES;
      auto convert = [&](int kb, const float4 (&cur)[8]) {
        const int s = (2 * kb) % S0, u = (2 * kb) / S0;
        if (u > 0) duo_wait(&empty_bar[s], (uint32_t)((u - 1) & 1), 121u, ab);
        uint8_t* tile = smem + s * sz0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = w8 * 16 + 2 * i + sub;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(cur[i].x, cur[i].y), h1 = __floats2bfloat162_rn(cur[i].z, cur[i].w);
          *reinterpret_cast<uint2*>(tile + (r >> 3) * 1024 + (r & 7) * 128 + ((((c4 >> 1) ^ (r & 7)) << 4) | ((c4 & 1) << 3))) =
              make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&a_bar[s])) : "memory");
      };
      for (int kb = 0; kb < nkb0; kb += 3) {
        if (kb + 2 < nkb0) load_tile(kb + 2, t2);
        convert(kb, t0);
        if (kb + 1 >= nkb0) break;
        if (kb + 3 < nkb0) load_tile(kb + 3, t0);
        convert(kb + 1, t1);
        if (kb + 2 >= nkb0) break;
        if (kb + 4 < nkb0) load_tile(kb + 4, t1);
        convert(kb + 2, t2);
      }
    }
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const bool leader = q == 0 && lane == 0;
    for (int l = 0; l < L; ++l)
      for (int n = 0; n < 2; ++n) {
        const ChainLayer& C = g.l[n][l];
        const int n_tile = C.n_tile, n0 = (int)crank * n_tile;
        const uint32_t acc = tmem + (uint32_t)(n * 256);
        uint8_t* buf = out_buf + half * (BM * 128);
        // all eight warps: the previous task's last stores have left the staging buffers (their leaders waited before
        // arriving here) and nobody reads the previous bias slice any more
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 64) duo_trace(g, 2 * l + n, DT_EPI_BEGIN);
        if (l + 1 < L) {
          // ---- hidden layer: bias + ELU -> bf16, 64-column sub-tiles through this half's staging buffer (mlp_chain_body) ----
          { const int e = tid - 64; if (e < n_tile) bias_s[e] = __ldg(C.bias + n0 + e); }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          duo_wait(&acc_bar[n], (uint32_t)(l & 1), 131u, ab);
          tc_fence_after();
          if (tid == 64) duo_trace(g, 2 * l + n, DT_ACC_COMPLETE);
          const int n_sub = n_tile >> 6, halves = n_sub >= 2 ? 2 : 1, per_half = n_sub / halves;
          if (half < halves) {
            const int cbeg = half * per_half * 64;
            const uint32_t taddr = acc + ((uint32_t)(q * 32) << 16) + (uint32_t)cbeg;
            uint8_t* rowp = buf + (row >> 3) * 1024 + (row & 7) * 128;
            uint32_t ra[32], rb[32], pk[32];
            auto act32 = [&](const uint32_t (&r)[32], int col, uint32_t* out16) {
              const float4* b4 = reinterpret_cast<const float4*>(bias_s + col);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 b = b4[i];
                __nv_bfloat162 h0 = __floats2bfloat162_rn(elu1(__uint_as_float(r[4 * i]) + b.x), elu1(__uint_as_float(r[4 * i + 1]) + b.y));
                __nv_bfloat162 h1 = __floats2bfloat162_rn(elu1(__uint_as_float(r[4 * i + 2]) + b.z), elu1(__uint_as_float(r[4 * i + 3]) + b.w));
                out16[2 * i] = *reinterpret_cast<uint32_t*>(&h0);
                out16[2 * i + 1] = *reinterpret_cast<uint32_t*>(&h1);
              }
            };
            tmem_ld32_issue(taddr, ra);
            for (int js = 0; js < per_half; ++js) {
              tmem_ld_wait(ra);
              tmem_ld32_issue(taddr + js * 64 + 32, rb);
              act32(ra, cbeg + js * 64, pk);
              tmem_ld_wait(rb);
              if (js + 1 < per_half) tmem_ld32_issue(taddr + js * 64 + 64, ra);
              act32(rb, cbeg + js * 64 + 32, pk + 16);
              if (js > 0) {                                       // the buffer fed the previous sub-tile's store
                if (leader) tma_store_wait_read();
                if (half) asm volatile("bar.sync 3, 128;" ::: "memory");
                else asm volatile("bar.sync 2, 128;" ::: "memory");
              }
#pragma unroll
              for (int c = 0; c < 8; ++c)
                *reinterpret_cast<uint4*>(rowp + ((c ^ (row & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
              fence_async_smem();
              if (half) asm volatile("bar.sync 3, 128;" ::: "memory");
              else asm volatile("bar.sync 2, 128;" ::: "memory");
              if (leader) {
                tma_store_2d(&C.map_y, buf, n0 + cbeg + js * 64, m0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
            }
            if (leader && half == 0) duo_trace(g, 2 * l + n, DT_COMPUTED);
            // the slices must have LANDED (not only left shared memory) before anybody is told to read them
            if (q == 0) {           // lane 0 reports to this CTA, lanes 1-4 to the four CTAs of the cluster in parallel
              if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
              if (leader && half == 0) duo_trace(g, 2 * l + n, DT_LANDED);
              __syncwarp();
              if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&own_bar[n][l])) : "memory");
              else if (lane <= CHAIN_CLUSTER) mbar_arrive_remote(&peer_bar[n][l], (uint32_t)(lane - 1));
            }
          }
          tc_fence_before();
          __syncwarp();
        } else {
          // ---- last layer: bias -> fp32 [M][N] ----
          duo_wait(&acc_bar[n], (uint32_t)(l & 1), 132u, ab);
          tc_fence_after();
          if (tid == 64) duo_trace(g, 2 * l + n, DT_ACC_COMPLETE);
          const int sub_cols = (C.epilogue == 0) ? 32 : 64;
          int cb = 0, ce = half ? 0 : n_tile;
          if (n_tile >= 2 * sub_cols) { cb = half ? n_tile / 2 : 0; ce = half ? n_tile : n_tile / 2; }
          mmb_mlp_layer_params p = {};
          p.M = g.M; p.N = C.N; p.epilogue = C.epilogue; p.bias = C.bias;
          if (cb < ce) {
            epilogue_row_staged(p, acc + ((uint32_t)(q * 32) << 16), row, n0, n_tile, cb, ce, buf, [&](int j) {
              fence_async_smem();
              if (half) asm volatile("bar.sync 3, 128;" ::: "memory");
              else asm volatile("bar.sync 2, 128;" ::: "memory");
              if (leader) {
                tma_store_2d(&C.map_y, buf, n0 + j * sub_cols, m0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                tma_store_wait_read();
              }
              if (half) asm volatile("bar.sync 3, 128;" ::: "memory");
              else asm volatile("bar.sync 2, 128;" ::: "memory");
            }, 0);
            if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            if (leader && half == 0) duo_trace(g, 2 * l + n, DT_LANDED);
          }
          tc_fence_before();
          __syncwarp();
        }
      }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // no CTA leaves while a peer's arrival may still be on its way to the barriers here
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// ---- host: 2-D tensor maps (rows x Kpad bf16, box = box_rows x 64, SWIZZLE_128B) through the driver entry point ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}
// Encoding a tensor map costs about a microsecond of host time per call and a layer needs two: a small direct-mapped
// cache keyed by (base, rows, kpad, box_rows) makes the steady-state launch as cheap as a plain one.
struct MapCacheEntry {
  const void* base = nullptr;
  uint64_t rows = 0, kpad = 0;
  uint32_t box_rows = 0;
  CUtensorMap map;
};
inline bool make_map_bf16_2d_uncached(CUtensorMap* map, const void* base, uint64_t rows, uint64_t kpad, uint32_t box_rows);
inline bool make_map_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t kpad, uint32_t box_rows) {
  constexpr int kEntries = 1024;
  static thread_local MapCacheEntry cache[kEntries];
  const uint64_t h = (reinterpret_cast<uintptr_t>(base) >> 8) * 0x9E3779B97F4A7C15ull + rows * 31 + kpad * 7 + box_rows;
  MapCacheEntry& e = cache[(h >> 32) % kEntries];
  if (e.base != base || e.rows != rows || e.kpad != kpad || e.box_rows != box_rows) {
    if (!make_map_bf16_2d_uncached(&e.map, base, rows, kpad, box_rows)) return false;
    e.base = base; e.rows = rows; e.kpad = kpad; e.box_rows = box_rows;
  }
  memcpy(map, &e.map, sizeof(CUtensorMap));
  return true;
}
// fp32 output of the last layer: [rows][cols] with row stride ld elements, box = 128 rows x 32 columns (128 bytes)
inline bool make_map_f32_out(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// fp32 operand [rows][cols], row pitch ld elements, box = box_rows x 32 columns (128 bytes), SWIZZLE_128B; columns and rows
// outside the matrix read as zeros (the tf32 chain pads nothing)
inline bool make_map_f32_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// [rows][kpad] bf16 seen as (64 elements, rows, kpad / 64): box = 64 x box_rows x group, i.e. `group` consecutive k-blocks
// of box_rows rows as consecutive SWIZZLE_128B tiles in shared memory - ONE tensor load instead of `group` (the TMA unit
// spends ~0.16 us per load whatever its size: mlp_chain_duo_kernel)
inline bool make_map_bf16_kgroup(CUtensorMap* map, const void* base, uint64_t rows, uint64_t kpad, uint32_t box_rows, uint32_t group) {
  constexpr int kEntries = 128;
  static thread_local MapCacheEntry cache[kEntries];
  const uint32_t key = box_rows | (group << 16);
  const uint64_t h = (reinterpret_cast<uintptr_t>(base) >> 8) * 0x9E3779B97F4A7C15ull + rows * 31 + kpad * 7 + key;
  MapCacheEntry& e = cache[(h >> 32) % kEntries];
  if (e.base != base || e.rows != rows || e.kpad != kpad || e.box_rows != key) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t)BK, rows, kpad / BK};
    cuuint64_t strides[2] = {kpad * 2, (cuuint64_t)BK * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, box_rows, group};
    cuuint32_t estr[3] = {1, 1, 1};
    if (fn(&e.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    e.base = base; e.rows = rows; e.kpad = kpad; e.box_rows = key;
  }
  memcpy(map, &e.map, sizeof(CUtensorMap));
  return true;
}
inline bool make_map_bf16_2d_uncached(CUtensorMap* map, const void* base, uint64_t rows, uint64_t kpad, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {kpad, rows};
  cuuint64_t strides[1] = {kpad * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// fp32 [M][K] -> optional LayerNorm (mlp.py:58-59 feature_norm) -> bf16 [Mpad][Kpad], zero padded: the first
// layer's A operand.  One warp per row.
__device__ __forceinline__ void ln_cast_body(const float* __restrict__ x, int M, int Mpad, int K, int Kpad,
                                             const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int use_ln,
                                             __nv_bfloat16* __restrict__ y) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  griddep_launch_dependents();  // a programmatically launched first layer may run its prologue / weight loads meanwhile
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
// Plain cast (no LayerNorm): fp32 [M][K] -> bf16 [Mpad][Kpad], zero padded.  One thread per 4 output columns: 128-bit loads where
// the row allows them, 64-bit stores; the one-warp-per-row kernel above spent 7.4 us on the PPO observation block (4096 x 388),
// a fifth of the whole single-launch forward.
__device__ __forceinline__ void cast_pad_body(const float* __restrict__ x, int M, int Mpad, int K, int Kpad, __nv_bfloat16* __restrict__ y) {
  griddep_launch_dependents();
  const int q_per_row = Kpad >> 2;                       // Kpad is a multiple of 64
  const int64_t total = (int64_t)Mpad * q_per_row;
  const bool vec = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / q_per_row), k = (int)(i - (int64_t)r * q_per_row) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < M && k < K) {
      const float* xr = x + (int64_t)r * K + k;
      if (vec) {
        v = __ldg(reinterpret_cast<const float4*>(xr));
      } else {
        v.x = __ldg(xr);
        if (k + 1 < K) v.y = __ldg(xr + 1);
        if (k + 2 < K) v.z = __ldg(xr + 2);
        if (k + 3 < K) v.w = __ldg(xr + 3);
      }
    }
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&a);
    pk.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(y + (int64_t)r * Kpad + k) = pk;
  }
}
__global__ void __launch_bounds__(256) cast_pad_kernel(const float* __restrict__ x, int M, int Mpad, int K, int Kpad,
                                                       __nv_bfloat16* __restrict__ y) {
  cast_pad_body(x, M, Mpad, K, Kpad, y);
}
__global__ void __launch_bounds__(256) ln_cast_kernel(const float* __restrict__ x, int M, int Mpad, int K, int Kpad,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      float eps, int use_ln, __nv_bfloat16* __restrict__ y) {
  ln_cast_body(x, M, Mpad, K, Kpad, gamma, beta, eps, use_ln, y);
}
__global__ void __launch_bounds__(256) ln_cast_group_kernel(const __grid_constant__ LnCastGroupArgs g, int M, int Mpad, int K, int Kpad,
                                                            float eps, int use_ln) {
  const int a = blockIdx.y;
  ln_cast_body(g.x[a], M, Mpad, K, Kpad, g.gamma[a], g.beta[a], eps, use_ln, g.y[a]);
}
__global__ void __launch_bounds__(256) cast_pad_group_kernel(const __grid_constant__ LnCastGroupArgs g, int M, int Mpad, int K, int Kpad) {
  const int a = blockIdx.y;
  cast_pad_body(g.x[a], M, Mpad, K, Kpad, g.y[a]);
}

}  // namespace
}  // namespace mmb

using namespace mmb;

static bool ln_deferred_ok(const mmb_mlp_layer_params& p) {     // the deferred LayerNorm fields of include/mmb.h
  if (p.ln_in_stats && (!p.ln_c || p.ln_in_parts <= 0 || p.ln_in_n <= 0 || (reinterpret_cast<uintptr_t>(p.ln_in_stats) & 7u))) return false;
  if (p.ln_out_stats && (p.epilogue != 1 || p.n_tile != 256 || p.Npad % 256 || (reinterpret_cast<uintptr_t>(p.ln_out_stats) & 7u))) return false;
  return true;
}

extern "C" __attribute__((visibility("default"))) int32_t mmb_mlp_debug_status(unsigned int* out4) {
  unsigned int z[4] = {0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(out4, g_pair_dbg, 16) != cudaSuccess) return MMB_ECUDA;
  return cudaMemcpyToSymbol(g_pair_dbg, z, 16) == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

extern "C" int32_t mmb_mlp_layer_group(const mmb_mlp_layer_params* params, int32_t count, void* stream) {
  if (!params || count <= 0 || count > MMB_MAX_GROUP) return MMB_EINVAL;
  // A rollout loop calls this with the same few parameter arrays over and over (one per layer and team): the validated
  // arguments and their 3 x count tensor maps are kept per distinct array (compared byte for byte), so the steady state is a
  // memcmp and a launch.  (The per-map cache alone thrashes: a team of ten has 240 maps in play, and a re-encode costs ~3 us.)
  struct GroupSlot { bool valid; int count, persist, smem; mmb_mlp_layer_params in[MMB_MAX_GROUP]; WsGroupArgs g; };
  constexpr int kSlots = 32;
  static thread_local GroupSlot* slots = nullptr;
  if (!slots) slots = static_cast<GroupSlot*>(calloc(kSlots, sizeof(GroupSlot)));
  GroupSlot* slot = nullptr;
  if (slots) {
    uint64_t h = 1469598103934665603ull;
    const unsigned char* bytes = reinterpret_cast<const unsigned char*>(params);
    for (size_t i = 0; i < (size_t)count * sizeof(mmb_mlp_layer_params); i += 8) {     // (the struct is a multiple of 8 bytes)
      uint64_t w;
      memcpy(&w, bytes + i, 8);
      h = (h ^ w) * 1099511628211ull;
    }
    slot = &slots[(h >> 20) % kSlots];
    if (slot->valid && slot->count == count && memcmp(slot->in, params, (size_t)count * sizeof(mmb_mlp_layer_params)) == 0) {
      const mmb_mlp_layer_params& q0 = params[0];
      LaunchScope ls(K_MLP_LAYER, (cudaStream_t)stream);
      cudaLaunchConfig_t cfg = {};
      const int tiles = count * (q0.Mpad / BM) * (q0.Npad / q0.n_tile), sms = sm_count();
      cfg.gridDim = slot->persist ? dim3(tiles < sms ? tiles : sms) : dim3(q0.Mpad / BM, q0.Npad / q0.n_tile, count);
      cfg.blockDim = dim3(slot->persist ? PERSIST_GROUP_THREADS : WS_THREADS);
      cfg.dynamicSmemBytes = slot->smem;
      cfg.stream = (cudaStream_t)stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = q0.overlap_prev ? 1 : 0;
      const cudaError_t le = slot->persist ? cudaLaunchKernelEx(&cfg, mlp_layer_ws_persist_group_kernel, slot->g, (int)count)
                                           : cudaLaunchKernelEx(&cfg, mlp_layer_ws_group_kernel, slot->g);
      if (le != cudaSuccess) { (void)cudaGetLastError(); return MMB_ECUDA; }
      return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
    }
  }
  static thread_local WsGroupArgs g;       // 8 KB: not on the stack; the launch copies it
  const mmb_mlp_layer_params& p0 = params[0];
  if (p0.n_tile <= 0 || p0.Kpad % BK || p0.Mpad % BM || p0.Npad % p0.n_tile || p0.n_tile % 32 || p0.n_tile > 512) return MMB_EINVAL;
  const int stage_bytes = A_STAGE_BYTES + p0.n_tile * BK * 2;
  int stages = SMEM_BUDGET / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return MMB_EUNSUPPORTED;
  int smem = stages * stage_bytes;
  // persistent, epilogue-overlapped variant over ALL problems' tiles (same conditions as the single-problem launch)
  static const int persist_pref = [] { const char* v = getenv("MMB_MLP_PERSIST"); return v ? atoi(v) : 1; }();
  const int num_tiles = count * (p0.Mpad / BM) * (p0.Npad / p0.n_tile);
  const int sub_cols_h = (p0.epilogue == 0) ? 32 : 64;
  bool persist = persist_pref && p0.epilogue != 2 && p0.n_tile <= 256 && num_tiles >= 2 * sm_count() &&
                 (p0.n_tile < 2 * sub_cols_h || (p0.n_tile / 2) % sub_cols_h == 0);
  for (int a = 0; a < count && persist; ++a)
    persist = (p0.epilogue == 0) ? (((params[a].y_stride & 3) | (reinterpret_cast<uintptr_t>(params[a].y) & 15u)) == 0) : (p0.n_tile % 64 == 0);
  if (persist) {
    stages = (SMEM_BUDGET + 24 * 1024 - PERSIST_GROUP_PARTS * BM * 128) / stage_bytes;      // ring + one 16 KB staging buffer per part within 224 KB
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    smem = stages * stage_bytes + PERSIST_GROUP_PARTS * BM * 128;
  }
  for (int a = 0; a < count; ++a) {
    mmb_mlp_layer_params p = params[a];
    if (p.M != p0.M || p.N != p0.N || p.K != p0.K || p.Mpad != p0.Mpad || p.Kpad != p0.Kpad || p.Npad != p0.Npad ||
        p.n_tile != p0.n_tile || p.epilogue != p0.epilogue || p.y_stride != p0.y_stride)
      return MMB_EINVAL;                   // one grid for all: identical geometry
    if (p.M <= 0 || p.N <= 0 || p.K <= 0 || !p.x || !p.w || !p.bias || !p.y || p.Kpad < p.K || p.Npad < p.N || p.Mpad < p.M) return MMB_EINVAL;
    if (p.epilogue < 0 || p.epilogue > 2) return MMB_EINVAL;
    if (p.epilogue == 2 && (p.n_tile != p.N || p.Npad != p.N || !p.ln_gamma || !p.ln_beta)) return MMB_EINVAL;
    if (!ln_deferred_ok(p) || (p.ln_in_stats == nullptr) != (p0.ln_in_stats == nullptr) || (p.ln_out_stats == nullptr) != (p0.ln_out_stats == nullptr))
      return MMB_EINVAL;
    if ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.w)) & 15u) return MMB_EALIGN;
    p.stages = stages;
    if (!make_map_bf16_2d(&g.map_x[a], p.x, (uint64_t)p.Mpad, (uint64_t)p.Kpad, BM) ||
        !make_map_bf16_2d(&g.map_w[a], p.w, (uint64_t)p.Npad, (uint64_t)p.Kpad, (uint32_t)(p.n_tile > 256 ? 256 : p.n_tile)))
      return MMB_ECUDA;
    if (p.epilogue == 0 && ((p.y_stride & 3) | (reinterpret_cast<uintptr_t>(p.y) & 15u)) == 0) {
      if (!make_map_f32_out(&g.map_y[a], p.y, (uint64_t)p.M, (uint64_t)p.N, (uint64_t)p.y_stride)) return MMB_ECUDA;
      if ((p.n_tile / 32) * BM * 128 > smem) return MMB_EUNSUPPORTED;
    } else if (p.epilogue != 0 && p.n_tile % 64 == 0) {
      if ((reinterpret_cast<uintptr_t>(p.y) & 15u) || (p.y_stride % 8)) return MMB_EALIGN;
      if (!make_map_bf16_2d(&g.map_y[a], p.y, (uint64_t)p.Mpad, (uint64_t)p.y_stride, BM)) return MMB_ECUDA;
    } else {
      g.map_y[a] = g.map_x[a];
    }
    g.p[a] = p;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= MMB_MAX_DEVICES) return MMB_EUNSUPPORTED;
  static bool attr_done[MMB_MAX_DEVICES] = {};
  if (!attr_done[dev]) {
    if (cudaFuncSetAttribute(mlp_layer_ws_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET) != cudaSuccess)
      return MMB_ECUDA;
    if (cudaFuncSetAttribute(mlp_layer_ws_persist_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024) != cudaSuccess)
      return MMB_ECUDA;
    attr_done[dev] = true;
  }
  if (slot) {                              // remember the validated arguments and their maps
    slot->valid = false;
    slot->count = count; slot->persist = persist ? 1 : 0; slot->smem = smem;
    memcpy(slot->in, params, (size_t)count * sizeof(mmb_mlp_layer_params));
    memcpy(&slot->g, &g, sizeof(WsGroupArgs));
    slot->valid = true;
  }
  if (persist) {
    LaunchScope ls(K_MLP_LAYER, (cudaStream_t)stream);
    cudaLaunchConfig_t cfg = {};
    const int sms = sm_count();
    cfg.gridDim = dim3(num_tiles < sms ? num_tiles : sms);
    cfg.blockDim = dim3(PERSIST_GROUP_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = p0.overlap_prev ? 1 : 0;
    if (cudaLaunchKernelEx(&cfg, mlp_layer_ws_persist_group_kernel, g, (int)count) != cudaSuccess) { (void)cudaGetLastError(); return MMB_ECUDA; }
    return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
  }
  {
    LaunchScope ls(K_MLP_LAYER, (cudaStream_t)stream);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p0.Mpad / BM, p0.Npad / p0.n_tile, count);
    cfg.blockDim = dim3(WS_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = p0.overlap_prev ? 1 : 0;
    if (cudaLaunchKernelEx(&cfg, mlp_layer_ws_group_kernel, g) != cudaSuccess) { (void)cudaGetLastError(); return MMB_ECUDA; }
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

extern "C" int32_t mmb_ln_cast_group(const float* const* x, int32_t count, int32_t M, int32_t Mpad, int32_t K, int32_t Kpad,
                                     const float* const* gamma, const float* const* beta, float eps, int32_t use_ln,
                                     void* const* y_bf16, void* stream) {
  if (!x || !y_bf16 || count <= 0 || count > MMB_MAX_GROUP || M <= 0 || Mpad < M || K <= 0 || Kpad < K) return MMB_EINVAL;
  if (use_ln && (!gamma || !beta)) return MMB_EINVAL;
  LnCastGroupArgs g = {};
  for (int a = 0; a < count; ++a) {
    if (!x[a] || !y_bf16[a] || (use_ln && (!gamma[a] || !beta[a]))) return MMB_EINVAL;
    g.x[a] = x[a]; g.y[a] = static_cast<__nv_bfloat16*>(y_bf16[a]);
    g.gamma[a] = use_ln ? gamma[a] : nullptr; g.beta[a] = use_ln ? beta[a] : nullptr;
  }
  {
    LaunchScope ls(K_LN_CAST, (cudaStream_t)stream);
    if (!use_ln && Kpad % 4 == 0) {
      int64_t bx = ((int64_t)Mpad * (Kpad / 4) + 255) / 256;
      const int64_t cap = (sm_count() * 8 + count - 1) / count;
      if (bx > cap) bx = cap;
      cast_pad_group_kernel<<<dim3((unsigned)bx, count), 256, 0, (cudaStream_t)stream>>>(g, M, Mpad, K, Kpad);
    } else {
      ln_cast_group_kernel<<<dim3((Mpad * 32 + 255) / 256, count), 256, 0, (cudaStream_t)stream>>>(g, M, Mpad, K, Kpad, eps, use_ln);
    }
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

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
  if (!ln_deferred_ok(p)) return MMB_EINVAL;
  if ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.w)) & 15u) return MMB_EALIGN;
  const int stage_bytes = A_STAGE_BYTES + p.n_tile * BK * 2;
  static const bool legacy = [] { const char* v = getenv("MMB_MLP_VARIANT"); return v && v[0] == 'l'; }();
  int stages = SMEM_BUDGET / stage_bytes;
  if (stages > (legacy ? 4 : MAX_STAGES)) stages = legacy ? 4 : MAX_STAGES;
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
  if (legacy) {
    LaunchScope ls(K_MLP_LAYER, (cudaStream_t)stream);
    mlp_layer_kernel<<<dim3(p.Mpad / BM, p.Npad / p.n_tile), MLP_THREADS, smem, (cudaStream_t)stream>>>(p);
    return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
  }
  static bool ws_attr_done[MMB_MAX_DEVICES] = {};
  if (!ws_attr_done[dev]) {
    if (cudaFuncSetAttribute(mlp_layer_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET) != cudaSuccess)
      return MMB_ECUDA;
    ws_attr_done[dev] = true;
  }
  // cluster size along M (MMB_MLP_CLUSTER = 2 / 4): the CTAs of a cluster share (multicast) each weight tile.  Off by
  // default: measured on B200 at M = 4096 it is 5-9 % SLOWER than independent CTAs (18.5 vs 17.6 us for 1024 x 1024),
  // i.e. the k-loop is not L2-read-bound at these sizes and the cluster launch / sync constraints cost more than the
  // saved L2 traffic.
  static const int cluster_pref = [] { const char* v = getenv("MMB_MLP_CLUSTER"); return v ? atoi(v) : 1; }();
  int cm = 1;
  for (int c = cluster_pref; c > 1; c >>= 1)
    if ((p.Mpad / BM) % c == 0 && p.n_tile <= 256 && (p.n_tile / c) % 8 == 0 && p.n_tile % c == 0) { cm = c; break; }
  // cta_group::2 (MMB_MLP_PAIR=1): CTA pairs along M run one 256-row MMA, each CTA holding half of the weight tile.
  // Parity-tested; measured on the 1024 x 1024 layer at M = 4096: k-loop 8.9 -> 6.1 us per tile, the layer launched alone
  // 17.8 -> 16.5 us, but the PDL-chained forward 45.1 -> 46.9 us (a pair can only start when BOTH SMs of its TPC are
  // free, which costs the overlap with the previous layer's drain) - hence opt-in.
  static const int pair_pref = [] { const char* v = getenv("MMB_MLP_PAIR"); return v ? atoi(v) : 0; }();
  const bool pair = pair_pref && (p.Mpad / BM) % 2 == 0 && p.n_tile <= 256 && p.n_tile % 32 == 0 && p.n_tile >= 64;
  if (p.operand_type != 0) return MMB_EUNSUPPORTED;   // tf32 operands: mmb_mlp_chain only
  if (pair) {
    cm = 2;
    const int sb = A_STAGE_BYTES + (p.n_tile / 2) * BK * 2;
    int st = SMEM_BUDGET / sb;
    if (st > MAX_STAGES) st = MAX_STAGES;
    p.stages = st;
  }
  // persistent, epilogue-overlapped variant: only worth it when a CTA gets more than one tile
  static const int persist_pref = [] { const char* v = getenv("MMB_MLP_PERSIST"); return v ? atoi(v) : 1; }();
  const int num_tiles = (p.Mpad / BM) * (p.Npad / p.n_tile);
  const bool y_tma = (p.epilogue == 0) ? (((p.y_stride & 3) | (reinterpret_cast<uintptr_t>(p.y) & 15u)) == 0) : (p.n_tile % 64 == 0);
  const int sub_cols_h = (p.epilogue == 0) ? 32 : 64;
  const bool persist = persist_pref && !pair && cm == 1 && p.epilogue != 2 && p.n_tile <= 256 && y_tma && num_tiles >= 2 * sm_count() &&
                       (p.n_tile < 2 * sub_cols_h || (p.n_tile / 2) % sub_cols_h == 0);
  if (persist) {
    const int sb = A_STAGE_BYTES + p.n_tile * BK * 2;
    int st = (SMEM_BUDGET + 24 * 1024 - 2 * BM * 128) / sb;      // ring + two 16 KB staging buffers within 224 KB
    if (st > MAX_STAGES) st = MAX_STAGES;
    p.stages = st;
  }
  CUtensorMap map_x, map_w, map_y;
  if (!make_map_bf16_2d(&map_x, p.x, (uint64_t)p.Mpad, (uint64_t)p.Kpad, BM) ||
      !make_map_bf16_2d(&map_w, p.w, (uint64_t)p.Npad, (uint64_t)p.Kpad, (uint32_t)((p.n_tile > 256 ? 256 : p.n_tile) / cm)))   // pair: cm == 2
    return MMB_ECUDA;
  if (p.epilogue == 0 && ((p.y_stride & 3) | (reinterpret_cast<uintptr_t>(p.y) & 15u)) == 0) {
    // fp32 [M][N] rows of y_stride floats with a 16-byte aligned base and pitch: staged in the operand ring, TMA-stored
    if (!make_map_f32_out(&map_y, p.y, (uint64_t)p.M, (uint64_t)p.N, (uint64_t)p.y_stride)) return MMB_ECUDA;
    if ((p.n_tile / 32) * BM * 128 > smem) return MMB_EUNSUPPORTED;
  } else if (p.epilogue == 0) {
    map_y = map_x;         // e.g. the critic's [M][1] value column: row-per-thread stores; the map is not used
  } else if (p.n_tile % 64 == 0) {   // bf16 [Mpad][y_stride]: the next layer's operand (columns >= N are written as zeros)
    if ((reinterpret_cast<uintptr_t>(p.y) & 15u) || (p.y_stride % 8)) return MMB_EALIGN;
    if (!make_map_bf16_2d(&map_y, p.y, (uint64_t)p.Mpad, (uint64_t)p.y_stride, BM)) return MMB_ECUDA;
  } else {
    map_y = map_x;         // narrow bf16 tiles keep the row-per-thread stores; the map is not used
  }
  {
    LaunchScope ls(K_MLP_LAYER, (cudaStream_t)stream);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.Mpad / BM, p.Npad / p.n_tile);
    cfg.blockDim = dim3(WS_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (cm > 1) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = cm; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
      ++na;
    }
    if (p.overlap_prev) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    static bool pair_attr_done[MMB_MAX_DEVICES] = {};
    if (pair && !pair_attr_done[dev]) {
      if (cudaFuncSetAttribute(mlp_layer_ws_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET) != cudaSuccess) return MMB_ECUDA;
      pair_attr_done[dev] = true;
    }
    cudaError_t le;
    if (persist) {
      static bool persist_attr_done[MMB_MAX_DEVICES] = {};
      const int psmem = p.stages * (A_STAGE_BYTES + p.n_tile * BK * 2) + 2 * BM * 128;
      if (!persist_attr_done[dev]) {
        if (cudaFuncSetAttribute(mlp_layer_ws_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024) != cudaSuccess) return MMB_ECUDA;
        persist_attr_done[dev] = true;
      }
      const int sms = sm_count();
      cfg.gridDim = dim3(num_tiles < sms ? num_tiles : sms);
      cfg.dynamicSmemBytes = psmem;
      le = cudaLaunchKernelEx(&cfg, mlp_layer_ws_persist_kernel, p, map_x, map_w, map_y);
    } else {
      le = pair ? cudaLaunchKernelEx(&cfg, mlp_layer_ws_pair_kernel, p, map_x, map_w, map_y)
                : cudaLaunchKernelEx(&cfg, mlp_layer_ws_kernel, p, map_x, map_w, map_y);
    }
    if (le != cudaSuccess) {
      if (getenv("MMB_DEBUG")) fprintf(stderr, "mmb_mlp_layer: launch failed: %s (cluster %d, pair %d, smem %d)\n", cudaGetErrorString(le), cm, (int)pair, smem);
      (void)cudaGetLastError();
      return MMB_ECUDA;
    }
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}


namespace {
unsigned long long* chain_trace_buffer() {          // one buffer on the device that first asks (diagnostics only)
  static unsigned long long* buf = [] {
    unsigned long long* b = nullptr;
    const char* v = getenv("MMB_CHAIN_TRACE");
#ifndef MMB_CHAIN_TRACE_BUILD
    v = nullptr;
#endif
    if (v && atoi(v) && cudaMalloc(&b, CHAIN_TRACE_WORDS * sizeof(unsigned long long)) == cudaSuccess)
      cudaMemset(b, 0, CHAIN_TRACE_WORDS * sizeof(unsigned long long));
    else b = nullptr;
    return b;
  }();
  return buf;
}
}  // namespace

extern "C" int32_t mmb_mlp_chain_trace(uint64_t* out, int32_t words) {
  unsigned long long* b = chain_trace_buffer();
  if (!b || !out || words < CHAIN_TRACE_WORDS) return MMB_EINVAL;
  if (cudaDeviceSynchronize() != cudaSuccess) return MMB_ECUDA;
  return cudaMemcpy(out, b, CHAIN_TRACE_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

extern "C" int32_t mmb_mlp_chain(const mmb_mlp_layer_params* layers, int32_t num_layers, int32_t count, const float* const* x_fp32,
                                 void* stream) {
  if (!layers || num_layers < 2 || num_layers > MMB_MLP_MAX_LAYERS || count < 1) return MMB_EINVAL;
  if (count > CHAIN_MAX_NETS) return MMB_EUNSUPPORTED;
  static thread_local ChainArgs g;
  const mmb_mlp_layer_params& f = layers[0];
  if (f.M <= 0 || f.Mpad % BM || f.Mpad < f.M) return MMB_EINVAL;
  const bool tf32 = f.operand_type == 1;
  if (f.operand_type != 0 && f.operand_type != 1) return MMB_EINVAL;
  if (tf32 && x_fp32) return MMB_EINVAL;               // the fp32 input IS layer 0's x
  const int kbe = tf32 ? 32 : BK;                      // elements per k-block (128 bytes)
  for (int a = 0; a < CHAIN_MAX_NETS; ++a) g.x32[a] = nullptr;
  if (x_fp32) {     // fp32 [M][K] input of layer 0, cast inside the kernel: rows must be 16-byte addressable
    if (f.K % 4) return MMB_EUNSUPPORTED;
    for (int a = 0; a < count; ++a) {
      if (!x_fp32[a] || (reinterpret_cast<uintptr_t>(x_fp32[a]) & 15u)) return MMB_EUNSUPPORTED;
      g.x32[a] = x_fp32[a];
    }
  }
  g.K0 = f.K;
  for (int a = 0; a < count; ++a) {
    for (int l = 0; l < num_layers; ++l) {
      const mmb_mlp_layer_params& p = layers[a * num_layers + l];
      const mmb_mlp_layer_params& r = layers[l];                       // network 0 defines the geometry
      if ((!p.x && !(l == 0 && x_fp32)) || !p.w || !p.bias || !p.y || p.M != f.M || p.Mpad != f.Mpad || p.N <= 0 || p.K <= 0) return MMB_EINVAL;
      if (p.N != r.N || p.K != r.K || p.Kpad != r.Kpad || p.epilogue != r.epilogue || p.y_stride != r.y_stride) return MMB_EINVAL;
      if (p.operand_type != f.operand_type) return MMB_EINVAL;
      if (p.ln_in_stats || p.ln_out_stats) return MMB_EUNSUPPORTED;   // deferred LayerNorm: per-layer launches only
      if (tf32 ? (p.Kpad % 4 || p.Kpad < p.K) : (p.Kpad % BK || p.Kpad < p.K)) return MMB_EINVAL;
      if ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.w) | reinterpret_cast<uintptr_t>(p.y)) & 15u) return MMB_EALIGN;
      const bool last = l == num_layers - 1;
      ChainLayer& c = g.l[a][l];
      c.bias = p.bias; c.N = p.N; c.nkb = tf32 ? (p.K + kbe - 1) / kbe : p.Kpad / BK; c.epilogue = p.epilogue;
      if (!last) {
        // hidden layer: bias + ELU -> bf16, the next layer's operand; its four column slices must be whole 64-column sub-tiles
        if (p.epilogue != 1 || p.N % (CHAIN_CLUSTER * 64) || p.N > CHAIN_CLUSTER * 256) return MMB_EUNSUPPORTED;
        const mmb_mlp_layer_params& nx = layers[a * num_layers + l + 1];
        if (nx.x != p.y || nx.Kpad != p.y_stride || p.y_stride < p.N || p.y_stride % 8) return MMB_EINVAL;
        if (p.y_stride != p.N) return MMB_EUNSUPPORTED;   // the next layer's k-blocks are exactly this layer's 64-column sub-tiles
        c.n_tile = p.N / CHAIN_CLUSTER;
        if (tf32 ? !make_map_f32_2d(&c.map_y, p.y, (uint64_t)p.Mpad, (uint64_t)p.N, (uint64_t)p.y_stride, BM)
                 : !make_map_bf16_2d(&c.map_y, p.y, (uint64_t)p.Mpad, (uint64_t)p.y_stride, BM)) return MMB_ECUDA;
      } else {
        // last layer: bias -> fp32 [M][N], TMA-addressable; the cluster splits round_up(N, 128) columns
        if (p.epilogue != 0 || (p.y_stride & 3) || p.y_stride < p.N) return MMB_EUNSUPPORTED;
        const int npad = (p.N + 127) / 128 * 128;
        if (npad > CHAIN_CLUSTER * 256) return MMB_EUNSUPPORTED;
        c.n_tile = npad / CHAIN_CLUSTER;
        if (!make_map_f32_out(&c.map_y, p.y, (uint64_t)p.M, (uint64_t)p.N, (uint64_t)p.y_stride)) return MMB_ECUDA;
      }
      // weights [Npad rows >= N][Kpad]: slices beyond the allocated rows are zero-filled by the TMA (out-of-bounds box rows)
      if (p.Npad < p.N) return MMB_EINVAL;
      if (tf32) {
        // fp32 operands as they are: w = the nn.Linear weight [N][K] (pitch Kpad), x = the observations [M][K] (layer 0) or the
        // previous layer's fp32 output [Mpad][K]; the K tail and missing rows are zero-filled by the TMA
        if (!make_map_f32_2d(&c.map_w, p.w, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)p.Kpad, (uint32_t)c.n_tile)) return MMB_ECUDA;
        if (!make_map_f32_2d(&c.map_x, p.x, (uint64_t)(l == 0 ? p.M : p.Mpad), (uint64_t)p.K, (uint64_t)p.Kpad, BM)) return MMB_ECUDA;
      } else {
        if (!make_map_bf16_2d(&c.map_w, p.w, (uint64_t)p.Npad, (uint64_t)p.Kpad, (uint32_t)c.n_tile)) return MMB_ECUDA;
        if (l == 0 && x_fp32) c.map_x = c.map_w;              // unused: layer 0's A tiles are cast in the kernel
        else if (!make_map_bf16_2d(&c.map_x, p.x, (uint64_t)p.Mpad, (uint64_t)p.Kpad, BM)) return MMB_ECUDA;
      }
    }
  }
  g.num_layers = num_layers; g.M = f.M; g.overlap_prev = f.overlap_prev;
  g.trace = chain_trace_buffer();
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= MMB_MAX_DEVICES) return MMB_EUNSUPPORTED;
  static bool attr_done[MMB_MAX_DEVICES] = {};
  if (!attr_done[dev]) {
    if (cudaFuncSetAttribute(mlp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHAIN_SMEM) != cudaSuccess) return MMB_ECUDA;
    if (cudaFuncSetAttribute(mlp_chain_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHAIN_SMEM) != cudaSuccess) return MMB_ECUDA;
    if (cudaFuncSetAttribute(mlp_chain_duo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHAIN_SMEM) != cudaSuccess) return MMB_ECUDA;
    attr_done[dev] = true;
  }
  for (int a = 0; a < count; ++a)
    for (int l = 0; l < num_layers; ++l) g.l[a][l].kgroup = 1;
  // two networks on the same fp32 rows (PPO act(): actor + critic): one CTA walks both, alternating (mlp_chain_duo_kernel)
  static const bool duo_on = [] { const char* v = getenv("MMB_MLP_DUO"); return !(v && v[0] == '0'); }();
  // (when the side-by-side launch fits one wave of clusters - M <= 2304 on 148 SMs - it is the faster one: twice the CTAs at work)
  static const bool duo_force = [] { const char* v = getenv("MMB_MLP_DUO"); return v && v[0] == '2'; }();
  const bool duo = duo_on && !tf32 && count == 2 && x_fp32 && x_fp32[0] == x_fp32[1] &&
                   (duo_force || 2 * (f.Mpad / BM) > sm_count() / CHAIN_CLUSTER);
  if (duo) {
    // layers >= 1: k-blocks in groups of two per tensor load (3-D maps) where the own / peer split and K allow it
    static const int kgroup_max = [] { const char* v = getenv("MMB_MLP_KGROUP"); return v ? atoi(v) : 2; }();
    for (int l = 1; l < num_layers; ++l) {
      const int n_own = g.l[0][l - 1].n_tile / 64, nkb = g.l[0][l].nkb;
      // (256-column slices keep single k-blocks: a 96 KB group leaves two stages, and the ring's round trip then shows)
      const int G = (kgroup_max >= 2 && n_own % 2 == 0 && nkb % 2 == 0 && (g.l[0][l].n_tile <= 128 || kgroup_max >= 3)) ? 2 : 1;
      for (int a = 0; a < count && G > 1; ++a) {
        const mmb_mlp_layer_params& p = layers[a * num_layers + l];
        ChainLayer& c = g.l[a][l];
        c.kgroup = G;
        if (!make_map_bf16_kgroup(&c.map_w, p.w, (uint64_t)p.Npad, (uint64_t)p.Kpad, (uint32_t)c.n_tile, (uint32_t)G)) return MMB_ECUDA;
        if (!make_map_bf16_kgroup(&c.map_x, p.x, (uint64_t)p.Mpad, (uint64_t)p.Kpad, BM, (uint32_t)G)) return MMB_ECUDA;
      }
    }
  }
  {
    LaunchScope ls(K_MLP_LAYER, (cudaStream_t)stream);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CHAIN_CLUSTER * (f.Mpad / BM), 1, duo ? 1 : count);
    cfg.blockDim = dim3(duo ? DUO_THREADS : WS_THREADS);
    cfg.dynamicSmemBytes = CHAIN_SMEM;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CHAIN_CLUSTER; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
    if (f.overlap_prev) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    const cudaError_t le = duo ? cudaLaunchKernelEx(&cfg, mlp_chain_duo_kernel, g)
                               : (tf32 ? cudaLaunchKernelEx(&cfg, mlp_chain_tf32_kernel, g) : cudaLaunchKernelEx(&cfg, mlp_chain_kernel, g));
    if (le != cudaSuccess) {
      if (getenv("MMB_DEBUG")) fprintf(stderr, "mmb_mlp_chain: launch failed: %s\n", cudaGetErrorString(le));
      (void)cudaGetLastError();
      return MMB_ECUDA;
    }
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}

extern "C" int32_t mmb_ln_cast(const float* x, int32_t M, int32_t Mpad, int32_t K, int32_t Kpad, const float* gamma,
                               const float* beta, float eps, int32_t use_ln, void* y_bf16, void* stream) {
  if (!x || !y_bf16 || M <= 0 || Mpad < M || K <= 0 || Kpad < K) return MMB_EINVAL;
  if (use_ln && (!gamma || !beta)) return MMB_EINVAL;
  {
    LaunchScope ls(K_LN_CAST, (cudaStream_t)stream);
    if (!use_ln && Kpad % 4 == 0) {
      int64_t bx = ((int64_t)Mpad * (Kpad / 4) + 255) / 256;
      if (bx > sm_count() * 8) bx = sm_count() * 8;
      cast_pad_kernel<<<(unsigned)bx, 256, 0, (cudaStream_t)stream>>>(x, M, Mpad, K, Kpad, static_cast<__nv_bfloat16*>(y_bf16));
    } else {
      ln_cast_kernel<<<(Mpad * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, M, Mpad, K, Kpad, gamma, beta, eps, use_ln,
                                                                              static_cast<__nv_bfloat16*>(y_bf16));
    }
  }
  return cudaGetLastError() == cudaSuccess ? MMB_OK : MMB_ECUDA;
}
