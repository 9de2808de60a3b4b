// tma_mma_rate.cu - what paces a TMA-fed tcgen05 k-loop on one SM?  One CTA per SM, warp 0 = TMA producer (one thread),
// warp 1 = MMA issuer (one thread), a full / empty mbarrier ring between them - the skeleton of mlp_chain_kernel's k-loop
// with everything else removed (no epilogue, no layer boundaries).  Swept: slice width (weight rows per k-block), ring
// depth, k-blocks per tensor load (3-D maps), with / without the MMAs, with / without the A or the W load.
// Prints one JSON line per case: ns per k-block (in-kernel %globaltimer, median CTA) and the implied GB/s into an SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe/tma_mma_rate tools/probe/tma_mma_rate.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int BM = 128, BK = 64, A_BYTES = BM * BK * 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ int g_abort;          // a wait that expired (~50 ms): every later wait returns at once, the host reports it
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!done && (*(volatile int*)&g_abort || clock64() - t0 > 100000000ll)) { g_abort = 1; return; }
  }
}
__device__ __forceinline__ void tma3(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
               "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t idesc_bf16(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24); }
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, bool acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc),
               "r"((uint32_t)acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// accumulate = 1 without a predicate computed per instruction
__device__ __forceinline__ void umma_acc(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mbar_wait_plain(uint64_t* bar, uint32_t parity) {   // the tight PTX loop (no clock, no flag)
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// The MMA issuer's loop with nothing in it that is not needed: stages unrolled (compile-time stage index: barrier addresses
// and the stage's two base descriptors are loop-invariant registers), one phase bit per round, accumulate flag constant
// after the first instruction, k offsets added to the descriptors' low words.
template <int S, int G>
__device__ __forceinline__ void mma_loop_lean(uint8_t* smem, uint64_t* full, uint64_t* empty, uint32_t tmem, int n_tile, int groups, bool do_mma) {
  const int stage_bytes = G * (A_BYTES + n_tile * BK * 2);
  const uint32_t idesc = idesc_bf16(n_tile);
  const uint32_t wstep = (uint32_t)(n_tile * BK * 2) >> 4;
  uint64_t da[S], db[S];
  uint32_t fb[S], eb[S];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    da[s] = desc_sw128(smem_u32(smem + s * stage_bytes));
    db[s] = desc_sw128(smem_u32(smem + s * stage_bytes + G * A_BYTES));
    fb[s] = smem_u32(&full[s]); eb[s] = smem_u32(&empty[s]);
  }
  uint32_t phase = 0;
  bool first = true;
  for (int i = 0; i < groups; i += S) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(fb[s]), "r"(phase) : "memory");
      if (do_mma) {
#pragma unroll
        for (int q = 0; q < G; ++q)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t a = da[s] + (uint64_t)(q * (A_BYTES >> 4) + j * 2), b = db[s] + (uint64_t)(q * wstep + j * 2);
            if (first) { umma(tmem, a, b, idesc, false); first = false; }
            else umma_acc(tmem, a, b, idesc);
          }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(eb[s]) : "memory");
    }
    phase ^= 1u;
  }
}

struct Args {
  CUtensorMap map_a, map_w;     // 3-D: (64, rows, K / 64), box (64, 128 | n_tile, G)
  int n_tile, stages, G, nkb, do_mma, do_a, do_w, fence, lean;   // fence: tcgen05.fence::after_thread_sync per k-block; commit_every: stages freed per tcgen05.commit
  unsigned long long* times;    // [grid][2]
};

__global__ void __launch_bounds__(128, 1) rate_kernel(const __grid_constant__ Args g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[16], empty[16], done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const int stage_bytes = g.G * (A_BYTES + g.n_tile * BK * 2);
  if (threadIdx.x == 0) {
    for (int i = 0; i < g.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int groups = g.nkb / g.G;
  unsigned long long t0 = 0, t1 = 0;
  if (warp == 0 && (threadIdx.x & 31) == 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    g.times[blockIdx.x * 2] = t0;
    int s = 0;
    uint32_t used = 0, par = 0;
    const uint32_t bytes = (uint32_t)(g.G * ((g.do_a ? A_BYTES : 0) + (g.do_w ? g.n_tile * BK * 2 : 0)));
    for (int i = 0; i < groups; ++i) {
      const uint32_t bit = 1u << s;
      if (used & bit) { mbar_wait(&empty[s], (par >> s) & 1u); par ^= bit; }
      used |= bit;
      if (bytes) {
        mbar_expect_tx(&full[s], bytes);
        if (g.do_a) tma3(smem + s * stage_bytes, &g.map_a, 0, (blockIdx.x / 4) * BM, i * g.G, &full[s]);   // four CTAs share a row block, as in the chain
        if (g.do_w) tma3(smem + s * stage_bytes + g.G * A_BYTES, &g.map_w, 0, (blockIdx.x % 4) * g.n_tile, i * g.G, &full[s]);
      } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[s])) : "memory");
      }
      s = (s + 1 == g.stages) ? 0 : s + 1;
    }
  } else if (warp == 1 && (threadIdx.x & 31) == 0) {
    int s = 0;
    uint32_t par = 0;
    const uint32_t idesc = idesc_bf16(g.n_tile);
    if (g.lean) {
      if (g.stages == 4 && g.G == 1) mma_loop_lean<4, 1>(smem, full, empty, tmem, g.n_tile, groups, g.do_mma != 0);
      else if (g.stages == 8 && g.G == 1) mma_loop_lean<8, 1>(smem, full, empty, tmem, g.n_tile, groups, g.do_mma != 0);
      else if (g.stages == 4 && g.G == 2) mma_loop_lean<4, 2>(smem, full, empty, tmem, g.n_tile, groups, g.do_mma != 0);
      else if (g.stages == 2 && g.G == 2) mma_loop_lean<2, 2>(smem, full, empty, tmem, g.n_tile, groups, g.do_mma != 0);
    } else
    for (int i = 0; i < groups; ++i) {
      mbar_wait(&full[s], (par >> s) & 1u);
      par ^= 1u << s;
      if (g.fence) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_addr = smem_u32(smem + s * stage_bytes), b_addr = a_addr + g.G * A_BYTES;
      if (g.do_mma)
        for (int q = 0; q < g.G; ++q)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            umma(tmem, desc_sw128(a_addr + q * A_BYTES + j * 32), desc_sw128(b_addr + q * g.n_tile * BK * 2 + j * 32), idesc, i > 0 || q > 0 || j > 0);
      commit(&empty[s]);
      s = (s + 1 == g.stages) ? 0 : s + 1;
    }
    commit(&done);
    mbar_wait(&done, 0);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    g.times[blockIdx.x * 2 + 1] = t1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn encode_fn() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
  return (EncodeFn)fn;
}
static bool make_map(CUtensorMap* m, void* base, uint64_t rows, uint64_t K, uint32_t box_rows, uint32_t G) {
  cuuint64_t dims[3] = {64, rows, K / 64};
  cuuint64_t strides[2] = {K * 2, 128};
  cuuint32_t box[3] = {64, box_rows, G};
  cuuint32_t es[3] = {1, 1, 1};
  return encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 128;
  const int K = 4096, nkb = K / BK;      // 64 k-blocks per launch; A: (grid / 4) row blocks x 128 rows x 8 KB = 32 MB at grid 128: L2-resident, like the chain's activations
  __nv_bfloat16 *a = nullptr, *w = nullptr;
  unsigned long long* times = nullptr;
  cudaMalloc(&a, (size_t)grid * BM * K * 2);
  cudaMalloc(&w, (size_t)1024 * K * 2);
  cudaMalloc(&times, grid * 2 * sizeof(unsigned long long));
  cudaMemset(a, 0, (size_t)grid * BM * K * 2);
  cudaMemset(w, 0, (size_t)1024 * K * 2);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Case { int n_tile, stages, G, mma, da, dw; int fence = 1, lean = 0; };
  std::vector<Case> cases;
  for (int nt : {256, 128, 32})
    for (int st : {2, 4, 8})
      for (int G : {1, 2})
        for (int mma : {1, 0}) {
          if (G * (A_BYTES + nt * 128) * st > 192 * 1024) continue;
          cases.push_back({nt, st, G, mma, 1, 1});
        }
  cases.push_back({256, 4, 1, 1, 0, 1});   // weights only
  cases.push_back({256, 4, 1, 1, 1, 0});   // activations only
  cases.push_back({256, 4, 1, 1, 0, 0});   // no loads at all: the MMA issue / commit loop alone
  cases.push_back({128, 4, 1, 1, 0, 0});
  cases.push_back({32, 4, 1, 1, 0, 0});
  cases.push_back({256, 4, 1, 0, 0, 0});   // neither: the barrier ring alone
  cases.push_back({256, 4, 1, 0, 0, 0, 0, 0});   // ... without the per-k-block tcgen05 fence
  cases.push_back({256, 4, 1, 0, 0, 0, 0, 1});   // lean MMA-issuer loop: the barrier ring alone
  cases.push_back({256, 4, 1, 1, 0, 0, 0, 1});   // lean, MMAs, no loads
  cases.push_back({32, 4, 1, 1, 0, 0, 0, 1});
  cases.push_back({256, 4, 1, 1, 1, 1, 0, 1});   // lean, the full loop
  cases.push_back({128, 4, 1, 1, 1, 1, 0, 1});
  cases.push_back({32, 8, 1, 1, 1, 1, 0, 1});
  cases.push_back({32, 4, 2, 1, 1, 1, 0, 1});
  cases.push_back({128, 2, 2, 1, 1, 1, 0, 1});
  cases.push_back({256, 2, 2, 1, 1, 1, 0, 1});
  for (const Case& c : cases) {
    Args g;
    if (!make_map(&g.map_a, a, (uint64_t)grid * BM, K, BM, c.G) || !make_map(&g.map_w, w, 1024, K, c.n_tile, c.G)) { printf("map failed\n"); return 1; }
    g.n_tile = c.n_tile; g.stages = c.stages; g.G = c.G; g.nkb = nkb; g.do_mma = c.mma; g.do_a = c.da; g.do_w = c.dw; g.times = times;
    g.fence = c.fence; g.lean = c.lean;
    const int smem = c.stages * c.G * (A_BYTES + c.n_tile * 128);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int r = 0; r < 3; ++r) rate_kernel<<<grid, 128, smem>>>(g);
    cudaEventRecord(e0);
    const int R = 10;
    for (int r = 0; r < R; ++r) rate_kernel<<<grid, 128, smem>>>(g);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    int ab = 0;
    cudaMemcpyFromSymbol(&ab, g_abort, sizeof(int));
    if (ab) { printf("{\"error\": \"a bounded wait expired\", \"n_tile\": %d, \"stages\": %d, \"kgroup\": %d}\n", c.n_tile, c.stages, c.G); return 1; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<unsigned long long> h(grid * 2);
    cudaMemcpy(h.data(), times, grid * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    std::vector<double> per;
    for (int b = 0; b < grid; ++b) per.push_back((double)(h[2 * b + 1] - h[2 * b]) / nkb);
    std::sort(per.begin(), per.end());
    const double bytes_kb = (c.da ? A_BYTES : 0) + (c.dw ? c.n_tile * 128 : 0);
    printf("{\"grid\": %d, \"n_tile\": %d, \"stages\": %d, \"kgroup\": %d, \"mma\": %d, \"load_a\": %d, \"load_w\": %d, \"fence\": %d, \"lean\": %d, \"ns_per_kblock_median\": %.1f, "
           "\"ns_per_kblock_max\": %.1f, \"kernel_us\": %.2f, \"gbs_into_sm\": %.1f, \"mma_ns_per_kblock_at_peak\": %.1f}\n",
           grid, c.n_tile, c.stages, c.G, c.mma, c.da, c.dw, c.fence, c.lean, per[grid / 2], per[grid - 1], ms / R * 1e3, bytes_kb / per[grid / 2],
           2.0 * BM * c.n_tile * BK / 8192.0 / 1.965);
    fflush(stdout);
  }
  return 0;
}
