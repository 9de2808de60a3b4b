// mmb_math.cuh - fp32 arithmetic that replays the reference's torch rounding sequence op for op.
//
// The reference's task functions are chains of separate torch ops, i.e. every add/sub/mul/div/sqrt
// is individually IEEE-rounded (no FMA contraction across ops).  The rewards are ill-conditioned in
// fp32 (SURVEY.md finding 11: 500*(l2_before - l2_now) amplifies one-ulp differences), so the
// kernels use the round-to-nearest intrinsics below in the reference's association order and plain
// libdevice transcendentals (never --use_fast_math).  Measured on B200 (profiles/r01_probe_torch_cuda.json):
// libdevice sinf/cosf/atanf/atan2f/fmodf and div.rn/sqrt.rn equal torch's CUDA kernels bit for bit.
//
// Where torch's CUDA and CPU kernels associate differently the `FLAVOR` template argument picks
// (MMB_FLAVOR_CUDA / MMB_FLAVOR_CPU, measured in the same probe):
//   sum(-1) over 3:  CUDA (a0+a2)+a1          CPU (a0+a1)+a2
//   sum(-1) over 8:  CUDA ((a0+a4)+(a2+a6))+((a1+a5)+(a3+a7))   CPU sequential
//   x / python_scalar: CUDA x * (1/s)          CPU x / s
// Same on both devices: torch.cross = fma(a1,b2,-(a2*b1)); 1x3.3x1 bmm = (a0b0+a1b1)+a2b2 unfused;
// norm over (x,y,0) = sqrt(x*x+y*y); remainder = fmod then +b on sign mismatch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mmb {

constexpr int FLAVOR_CUDA = 0;
constexpr int FLAVOR_CPU = 1;

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }

struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };  // quaternion (x,y,z,w)

// torch.clamp(x, lo, hi) = min(max(x, lo), hi) with NaN propagation: two FMNMX.NAN instead of compare+select pairs
__device__ __forceinline__ float clampf(float x, float lo, float hi) {
  float r;
  asm("{\n\t.reg .f32 t;\n\tmax.NaN.f32 t, %1, %2;\n\tmin.NaN.f32 %0, t, %3;\n\t}" : "=f"(r) : "f"(x), "f"(lo), "f"(hi));
  return r;
}

template <int FLAVOR>
__device__ __forceinline__ float sum3(float a0, float a1, float a2) {
  if (FLAVOR == FLAVOR_CUDA) return fadd(fadd(a0, a2), a1);
  return fadd(fadd(a0, a1), a2);
}

template <int FLAVOR>
__device__ __forceinline__ float sum8(const float* e) {
  if (FLAVOR == FLAVOR_CUDA) {
    float t0 = fadd(e[0], e[4]), t1 = fadd(e[1], e[5]), t2 = fadd(e[2], e[6]), t3 = fadd(e[3], e[7]);
    return fadd(fadd(t0, t2), fadd(t1, t3));
  }
  float s = e[0];
#pragma unroll
  for (int j = 1; j < 8; ++j) s = fadd(s, e[j]);
  return s;
}

// isaacgym.torch_utils.quat_mul (oracle/isaac_torch_utils.py)
__device__ __forceinline__ f4 quat_mul(f4 a, f4 b) {
  float ww = fmul(fadd(a.z, a.x), fadd(b.x, b.y));
  float yy = fmul(fsub(a.w, a.y), fadd(b.w, b.z));
  float zz = fmul(fadd(a.w, a.y), fsub(b.w, b.z));
  float xx = fadd(fadd(ww, yy), zz);
  float qq = fmul(0.5f, fadd(xx, fmul(fsub(a.z, a.x), fsub(b.x, b.y))));
  f4 r;
  r.w = fadd(fsub(qq, ww), fmul(fsub(a.z, a.y), fsub(b.y, b.z)));
  r.x = fadd(fsub(qq, xx), fmul(fadd(a.x, a.w), fadd(b.x, b.w)));
  r.y = fadd(fsub(qq, yy), fmul(fsub(a.w, a.x), fadd(b.y, b.z)));
  r.z = fadd(fsub(qq, zz), fmul(fadd(a.z, a.y), fsub(b.w, b.x)));
  return r;
}

// torch.cross(a, b, dim=-1): each component is fma(a_i, b_j, -(a_j * b_i)) on CUDA and CPU
__device__ __forceinline__ f3 cross3(f3 a, f3 b) {
  f3 r;
  r.x = __fmaf_rn(a.y, b.z, -fmul(a.z, b.y));
  r.y = __fmaf_rn(a.z, b.x, -fmul(a.x, b.z));
  r.z = __fmaf_rn(a.x, b.y, -fmul(a.y, b.x));
  return r;
}

// torch.bmm of [1x3]x[3x1]
__device__ __forceinline__ float dot3(f3 a, f3 b) { return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z)); }

// quat_rotate(q, v) = a + b + c   /  quat_rotate_inverse = a - b + c
template <bool INVERSE>
__device__ __forceinline__ f3 quat_rot(f4 q, f3 v) {
  float s = fsub(fmul(2.0f, fmul(q.w, q.w)), 1.0f);
  f3 qv = {q.x, q.y, q.z};
  f3 cr = cross3(qv, v);
  float d = dot3(qv, v);
  f3 r;
  {
    float a = fmul(v.x, s), b = fmul(fmul(cr.x, q.w), 2.0f), c = fmul(fmul(qv.x, d), 2.0f);
    r.x = fadd(INVERSE ? fsub(a, b) : fadd(a, b), c);
  }
  {
    float a = fmul(v.y, s), b = fmul(fmul(cr.y, q.w), 2.0f), c = fmul(fmul(qv.y, d), 2.0f);
    r.y = fadd(INVERSE ? fsub(a, b) : fadd(a, b), c);
  }
  {
    float a = fmul(v.z, s), b = fmul(fmul(cr.z, q.w), 2.0f), c = fmul(fmul(qv.z, d), 2.0f);
    r.z = fadd(INVERSE ? fsub(a, b) : fadd(a, b), c);
  }
  return r;
}

// torch.remainder(a, b) for float: fmod, then + b when non-zero and of opposite sign
__device__ __forceinline__ float remainderf_torch(float a, float b) {
  float m = fmodf(a, b);
  if (m != 0.0f && ((b < 0.0f) != (m < 0.0f))) m = fadd(m, b);
  return m;
}

#define MMB_TWO_PI_F 6.28318530717958647692f  // float(2*np.pi): Tensor % python_float casts the scalar to fp32

// remainder(a, 2*pi) for a = atan2f(..) in [-pi, pi]: |a| < 2*pi, so fmod(a, 2*pi) == a exactly and the
// torch rule reduces to "add 2*pi when negative" (NaN stays NaN: both comparisons are false).
__device__ __forceinline__ float wrap_angle_2pi(float a) {
  return (a != 0.0f && a < 0.0f) ? fadd(a, MMB_TWO_PI_F) : a;
}

// roll and yaw of get_euler_xyz (pitch is computed by the reference but never used on the path)
__device__ __forceinline__ void euler_roll_yaw(f4 q, float& roll, float& yaw) {
  float ww = fmul(q.w, q.w), xx = fmul(q.x, q.x), yy = fmul(q.y, q.y), zz = fmul(q.z, q.z);
  float sinr = fmul(2.0f, fadd(fmul(q.w, q.x), fmul(q.y, q.z)));
  float cosr = fadd(fsub(fsub(ww, xx), yy), zz);
  roll = wrap_angle_2pi(atan2f(sinr, cosr));
  float siny = fmul(2.0f, fadd(fmul(q.w, q.z), fmul(q.x, q.y)));
  float cosy = fsub(fsub(fadd(ww, xx), yy), zz);
  yaw = wrap_angle_2pi(atan2f(siny, cosy));
}

// l2_dist (ten_ant.py:975-985): sqrt((a-b)_x^2 + (a-b)_y^2)
__device__ __forceinline__ float l2_dist2(float ax, float ay, float bx, float by) {
  float c1 = fsub(ax, bx), c2 = fsub(ay, by);
  return fsqrt(fadd(fmul(c1, c1), fmul(c2, c2)));
}

// compute_box_quat + compute_box_quat_dist (ten_ant.py:951-973)
__device__ __forceinline__ float box_quat_dist(f4 q, float xg, float yg, float zg) {
  float x = fmul(2.0f, fadd(fmul(q.x, q.y), fmul(q.w, q.z)));
  float y = fsub(1.0f, fmul(2.0f, fadd(fmul(q.x, q.x), fmul(q.z, q.z))));
  float z = fmul(2.0f, fsub(fmul(q.y, q.z), fmul(q.w, q.x)));
  float num = fadd(fadd(fmul(x, xg), fmul(y, yg)), fmul(z, zg));
  float den = fsqrt(fadd(fadd(fmul(x, x), fmul(y, y)), fmul(z, z)));
  // second divisor is a Python float: sqrt(xg^2+yg^2+zg^2) evaluated in double, then cast
  float gn = (float)sqrt((double)xg * xg + (double)yg * yg + (double)zg * zg);
  return fdiv(fdiv(num, den), gn);
}

// The 38-wide per-ant observation core shared by TenAnt (ten_ant.py:1304-1350) and OneAnt
// (one_ant.py:563-618): everything except the position prefix / sensors / actions.
struct AntCore {
  f3 vel_loc, angvel_loc;
  float yaw, roll, angle_to_target, up_proj, heading_proj;
  f3 up_vec, heading_vec;
};

template <int FLAVOR>
__device__ __forceinline__ AntCore ant_core(f3 p, f4 q, f3 v, f3 w, f4 inv_start_rot) {
  AntCore o;
  // to_target = targets(0,0,0) - torso_position ; z := 0
  f3 tt = {fsub(0.0f, p.x), fsub(0.0f, p.y), 0.0f};
  // normalize: x / clamp(norm, 1e-9)
  float nrm = fsqrt(fadd(fmul(tt.x, tt.x), fmul(tt.y, tt.y)));
  if (FLAVOR == FLAVOR_CPU) nrm = fsqrt(__fmaf_rn(tt.y, tt.y, fmul(tt.x, tt.x)));  // CPU vectorised norm (99.3% of rows)
  nrm = nrm < 1e-9f ? 1e-9f : nrm;
  // dir.z = 0 / nrm = 0, so the z term of the heading dot product is (+-)0 and drops out
  const float dirx = fdiv(tt.x, nrm), diry = fdiv(tt.y, nrm);
  f4 tq = quat_mul(q, inv_start_rot);
  // quat_rotate(tq, e_z) and quat_rotate(tq, e_x) with the products by the basis vector's 0s and 1s folded:
  // x*1 = x and y + (+-0) = y are exact, so for finite inputs every surviving operation is the reference's
  // (only the sign of an exact-zero result can differ).
  {
    const float s = fsub(fmul(2.0f, fmul(tq.w, tq.w)), 1.0f);
    o.up_vec.x = fadd(fmul(fmul(tq.y, tq.w), 2.0f), fmul(fmul(tq.x, tq.z), 2.0f));
    o.up_vec.y = fadd(-fmul(fmul(tq.x, tq.w), 2.0f), fmul(fmul(tq.y, tq.z), 2.0f));
    o.up_vec.z = fadd(s, fmul(fmul(tq.z, tq.z), 2.0f));
    o.heading_vec.x = fadd(s, fmul(fmul(tq.x, tq.x), 2.0f));
    o.heading_vec.y = fadd(fmul(fmul(tq.z, tq.w), 2.0f), fmul(fmul(tq.y, tq.x), 2.0f));
    o.heading_vec.z = fadd(-fmul(fmul(tq.y, tq.w), 2.0f), fmul(fmul(tq.z, tq.x), 2.0f));
  }
  o.up_proj = o.up_vec.z;
  o.heading_proj = fadd(fmul(o.heading_vec.x, dirx), fmul(o.heading_vec.y, diry));
  o.vel_loc = quat_rot<true>(tq, v);
  o.angvel_loc = quat_rot<true>(tq, w);
  euler_roll_yaw(tq, o.roll, o.yaw);
  float walk = atan2f(fsub(0.0f, p.z), fsub(0.0f, p.x));
  o.angle_to_target = fsub(walk, o.yaw);
  return o;
}

// ant_core without the roll angle, and the roll angle on its own (recomputes tq = q * inv_start_rot, 28 flops), so
// that two threads of different warps can share one ant: identical operations in identical order, identical bits.
struct AntCoreNoRoll {
  f3 vel_loc, angvel_loc;
  float yaw, angle_to_target, up_proj, heading_proj;
};
template <int FLAVOR>
__device__ __forceinline__ AntCoreNoRoll ant_core_no_roll(f3 p, f4 q, f3 v, f3 w, f4 inv_start_rot) {
  AntCoreNoRoll o;
  f3 tt = {fsub(0.0f, p.x), fsub(0.0f, p.y), 0.0f};
  float nrm = fsqrt(fadd(fmul(tt.x, tt.x), fmul(tt.y, tt.y)));
  if (FLAVOR == FLAVOR_CPU) nrm = fsqrt(__fmaf_rn(tt.y, tt.y, fmul(tt.x, tt.x)));
  nrm = nrm < 1e-9f ? 1e-9f : nrm;
  const float dirx = fdiv(tt.x, nrm), diry = fdiv(tt.y, nrm);
  f4 tq = quat_mul(q, inv_start_rot);
  const float s = fsub(fmul(2.0f, fmul(tq.w, tq.w)), 1.0f);
  o.up_proj = fadd(s, fmul(fmul(tq.z, tq.z), 2.0f));
  const float hx = fadd(s, fmul(fmul(tq.x, tq.x), 2.0f));
  const float hy = fadd(fmul(fmul(tq.z, tq.w), 2.0f), fmul(fmul(tq.y, tq.x), 2.0f));
  o.heading_proj = fadd(fmul(hx, dirx), fmul(hy, diry));
  o.vel_loc = quat_rot<true>(tq, v);
  o.angvel_loc = quat_rot<true>(tq, w);
  {  // yaw of get_euler_xyz
    float ww = fmul(tq.w, tq.w), xx = fmul(tq.x, tq.x), yy = fmul(tq.y, tq.y), zz = fmul(tq.z, tq.z);
    float siny = fmul(2.0f, fadd(fmul(tq.w, tq.z), fmul(tq.x, tq.y)));
    float cosy = fsub(fsub(fadd(ww, xx), yy), zz);
    o.yaw = wrap_angle_2pi(atan2f(siny, cosy));
  }
  float walk = atan2f(fsub(0.0f, p.z), fsub(0.0f, p.x));
  o.angle_to_target = fsub(walk, o.yaw);
  return o;
}
__device__ __forceinline__ float ant_roll(f4 q, f4 inv_start_rot) {
  f4 tq = quat_mul(q, inv_start_rot);
  float ww = fmul(tq.w, tq.w), xx = fmul(tq.x, tq.x), yy = fmul(tq.y, tq.y), zz = fmul(tq.z, tq.z);
  float sinr = fmul(2.0f, fadd(fmul(tq.w, tq.x), fmul(tq.y, tq.z)));
  float cosr = fadd(fsub(fsub(ww, xx), yy), zz);
  return wrap_angle_2pi(atan2f(sinr, cosr));
}

// unscale(x, lower, upper) = (2x - upper - lower) / (upper - lower)
__device__ __forceinline__ float unscale(float x, float lo, float hi) {
  return fdiv(fsub(fsub(fmul(2.0f, x), hi), lo), fsub(hi, lo));
}

// ---- 128-bit global access helpers ---------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stg4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// cooperative copy global -> shared of n floats (n need not be a multiple of 4)
__device__ __forceinline__ void tile_load(float* __restrict__ s, const float* __restrict__ g, int n, int tid, int nthreads) {
  if (aligned16(g)) {
    int n4 = n >> 2;
    for (int i = tid; i < n4; i += nthreads) reinterpret_cast<float4*>(s)[i] = ldg4(g + 4 * i);
    for (int i = (n4 << 2) + tid; i < n; i += nthreads) s[i] = __ldg(g + i);
  } else {
    for (int i = tid; i < n; i += nthreads) s[i] = __ldg(g + i);
  }
}


// ---- 1-D TMA (cp.async.bulk) + mbarrier helpers: contiguous tiles move global <-> shared without touching the
// LSU instruction stream (SASS: UBLKCP / SYNCS) ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// generic-proxy writes to shared memory must be fenced before the async proxy (TMA) reads them
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// Philox4x32-10 (Salmon et al. 2011), the counter-based generator torch's CUDA RNG also uses; keyed by
// (seed), counter = (env, step, block, 0) so that an env-sharded multi-GPU run draws the same numbers as
// the single-GPU run on the concatenated envs.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// programmatic dependent launch (PTX griddepcontrol): let the next kernel in the stream start early / wait until the
// previous kernel in the stream has completed and its writes are visible
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// L2 prefetch of a contiguous global range by the TMA engine (no destination, no registers, no smem): 16 B aligned
// address, size a multiple of 16 B.
__device__ __forceinline__ void tma_prefetch_l2(const void* src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
// prefetch [p, p + bytes) into L2 when the range is 16-byte aligned and sized (no-op otherwise)
__device__ __forceinline__ void prefetch_range_l2(const void* p, int64_t bytes) {
  if (((reinterpret_cast<uintptr_t>(p) | (uintptr_t)bytes) & 15u) == 0 && bytes > 0) tma_prefetch_l2(p, (uint32_t)bytes);
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

}  // namespace mmb
