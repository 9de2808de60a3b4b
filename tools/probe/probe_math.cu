// Probe: libdevice transcendental results for comparison against torch CUDA eager ops.
#include <cuda_runtime.h>
#include <stdint.h>
extern "C" __global__ void k_unary(const float* x, float* o_sin, float* o_cos, float* o_atan, float* o_asin, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { o_sin[i] = sinf(x[i]); o_cos[i] = cosf(x[i]); o_atan[i] = atanf(x[i]); o_asin[i] = asinf(x[i]); }
}
extern "C" __global__ void k_binary(const float* a, const float* b, float* o_atan2, float* o_fmod, float* o_div, float* o_sqrt, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { o_atan2[i] = atan2f(a[i], b[i]); o_fmod[i] = fmodf(a[i], b[i]); o_div[i] = __fdiv_rn(a[i], b[i]); o_sqrt[i] = __fsqrt_rn(fabsf(a[i])); }
}
extern "C" int run_unary(const float* x, float* s, float* c, float* at, float* as, int n, void* stream) {
  k_unary<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, s, c, at, as, n); return (int)cudaGetLastError();
}
extern "C" int run_binary(const float* a, const float* b, float* o1, float* o2, float* o3, float* o4, int n, void* stream) {
  k_binary<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a, b, o1, o2, o3, o4, n); return (int)cudaGetLastError();
}
