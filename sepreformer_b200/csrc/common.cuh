// Shared device helpers for the sepref kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sepref {

constexpr float kLnEps = 1e-5f;   // torch.nn.LayerNorm default (reference network.py:50,81,133,162)
constexpr float kBnEps = 1e-5f;   // torch.nn.BatchNorm1d default (network.py:167, module.py:69)
constexpr float kGnEps = 1e-8f;   // module.py:117

__device__ __forceinline__ float gelu_erf(float x) {            // GELU(approximate='none')
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
// sigmoid through one MUFU.TANH: 0.5*tanh(0.5x)+0.5 (abs err ~1e-3 rel on tanh -> ~5e-4 abs; used on the TC path only)
__device__ __forceinline__ float sigmoid_fast(float x) {
  float e = __expf(-x);
  return __fdividef(1.0f, 1.0f + e);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t f32_to_tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float tf32_round(float x) { return __uint_as_float(f32_to_tf32_rna(x)); }

}  // namespace sepref
