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
// GELU with erf from Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, far below the 11-bit operand rounding that
// follows on the tensor-core paths): two MUFU ops and six FMAs instead of erff's ~25 instructions.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-z * z);          // erf(|x| / sqrt(2))
  return 0.5f * x + 0.5f * fabsf(x) * e;                  // 0.5 x (1 + sign(x) e)
}
// GELU through one MUFU.TANH: erf(x/sqrt2) = tanh(x * P(x^2)) with a fitted cubic P (|GELU error| < 3e-5 from the fit,
// plus tanh.approx's 2^-11 relative error - the size of the FP16/TF32 operand rounding that follows it on the
// tensor-core paths, the only place it is used): 7 instructions instead of 18.
__device__ __forceinline__ float gelu_tanh_fit(float x) {
  const float t = x * x;
  float p = fmaf(-0.00035873236f, t, 0.037050345f);
  p = fmaf(p, t, 0.79745847f);
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(x * p));
  const float hx = 0.5f * x;
  return fmaf(hx, th, hx);
}
// 2^x in one MUFU.EX2 (flush-to-zero; 2^-inf = 0): softmax scores arrive pre-multiplied by log2(e) (pack_mha)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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

// fp32 -> fp16 with round-to-nearest and saturation to the largest finite value (operands of kind::f16 / m16n8k16)
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint16_t f16_sat(float x) {
  uint16_t r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(x));
  return r;
}

}  // namespace sepref
