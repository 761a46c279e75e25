// tcgen05 / TMEM / TMA kernels of the separator path (GEMM_PATH 1), sm_100a only.
//
// k_gcfn<F>: the whole GCFN block (reference network.py:60-66, 56 calls and 62% of the separator's FLOPs per
// forward) in one persistent, warp-specialised kernel:
//
//      y = x + W2' . GLU( dw3( W1' . norm(x) + b1' ) ) + b2'
//
// (LayerNorm's affine is folded into W1'/b1', LayerScale into W2'/b2' at pack time.)
//
// Orientation: output CHANNELS are the MMA M dimension (A operand = weights, streamed by TMA from L2), TOKENS are
// N (B operand = activations, produced on chip).  So in TMEM an accumulator lane is a channel and a column is a
// token, which makes every per-channel quantity (bias, depthwise taps) a per-thread scalar, turns the depthwise
// time convolution into register arithmetic along the columns a thread owns, and makes channels-last global
// accesses coalesced (32 lanes = 32 consecutive channels of one token).
//
// One CTA = one token tile of NTOK frames (1 halo frame each side, recomputed) of one utterance row:
//   warp 0      TMA producer: streams W1'/W2' k-slabs ([128 rows x 32 fp32], SWIZZLE_128B) through an NST-deep ring
//   warp 1      MMA issuer (one elected lane): tcgen05.mma kind::tf32, M=128, N=NTOK, K=8; owns TMEM alloc/dealloc
//   warps 2-5   producer group: reads x, normalises (two-pass, fp32), rounds to TF32 (rna) and writes the B operand
//               tile in the SWIZZLE_128B K-major layout; then drains the finished Y accumulator of the PREVIOUS
//               tile: y = x + Y + b2' (coalesced stores)
//   warps 6-13  two epilogue groups, each owning one TMEM (value,gate) accumulator pair and one stage-2 operand
//               buffer: h = D + b1' (zero outside the utterance = the conv's zero padding), depthwise k=3 along
//               the columns, GLU, round to TF32, write as the B operand of stage 2
// Stage 1 (GEMM1) runs one (value tile, gate tile) pair of 128 channels at a time into a double-buffered TMEM
// pair; stage 2 (GEMM2) accumulates Y over the 128-channel chunks.  All hand-offs are mbarriers; tcgen05.commit
// releases smem/TMEM back to the producers.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>

#include "common.cuh"

namespace sepref {
namespace tc {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.  The slow path is kept out of
// line - it is inlined at ~20 wait sites otherwise, and instruction-cache footprint is what limits these kernels.
// The poll loop is the most-executed code of these kernels (ncu: 47 % of all warp instructions of a k_gcfn launch were
// its 11 instructions, issued by waiting warps that share their scheduler with the epilogue warps), so the
// wall-clock bound is checked once per 64 polls only.  SEPREF_MBAR_SUSPEND_NS (opt-in, unmeasured) additionally
// passes a suspend-time hint so that the hardware parks the warp instead of returning after ~30 clk.
#ifdef SEPREF_MBAR_SUSPEND_NS
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)SEPREF_MBAR_SUSPEND_NS)
      : "memory");
  return ok != 0;
}
#endif
// -DSEPREF_INLINE_WAITS (profiling builds) inlines the loop at every wait site, so that ncu's source view attributes the
// polling to the barrier being waited for; the default keeps one out-of-line copy (instruction-cache footprint).
#ifdef SEPREF_INLINE_WAITS
__device__ __forceinline__
#else
__device__ __noinline__
#endif
void mbar_wait_slow(uint64_t* bar, uint32_t parity, int tag) {
  const long long t0 = clock64();
  for (;;) {
#pragma unroll 1
    for (int n = 0; n < 64; ++n) {
#ifdef SEPREF_MBAR_SUSPEND_NS
      if (mbar_try_wait_hint(bar, parity)) return;
#else
      if (mbar_try_wait(bar, parity)) return;
#endif
    }
    if (clock64() - t0 > 4000000000LL) {
      printf("sepref: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, (int)blockIdx.x, (int)threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity, tag);
}
// Wait of a role that is far off the critical path (operand producers a tile ahead, the issuing thread of an
// arithmetic-bound kernel): the hardware may park the warp for up to ~1 us per try instead of returning after ~30 clk,
// so its polling stops competing for issue slots with the warps that do the arithmetic (k_cla_front issues at 74 %).
__device__ __noinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, int tag) {
  const long long t0 = clock64();
  for (;;) {
#pragma unroll 1
    for (int n = 0; n < 16; ++n) {
      uint32_t ok;
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
          : "=r"(ok)
          : "r"(smem_u32(bar)), "r"(parity), "r"(1000u)
          : "memory");
      if (ok) return;
    }
    if (clock64() - t0 > 4000000000LL) {
      printf("sepref: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, (int)blockIdx.x, (int)threadIdx.x, parity);
      __trap();
    }
  }
}
// Programmatic dependent launch: the tensor-core kernels are launched with programmatic stream serialization, so a
// kernel's CTAs may start (barrier init, TMEM allocation, tensor-map prefetch, weight slabs - none of which depend on
// the previous kernel) while the previous kernel's last CTAs are still running.  pdl_wait() blocks until the previous
// grid has completed and its memory is visible: every warp that touches activations calls it before its first access.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// A 128-thread role group (the producers, one epilogue group) waits on an mbarrier.  With -DSEPREF_FANOUT_WAITS only the
// group's first warp polls; the other three park on a named barrier (no issue slots while parked) that the first warp
// joins once the phase has completed.  Opt-in until measured: polling was 47 % of k_gcfn's executed instructions.
__device__ __forceinline__ void mbar_wait_group(uint64_t* bar, uint32_t parity, int tag, int barid, bool leader) {
#ifdef SEPREF_FANOUT_WAITS
  if (leader) mbar_wait(bar, parity, tag);
  asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory");
#else
  (void)barid; (void)leader;
  mbar_wait(bar, parity, tag);
#endif
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// UMMA shared-memory descriptor: K-major operand tile in the canonical SWIZZLE_128B layout (rows of 128 B, groups of
// 8 rows = 1024 B).  Fields as in the sm_100 descriptor format: start>>4 [0,14), LBO>>4 [16,30) (unused for swizzled
// K-major), SBO>>4 [32,46) = 1024 B between 8-row groups, version=1 [46,48), layout type SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::tf32, fp32 accumulate, both operands K-major: c_format F32 (1) at [4,6),
// a/b format TF32 (2) at [7,10)/[10,13), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// ---- operand kinds ------------------------------------------------------------------------------------------------
// KIND_TF32: fp32 containers, 19-bit tf32 operands (rounded to nearest by the producers / at pack time).
// KIND_F16 : fp16 operands - the same 11-bit significand as TF32 in half the bytes.  The kernels are bound by the
//            per-SM operand ingest (measured 37.8 B/clk/SM, tools/microbench/tma_ingest.cu), so halving the weight
//            bytes is worth 2x where TF32 is not.  Range is kept safe by per-row power-of-two weight scaling (undone
//            exactly in the epilogue) and saturating activation conversion (cvt.rn.satfinite).
enum Kind { KIND_TF32 = 0, KIND_F16 = 1 };
template <int KIND> struct KindT;
template <> struct KindT<KIND_TF32> { static constexpr int ES = 4, KSLAB = 32; static constexpr uint32_t FMT = 2; };
template <> struct KindT<KIND_F16> { static constexpr int ES = 2, KSLAB = 64; static constexpr uint32_t FMT = 0; };

template <int KIND>
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (KindT<KIND>::FMT << 7) | (KindT<KIND>::FMT << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int KIND>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if (KIND == KIND_TF32) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
  } else {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
  }
}
// Producer side: channels 4*lane + 128*k .. +3 of tile row r (values v * scale) into a K-major SWIZZLE_128B operand
// tile whose 128-byte k slabs of all rows are `atom_b` bytes apart.
template <int KIND>
__device__ __forceinline__ void store_row4(unsigned char* buf, int atom_b, int r, int lane, int k, float4 v, float scale) {
  const uint32_t row = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
  if (KIND == KIND_TF32) {
    uint4 o;
    o.x = f32_to_tf32_rna(v.x * scale); o.y = f32_to_tf32_rna(v.y * scale);
    o.z = f32_to_tf32_rna(v.z * scale); o.w = f32_to_tf32_rna(v.w * scale);
    *reinterpret_cast<uint4*>(buf + ((lane >> 3) + 4 * k) * atom_b + row + (uint32_t)(((lane & 7) ^ (r & 7)) << 4)) = o;
  } else {
    uint2 o;
    o.x = pack_f16x2_sat(v.x * scale, v.y * scale);
    o.y = pack_f16x2_sat(v.z * scale, v.w * scale);
    *reinterpret_cast<uint2*>(buf + ((lane >> 4) + 2 * k) * atom_b + row + (uint32_t)((((lane & 15) >> 1) ^ (r & 7)) << 4) + (uint32_t)(lane & 1) * 8u) = o;
  }
}
// Epilogue side: thread (quarter q, lane) owns channel 32q+lane of a 128-channel chunk; element (row c) lives at
// sbase[c & 7] + (c >> 3) * 1024.
template <int KIND>
__device__ __forceinline__ void make_sbase(unsigned char* (&sbase)[8], unsigned char* buf, int atom_b, int q, int lane) {
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    if (KIND == KIND_TF32) sbase[m] = buf + q * atom_b + m * 128 + (((lane >> 2) ^ m) << 4) + (lane & 3) * 4;
    else sbase[m] = buf + (q >> 1) * atom_b + m * 128 + ((((q & 1) * 4 + (lane >> 3)) ^ m) << 4) + (lane & 7) * 2;
  }
}
template <int KIND>
__device__ __forceinline__ void store_elem(unsigned char* p, float v) {
  if (KIND == KIND_TF32) *reinterpret_cast<uint32_t*>(p) = f32_to_tf32_rna(v);
  else *reinterpret_cast<uint16_t*>(p) = f16_sat(v);
}

// The same store through a 32-bit shared-memory address: with a generic pointer every 2-byte store cost three extra
// address instructions (64-bit add with carry) and went out as ST.E.U16; here constant offsets fold into STS [R + imm].
template <int KIND>
__device__ __forceinline__ void sts_elem(uint32_t addr, float v) {
  if (KIND == KIND_TF32) asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(f32_to_tf32_rna(v)) : "memory");
  else asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(f16_sat(v)) : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- stage-1 operand producer ---------------------------------------------------------------------------------------
// float4 number c4 (channels 4*c4 .. 4*c4+3) of tile row r, scaled, into a K-major SWIZZLE_128B operand tile
template <int KIND>
__device__ __forceinline__ void store_c4(unsigned char* buf, int atom_b, int r, int c4, float4 v, float scale) {
  const uint32_t row = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
  if (KIND == KIND_TF32) {
    uint4 o;
    o.x = f32_to_tf32_rna(v.x * scale); o.y = f32_to_tf32_rna(v.y * scale);
    o.z = f32_to_tf32_rna(v.z * scale); o.w = f32_to_tf32_rna(v.w * scale);
    *reinterpret_cast<uint4*>(buf + (c4 >> 3) * atom_b + row + (uint32_t)(((c4 & 7) ^ (r & 7)) << 4)) = o;
  } else {
    uint2 o;
    o.x = pack_f16x2_sat(v.x * scale, v.y * scale);
    o.y = pack_f16x2_sat(v.z * scale, v.w * scale);
    *reinterpret_cast<uint2*>(buf + (c4 >> 4) * atom_b + row + (uint32_t)((((c4 & 15) >> 1) ^ (r & 7)) << 4) + (uint32_t)(c4 & 1) * 8u) = o;
  }
}
// One producer warp (pw of 4) fills rows 4*(pw + 4*i) + (lane >> 3) of an [NTOK x F_IN] operand tile.  Eight lanes
// share a row (each holds F_IN/8 channels), so a LayerNorm reduction is 3 shuffle steps shared by 4 rows instead of
// 5 steps per row; G row groups (4*G rows per warp) are kept in flight.  load(r, c4) returns float4 c4 of row r.
// `passes` > 1 averages that many source rows per tile row (load(r, c4, pass)): EGA's adaptive_avg_pool1d.
// LayerNorm (optional) + operand store of G row groups held in registers: v[g][k] is float4 j + 8k of row 4*(pw + 4*(i0+g)) + sub
template <int KIND, int F_IN, bool NORM, int G>
__device__ __forceinline__ void finish_rows(unsigned char* buf, int atom_b, int pw, int lane, int i0, float4 (&v)[G][F_IN / 32],
                                            float* amax = nullptr) {
  constexpr int NV4 = F_IN / 32;
  const int sub = lane >> 3, j = lane & 7;
  // range tracking of raw-stream operands (TokParams::range_flag) happens HERE, on values that have arrived - an
  // abs-max placed next to each load made every load wait for the previous one (the predicated-off instruction still
  // carries the scoreboard wait): 60 k instead of 10 k clk per tile
  if (!NORM && amax != nullptr) {
    float a = *amax;
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int k = 0; k < NV4; ++k) a = fmaxf(fmaxf(a, fmaxf(fabsf(v[g][k].x), fabsf(v[g][k].y))), fmaxf(fabsf(v[g][k].z), fabsf(v[g][k].w)));
    *amax = a;
  }
  float sc[G];
  if (NORM) {
    float m[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < NV4; ++k) s += v[g][k].x + v[g][k].y + v[g][k].z + v[g][k].w;
      m[g] = s;
    }
#pragma unroll
    for (int o = 1; o <= 4; o <<= 1)
#pragma unroll
      for (int g = 0; g < G; ++g) m[g] += __shfl_xor_sync(0xffffffffu, m[g], o);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float mean = m[g] * (1.0f / F_IN);
      float qq = 0.f;
#pragma unroll
      for (int k = 0; k < NV4; ++k) {
        v[g][k].x -= mean; v[g][k].y -= mean; v[g][k].z -= mean; v[g][k].w -= mean;
        qq += v[g][k].x * v[g][k].x + v[g][k].y * v[g][k].y + v[g][k].z * v[g][k].z + v[g][k].w * v[g][k].w;
      }
      sc[g] = qq;
    }
#pragma unroll
    for (int o = 1; o <= 4; o <<= 1)
#pragma unroll
      for (int g = 0; g < G; ++g) sc[g] += __shfl_xor_sync(0xffffffffu, sc[g], o);
#pragma unroll
    for (int g = 0; g < G; ++g) sc[g] = rsqrtf(sc[g] * (1.0f / F_IN) + kLnEps);
  } else {
#pragma unroll
    for (int g = 0; g < G; ++g) sc[g] = 1.0f;
  }
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int r = 4 * (pw + 4 * (i0 + g)) + sub;
#pragma unroll
    for (int k = 0; k < NV4; ++k) store_c4<KIND>(buf, atom_b, r, j + 8 * k, v[g][k], sc[g]);
  }
}
template <int KIND, int F_IN, int NTOK, bool NORM, class LoadFn>
__device__ __forceinline__ void produce_rows(unsigned char* buf, int atom_b, int pw, int lane, int passes, LoadFn load,
                                             float* amax = nullptr) {
  constexpr int NV4 = F_IN / 32;                 // float4 per lane per row
  constexpr int GI = NTOK / 16;                  // row groups per warp
  constexpr int CAP = (20 / NV4 > 0) ? 20 / NV4 : 1;
  constexpr int G = (GI % 5 == 0 && CAP >= 5) ? 5 : (GI % 4 == 0 && CAP >= 4) ? 4 : (GI % 3 == 0 && CAP >= 3) ? 3
                  : (GI % 2 == 0 && CAP >= 2) ? 2 : 1;
  static_assert(NTOK % 16 == 0 && GI % G == 0, "producer tiling");
  const int sub = lane >> 3, j = lane & 7;
  if (passes == 1) {
    // -DSEPREF_PRODUCER_PIPE (opt-in): software-pipelined over batches of HB row groups (two batches = the same registers
    // the plain loop keeps in flight), the loads of batch b+1 requested before batch b is normalised and stored.
    // Measured on the B = 32 forward: no gain (gate 1.11 -> 1.13 ms, qkv 0.77 -> 0.80, k_gcfn 7.04 -> 7.07) - the
    // producers already run a tile ahead of the MMA, so their exposed load latency is not on the critical path.
    constexpr int HB = (CAP >= 4 && GI % 2 == 0) ? 2 : 1;
#ifdef SEPREF_PRODUCER_PIPE
    constexpr bool PIPE = CAP >= 2 && (GI / HB) % 2 == 0;
#else
    constexpr bool PIPE = false;
#endif
    if constexpr (PIPE) {
      constexpr int NBATCH = GI / HB;
      float4 va[HB][NV4], vb[HB][NV4];
      auto fetch = [&](int b, float4 (&v)[HB][NV4]) {
#pragma unroll
        for (int g = 0; g < HB; ++g) {
          const int r = 4 * (pw + 4 * (b * HB + g)) + sub;
#pragma unroll
          for (int k = 0; k < NV4; ++k) v[g][k] = load(r, j + 8 * k, 0);
        }
      };
      fetch(0, va);
#pragma unroll 1
      for (int b = 0; b < NBATCH; b += 2) {
        fetch(b + 1, vb);
        finish_rows<KIND, F_IN, NORM, HB>(buf, atom_b, pw, lane, b * HB, va, amax);
        if (b + 2 < NBATCH) fetch(b + 2, va);
        finish_rows<KIND, F_IN, NORM, HB>(buf, atom_b, pw, lane, (b + 1) * HB, vb, amax);
      }
      return;
    }
  }
#pragma unroll 1
  for (int i0 = 0; i0 < GI; i0 += G) {
    float4 v[G][NV4];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int k = 0; k < NV4; ++k) v[g][k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (passes == 1) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int r = 4 * (pw + 4 * (i0 + g)) + sub;
#pragma unroll
        for (int k = 0; k < NV4; ++k) v[g][k] = load(r, j + 8 * k, 0);
      }
    } else {
      // pooled rows: per row group, four source rows (passes) are requested before they are summed
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int r = 4 * (pw + 4 * (i0 + g)) + sub;
#pragma unroll 1
        for (int ps = 0; ps < passes; ps += 4) {
          float4 t[4][NV4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < NV4; ++k)
              t[u][k] = (ps + u < passes) ? load(r, j + 8 * k, ps + u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < NV4; ++k) { v[g][k].x += t[u][k].x; v[g][k].y += t[u][k].y; v[g][k].z += t[u][k].z; v[g][k].w += t[u][k].w; }
        }
      }
      const float inv = 1.0f / (float)passes;
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int k = 0; k < NV4; ++k) { v[g][k].x *= inv; v[g][k].y *= inv; v[g][k].z *= inv; v[g][k].w *= inv; }
    }
    finish_rows<KIND, F_IN, NORM, G>(buf, atom_b, pw, lane, i0, v, amax);
  }
}

// Speaker-attention producer (SpkAttention's 2-token MultiHeadAttention, network.py:240-246 with 106-122, pos_k=None).
// A token tile of this kernel is a PAIR tile: rows [0,64) are frames t0..t0+63 of speaker 0 of one utterance, rows
// [64,128) the same frames of speaker 1, so both attention partners of a frame sit in one tile and every q|k|v row
// is read once.  qkv rows are [q | k | v] (3F values, q pre-scaled by log2(e)/sqrt(dk), softmax in base 2; FP16 when IN16).  With 8 heads the
// 8 lanes that share a frame each own exactly one head (F/8 = dk channels), so the four scores, the two 2-way
// softmaxes and the weighted sums of the two value vectors are thread-local; the attention output goes straight
// into the out-projection's B operand.  mb0 = first row of speaker 0's frames, nh = valid frames (<= 64).
template <int KIND, int F, int NTOK, bool IN16>
__device__ __forceinline__ void produce_spk_pair(unsigned char* buf, int atom_b, int pw, int lane, const float* qkv,
                                                 int T, long long mb0, int nh) {
  static_assert(NTOK == 128, "pair tiles are 2 x 64 frames");
  constexpr int DK = F / 8;
  constexpr int EPV = IN16 ? 8 : 4;              // elements per 16-byte vector
  constexpr int NV = DK / EPV;                   // vectors per head slice
  constexpr int G = (IN16 && DK == 16) ? 2 : 1;  // frame groups in flight per warp
  const int sub = lane >> 3, j = lane & 7;
  auto ldv = [&](long long row, int part, int k) -> uint4 {
    if (IN16) return __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(qkv) + row * (3 * F) + part * F + j * DK) + k);
    return __ldg(reinterpret_cast<const uint4*>(qkv + row * (3 * F) + part * F + j * DK) + k);
  };
  auto cvt = [&](const uint4& v, float (&f)[EPV]) {
    if (IN16) {
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
      const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&v.z)), d = __half22float2(*reinterpret_cast<const __half2*>(&v.w));
      f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
      if (EPV == 8) { f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y; }
    } else {
      f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
    }
  };
#pragma unroll 1
  for (int i0 = 0; i0 < 4; i0 += G) {
    uint4 raw[G][6][NV];                         // [q0 k0 v0 q1 k1 v1]
    int tr[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      tr[g] = 4 * (pw + 4 * (i0 + g)) + sub;
      const long long r0 = mb0 + (tr[g] < nh ? tr[g] : nh - 1);
#pragma unroll
      for (int pt = 0; pt < 3; ++pt)
#pragma unroll
        for (int k = 0; k < NV; ++k) { raw[g][pt][k] = ldv(r0, pt, k); raw[g][3 + pt][k] = ldv(r0 + T, pt, k); }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        float q0[EPV], k0[EPV], q1[EPV], k1[EPV];
        cvt(raw[g][0][k], q0); cvt(raw[g][1][k], k0); cvt(raw[g][3][k], q1); cvt(raw[g][4][k], k1);
#pragma unroll
        for (int e = 0; e < EPV; ++e) {
          s00 = fmaf(q0[e], k0[e], s00); s01 = fmaf(q0[e], k1[e], s01);
          s10 = fmaf(q1[e], k0[e], s10); s11 = fmaf(q1[e], k1[e], s11);
        }
      }
      const bool valid = tr[g] < nh;
      const float m0 = fmaxf(s00, s01), m1 = fmaxf(s10, s11);
      float p00 = exp2f(s00 - m0), p01 = exp2f(s01 - m0), p10 = exp2f(s10 - m1), p11 = exp2f(s11 - m1);   // q carries log2(e)
      const float i0v = valid ? 1.0f / (p00 + p01) : 0.f, i1v = valid ? 1.0f / (p10 + p11) : 0.f;
      p00 *= i0v; p01 *= i0v; p10 *= i1v; p11 *= i1v;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        float v0[EPV], v1[EPV];
        cvt(raw[g][2][k], v0); cvt(raw[g][5][k], v1);
#pragma unroll
        for (int e4 = 0; e4 < EPV; e4 += 4) {
          const float4 o0 = make_float4(p00 * v0[e4] + p01 * v1[e4], p00 * v0[e4 + 1] + p01 * v1[e4 + 1],
                                        p00 * v0[e4 + 2] + p01 * v1[e4 + 2], p00 * v0[e4 + 3] + p01 * v1[e4 + 3]);
          const float4 o1 = make_float4(p10 * v0[e4] + p11 * v1[e4], p10 * v0[e4 + 1] + p11 * v1[e4 + 1],
                                        p10 * v0[e4 + 2] + p11 * v1[e4 + 2], p10 * v0[e4 + 3] + p11 * v1[e4 + 3]);
          const int c4 = (j * DK + k * EPV + e4) >> 2;
          store_c4<KIND>(buf, atom_b, tr[g], c4, o0, 1.0f);
          store_c4<KIND>(buf, atom_b, 64 + tr[g], c4, o1, 1.0f);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ cluster helpers
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one slab part, delivered to the same smem offset (and signalling the mbarrier at the same offset) in every CTA of `mask`
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// plain (non-tensor) bulk copy global -> shared, completion counted in bytes on an mbarrier; 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// A read-only load the compiler must issue where it is written: plain __ldg()s placed ahead of an mbarrier wait were
// sunk next to their first use by the scheduler, which put the full memory latency back on the critical path.
__device__ __forceinline__ float ldg_now(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ------------------------------------------------------------------------------------------------ configuration
// WIDE (fp16, F = 128 only): one 160-frame tile with SINGLE-buffered (value,gate) accumulators (2*160 + 160 TMEM
// columns).  The weight slabs are then fetched once per 158 instead of once per 94 frames and every tcgen05.mma
// carries N = 160 (the per-instruction re-read of the 4 KB A slice amortises over 1.7x the work); the price is that
// GEMM1 of chunk j+1 cannot overlap the epilogue of chunk j, so both epilogue groups share every chunk (80 columns
// each) instead of alternating chunks.
template <int F, int KIND, bool WIDE = false>
struct GcfnTraits {
  static_assert(!WIDE || (F == 128 && KIND == KIND_F16), "wide tiles: fp16, F = 128");
  using KT = KindT<KIND>;
  // frames per tile incl. 2 halo frames (TMEM: 5N resp. 6N <= 512 columns)
  static constexpr int NTOK = WIDE ? 160 : (KIND == KIND_F16 && F == 128) ? 96 : 80;
  static constexpr int NV = NTOK - 2;                    // frames a tile produces
  static constexpr int NST = WIDE ? 6 : (KIND == KIND_F16) ? (F == 128 ? 8 : 6) : 4;   // weight ring depth
  static constexpr int NB1 = (F == 128 || KIND == KIND_F16) ? 2 : 1;   // stage-1 operand buffers (next tile's LayerNorm overlaps)
  static constexpr int K1A = F / KT::KSLAB;              // 128-byte k slabs of GEMM1
  static constexpr int K2A = 128 / KT::KSLAB;            // k slabs per 128-channel chunk of GEMM2
  static constexpr int NCH = 3 * F / 128;                // (value,gate) tile pairs == 128-wide k chunks of GEMM2
  static constexpr int M2 = F / 128;                     // output-channel tiles of GEMM2
  static constexpr int ATOM_B = NTOK * 128;              // one [NTOK rows x 128 B] swizzled slab
  static constexpr int B1_BYTES = K1A * ATOM_B;
  static constexpr int B2_BYTES = K2A * ATOM_B;
  static constexpr int A_BYTES = 128 * 128;              // one weight slab [128 rows x 128 B]
  static constexpr int BAR_BYTES = 512;
  static constexpr int NB2 = WIDE ? 1 : 2;               // stage-2 operand buffers
  static constexpr int SMEM_BYTES = 1024 + NST * A_BYTES + NB1 * B1_BYTES + NB2 * B2_BYTES + BAR_BYTES;
  // SPLIT: GEMM1 and GEMM2 are issued by two different threads, each fed by its own TMA thread through its own slab
  // ring (NST1 + NST2 = NST slots).  With ONE issuing thread the in-kernel timeline showed the tensor pipe idle ~78 % of
  // a tile although MMAs execute at their nominal ~50 clk: the thread issues strictly in order S1(g); S2(g-1), so a
  // GEMM2 step that waits for its epilogue (b2_full) also holds back the GEMM1 steps behind it (and the slabs queued
  // behind its slabs in the shared ring), and every slab / accumulator hand-off costs a ~90 clk mbarrier round trip on
  // that one thread.  Two threads halve the serial work and remove the head-of-line blocking.
  static constexpr bool SPLIT = (KIND == KIND_F16) && !WIDE && F == 128;
  static constexpr int NST2 = SPLIT ? M2 * K2A : 0;      // GEMM2 ring: one step's slabs (steps are a chunk period apart)
  static constexpr int NST1 = NST - NST2;
  static constexpr int THREADS = (SPLIT ? 16 : 14) * 32;
  static constexpr int RB = NTOK / 8;                    // rows a producer warp keeps in flight (half of its NTOK/4 rows)
  __host__ __device__ static constexpr int tm_pair(int buf, int half) { return ((WIDE ? 0 : buf) * 2 + half) * NTOK; }
  __host__ __device__ static constexpr int tm_y(int m2) { return (WIDE ? 2 : 4) * NTOK + m2 * NTOK; }
  static_assert(NTOK % 16 == 0 && (NTOK / 4) % RB == 0, "tile shape");
  static_assert((WIDE ? 2 : 4) * NTOK + M2 * NTOK <= 512, "TMEM columns");
  static_assert(SMEM_BYTES <= 232448, "shared memory");
};

struct GcfnPack {
  // shared by both operand kinds (packed row order: pair j -> value tile 2j, gate tile 2j+1)
  const float *b1 = nullptr;                  // [6F] GEMM1 bias
  const float *dw = nullptr, *dwb = nullptr;  // tap-major [3][6F] / [6F], pre-scaled by 1/2 (tanh form of the gate, see epilogue)
  const float *cb = nullptr;                  // [6F] interior conv constant (dwb + b1 * (w0+w1+w2)) / 2
  const float *b2 = nullptr;                  // [F]
  // per kind: operands (w1 [6F, F], w2 [F, 3F]; LN affine / LayerScale folded, per-row power-of-two scaled for fp16),
  // the inverse row scales and the interior-path taps with the inverse scale folded in
  const void *w1[2] = {nullptr, nullptr}, *w2[2] = {nullptr, nullptr};
  const float *s1inv[2] = {nullptr, nullptr}, *s2inv[2] = {nullptr, nullptr};
  const float *dwf[2] = {nullptr, nullptr};   // tap-major [3][6F]: dw * s1inv
  alignas(64) CUtensorMap map_w1[2][3];       // [kind][cluster 1/2/4]: box rows 128 / 64 / 32
  alignas(64) CUtensorMap map_w2[2][3];
};

struct GcfnParams {
  const float* x;
  float* y;
  const float *b1, *dw, *dwb, *cb, *b2, *dwf, *s1inv, *s2inv;
  int rows, T, tiles_per_row, num_tiles, iters;
  float* dbg_h;   // optional [rows*T, 6F] dump of h = W1'.norm(x)+b1' in the reference's channel order (tests)
  long long* dbg_clk;   // optional [8][64] clock64 stamps of block 0's first 8 tiles (pipeline timeline, tools/)
};

// ------------------------------------------------------------------------------------------------ the kernel
// CL = cluster size: the CL CTAs of a cluster walk their tiles in lockstep and share every weight slab - each CTA
// fetches 1/CL of the slab's rows and TMA-multicasts it into all CL shared memories, so L2->SM weight traffic (the
// measured limiter: ~28 B/clk/SM chip-wide) drops by CL.  A ring slot is recycled when all CL MMA issuers released it.
template <int F, int CL, int KIND, bool WIDE = false>
__global__ void __launch_bounds__(GcfnTraits<F, KIND, WIDE>::THREADS, 1)
k_gcfn(const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2, const GcfnParams p) {
  using TR = GcfnTraits<F, KIND, WIDE>;
  constexpr int NTOK = TR::NTOK, NV = TR::NV, NST = TR::NST, K1A = TR::K1A, K2A = TR::K2A, NCH = TR::NCH, M2 = TR::M2, NB1 = TR::NB1;
  constexpr int ATOM_B = TR::ATOM_B, A_BYTES = TR::A_BYTES, B2_BYTES = TR::B2_BYTES, B1_BYTES = TR::B1_BYTES;
  constexpr int KSLAB = TR::KT::KSLAB;
  constexpr uint32_t IDESC = make_idesc<KIND>(128, NTOK);
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1);
  constexpr int PART_ROWS = 128 / CL;
  constexpr bool SPLIT = TR::SPLIT;
  constexpr int NST1 = TR::NST1, NST2 = TR::NST2;          // slab ring of GEMM1 / of GEMM2 (SPLIT; else one ring of NST)

  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = sm;
  unsigned char* sA2 = sA + NST1 * A_BYTES;                // second ring (SPLIT): the last NST2 slots
  unsigned char* sB1 = sA + NST * A_BYTES;
  unsigned char* sB2 = sB1 + NB1 * B1_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB2 + TR::NB2 * B2_BYTES);
  uint64_t* a_full = bars;                 // [NST]   (SPLIT: [0,NST1) ring 1, [NST1,NST) ring 2)
  uint64_t* a_empty = a_full + NST;        // [NST]
  uint64_t* b1_full = a_empty + NST;       // [2]
  uint64_t* b1_empty = b1_full + 2;        // [2]
  uint64_t* tm_full = b1_empty + 2;        // [2]
  uint64_t* tm_empty = tm_full + 2;        // [2]
  uint64_t* b2_full = tm_empty + 2;        // [2]
  uint64_t* b2_empty = b2_full + 2;        // [2]
  uint64_t* y_full = b2_empty + 2;
  uint64_t* y_empty = y_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y_empty + 1);

  // Roles are assigned from the top warp id down: the SM's arbiter prefers the highest warp id among eligible warps,
  // so the latency-critical single-lane roles (TMA, MMA issue) and the producers must outrank the ALU-heavy epilogue.
  const int pwarp = threadIdx.x >> 5, lane = threadIdx.x & 31;     // physical warp: fixes the TMEM lane quarter
  // role index: 0 TMA, 1 MMA issue, 2-5 producers, 6-13 epilogue; SPLIT adds 14 = TMA of GEMM2's ring, 15 = GEMM2 issue.
  // The four single-lane roles sit on the four highest physical warps (one per scheduler, highest arbitration priority).
  const int warp = !SPLIT ? 13 - pwarp : (pwarp >= 12 ? (pwarp == 15 ? 0 : pwarp == 14 ? 1 : pwarp == 13 ? 14 : 15) : 13 - pwarp);
  const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
#define STAMP(itv, slot) do { if (p.dbg_clk != nullptr && blockIdx.x == 0 && (itv) < 8) p.dbg_clk[(itv) * 64 + (slot)] = clock64(); } while (0)

  // ---- one-time setup
  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], CL); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&b1_full[i], 128); mbar_init(&b1_empty[i], 1);
      mbar_init(&tm_full[i], 1); mbar_init(&tm_empty[i], WIDE ? 256 : 128);
      mbar_init(&b2_full[i], WIDE ? 256 : 128); mbar_init(&b2_empty[i], 1);
    }
    mbar_init(y_full, 1); mbar_init(y_empty, WIDE ? 256 : 128);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_w1); tma_prefetch_desc(&map_w2); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // stage-2 operand buffers: halo rows are never written by the epilogue; keep them finite
  for (int i = threadIdx.x; i < (TR::NB2 * B2_BYTES) / 16; i += TR::THREADS) reinterpret_cast<uint4*>(sB2)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();            // peers' barriers are initialised before anyone multicasts into them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  if (warp != 0 && warp != 14) pdl_wait();   // the TMA threads only stream weights, which no kernel writes

  // every CTA runs p.iters iterations (lockstep inside a cluster); iterations past the last tile are dummies
  auto tile_of = [&](int it) { return (int)blockIdx.x + it * (int)gridDim.x; };

  // =============================================================================== warp 0 (and 14): weight slabs via TMA
  if (warp == 0 || (SPLIT && warp == 14)) {
    if (lane == 0) {
      // ring of this thread: the whole ring, or (SPLIT) GEMM1's slots [0, NST1) for warp 0 and GEMM2's [NST1, NST) for warp 14
      const int r0 = (SPLIT && warp == 14) ? NST1 : 0, rn = !SPLIT ? NST : (warp == 14 ? NST2 : NST1);
      int st = 0; uint32_t ph = 0;
      auto load = [&](const CUtensorMap* map, int c0, int c1) {
        mbar_wait(&a_empty[r0 + st], ph ^ 1, 100);
        mbar_arrive_expect_tx(&a_full[r0 + st], A_BYTES);
        if (CL == 1) tma_load_2d(map, &a_full[r0 + st], sA + (r0 + st) * A_BYTES, c0, c1);
        else tma_load_2d_mc(map, &a_full[r0 + st], sA + (r0 + st) * A_BYTES + crank * (PART_ROWS * 128), c0, c1 + (int)crank * PART_ROWS, MC_MASK);
        if (++st == rn) { st = 0; ph ^= 1; }
      };
      auto s1 = [&](int j) {
        for (int half = 0; half < 2; ++half)
          for (int ka = 0; ka < K1A; ++ka) load(&map_w1, ka * KSLAB, (2 * j + half) * 128);
      };
      auto s2 = [&](int j) {
        for (int m2 = 0; m2 < M2; ++m2)
          for (int ka = 0; ka < K2A; ++ka) load(&map_w2, j * 128 + ka * KSLAB, m2 * 128);
      };
      const int total = p.iters * NCH;
      if (SPLIT) {
        if (warp == 0) { for (int g = 0; g < total; ++g) s1(g % NCH); }
        else { for (int g = 0; g < total; ++g) s2(g % NCH); }
      } else if (WIDE) {
        for (int g = 0; g < total; ++g) { s1(g % NCH); s2(g % NCH); }       // S1(g); S2(g): nothing to pipeline across
      } else {
        // issue order, identical in the MMA warp: S1(g); S2(g-1) over the running chunk index g, across tile boundaries
        for (int g = 0; g < total; ++g) {
          s1(g % NCH);
          if (g >= 1) s2((g - 1) % NCH);
        }
        if (total > 0) s2(NCH - 1);
      }
    }
  }
  // =============================================================================== warp 1 (and 15): MMA issue
  else if (warp == 1 || (SPLIT && warp == 15)) {
    if (lane == 0) {
      const int r0 = (SPLIT && warp == 15) ? NST1 : 0, rn = !SPLIT ? NST : (warp == 15 ? NST2 : NST1);
      int st = 0; uint32_t ph = 0;
      int it = 0;       // tile index of the GEMM2 chunk being issued (y_empty parity)
      auto release = [&](uint64_t* bar) { if (CL == 1) umma_commit(bar); else umma_commit_mc(bar, MC_MASK); };
      auto s1 = [&](uint32_t gj, const unsigned char* b1buf) {
        const uint32_t b = WIDE ? 0u : (gj & 1), n = WIDE ? gj : (gj >> 1);
        mbar_wait(&tm_empty[b], (n & 1) ^ 1, 200);
        tcgen05_fence_after();
        for (int half = 0; half < 2; ++half) {
          const uint32_t d = tmem_base + TR::tm_pair(b, half);
          for (int ka = 0; ka < K1A; ++ka) {
            mbar_wait(&a_full[r0 + st], ph, 201);
            tcgen05_fence_after();
            const uint64_t ad = make_sdesc(smem_u32(sA + (r0 + st) * A_BYTES));
            const uint64_t bd = make_sdesc(smem_u32(b1buf + ka * ATOM_B));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma<KIND>(d, ad + 2 * k, bd + 2 * k, IDESC, (ka | k) != 0);
            release(&a_empty[r0 + st]);
            if (++st == rn) { st = 0; ph ^= 1; }
          }
        }
        umma_commit(&tm_full[b]);
      };
      auto s2 = [&](int j, uint32_t gj) {
        const uint32_t b = WIDE ? 0u : (gj & 1), n = WIDE ? gj : (gj >> 1);
        mbar_wait(&b2_full[b], n & 1, 202);
        if (j == 0) mbar_wait(y_empty, (it & 1) ^ 1, 203);
        tcgen05_fence_after();
        for (int m2 = 0; m2 < M2; ++m2) {
          const uint32_t d = tmem_base + TR::tm_y(m2);
          for (int ka = 0; ka < K2A; ++ka) {
            mbar_wait(&a_full[r0 + st], ph, 204);
            tcgen05_fence_after();
            const uint64_t ad = make_sdesc(smem_u32(sA + (r0 + st) * A_BYTES));
            const uint64_t bd = make_sdesc(smem_u32(sB2 + b * B2_BYTES + ka * ATOM_B));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma<KIND>(d, ad + 2 * k, bd + 2 * k, IDESC, (j | ka | k) != 0);
            release(&a_empty[r0 + st]);
            if (++st == rn) { st = 0; ph ^= 1; }
          }
        }
        umma_commit(&b2_empty[b]);
      };
      const int total = p.iters * NCH;
      auto do_s2 = [&](int gprev) {
        const int jj = gprev % NCH;
        it = gprev / NCH;                               // s2() reads `it` for the y_empty parity
        s2(jj, (uint32_t)gprev);
        STAMP(it, 8 + jj);
        if (jj == NCH - 1) umma_commit(y_full);
      };
      auto do_s1 = [&](int gg) {
        const int ti = gg / NCH, j = gg % NCH;
        const int bb = (NB1 == 2) ? (ti & 1) : 0;
        const uint32_t bpar = (NB1 == 2) ? ((ti >> 1) & 1) : (ti & 1);
        const unsigned char* b1buf = sB1 + bb * B1_BYTES;
        if (j == 0) {
          mbar_wait(&b1_full[bb], bpar, 205);
          tcgen05_fence_after();
          STAMP(ti, 0);
        }
        s1((uint32_t)gg, b1buf);
        STAMP(ti, 1 + j);
        if (j == NCH - 1) umma_commit(&b1_empty[bb]);      // every GEMM1 MMA of this tile has been issued
      };
      if (SPLIT) {
        // each GEMM has its own issuing thread: a GEMM2 step that waits for its epilogue no longer holds back GEMM1
        if (warp == 1) { for (int gg = 0; gg < total; ++gg) do_s1(gg); }
        else { for (int gg = 0; gg < total; ++gg) do_s2(gg); }
      } else {
        // Software-pipelined across tiles: S1(g); S2(g-1).  The last GEMM2 chunk of tile i is issued after the first
        // GEMM1 pair of tile i+1, so the tensor pipe never waits for an epilogue at a tile boundary.
        for (int gg = 0; gg < total; ++gg) {
          do_s1(gg);
          if (WIDE) do_s2(gg);
          else if (gg >= 1) do_s2(gg - 1);
        }
        if (!WIDE && total > 0) do_s2(total - 1);
      }
    }
  }
  // =============================================================================== warps 2-5: stage-1 operand producer
  else if (warp < 6) {
    const int pw = warp - 2;
    for (int it = 0; it < p.iters; ++it) {
      const int tile = tile_of(it);
      const bool live = tile < p.num_tiles;
      const int n = tile / p.tiles_per_row, t0 = (tile % p.tiles_per_row) * NV;
      const int bb = (NB1 == 2) ? (it & 1) : 0;
      const uint32_t bpar = (NB1 == 2) ? ((it >> 1) & 1) : (it & 1);
      unsigned char* b1buf = sB1 + bb * B1_BYTES;
      mbar_wait_group(&b1_empty[bb], bpar ^ 1, 301, 1, pw == 0);
      if (warp == 2 && lane == 0) STAMP(it, 16);
      {
        const float4* x4 = reinterpret_cast<const float4*>(p.x) + (size_t)(live ? n : 0) * p.T * (F / 4);
        const int T = p.T;
        produce_rows<KIND, F, NTOK, true>(b1buf, ATOM_B, pw, lane, 1, [&](int r, int c4, int) {
          const int t = t0 - 1 + r;                        // frames outside the utterance are zero rows
          return (live && t >= 0 && t < T) ? __ldg(x4 + (size_t)t * (F / 4) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        });
      }
      fence_proxy_async();
      mbar_arrive(&b1_full[bb]);
      if (warp == 2 && lane == 0) STAMP(it, 17);
    }
  }
  // =============================================================================== warps 6-13: gated-conv epilogue
  else {
    const int eg = (warp - 6) >> 2;               // epilogue group == TMEM pair == stage-2 buffer
    const bool gl = ((warp - 6) & 3) == 0;        // first warp of the group (polls for the group with SEPREF_FANOUT_WAITS)
    const int q = pwarp & 3;
    const int ch = q * 32 + lane;                 // channel within the 128-chunk; its k slab is q, k index is lane
    // per-thread store bases for the 8 possible (column & 7): the swizzle XOR is folded in, the rest is an immediate
    unsigned char* sbase[8];
    make_sbase<KIND>(sbase, sB2 + (WIDE ? 0 : eg) * B2_BYTES, ATOM_B, q, lane);
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    // WIDE: both groups work on every chunk, group eg on columns [c0, c0 + EC); otherwise a group owns whole chunks
    constexpr int EC = WIDE ? NTOK / 2 : NTOK;
    const int c0 = WIDE ? eg * EC : 0;
    const int bi = WIDE ? 0 : eg;                 // accumulator pair / stage-2 buffer / barrier index of this group
    // y = x + Y * s2inv + b2' for the tile whose GEMM2 finished.  Tile i is drained by group (i & 1) at the start of
    // iteration i+1, where that group owns the smaller share of the (value,gate) chunks; 32 columns are in flight.
    auto drain = [&](int tile, int it) {
      const bool live = tile < p.num_tiles;
      const int n = tile / p.tiles_per_row, t0 = (tile % p.tiles_per_row) * NV;
      // columns 1 .. cmax of the tile are frames of the utterance (column c is frame t0 - 1 + c)
      const int cmax = live ? min(NTOK - 2, p.T - t0) : 0;
      const float* xcol = p.x + (((long long)n * p.T + t0 - 1) * F + ch);
      float* ycol = p.y + (((long long)n * p.T + t0 - 1) * F + ch);
      float xin[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) xin[i] = (c0 + i >= 1 && c0 + i <= cmax) ? ldg_now(xcol + (c0 + i) * F) : 0.f;   // batch 0 of output tile 0
      const float bias0 = ldg_now(p.b2 + ch), s2i0 = ldg_now(p.s2inv + ch);      // ahead of the wait
      mbar_wait_group(y_full, it & 1, 300, 2 + eg, gl);
      tcgen05_fence_after();
      if ((warp == 6 || warp == 10) && lane == 0) STAMP(it + 1, 19);
#pragma unroll
      for (int m2 = 0; m2 < M2; ++m2) {
        const float bias = m2 == 0 ? bias0 : __ldg(p.b2 + m2 * 128 + ch), s2i = m2 == 0 ? s2i0 : __ldg(p.s2inv + m2 * 128 + ch);
#pragma unroll 1
        for (int cb = c0; cb < c0 + EC; cb += 32) {
          uint32_t ra[16], rb[16];
          tmem_ld16(tmem_base + tlane + TR::tm_y(m2) + cb, ra);
          if (cb + 16 < c0 + EC) tmem_ld16(tmem_base + tlane + TR::tm_y(m2) + cb + 16, rb);
          const float* xc = xcol + m2 * 128 + cb * F;
          float* yc = ycol + m2 * 128 + cb * F;
          if (m2 > 0 || cb > c0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) xin[i] = (cb + i >= 1 && cb + i <= cmax && cb + i < c0 + EC) ? ldg_now(xc + i * F) : 0.f;
          }
          tmem_wait_ld();
          if (m2 == M2 - 1 && cb + 32 >= c0 + EC) { tcgen05_fence_before(); mbar_arrive(y_empty); }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const uint32_t rr = i < 16 ? ra[i & 15] : rb[i & 15];
            if (cb + i >= 1 && cb + i <= cmax && cb + i < c0 + EC) yc[i * F] = fmaf(__uint_as_float(rr), s2i, xin[i] + bias);
          }
        }
      }
      if ((warp == 6 || warp == 10) && lane == 0) STAMP(it + 1, 18);
    };
    for (int it = 0; it <= p.iters; ++it) {
      if (it > 0 && (WIDE || eg == ((it - 1) & 1))) drain(tile_of(it - 1), it - 1);   // single call site: one inlined copy
      if (it == p.iters) break;
      const int tile = tile_of(it);
      const bool live = tile < p.num_tiles;
      const int n = tile / p.tiles_per_row, t0 = (tile % p.tiles_per_row) * NV;
      const int tcol0 = t0 - 1;
      const bool edge = (tcol0 < 0) || (tcol0 + NTOK > p.T);     // some column lies outside the utterance (zero padding)
#pragma unroll 1
      for (int j = 0; j < NCH; ++j) {
        const uint32_t gj = (uint32_t)it * NCH + j;
        if (!WIDE && (int)(gj & 1) != eg) continue;
        const uint32_t nuse = WIDE ? gj : (gj >> 1);
        const int rv = (2 * j) * 128 + ch, rg = rv + 128;       // packed GEMM1 rows of this thread's value / gate channel
        // depthwise taps pre-scaled by 1/2:  u = dv * sigmoid(dg) = (dv/2) * (1 + tanh(dg/2))
        // The per-channel constants of the interior path are requested BEFORE waiting for the accumulator (ldg_now is
        // issued where it is written), so their L2 latency hides behind the wait instead of opening every chunk.
        const bool interior = !edge && p.dbg_h == nullptr;
        float cv = 0.f, cg = 0.f, wv0 = 0.f, wv1 = 0.f, wv2 = 0.f, wg0 = 0.f, wg1 = 0.f, wg2 = 0.f;
        if (interior) {
          cv = ldg_now(p.cb + rv); cg = ldg_now(p.cb + rg);
          wv0 = ldg_now(p.dwf + rv); wv1 = ldg_now(p.dwf + 6 * F + rv); wv2 = ldg_now(p.dwf + 12 * F + rv);
          wg0 = ldg_now(p.dwf + rg); wg1 = ldg_now(p.dwf + 6 * F + rg); wg2 = ldg_now(p.dwf + 12 * F + rg);
        }
        mbar_wait_group(&tm_full[bi], nuse & 1, 400, 2 + eg, gl);
        if ((warp == 6 || warp == 10) && lane == 0) STAMP(it, 24 + j * 4);
        mbar_wait_group(&b2_empty[bi], (nuse & 1) ^ 1, 401, 2 + eg, gl);
        if ((warp == 6 || warp == 10) && lane == 0) STAMP(it, 25 + j * 4);
        tcgen05_fence_after();
        const uint32_t tv = tmem_base + tlane + TR::tm_pair(bi, 0) + c0, tg = tmem_base + tlane + TR::tm_pair(bi, 1) + c0;
        if (interior) {
          // ---- interior tile: h = D + b1 everywhere, so b1 folds into the conv constant and D is used raw.
          // The 16-column batch loop is deliberately NOT unrolled: the kernel's warps run five different code regions
          // and the instruction cache, not the ALUs, was the limiter when this body was replicated NTOK/16 times.
#ifndef SEPREF_EPI_SCALAR
          // Packed form: the outer taps of the conv and the gate product are fma.rn.f32x2 over two neighbouring columns
          // (the FP32 pipe executes a packed FMA at the scalar rate, but it takes ONE issue slot instead of two, and issue
          // slots and latency - not FMA throughput - bound this loop).  A packed operand must sit in an aligned register
          // pair: for the output pair (cb-1+2m, cb+2m) the taps w0 and w2 read the aligned pairs (e[2m-2], e[2m-1]) and
          // (e[2m], e[2m+1]) of the 16 columns just loaded; the middle tap needs the misaligned pair (e[2m-1], e[2m]) and
          // stays two scalar FMAs (reading the accumulator a second time at a one-column offset would supply it aligned,
          // but costs a second TMEM read per batch - measured slower once the loads were pipelined).
          // The TMEM reads are software-pipelined: the loads of batch b+1 are in flight while batch b is computed, so
          // the ~200 clk tcgen05.ld latency no longer sits between every pair of batches.
          constexpr bool PIPE = !WIDE && (EC / 16) % 2 == 0;    // the pipelined loop walks two batches per iteration
          float pv0 = 0.f, pv1 = 0.f, pg0 = 0.f, pg1 = 0.f;     // D of columns cb-2, cb-1
          if (WIDE && c0 > 0) {                                  // the second group starts mid-tile
            uint32_t x[16], y[16];
            tmem_ld16(tv - 16, x); tmem_ld16(tg - 16, y);
            tmem_wait_ld();
            pv0 = __uint_as_float(x[14]); pv1 = __uint_as_float(x[15]); pg0 = __uint_as_float(y[14]); pg1 = __uint_as_float(y[15]);
          }
          const float2 wv0p = make_float2(wv0, wv0), wv2p = make_float2(wv2, wv2), wg0p = make_float2(wg0, wg0), wg2p = make_float2(wg2, wg2);
          const float2 cvp = make_float2(cv, cv), cgp = make_float2(cg, cg);
          uint32_t sb[8];                                        // running 32-bit store bases: row block of the current batch
#pragma unroll
          for (int k = 0; k < 8; ++k) sb[k] = smem_u32(sbase[k]) + (uint32_t)((c0 >> 3) * 1024);
          // one 16-column batch: output columns cb-1 .. cb+14 (column -1 of the tile is skipped)
          auto batch = [&](const uint32_t (&ev)[16], const uint32_t (&eg)[16], bool first) {
#pragma unroll
            for (int m = 0; m < 8; ++m) {
              const float2 av = m == 0 ? make_float2(pv0, pv1) : make_float2(__uint_as_float(ev[2 * m - 2]), __uint_as_float(ev[2 * m - 1]));
              const float2 ag = m == 0 ? make_float2(pg0, pg1) : make_float2(__uint_as_float(eg[2 * m - 2]), __uint_as_float(eg[2 * m - 1]));
              const float2 cvv = make_float2(__uint_as_float(ev[2 * m]), __uint_as_float(ev[2 * m + 1]));
              const float2 cgg = make_float2(__uint_as_float(eg[2 * m]), __uint_as_float(eg[2 * m + 1]));
              float2 dv = __ffma2_rn(wv2p, cvv, __ffma2_rn(wv0p, av, cvp));
              float2 dg = __ffma2_rn(wg2p, cgg, __ffma2_rn(wg0p, ag, cgp));
              dv.x = fmaf(wv1, av.y, dv.x); dv.y = fmaf(wv1, cvv.x, dv.y);     // middle tap: columns cb-1+2m and cb+2m
              dg.x = fmaf(wg1, ag.y, dg.x); dg.y = fmaf(wg1, cgg.x, dg.y);
              const float2 th = make_float2(tanh_approx(dg.x), tanh_approx(dg.y));
              const float2 u = __ffma2_rn(dv, th, dv);
              // (column & 7) of x is (2m+7)&7, of y is (2m)&7 because cb % 16 == 0.  Halo rows (columns 0, NTOK-1) are
              // written too: they only feed Y's halo columns, which are never stored.
              if (m > 0 || !first) sts_elem<KIND>(sb[(2 * m + 7) & 7] + (uint32_t)(((2 * m - 1) >> 3) * 1024), u.x);
              sts_elem<KIND>(sb[(2 * m) & 7] + (uint32_t)(((2 * m) >> 3) * 1024), u.y);
            }
            pv0 = __uint_as_float(ev[14]); pv1 = __uint_as_float(ev[15]); pg0 = __uint_as_float(eg[14]); pg1 = __uint_as_float(eg[15]);
#pragma unroll
            for (int k = 0; k < 8; ++k) sb[k] += 2048u;
          };
          auto release_acc = [&]() {
            tcgen05_fence_before(); mbar_arrive(&tm_empty[bi]);
            if ((warp == 6 || warp == 10) && lane == 0) STAMP(it, 26 + j * 4);
          };
          if (!PIPE) {
#pragma unroll 1
            for (int cb = 0; cb < EC; cb += 16) {
              uint32_t ev[16], eg[16];
              tmem_ld16(tv + cb, ev); tmem_ld16(tg + cb, eg);
              tmem_wait_ld();
              if (cb + 16 == EC) release_acc();
              batch(ev, eg, cb == 0 && c0 == 0);
            }
          } else {
            uint32_t av[16], ag[16], bv[16], bg[16];
            tmem_ld16(tv, av); tmem_ld16(tg, ag);
            tmem_wait_ld();
#pragma unroll 1
            for (int cb = 0; cb < EC; cb += 32) {
              tmem_ld16(tv + cb + 16, bv); tmem_ld16(tg + cb + 16, bg);      // in flight while batch cb is computed
              batch(av, ag, cb == 0);
              tmem_wait_ld();
              if (cb + 32 < EC) { tmem_ld16(tv + cb + 32, av); tmem_ld16(tg + cb + 32, ag); }
              else release_acc();                                            // every column of this accumulator pair has been read
              batch(bv, bg, false);
              if (cb + 32 < EC) tmem_wait_ld();
            }
          }
#else
          float pv0 = 0.f, pv1 = 0.f, pg0 = 0.f, pg1 = 0.f;     // D of the two columns before the current batch
          if (WIDE && c0 > 0) {                                  // the second group starts mid-tile: columns c0-2, c0-1
            uint32_t rvv[16], rgg[16];
            tmem_ld16(tv - 16, rvv);
            tmem_ld16(tg - 16, rgg);
            tmem_wait_ld();
            pv0 = __uint_as_float(rvv[14]); pv1 = __uint_as_float(rvv[15]);
            pg0 = __uint_as_float(rgg[14]); pg1 = __uint_as_float(rgg[15]);
          }
#pragma unroll 1
          for (int cb = 0; cb < EC; cb += 16) {
            uint32_t rvv[16], rgg[16];
            tmem_ld16(tv + cb, rvv);
            tmem_ld16(tg + cb, rgg);
            tmem_wait_ld();
            if (cb + 16 == EC) { tcgen05_fence_before(); mbar_arrive(&tm_empty[bi]); if ((warp == 6 || warp == 10) && lane == 0) STAMP(it, 26 + j * 4); }
            float hv[18], hg[18];
            hv[0] = pv0; hv[1] = pv1; hg[0] = pg0; hg[1] = pg1;
#pragma unroll
            for (int i = 0; i < 16; ++i) { hv[2 + i] = __uint_as_float(rvv[i]); hg[2 + i] = __uint_as_float(rgg[i]); }
            // output columns c = cb - 2 + i, i = 1..16; c & 7 == (i + 6) & 7 because cb is a multiple of 16.  Halo rows
            // (c = 0, NTOK-1) are written too: they only feed Y's halo columns, which are never stored.  c = -1 is skipped.
            const int rowblk = ((c0 + cb) >> 3) * 1024;
#pragma unroll
            for (int i = 1; i <= 16; ++i) {
              const float dv = fmaf(wv2, hv[i + 1], fmaf(wv1, hv[i], fmaf(wv0, hv[i - 1], cv)));
              const float dg = fmaf(wg2, hg[i + 1], fmaf(wg1, hg[i], fmaf(wg0, hg[i - 1], cg)));
              const float u = fmaf(dv, tanh_approx(dg), dv);
              unsigned char* dst = sbase[(i + 6) & 7] + rowblk + ((i - 2) >> 3) * 1024;   // (i-2)>>3 is -1 for i=1, else 0/1
              if (i > 1 || cb > 0 || c0 > 0) store_elem<KIND>(dst, u);
            }
            pv0 = hv[16]; pv1 = hv[17]; pg0 = hg[16]; pg1 = hg[17];
          }
#endif
        } else {
          // ---- edge tile (or debug dump): columns outside the utterance are the conv's zero padding
          const float b1v = __ldg(p.b1 + rv), b1g = __ldg(p.b1 + rg), s1v = __ldg(p.s1inv + rv), s1g = __ldg(p.s1inv + rg);
          const float wv0 = __ldg(p.dw + rv), wv1 = __ldg(p.dw + 6 * F + rv), wv2 = __ldg(p.dw + 12 * F + rv);
          const float wg0 = __ldg(p.dw + rg), wg1 = __ldg(p.dw + 6 * F + rg), wg2 = __ldg(p.dw + 12 * F + rg);
          const float dbv = __ldg(p.dwb + rv), dbg = __ldg(p.dwb + rg);
          float cv0 = 0.f, cv1 = 0.f, cg0 = 0.f, cg1 = 0.f;     // h of the two columns before the current batch
          if (WIDE && c0 > 0) {                                  // the second group starts mid-tile: columns c0-2, c0-1
            uint32_t rvv[16], rgg[16];
            tmem_ld16(tv - 16, rvv);
            tmem_ld16(tg - 16, rgg);
            tmem_wait_ld();
            const bool ok0 = (unsigned)(tcol0 + c0 - 2) < (unsigned)p.T, ok1 = (unsigned)(tcol0 + c0 - 1) < (unsigned)p.T;
            cv0 = ok0 ? fmaf(__uint_as_float(rvv[14]), s1v, b1v) : 0.f; cv1 = ok1 ? fmaf(__uint_as_float(rvv[15]), s1v, b1v) : 0.f;
            cg0 = ok0 ? fmaf(__uint_as_float(rgg[14]), s1g, b1g) : 0.f; cg1 = ok1 ? fmaf(__uint_as_float(rgg[15]), s1g, b1g) : 0.f;
          }
#pragma unroll 1
          for (int cb = 0; cb < EC; cb += 16) {
            uint32_t rvv[16], rgg[16];
            tmem_ld16(tv + cb, rvv);
            tmem_ld16(tg + cb, rgg);
            tmem_wait_ld();
            if (cb + 16 == EC) { tcgen05_fence_before(); mbar_arrive(&tm_empty[bi]); if ((warp == 6 || warp == 10) && lane == 0) STAMP(it, 26 + j * 4); }
            float hv[18], hg[18];
            hv[0] = cv0; hv[1] = cv1; hg[0] = cg0; hg[1] = cg1;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int t = tcol0 + c0 + cb + i;
              const bool ok = (unsigned)t < (unsigned)p.T;
              hv[2 + i] = ok ? fmaf(__uint_as_float(rvv[i]), s1v, b1v) : 0.f;
              hg[2 + i] = ok ? fmaf(__uint_as_float(rgg[i]), s1g, b1g) : 0.f;
            }
            if (p.dbg_h != nullptr && live) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int c = c0 + cb + i, t = tcol0 + c;
                if (c >= 1 && c <= NTOK - 2 && t < p.T) {
                  float* d = p.dbg_h + ((size_t)n * p.T + t) * 6 * F + j * 128 + ch;
                  d[0] = hv[2 + i];
                  d[3 * F] = hg[2 + i];
                }
              }
            }
#pragma unroll
            for (int i = 1; i <= 16; ++i) {
              const int c = c0 + cb - 2 + i;
              if (c >= 1 && c <= NTOK - 2) {
                const float dv = fmaf(wv2, hv[i + 1], fmaf(wv1, hv[i], fmaf(wv0, hv[i - 1], dbv)));
                const float dg = fmaf(wg2, hg[i + 1], fmaf(wg1, hg[i], fmaf(wg0, hg[i - 1], dbg)));
                const float u = fmaf(dv, tanh_approx(dg), dv);
                store_elem<KIND>(sbase[(i + 6) & 7] + (c >> 3) * 1024, u);   // cb % 16 == 0: (cb - 2 + i) & 7
              }
            }
            cv0 = hv[16]; cv1 = hv[17]; cg0 = hg[16]; cg1 = hg[17];
          }
        }
        fence_proxy_async();
        mbar_arrive(&b2_full[bi]);
        if ((warp == 6 || warp == 10) && lane == 0) STAMP(it, 27 + j * 4);
      }
    }
  }

  // ---- teardown
  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();            // nobody leaves while a peer may still multicast into / signal this CTA
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
#undef STAMP
}

// =================================================================================================================
// k_tok<Cfg>: the other token-wise linear maps of the separator (CLA, EGA gate, attention projections, speaker
// split, fusion), same machinery as k_gcfn but on a flat token axis (no halo):
//
//   producer (warps 2-5)   builds the [NTOK x F_IN] TF32 B-operand tile: LayerNorm (PRO_LN), average-pool + LayerNorm
//                          (PRO_POOL_LN: EGA's adaptive_avg_pool1d, network.py:146), plain rounding (PRO_RAW) or the
//                          [up2(low) || skip] concatenation of the fusion conv (PRO_CONCAT, module.py:212-213)
//   stage 1                N1 steps of 128 output channels (PAIR: a value tile and a gate tile per step -> GLU)
//   single-stage kernels   the two epilogue groups apply the output op and store straight to global (coalesced:
//                          32 lanes = 32 consecutive channels of one token)
//   two-stage kernels      the epilogue groups apply the middle op (GLU / GELU), round to TF32 and write the
//                          stage-2 operand; stage 2 accumulates Y over the N1 chunks; the producer group drains Y
// -----------------------------------------------------------------------------------------------------------------
enum TokPro { PRO_LN = 0, PRO_POOL_LN = 1, PRO_RAW = 2, PRO_CONCAT = 3, PRO_SPKATTN = 4 };
enum TokOp {
  OP_BIAS = 0,    // single stage: out = D + b
  OP_GLU = 1,     // PAIR: (Dv + bv) * sigmoid(Dg + bg)         (single stage: -> global; two stage: -> stage-2 operand)
  OP_GELU = 2,    // gelu(D + b)                                 (two stage middle op)
  OP_RES = 3,     // single stage: out = res + D + b
  OP_GATE = 4,    // single stage: out = res + sigmoid(D + b) * up[m >> up_shift]  (EGA, network.py:153)
};
enum TokDrain { DRAIN_RES = 0, DRAIN_BIAS = 1 };

template <int F_IN_, int PRO_, bool PAIR_, int N1_, bool STAGE2_, int M2_, int OP_, int DRAIN_, int NTOK_, int NST_, int KIND_,
          int IO16_ = 0, int RAW_ = 0>
struct TokCfg {
  static constexpr int KIND = KIND_;
  // FP16 q|k|v rows between the projection and the attention that consumes them (both round to FP16 anyway on the
  // kind::f16 path, so this only halves the bytes): OP_BIAS kernels store halves, PRO_SPKATTN reads halves
  static constexpr bool OUT16 = (IO16_ & 1) && KIND_ == KIND_F16;   // IO16_: 1 = half output rows, 2 = half input rows
  static constexpr bool IN16 = (IO16_ & 2) && KIND_ == KIND_F16;
  using KT = KindT<KIND_>;
  static constexpr int F_IN = F_IN_, PRO = PRO_, N1 = N1_, M2 = STAGE2_ ? M2_ : 0, OP = OP_, DRAIN = DRAIN_;
  static constexpr bool PAIR = PAIR_, STAGE2 = STAGE2_;
  static constexpr int NTOK = NTOK_, NST = NST_;
  static constexpr int K1A = F_IN / KT::KSLAB;
  static constexpr int K2A = 128 / KT::KSLAB;
  static constexpr int ACC = PAIR ? 2 : 1;                 // accumulator tiles per stage-1 step
  static constexpr int ATOM_B = NTOK * 128;
  static constexpr int B1_BYTES = K1A * ATOM_B;
  static constexpr int B2_BYTES = STAGE2 ? K2A * ATOM_B : 0;
  static constexpr int A_BYTES = 128 * 128;
  // RAW > 0: the fp32 source rows of a token tile are brought into shared memory by one bulk copy (RAW buffers deep)
  // issued by the TMA warp ahead of time; the producer warps then read shared memory instead of waiting on global
  // loads (they can only keep ~16 float4 per lane in flight), and with RAW == 2 the OP_GATE epilogue takes its
  // residual rows from the same tile instead of reading them from global memory a second time.
  static constexpr int RAW = RAW_;
  // PRO_POOL_LN streams its r x NTOK source rows through a ring of RAW chunks of RAW_ROWS rows instead of one tile
  static constexpr int RAW_ROWS = (PRO_ == 1) ? 64 : NTOK_;
  static constexpr int RAW_BYTES = RAW ? RAW_ROWS * F_IN * 4 : 0;
  // the stage-1 operand is double-buffered (next tile's producer work overlaps this tile) whenever it fits
  static constexpr int NB1 = (1024 + NST * A_BYTES + 2 * B1_BYTES + 2 * B2_BYTES + RAW * RAW_BYTES + 512 <= 232448) ? 2 : 1;
  static constexpr int SMEM_BYTES = 1024 + NST * A_BYTES + NB1 * B1_BYTES + 2 * B2_BYTES + RAW * RAW_BYTES + 512;
  // weights that fit the slab ring stay resident (fetched once per CTA, see k_tok); a two-stage kernel with resident
  // weights gets a SECOND issuing thread (DUAL): stage 1 and stage 2 are then issued independently, each as soon as its own
  // operands are ready.  With one thread in the order S1(t); S2(t); S1(t+1) the stage-1 MMAs of tile t+1 waited behind
  // stage 2 of tile t, which waits for the epilogues: cla_b ran an 8.3 k clk tile of which the epilogues filled 3.3 k.
  static constexpr int S1SLABS = N1_ * (PAIR_ ? 2 : 1) * K1A, S2SLABS = STAGE2_ ? N1_ * M2_ * K2A : 0;
  static constexpr bool RESIDENT = S1SLABS + S2SLABS <= NST_;
  static constexpr bool DUAL = STAGE2_ && RESIDENT;
  static constexpr int THREADS = (DUAL ? 15 : 14) * 32;
  // stage-2 accumulators are double-buffered when TMEM has room: the drain of tile i then overlaps tile i+1's GEMM2
  static constexpr int NY = (STAGE2 && 2 * ACC * NTOK + 2 * M2 * NTOK <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS = 2 * ACC * NTOK + NY * M2 * NTOK;
  __host__ __device__ static constexpr int tm_acc(int buf, int half) { return (buf * ACC + half) * NTOK; }
  __host__ __device__ static constexpr int tm_y(int yb, int m2) { return 2 * ACC * NTOK + (yb * M2 + m2) * NTOK; }
  static_assert(NTOK % 16 == 0 && NTOK <= 256, "tile shape");
  static_assert(TMEM_COLS <= 512, "TMEM columns");
  static_assert(SMEM_BYTES <= 232448, "shared memory");
  static_assert(RAW == 0 || ((PRO == 0 || PRO == 1 || PRO == 2) && !STAGE2 && RAW <= 4), "raw tiles: PRO_LN / PRO_POOL_LN / PRO_RAW single-GEMM kernels");
};

struct TokParams {
  const float* a0;        // producer source rows [M(*pool_r), F_IN] (PRO_CONCAT: low-rate rows [M/2, F_IN/2])
  const float* a1;        // PRO_CONCAT: skip rows [M, F_IN/2]
  int pool_r;             // PRO_POOL_LN: input rows averaged per token
  int spk_T;              // PRO_SPKATTN: frames per speaker row (a0 = q|k|v rows [M, 3*F_IN], partner row = m +- spk_T)
  int tiles_per_pair;     // PRO_SPKATTN: ceil(spk_T / 64) pair tiles per utterance (set by launch_tok)
  float* out;             // [M, ld_out]
  int ld_out;
  const float* b1;        // stage-1 bias in packed row order
  const float* b2;        // stage-2 bias
  const float* s1inv;     // inverse row scales of the stage-1 / stage-2 weights (all ones for TF32)
  const float* s2inv;
  const float* res;       // residual rows [M, ld_out]
  const float* up;        // OP_GATE: pooled attention rows [M >> up_shift, ld_out] (nearest upsample by 2^up_shift)
  int up_shift;
  long long M;            // tokens
  int num_tiles;
  long long* dbg_clk;     // optional [8][64] clock64 stamps of block 0's first 8 tiles (tools/tok_timeline.py)
  int dbg_flags;          // tuning experiments: 1 = skip residual loads, 2 = skip global stores
  int out_ch;             // two-stage kernels: store only output channels [0, out_ch) of each 128-wide row (0 = all)
  // FP16 operands for GEMMs fed by the un-normalised residual stream (no pack-time range bound exists for them): the
  // FP16 launch sets *range_flag when a source value or a stage-2 operand exceeds the FP16 range (the conversions
  // saturate, so its result is finite but clamped); the TF32 launch that follows runs only if *only_if is set.
  int* range_flag;
  const int* only_if;
  int* rerun_count;       // incremented once by a launch that does run because of only_if (tests)
};

// OP_GATE: the pooled-attention values of a 16-column batch when 2^S consecutive tokens share a pooled row (S = 1..3):
// 16 >> S loads instead of 16 (the per-column form re-read every pooled row 2^S times: the decoder's T = 4000 gate
// launch took 80 us against 55 us for the same number of tokens at S = 4).  mcol0 = global token index of the tile
// half's first column; columns past `lastc` are clamped to it (their values are never stored).
template <int S>
__device__ __forceinline__ void gate_up_batch(const float* ucol, long long mcol0, int cb, int lastc, size_t ld, float (&up)[16]) {
  constexpr int n = 16 >> S;
  float u[n];
#pragma unroll
  for (int j = 0; j < n; ++j) {
    const int cj = cb + (j << S);
    u[j] = ldg_now(ucol + (size_t)((mcol0 + (cj < lastc ? cj : lastc)) >> S) * ld);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) up[i] = u[i >> S];
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1)
k_tok(const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2, const TokParams p) {
  constexpr int NTOK = C::NTOK, NST = C::NST, K1A = C::K1A, N1 = C::N1, M2 = C::M2, ACC = C::ACC;
  constexpr int ATOM_B = C::ATOM_B, A_BYTES = C::A_BYTES, B2_BYTES = C::B2_BYTES, F_IN = C::F_IN;
  constexpr int KIND = C::KIND, K2A = C::K2A, KSLAB = C::KT::KSLAB, NB1 = C::NB1, B1_BYTES = C::B1_BYTES;
  constexpr uint32_t IDESC = make_idesc<KIND>(128, NTOK);
  // single-GEMM kernels with one output tile: both epilogue groups share every token tile (half the columns each)
  constexpr bool SPLIT = !C::STAGE2 && N1 == 1;
  constexpr int NY = C::NY;
  // kernels whose weight matrices fit the slab ring keep them resident: fetched once per CTA instead of once per
  // token tile.  For the two-stage cla_b this removes the slab waits from the per-tile critical path (the in-order
  // ring could not prefetch the stage-2 slabs while the MMA warp waited for the epilogue's operand)
  constexpr int S1SLABS = C::S1SLABS, S2SLABS = C::S2SLABS;
  constexpr bool RESIDENT = C::RESIDENT, DUAL = C::DUAL;
  if (p.only_if != nullptr) {                // conditional re-run: decided before any setup, uniformly by every thread
    pdl_wait();
    if (*reinterpret_cast<const volatile int*>(p.only_if) == 0) return;
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.rerun_count != nullptr) atomicAdd(p.rerun_count, 1);
  }
  constexpr int RAW = C::RAW;
  constexpr bool RES_RAW = RAW == 2 && SPLIT && C::OP == OP_GATE;     // residual rows = the raw source tile (p.res == p.a0)

  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = sm;
  unsigned char* sB1 = sA + NST * A_BYTES;
  unsigned char* sB2 = sB1 + C::NB1 * C::B1_BYTES;
  unsigned char* sRaw = sB2 + 2 * B2_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRaw + C::RAW * C::RAW_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + NST;
  uint64_t* b1_full = a_empty + NST;       // [2]
  uint64_t* b1_empty = b1_full + 2;        // [2]
  uint64_t* tm_full = b1_empty + 2;
  uint64_t* tm_empty = tm_full + 2;
  uint64_t* b2_full = tm_empty + 2;
  uint64_t* b2_empty = b2_full + 2;
  uint64_t* y_full = b2_empty + 2;         // [2]
  uint64_t* y_empty = y_full + 2;          // [2]
  uint64_t* raw_full = y_empty + 2;        // [4]
  uint64_t* raw_empty = raw_full + 4;      // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + 4);

  const int pwarp = threadIdx.x >> 5, lane = threadIdx.x & 31;     // physical warp: fixes the TMEM lane quarter
  const int warp = (C::DUAL && pwarp == 14) ? 14 : 13 - pwarp;     // role index: critical roles get the top warp ids (14: stage-2 issue, DUAL)
#define TSTAMP(itv, slot) do { if (p.dbg_clk != nullptr && blockIdx.x == 0 && (itv) < 8) p.dbg_clk[(itv) * 64 + (slot)] = clock64(); } while (0)

  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&b1_full[i], 128); mbar_init(&b1_empty[i], 1);
      mbar_init(&tm_full[i], 1); mbar_init(&tm_empty[i], SPLIT ? 256 : 128);
      mbar_init(&b2_full[i], 128); mbar_init(&b2_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], 256); }
    for (int i = 0; i < 4; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], RES_RAW ? 384 : 128); }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_w1); if (C::STAGE2) tma_prefetch_desc(&map_w2); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  if (warp != 0) pdl_wait();                 // warp 0 waits only before it copies activation rows (RAW tiles)
  const int my_iters = ((int)blockIdx.x < p.num_tiles) ? (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  // =============================================================================== warp 0: weight slabs via TMA
  if (warp == 0) {
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      auto load = [&](const CUtensorMap* map, int c0, int c1) {
        mbar_wait(&a_empty[st], ph ^ 1, 500);
        mbar_arrive_expect_tx(&a_full[st], A_BYTES);
        tma_load_2d(map, &a_full[st], sA + st * A_BYTES, c0, c1);
        if (++st == NST) { st = 0; ph ^= 1; }
      };
      auto s1 = [&](int j) {
        for (int half = 0; half < ACC; ++half)
          for (int ka = 0; ka < K1A; ++ka) load(&map_w1, ka * KSLAB, (ACC * j + half) * 128);
      };
      auto s2 = [&](int j) {
        for (int m2 = 0; m2 < M2; ++m2)
          for (int ka = 0; ka < K2A; ++ka) load(&map_w2, j * 128 + ka * KSLAB, m2 * 128);
      };
      // issue order, identical in the MMA warp: S1(g); S2(g-1) over the running chunk index g, across tile boundaries
      const int total = RESIDENT ? (my_iters > 0 ? N1 : 0) : my_iters * N1;
      for (int g = 0; g < total; ++g) {
        s1(g % N1);
        if (C::STAGE2 && g >= 1) s2((g - 1) % N1);
      }
      if (C::STAGE2 && total > 0) s2(N1 - 1);
      if (RAW > 0) {
        static_assert(RAW == 0 || RESIDENT, "raw tiles need resident weights (this thread must be free to run ahead)");
        pdl_wait();
        uint32_t gc = 0;                                   // running chunk counter (PRO_POOL_LN)
        for (int it = 0; it < my_iters; ++it) {
          const long long m0 = (long long)((int)blockIdx.x + it * (int)gridDim.x) * NTOK;
          const int nvalid = (int)((p.M - m0) < (long long)NTOK ? (p.M - m0) : (long long)NTOK);
          if (C::PRO == PRO_POOL_LN) {
            // the tile's source rows [m0*r, (m0+nvalid)*r) are contiguous: stream them in chunks of RAW_ROWS rows
            const int rows_valid = nvalid * p.pool_r;
            for (int r0 = 0; r0 < rows_valid; r0 += C::RAW_ROWS, ++gc) {
              const uint32_t slot = gc % RAW, use = gc / RAW;
              mbar_wait(&raw_empty[slot], (use & 1) ^ 1, 511);
              const int rows = (rows_valid - r0) < C::RAW_ROWS ? (rows_valid - r0) : C::RAW_ROWS;
              const uint32_t bytes = (uint32_t)rows * (F_IN * 4);
              mbar_arrive_expect_tx(&raw_full[slot], bytes);
              bulk_load(sRaw + slot * C::RAW_BYTES, p.a0 + (m0 * p.pool_r + r0) * F_IN, bytes, &raw_full[slot]);
            }
          } else {
            const int rb = (RAW == 2) ? (it & 1) : 0;
            const uint32_t ruse = (RAW == 2) ? ((uint32_t)it >> 1) : (uint32_t)it;
            mbar_wait(&raw_empty[rb], (ruse & 1) ^ 1, 510);
            const uint32_t bytes = (uint32_t)nvalid * (F_IN * 4);
            mbar_arrive_expect_tx(&raw_full[rb], bytes);
            bulk_load(sRaw + rb * C::RAW_BYTES, p.a0 + m0 * F_IN, bytes, &raw_full[rb]);
          }
        }
      }
    }
  }
  // =============================================================================== warp 1 (and 14 with DUAL): MMA issue
  else if (warp == 1 || (DUAL && warp == 14)) {
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      int it = 0;
      auto s1 = [&](uint32_t gj, const unsigned char* b1buf) {
        const uint32_t b = gj & 1, n = gj >> 1;
        mbar_wait(&tm_empty[b], (n & 1) ^ 1, 600);
        tcgen05_fence_after();
        for (int half = 0; half < ACC; ++half) {
          const uint32_t d = tmem_base + C::tm_acc(b, half);
          for (int ka = 0; ka < K1A; ++ka) {
            const int sx = RESIDENT ? (ACC * (int)(gj % N1) + half) * K1A + ka : st;
            mbar_wait(&a_full[sx], RESIDENT ? 0u : ph, 601);
            tcgen05_fence_after();
            const uint64_t ad = make_sdesc(smem_u32(sA + sx * A_BYTES));
            const uint64_t bd = make_sdesc(smem_u32(b1buf + ka * ATOM_B));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma<KIND>(d, ad + 2 * k, bd + 2 * k, IDESC, (ka | k) != 0);
            if (!RESIDENT) {
              umma_commit(&a_empty[st]);
              if (++st == NST) { st = 0; ph ^= 1; }
            }
          }
        }
        umma_commit(&tm_full[b]);
      };
      auto s2 = [&](int j, uint32_t gj) {
        const uint32_t b = gj & 1, n = gj >> 1;
        TSTAMP(it, 2 + 3 * j);
        mbar_wait(&b2_full[b], n & 1, 602);
        TSTAMP(it, 3 + 3 * j);
        const int yb = (NY == 2) ? (it & 1) : 0;
        const uint32_t yuse = (NY == 2) ? ((uint32_t)it >> 1) : (uint32_t)it;
        if (j == 0) mbar_wait(&y_empty[yb], (yuse & 1) ^ 1, 603);
        tcgen05_fence_after();
        for (int m2 = 0; m2 < M2; ++m2) {
          const uint32_t d = tmem_base + C::tm_y(yb, m2);
          for (int ka = 0; ka < K2A; ++ka) {
            const int sx = RESIDENT ? S1SLABS + (j * M2 + m2) * K2A + ka : st;
            mbar_wait(&a_full[sx], RESIDENT ? 0u : ph, 604);
            tcgen05_fence_after();
            const uint64_t ad = make_sdesc(smem_u32(sA + sx * A_BYTES));
            const uint64_t bd = make_sdesc(smem_u32(sB2 + b * B2_BYTES + ka * ATOM_B));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma<KIND>(d, ad + 2 * k, bd + 2 * k, IDESC, (j | ka | k) != 0);
            if (!RESIDENT) {
              umma_commit(&a_empty[st]);
              if (++st == NST) { st = 0; ph ^= 1; }
            }
          }
        }
        umma_commit(&b2_empty[b]);
        TSTAMP(it, 4 + 3 * j);
      };
      const int total = my_iters * N1;
      auto do_s2 = [&](int gprev) {
        const int jj = gprev % N1;
        it = gprev / N1;                                 // s2() reads `it` for the y_empty parity
        s2(jj, (uint32_t)gprev);
        if (jj == N1 - 1) { umma_commit(&y_full[(NY == 2) ? (it & 1) : 0]); TSTAMP(it, 1); }
      };
      auto s1_step = [&](int gg) {
        const int ti = gg / N1, j = gg % N1;
        const int bb = (NB1 == 2) ? (ti & 1) : 0;
        const uint32_t bpar = (NB1 == 2) ? ((ti >> 1) & 1) : (ti & 1);
        const unsigned char* b1buf = sB1 + bb * B1_BYTES;
        if (j == 0) {
          mbar_wait(&b1_full[bb], bpar, 605);
          tcgen05_fence_after();
          TSTAMP(ti, 0);
        }
        s1((uint32_t)gg, b1buf);
        if (j == N1 - 1) { umma_commit(&b1_empty[bb]); if (!C::STAGE2) TSTAMP(ti, 1); }
      };
      if (DUAL) {
        // two issuing threads: this one all stage-1 steps, warp 14 all stage-2 steps, each in chunk order
        if (warp == 1) { for (int gg = 0; gg < total; ++gg) s1_step(gg); }
        else { for (int gg = 0; gg < total; ++gg) do_s2(gg); }
      } else if (C::STAGE2 && RESIDENT) {
        // resident weights leave the issue order free: all stage-2 steps of tile t first (its Y completes as early as
        // possible, so the drain starts while the next tile's stage-1 MMAs run), then stage 1 of tile t+1.  (With a
        // streamed ring the order must match the TMA warp's: S1(g); S2(g-1).)
        for (int j = 0; j < N1 && j < total; ++j) s1_step(j);
        for (int t = 0; t < my_iters; ++t) {
          for (int j = 0; j < N1; ++j) do_s2(t * N1 + j);
          if (t + 1 < my_iters)
            for (int j = 0; j < N1; ++j) s1_step((t + 1) * N1 + j);
        }
      } else {
        for (int gg = 0; gg < total; ++gg) {
          s1_step(gg);
          if (C::STAGE2 && gg >= 1) do_s2(gg - 1);
        }
        if (C::STAGE2 && total > 0) do_s2(total - 1);
      }
    }
  }
  // =============================================================================== warps 2-5: producer (+ drain)
  else if (warp < 6) {
    const int pw = warp - 2;
    int it = 0;
    uint32_t pool_gc = 0;       // running chunk counter of the pooled raw ring (same sequence as the TMA thread's)
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const long long m0 = (long long)tile * NTOK;
      const int bb = (NB1 == 2) ? (it & 1) : 0;
      const uint32_t bpar = (NB1 == 2) ? ((it >> 1) & 1) : (it & 1);
      unsigned char* b1buf = sB1 + bb * B1_BYTES;
      mbar_wait_group(&b1_empty[bb], bpar ^ 1, 701, 1, pw == 0);
      if (warp == 2 && lane == 0) TSTAMP(it, 16);
      {
        const long long M = p.M;
        const bool track = KIND == KIND_F16 && p.range_flag != nullptr;      // raw-stream GEMMs: see TokParams::range_flag
        float amax = 0.f;
        (void)track; (void)amax;
        if constexpr (C::PRO == PRO_POOL_LN && RAW > 0) {
          // pooled rows from the chunk ring: a warp owns one token at a time (lane = float4 column, so a source row is
          // one conflict-free 512-byte LDS.128 per 128 channels), sums its r rows, LayerNorm over the warp
          constexpr int NV = F_IN / 128;
          const int pr = p.pool_r, tpc = C::RAW_ROWS / pr;          // tokens per chunk (host guarantees RAW_ROWS % r == 0)
          const int nv = (int)((M - m0) < (long long)NTOK ? (M - m0) : (long long)NTOK);
          const float inv = 1.0f / (float)pr;
          for (int t0 = 0; t0 < nv; t0 += tpc, ++pool_gc) {
            const uint32_t slot = pool_gc % RAW, use = pool_gc / RAW;
            mbar_wait_group(&raw_full[slot], use & 1, 703, 1, pw == 0);
            const float4* raw = reinterpret_cast<const float4*>(sRaw + slot * C::RAW_BYTES);
            // U tokens per warp in flight: one token at a time was a ~700 clk dependent chain (shared loads, two warp
            // reductions, store) - 22 k clk for a 128-token tile even without pooling (r = 1), 2/3 of the smallest launches
            constexpr int U = 4;
            for (int tk0 = pw; tk0 < tpc; tk0 += 4 * U) {
              float4 a[U][NV];
              bool live[U];
#pragma unroll
              for (int u = 0; u < U; ++u) {
                const int tk = tk0 + 4 * u;
                live[u] = tk < tpc && t0 + tk < nv;
#pragma unroll
                for (int k = 0; k < NV; ++k) a[u][k] = make_float4(0.f, 0.f, 0.f, 0.f);
              }
              for (int i = 0; i < pr; ++i) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                  if (live[u]) {
#pragma unroll
                    for (int k = 0; k < NV; ++k) {
                      const float4 v = raw[((tk0 + 4 * u) * pr + i) * (F_IN / 4) + lane + 32 * k];
                      a[u][k].x += v.x; a[u][k].y += v.y; a[u][k].z += v.z; a[u][k].w += v.w;
                    }
                  }
                }
              }
              float sum[U];
#pragma unroll
              for (int u = 0; u < U; ++u) {
                sum[u] = 0.f;
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                  a[u][k].x *= inv; a[u][k].y *= inv; a[u][k].z *= inv; a[u][k].w *= inv;
                  sum[u] += a[u][k].x + a[u][k].y + a[u][k].z + a[u][k].w;
                }
              }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int u = 0; u < U; ++u) sum[u] += __shfl_xor_sync(0xffffffffu, sum[u], o);
              float qq[U];
#pragma unroll
              for (int u = 0; u < U; ++u) {
                const float mean = sum[u] * (1.0f / F_IN);
                qq[u] = 0.f;
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                  a[u][k].x -= mean; a[u][k].y -= mean; a[u][k].z -= mean; a[u][k].w -= mean;
                  qq[u] += a[u][k].x * a[u][k].x + a[u][k].y * a[u][k].y + a[u][k].z * a[u][k].z + a[u][k].w * a[u][k].w;
                }
              }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int u = 0; u < U; ++u) qq[u] += __shfl_xor_sync(0xffffffffu, qq[u], o);
#pragma unroll
              for (int u = 0; u < U; ++u) {
                const int tk = tk0 + 4 * u;
                if (tk < tpc) {             // tokens of this chunk past the end of the tensor are written as zero rows
                  const float rstd = live[u] ? rsqrtf(qq[u] * (1.0f / F_IN) + kLnEps) : 0.f;
#pragma unroll
                  for (int k = 0; k < NV; ++k) store_c4<KIND>(b1buf, ATOM_B, t0 + tk, lane + 32 * k, a[u][k], rstd);
                }
              }
            }
            mbar_arrive(&raw_empty[slot]);
          }
        } else if constexpr (C::PRO == PRO_POOL_LN) {
          const float4* x4 = reinterpret_cast<const float4*>(p.a0);
          const int pr = p.pool_r;
          produce_rows<KIND, F_IN, NTOK, true>(b1buf, ATOM_B, pw, lane, pr, [&](int r, int c4, int ps) {
            const long long m = m0 + r;
            return (m < M) ? __ldg(x4 + ((size_t)m * pr + ps) * (F_IN / 4) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
          });
        } else if constexpr (C::PRO == PRO_SPKATTN) {
          const int pb = tile / p.tiles_per_pair, t0 = (tile - pb * p.tiles_per_pair) * 64;
          const int nh = (p.spk_T - t0) < 64 ? (p.spk_T - t0) : 64;
          produce_spk_pair<KIND, F_IN, NTOK, C::IN16>(b1buf, ATOM_B, pw, lane, p.a0, p.spk_T,
                                                      (long long)(2 * pb) * p.spk_T + t0, nh);
        } else if constexpr (C::PRO == PRO_CONCAT) {
          // channels [0, F_IN/2) come from the half-rate tensor (nearest upsample), [F_IN/2, F_IN) from the skip
          constexpr int H4 = F_IN / 8;
          const float4* lo = reinterpret_cast<const float4*>(p.a0);
          const float4* sk = reinterpret_cast<const float4*>(p.a1);
          produce_rows<KIND, F_IN, NTOK, false>(b1buf, ATOM_B, pw, lane, 1, [&](int r, int c4, int) {
            const long long m = m0 + r;
            if (m >= M) return make_float4(0.f, 0.f, 0.f, 0.f);
            return c4 < H4 ? __ldg(lo + (size_t)(m >> 1) * H4 + c4) : __ldg(sk + (size_t)m * H4 + (c4 - H4));
          }, track ? &amax : nullptr);
        } else if constexpr (RAW > 0) {
          const int rb = (RAW == 2) ? (it & 1) : 0;
          const uint32_t ruse = (RAW == 2) ? ((uint32_t)it >> 1) : (uint32_t)it;
          mbar_wait_group(&raw_full[rb], ruse & 1, 702, 1, pw == 0);
          const float4* x4 = reinterpret_cast<const float4*>(sRaw + rb * C::RAW_BYTES);
          const int nv = (int)((M - m0) < (long long)NTOK ? (M - m0) : (long long)NTOK);
          produce_rows<KIND, F_IN, NTOK, C::PRO == PRO_LN>(b1buf, ATOM_B, pw, lane, 1, [&](int r, int c4, int) {
            return (r < nv) ? x4[r * (F_IN / 4) + c4] : make_float4(0.f, 0.f, 0.f, 0.f);
          });
          mbar_arrive(&raw_empty[rb]);          // generic-proxy reads are done (values are in registers / stored)
        } else if constexpr (C::IN16) {
          // FP16 source rows (k_cla_front's output): widened here and narrowed again by the operand store - exact
          const uint2* xh = reinterpret_cast<const uint2*>(p.a0);
          produce_rows<KIND, F_IN, NTOK, C::PRO == PRO_LN>(b1buf, ATOM_B, pw, lane, 1, [&](int r, int c4, int) {
            const long long m = m0 + r;
            if (m >= M) return make_float4(0.f, 0.f, 0.f, 0.f);
            const uint2 hq = __ldg(xh + (size_t)m * (F_IN / 4) + c4);
            const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&hq.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&hq.y));
            return make_float4(lo.x, lo.y, hi.x, hi.y);
          });
        } else {
          const float4* x4 = reinterpret_cast<const float4*>(p.a0);
          produce_rows<KIND, F_IN, NTOK, C::PRO == PRO_LN>(b1buf, ATOM_B, pw, lane, 1, [&](int r, int c4, int) {
            const long long m = m0 + r;
            return (m < M) ? __ldg(x4 + (size_t)m * (F_IN / 4) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
          }, (C::PRO == PRO_RAW && track) ? &amax : nullptr);
        }
        // a raw source value beyond the FP16 range was clamped by the operand store (NaN compares false: not > - but a
        // NaN source poisons the result on every path alike)
        if (KIND == KIND_F16 && track && !(amax <= 65504.0f)) atomicOr(p.range_flag, 1);
      }
      fence_proxy_async();
      mbar_arrive(&b1_full[bb]);
      if (warp == 2 && lane == 0) TSTAMP(it, 17);
    }
  }
  // =============================================================================== warps 6-13: epilogue groups
  else {
    const int eg = (warp - 6) >> 2;
    const bool gl = ((warp - 6) & 3) == 0;        // first warp of the group (polls for the group with SEPREF_FANOUT_WAITS)
    const int q = pwarp & 3;
    const int ch = q * 32 + lane;
    unsigned char* sbase[8];
    make_sbase<KIND>(sbase, sB2 + eg * B2_BYTES, ATOM_B, q, lane);
    uint32_t sbase32[8];                          // the same as 32-bit shared addresses (constant offsets fold into STS [R + imm])
#pragma unroll
    for (int m = 0; m < 8; ++m) sbase32[m] = smem_u32(sbase[m]);
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    // the row stride of the output / residual / pooled rows is a compile-time constant (address math folds into
    // immediates; with a run-time stride the compiler re-derived 64-bit addresses from the parameter bank per store)
    constexpr int ld = (C::STAGE2 ? M2 : N1) * 128;
    // two-stage kernels: both groups drain tile i's Y (half the columns each) at the start of iteration i+1; the
    // residual values of the first output tile are requested before waiting for the accumulator (single call site)
    constexpr int SP = ((NTOK / 2 + 15) / 16) * 16;           // group 0: columns [0, SP), group 1: [SP, NTOK)
    auto drain = [&](int tile, int it) {
      const long long m0 = (long long)tile * NTOK;
      const int nvalid = (int)((p.M - m0) < (long long)NTOK ? (p.M - m0) : (long long)NTOK);
      const int c0 = eg ? SP : 0, nc = eg ? (NTOK - SP) : SP;
      const int yb = (NY == 2) ? (it & 1) : 0;
      const uint32_t yuse = (NY == 2) ? ((uint32_t)it >> 1) : (uint32_t)it;
      // last valid column relative to c0; negative when this group's half lies wholly past the end of the token axis
      // (the clamped address then points at the tile's last valid row, still inside the tensor)
      const int lastc = nvalid - 1 - c0;
      float* ocol0 = p.out + (m0 * ld + ch) + c0 * ld;
      const bool chok = p.out_ch == 0 || ch < p.out_ch;       // padded output rows (fused decoder: 16 of 128 channels)
      const float* rcol0 = (C::DRAIN == DRAIN_RES) ? p.res + (m0 * ld + ch) + c0 * ld : nullptr;
      float xin[SP];
      if (C::DRAIN == DRAIN_RES) {
#pragma unroll
        for (int i = 0; i < SP; ++i) xin[i] = ldg_now(rcol0 + (i < lastc ? i : lastc) * ld);
      }
      const float bias0 = ldg_now(p.b2 + ch), s2i0 = ldg_now(p.s2inv + ch);      // ahead of the wait
      mbar_wait_group(&y_full[yb], yuse & 1, 700, 2 + eg, gl);
      tcgen05_fence_after();
#pragma unroll
      for (int m2 = 0; m2 < (C::STAGE2 ? M2 : 1); ++m2) {
        const float bias = m2 == 0 ? bias0 : __ldg(p.b2 + m2 * 128 + ch), s2i = m2 == 0 ? s2i0 : __ldg(p.s2inv + m2 * 128 + ch);
        float* ocol = ocol0 + m2 * 128;
        if (C::DRAIN == DRAIN_RES && m2 > 0) {
#pragma unroll
          for (int i = 0; i < SP; ++i) xin[i] = ldg_now(rcol0 + m2 * 128 + (i < lastc ? i : lastc) * ld);
        }
#pragma unroll
        for (int cb = 0; cb < SP; cb += 16) {
          if (cb < nc) {
            uint32_t r[16];
            tmem_ld16(tmem_base + tlane + C::tm_y(yb, m2) + c0 + cb, r);
            tmem_wait_ld();
            if (m2 == M2 - 1 && cb + 16 >= nc) { tcgen05_fence_before(); mbar_arrive(&y_empty[yb]); }
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c0 + cb + i < nvalid && chok)
                ocol[(cb + i) * ld] = fmaf(__uint_as_float(r[i]), s2i, (C::DRAIN == DRAIN_RES ? xin[cb + i] : 0.f) + bias);
          }
        }
      }
      if ((warp == 6 || warp == 10) && lane == 0) TSTAMP(it + 1, 18);
    };
    // LATE (two issuing threads and a double-buffered Y): tile t-1 is drained AFTER this group's chunk of tile t instead
    // of before it.  Draining first put the wait for Y(t-1) - i.e. for stage 2 of tile t-1, which itself waits for the
    // epilogues of t-1 - in front of every tile: epilogue -> stage 2 -> drain -> epilogue ran strictly one after the
    // other (8.8 k clk per cla_b tile with 3.3 k of epilogue work).  Y(t-1) is complete long before its late drain, and
    // stage 2 of tile t accumulates into the other Y buffer.  Measured: cla_b 1.55 -> 1.40 ms per forward (7.5 k clk per
    // tile; the drain's residual loads are now exposed.  Also measured, none of them a gain: an L2 prefetch of those rows
    // by the producers (1.40), requesting them into registers before the chunk epilogue (1.54 - the epilogue itself slows
    // down by more than the drain gains), bounded-suspension waits for the two issuing threads (1.40)).
    constexpr bool LATE = DUAL && NY == 2;
    for (int it = 0; it <= my_iters; ++it) {
      if (C::STAGE2 && !LATE && it > 0) drain((int)blockIdx.x + (it - 1) * (int)gridDim.x, it - 1);
      if (!LATE && it == my_iters) break;
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      long long m0 = (long long)tile * NTOK;
      int nvalid = (int)((p.M - m0) < (long long)NTOK ? (p.M - m0) : (long long)NTOK);   // columns of this tile that are tokens
      long long mb0 = 0;      // pair tiles (PRO_SPKATTN): columns [0,64) are rows mb0.. of speaker 0, [64,128) rows mb0+T..
      int nh = 0;
      if (C::PRO == PRO_SPKATTN) {
        const int pb = tile / p.tiles_per_pair, t0 = (tile - pb * p.tiles_per_pair) * 64;
        nh = (p.spk_T - t0) < 64 ? (p.spk_T - t0) : 64;
        mb0 = (long long)(2 * pb) * p.spk_T + t0;
        // split epilogue: group eg owns columns [64 eg, 64 eg + 64) = speaker eg; shift the row base accordingly
        m0 = eg ? mb0 + p.spk_T - 64 : mb0;
        nvalid = eg ? 64 + nh : nh;
      }
      if (SPLIT) {
        // ---- one output tile per token tile: this group owns columns [c0, c0 + NTOK/2); the residual values of the
        // whole half are requested before waiting for the accumulator, so their latency overlaps the GEMM
        constexpr int HC = NTOK / 2;
        const int c0 = eg * HC;
        const uint32_t b = (uint32_t)it & 1, nuse = (uint32_t)it >> 1;
        const float bv = ldg_now(p.b1 + ch), sv = ldg_now(p.s1inv + ch);       // issued here, ahead of the accumulator wait
        const float bg = C::PAIR ? ldg_now(p.b1 + 128 + ch) : 0.f, sg = C::PAIR ? ldg_now(p.s1inv + 128 + ch) : 0.f;
        float* ocol = p.out + (m0 * ld + ch) + c0 * ld;
        // Prefetched operands: addresses of columns past the end of the token axis are clamped to the last valid
        // column instead of predicated - a select after each load would make the scheduler wait for it before issuing
        // the next one (measured); the clamped values are never stored.
          // last valid column relative to c0; negative when this group's half lies wholly past the end of the token axis
        // (the clamped address then points at the tile's last valid row, still inside the tensor)
        const int lastc = nvalid - 1 - c0;
        float res[HC];
        if (!RES_RAW && (C::OP == OP_RES || C::OP == OP_GATE)) {
          const float* rcol = p.res + (m0 * ld + ch) + c0 * ld;
#pragma unroll
          for (int i = 0; i < HC; ++i) res[i] = ldg_now(rcol + (i < lastc ? i : lastc) * ld);
        }
        const bool shared_up = (C::OP == OP_GATE) && p.up_shift >= 4;   // 16 consecutive tokens share one pooled row
        const float* ucol = (C::OP == OP_GATE) ? p.up + ch : nullptr;
        float upre[HC / 16];
        if (shared_up) {
#pragma unroll
          for (int i = 0; i < HC / 16; ++i) {
            const int cc = 16 * i < lastc ? 16 * i : lastc;
            upre[i] = ldg_now(ucol + (size_t)((m0 + c0 + cc) >> p.up_shift) * ld);
          }
        }
        mbar_wait_group(&tm_full[b], nuse & 1, 800, 2 + eg, gl);
        if ((warp == 6 || warp == 10) && lane == 0) TSTAMP(it, 24 + 4 * eg);
        tcgen05_fence_after();
        // residual rows straight from the raw source tile the bulk copy left in shared memory (thread = channel:
        // a warp reads 128 contiguous bytes per column, conflict-free)
        const int rb = (RAW == 2) ? (it & 1) : 0;
        const float* rawcol = nullptr;
        if (RES_RAW) {
          mbar_wait_group(&raw_full[rb], ((uint32_t)it >> 1) & 1, 803, 2 + eg, gl);
          rawcol = reinterpret_cast<const float*>(sRaw + rb * C::RAW_BYTES) + c0 * F_IN + ch;
        }
        const uint32_t tv = tmem_base + tlane + C::tm_acc(b, 0) + c0, tg = tmem_base + tlane + C::tm_acc(b, C::PAIR ? 1 : 0) + c0;
#pragma unroll
        for (int cb = 0; cb < HC; cb += 16) {
          uint32_t rv[16], rg[16];
          tmem_ld16(tv + cb, rv);
          if (C::PAIR) tmem_ld16(tg + cb, rg);
          float up[16];
          if (C::OP == OP_GATE) {
            if (shared_up) {
#pragma unroll
              for (int i = 0; i < 16; ++i) up[i] = upre[cb / 16];
            } else if (p.up_shift == 3) {
              gate_up_batch<3>(ucol, m0 + c0, cb, lastc, (size_t)ld, up);
            } else if (p.up_shift == 2) {
              gate_up_batch<2>(ucol, m0 + c0, cb, lastc, (size_t)ld, up);
            } else if (p.up_shift == 1) {
              gate_up_batch<1>(ucol, m0 + c0, cb, lastc, (size_t)ld, up);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int cc = (cb + i) < lastc ? (cb + i) : lastc;
                up[i] = ldg_now(ucol + (size_t)((m0 + c0 + cc) >> p.up_shift) * ld);
              }
            }
          }
          tmem_wait_ld();
          if (cb + 16 == HC) { tcgen05_fence_before(); mbar_arrive(&tm_empty[b]); }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float val = fmaf(__uint_as_float(rv[i]), sv, bv);
            if (C::OP == OP_GLU) {              // v * sigmoid(g) = (v/2) * (1 + tanh(g/2))
              const float gt = fmaf(__uint_as_float(rg[i]), sg, bg);
              const float hv = 0.5f * val;
              val = fmaf(hv, tanh_approx(0.5f * gt), hv);
            } else if (C::OP == OP_GELU) {
              val = gelu_tanh_fit(val);
            } else if (C::OP == OP_RES) {
              val += res[cb + i];
            } else if (C::OP == OP_GATE) {      // res + sigmoid(val) * up
              const float hu = 0.5f * up[i];
              val = (RES_RAW ? rawcol[(cb + i) * F_IN] : res[cb + i]) + fmaf(hu, tanh_approx(0.5f * val), hu);
            }
            if (c0 + cb + i < nvalid) ocol[(cb + i) * ld] = val;
          }
        }
        if (RES_RAW) mbar_arrive(&raw_empty[rb]);
        if ((warp == 6 || warp == 10) && lane == 0) TSTAMP(it, 27 + 4 * eg);
        continue;
      }
#pragma unroll 1
      for (int j = 0; j < N1; ++j) {
        if (LATE && it == my_iters) break;               // LATE: the extra iteration only drains the last tile
        const uint32_t gj = (uint32_t)it * N1 + j;
        if ((int)(gj & 1) != eg) continue;
        const uint32_t nuse = gj >> 1;
        const float bv = ldg_now(p.b1 + (ACC * j) * 128 + ch), sv = ldg_now(p.s1inv + (ACC * j) * 128 + ch);   // ahead of the wait
        const float bg = C::PAIR ? ldg_now(p.b1 + (ACC * j + 1) * 128 + ch) : 0.f;
        const float sg = C::PAIR ? ldg_now(p.s1inv + (ACC * j + 1) * 128 + ch) : 0.f;
        mbar_wait_group(&tm_full[eg], nuse & 1, 800, 2 + eg, gl);
        if ((warp == 6 || warp == 10) && lane == 0 && j < 4) TSTAMP(it, 24 + j * 4);
        if (C::STAGE2) mbar_wait_group(&b2_empty[eg], (nuse & 1) ^ 1, 801, 2 + eg, gl);
        tcgen05_fence_after();
        const uint32_t tv = tmem_base + tlane + C::tm_acc(eg, 0), tg = tmem_base + tlane + C::tm_acc(eg, C::PAIR ? 1 : 0);
        // single stage: this thread's column of the output / residual (channel j*128+ch), 32-bit offsets from here
        float* ocol = p.out + (m0 * ld + j * 128 + ch);
        const float* rcol = (C::OP == OP_RES || C::OP == OP_GATE) ? p.res + (m0 * ld + j * 128 + ch) : nullptr;
        (void)mb0; (void)nh;
        const float* ucol = (C::OP == OP_GATE) ? p.up + (j * 128 + ch) : nullptr;
        constexpr bool TRACK2 = C::STAGE2 && C::OP == OP_GLU && KIND == KIND_F16;   // stage-2 operand of a raw-stream GEMM
        float umax = 0.f;
        (void)umax;
#pragma unroll 1
        for (int cb = 0; cb < NTOK; cb += 16) {
          uint32_t rv[16], rg[16];
          tmem_ld16(tv + cb, rv);
          if (C::PAIR) tmem_ld16(tg + cb, rg);
          float aux[16], aux2[16];
          if (C::PRO == PRO_SPKATTN) {   // pair tile: this batch of 16 columns belongs to speaker cb / 64
            const long long mrow = (cb < 64) ? mb0 : mb0 + p.spk_T - 64;
            ocol = p.out + (mrow * ld + j * 128 + ch);
            rcol = p.res + (mrow * ld + j * 128 + ch);
            nvalid = (cb < 64) ? nh : 64 + nh;
          }
          if (!C::STAGE2 && (C::OP == OP_RES || C::OP == OP_GATE)) {
#pragma unroll
            for (int i = 0; i < 16; ++i) aux[i] = (cb + i < nvalid) ? ldg_now(rcol + (cb + i) * ld) : 0.f;
            if (C::OP == OP_GATE) {
              if (p.up_shift >= 4) {
                const float u = (cb < nvalid) ? __ldg(ucol + (size_t)((m0 + cb) >> p.up_shift) * ld) : 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) aux2[i] = u;
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) aux2[i] = (cb + i < nvalid) ? __ldg(ucol + (size_t)((m0 + cb + i) >> p.up_shift) * ld) : 0.f;
              }
            }
          }
          tmem_wait_ld();
          if (cb + 16 == NTOK) { tcgen05_fence_before(); mbar_arrive(&tm_empty[eg]); }
          const int rowblk = (cb >> 3) * 1024;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float val = fmaf(__uint_as_float(rv[i]), sv, bv);
            if (C::OP == OP_GLU) {
              const float gt = fmaf(__uint_as_float(rg[i]), sg, bg);
              const float hv = 0.5f * val;
              val = fmaf(hv, tanh_approx(0.5f * gt), hv);
            } else if (C::OP == OP_GELU) {
              val = gelu_tanh_fit(val);
            } else if (C::OP == OP_RES) {
              val += aux[i];
            } else if (C::OP == OP_GATE) {
              const float hu = 0.5f * aux2[i];
              val = aux[i] + fmaf(hu, tanh_approx(0.5f * val), hu);
            }
            if (C::STAGE2) {
              if (TRACK2) umax = fmaxf(umax, fabsf(val));
              sts_elem<KIND>(sbase32[i & 7] + (uint32_t)(rowblk + (i >> 3) * 1024), val);   // cb is a multiple of 16: (cb + i) & 7 == i & 7
            } else {
              if (cb + i < nvalid) {
                if (C::OUT16) reinterpret_cast<uint16_t*>(p.out)[(m0 * ld + j * 128 + ch) + (cb + i) * ld] = f16_sat(val);
                else ocol[(cb + i) * ld] = val;
              }
            }
          }
        }
        if (TRACK2 && p.range_flag != nullptr && !(umax <= 65504.0f)) atomicOr(p.range_flag, 1);
        if (C::STAGE2) { fence_proxy_async(); mbar_arrive(&b2_full[eg]); }
        if ((warp == 6 || warp == 10) && lane == 0 && j < 4) TSTAMP(it, 27 + j * 4);
        if (lane == 0) TSTAMP(it, 48 + (warp - 6));       // per-warp completion (spread inside a group)
      }
      if (LATE && it > 0) drain((int)blockIdx.x + (it - 1) * (int)gridDim.x, it - 1);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
#undef TSTAMP
}

// ------------------------------------------------------------------------------------------------ host side
static thread_local char g_tc_err[256] = "";
inline const char* last_error() { return g_tc_err; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

inline int init(int /*F*/) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      snprintf(g_tc_err, sizeof(g_tc_err), "cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(e));
      return -1;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  return 0;
}

// 2-D row-major [rows, cols] weight matrix of `kind`; box = [box_rows x one 128-byte k slab], SWIZZLE_128B
inline int make_weight_map(CUtensorMap* map, const void* ptr, int kind, int rows, int cols, int box_rows = 128) {
  const int es = kind == KIND_F16 ? 2 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * es};
  cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, kind == KIND_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                        const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(g_tc_err, sizeof(g_tc_err), "cuTensorMapEncodeTiled failed (%d)", (int)r); return -1; }
  return 0;
}

inline int prepare_gcfn(GcfnPack& g, int F) {
  for (int kind = 0; kind < 2; ++kind)
    for (int i = 0; i < 3; ++i) {
      if (make_weight_map(&g.map_w1[kind][i], g.w1[kind], kind, 6 * F, F, 128 >> i)) return -1;
      if (make_weight_map(&g.map_w2[kind][i], g.w2[kind], kind, F, 3 * F, 128 >> i)) return -1;
    }
  return 0;
}

template <int F, int CL, int KIND, bool WIDE = false>
inline int launch_gcfn_t(const GcfnPack& g, GcfnParams p, int sm_count, cudaStream_t st) {
  using TR = GcfnTraits<F, KIND, WIDE>;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaError_t e;
  static bool attr_set[16] = {};      // the opt-in shared-memory size is a per-device function attribute: set it once
  if (!attr_set[dev & 15]) {
    e = cudaFuncSetAttribute(k_gcfn<F, CL, KIND, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TR::SMEM_BYTES);
    if (e != cudaSuccess) { snprintf(g_tc_err, sizeof(g_tc_err), "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
    attr_set[dev & 15] = true;
  }
  p.tiles_per_row = (p.T + TR::NV - 1) / TR::NV;
  p.num_tiles = p.rows * p.tiles_per_row;
  p.dwf = g.dwf[KIND]; p.s1inv = g.s1inv[KIND]; p.s2inv = g.s2inv[KIND];
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((sm_count / CL) * CL);
  cfg.blockDim = dim3(TR::THREADS);
  cfg.dynamicSmemBytes = TR::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  // persistent grid = clusters that can be co-resident (a cluster of 4 cannot use every SM of the 148)
  static thread_local int max_clusters[16] = {0};
  if (max_clusters[dev & 15] == 0) {
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, k_gcfn<F, CL, KIND, WIDE>, &cfg);
    if (e != cudaSuccess || n <= 0) { cudaGetLastError(); n = sm_count / CL; }
    max_clusters[dev & 15] = n;
  }
  int grid = max_clusters[dev & 15] * CL;
  const int want = ((p.num_tiles + CL - 1) / CL) * CL;
  if (grid > want) grid = want;
  p.iters = (p.num_tiles + grid - 1) / grid;
  cfg.gridDim = dim3(grid);
  constexpr int mi = (CL == 1) ? 0 : (CL == 2 ? 1 : 2);
  e = cudaLaunchKernelEx(&cfg, k_gcfn<F, CL, KIND, WIDE>, g.map_w1[KIND][mi], g.map_w2[KIND][mi], p);
  if (e != cudaSuccess) { snprintf(g_tc_err, sizeof(g_tc_err), "k_gcfn launch: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

template <int F, int KIND, bool WIDE = false>
inline int launch_gcfn_c(const GcfnPack& g, const GcfnParams& p, int sm_count, cudaStream_t st, int cluster) {
  if (cluster == 1) return launch_gcfn_t<F, 1, KIND, WIDE>(g, p, sm_count, st);
  if (cluster == 2) return launch_gcfn_t<F, 2, KIND, WIDE>(g, p, sm_count, st);
  return launch_gcfn_t<F, 4, KIND, WIDE>(g, p, sm_count, st);
}

inline int launch_gcfn(const GcfnPack& g, const float* x, float* y, int rows, int T, int F, int sm_count, cudaStream_t st,
                       float* dbg_h = nullptr, long long* dbg_clk = nullptr, int cluster = 2, int kind = KIND_TF32,
                       bool wide = false) {
  GcfnParams p{};
  p.x = x; p.y = y; p.b1 = g.b1; p.dw = g.dw; p.dwb = g.dwb; p.cb = g.cb; p.b2 = g.b2;
  p.rows = rows; p.T = T; p.dbg_h = dbg_h; p.dbg_clk = dbg_clk;
  if (wide && F == 128 && kind == KIND_F16) return launch_gcfn_c<128, KIND_F16, true>(g, p, sm_count, st, cluster);
  if (F == 128) return kind == KIND_F16 ? launch_gcfn_c<128, KIND_F16>(g, p, sm_count, st, cluster)
                                        : launch_gcfn_c<128, KIND_TF32>(g, p, sm_count, st, cluster);
  return kind == KIND_F16 ? launch_gcfn_c<256, KIND_F16>(g, p, sm_count, st, cluster)
                          : launch_gcfn_c<256, KIND_TF32>(g, p, sm_count, st, cluster);
}

// ---- generic token-GEMM launchers ---------------------------------------------------------------------------------
struct TcLin {          // one weight matrix [rows, cols] per operand kind (+ bias / inverse row scale in the same row order)
  const void* w[2] = {nullptr, nullptr};
  const float* sinv[2] = {nullptr, nullptr};
  const float* b = nullptr;
  int rows = 0, cols = 0;
  alignas(64) CUtensorMap map[2];
};
inline int prepare_lin(TcLin& l) {
  for (int kind = 0; kind < 2; ++kind)
    if (make_weight_map(&l.map[kind], l.w[kind], kind, l.rows, l.cols)) return -1;
  return 0;
}

template <int F, int K> using CfgClaA = TokCfg<F, PRO_LN, true, F / 128, false, 0, OP_GLU, 0, 128, (F == 128 ? (K == KIND_F16 ? 4 : 6) : 5), K, 0, (F == 128 && K == KIND_F16 ? 1 : 0)>;
template <int F, int K> using CfgClaB = TokCfg<F, PRO_RAW, false, 2 * F / 128, true, F / 128, OP_GELU, DRAIN_RES, (F == 128 ? 96 : 80), (F == 128 ? (K == KIND_F16 ? 8 : 5) : 4), K>;
// cla_b reading FP16 rows (the fused k_cla_front writes d as FP16)
template <int F, int K> using CfgClaB16 = TokCfg<F, PRO_RAW, false, 2 * F / 128, true, F / 128, OP_GELU, DRAIN_RES, (F == 128 ? 96 : 80), (F == 128 ? (K == KIND_F16 ? 8 : 5) : 4), K, 2>;
template <int F, int K> using CfgGate = TokCfg<F, PRO_LN, false, F / 128, false, 0, OP_GATE, 0, 128, (F == 128 ? (K == KIND_F16 ? 2 : 4) : 5), K, 0, (F == 128 && K == KIND_F16 ? 2 : 0)>;
template <int F, int K> using CfgQkvPool = TokCfg<F, PRO_POOL_LN, false, 3 * F / 128, false, 0, OP_BIAS, 0, 128, (F == 128 ? 6 : 5), K>;
template <int F, int K> using CfgQkv = TokCfg<F, PRO_LN, false, 3 * F / 128, false, 0, OP_BIAS, 0, 128, (F == 128 ? 6 : 5), K>;
template <int F, int K> using CfgProj = TokCfg<F, PRO_RAW, false, F / 128, false, 0, OP_BIAS, 0, 128, (F == 128 ? 6 : 5), K>;
template <int F, int K> using CfgProjRes = TokCfg<F, PRO_RAW, false, F / 128, false, 0, OP_RES, 0, 128, (F == 128 ? 6 : 5), K>;
template <int F, int K> using CfgSpkProj = TokCfg<F, PRO_SPKATTN, false, F / 128, false, 0, OP_RES, 0, 128, (F == 128 ? 6 : 5), K, 2>;
template <int F, int K> using CfgQkvPool16 = TokCfg<F, PRO_POOL_LN, false, 3 * F / 128, false, 0, OP_BIAS, 0, 128, (F == 128 ? 6 : 5), K, 1>;
// the same with the source rows streamed through a ring of three 64-row bulk copies (pool factors that divide 64)
template <int F, int K> using CfgQkvPool16R = TokCfg<F, PRO_POOL_LN, false, 3 * F / 128, false, 0, OP_BIAS, 0, 128, (F == 128 ? 6 : 5), K, 1, (F == 128 && K == KIND_F16 ? 3 : 0)>;
template <int F, int K> using CfgQkv16 = TokCfg<F, PRO_LN, false, 3 * F / 128, false, 0, OP_BIAS, 0, 128, (F == 128 ? 6 : 5), K, 1, (F == 128 && K == KIND_F16 ? 1 : 0)>;
template <int F, int K> using CfgSplit = TokCfg<F, PRO_RAW, true, 4 * F / 128, true, 2 * F / 128, OP_GLU, DRAIN_BIAS, (F == 128 ? 80 : 64), (F == 128 ? 6 : 5), K>;
template <int F, int K> using CfgFuse = TokCfg<2 * F, PRO_CONCAT, false, F / 128, false, 0, OP_BIAS, 0, (F == 128 ? 128 : 64), 5, K>;
// model shell (kernels_shell.cuh): FeatureProjector's 1x1 conv on the normalised encoder rows (K = 256 -> F) ...
template <int F, int K> using CfgEncProj = TokCfg<256, PRO_RAW, false, F / 128, false, 0, OP_BIAS, 0, 128, 5, K>;
// ... and OutputLayer (Linear F -> 4F, GLU, Linear 2F -> 256) with the AudioDecoder folded into the second matrix (16 rows of 128)
template <int F, int K> using CfgOutDec = TokCfg<F, PRO_RAW, true, 2 * F / 128, true, 1, OP_GLU, DRAIN_BIAS, (F == 128 ? 80 : 64), (F == 128 ? 6 : 5), K>;

template <class C>
inline int launch_tok(const TcLin& l1, const TcLin* l2, TokParams p, int sm_count, cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  cudaError_t e;
  static bool attr_set[16] = {};      // per-device function attribute: set once per configuration
  if (!attr_set[dev & 15]) {
    e = cudaFuncSetAttribute(k_tok<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) { snprintf(g_tc_err, sizeof(g_tc_err), "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
    attr_set[dev & 15] = true;
  }
  p.num_tiles = (int)((p.M + C::NTOK - 1) / C::NTOK);
  if (C::PRO == PRO_SPKATTN) {      // pair tiles: 64 frames x both speakers of one utterance
    p.tiles_per_pair = (p.spk_T + 63) / 64;
    p.num_tiles = (int)(p.M / (2 * p.spk_T)) * p.tiles_per_pair;
  }
  p.b1 = l1.b; p.s1inv = l1.sinv[C::KIND];
  p.b2 = l2 ? l2->b : l1.b; p.s2inv = l2 ? l2->sinv[C::KIND] : l1.sinv[C::KIND];
  const int grid = p.num_tiles < sm_count ? p.num_tiles : sm_count;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, k_tok<C>, l1.map[C::KIND], l2 ? l2->map[C::KIND] : l1.map[C::KIND], p);
  if (e != cudaSuccess) { snprintf(g_tc_err, sizeof(g_tc_err), "k_tok launch: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}
// runtime dispatch on F and operand kind for a config family
#define SEPREF_TOK_DISPATCH(FAMILY, F, KIND, ...)                                                                     \
  ((F) == 128 ? ((KIND) == ::sepref::tc::KIND_F16 ? ::sepref::tc::launch_tok<FAMILY<128, ::sepref::tc::KIND_F16>>(__VA_ARGS__)   \
                                                   : ::sepref::tc::launch_tok<FAMILY<128, ::sepref::tc::KIND_TF32>>(__VA_ARGS__)) \
              : ((KIND) == ::sepref::tc::KIND_F16 ? ::sepref::tc::launch_tok<FAMILY<256, ::sepref::tc::KIND_F16>>(__VA_ARGS__)   \
                                                   : ::sepref::tc::launch_tok<FAMILY<256, ::sepref::tc::KIND_TF32>>(__VA_ARGS__)))

}  // namespace tc
}  // namespace sepref
