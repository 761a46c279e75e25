// Microbenchmarks behind two design questions of the GCFN / token kernels (results: profiles/r1_pipe_rates.md):
//   1. FP32 FMA issue rate per SM for the instruction forms the epilogues use: three-register FFMA (all operands
//      distinct registers), FFMA with a shared (reused) multiplicand, and the packed fma.rn.f32x2.
//   2. tcgen05.mma pacing (clk per instruction as seen by the issuing thread) for M = 128, K = 16 (fp16), both operands
//      in shared memory, as a function of N - the per-instruction re-read of the 4 KB A slice is what should make
//      small-N instructions slow.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench/pipe_rates tools/microbench/pipe_rates.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../sepreformer_b200/csrc/kernels_tc.cuh"
using namespace sepref::tc;

// ------------------------------------------------------------------------------------------------ 1. FMA forms
// MODE 0: acc[i] = fma(a[i], b[i], acc[i])   16 independent chains, three distinct registers per instruction
// MODE 1: acc[i] = fma(a,    b[i], acc[i])   one multiplicand shared by all 16
// MODE 2: acc2[i] = ffma2(a2[i], b2[i], acc2[i])   8 packed chains (16 FMAs per 8 instructions)
template <int MODE>
__global__ void __launch_bounds__(128) k_fma(float* out, int iters, float seed) {
  float acc[16], a[16], b[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { acc[i] = seed + i; a[i] = 1.0f + 1e-7f * (threadIdx.x + i); b[i] = 1e-9f * (i + 1); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 8; ++rep) {
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(a[i], acc[i], b[i]);
      } else if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(a[0], acc[i], b[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const float2 r = __ffma2_rn(make_float2(a[i], a[i + 1]), make_float2(acc[i], acc[i + 1]), make_float2(b[i], b[i + 1]));
          acc[i] = r.x; acc[i + 1] = r.y;
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 12345.678f) out[0] = s;
}

template <int MODE>
static void run_fma(const char* name, int sms, float* d_out) {
  const int iters = 4096, blocks = sms * 4;               // 4 blocks x 4 warps per SM = 4 warps per sub-partition
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_fma<MODE><<<blocks, 128>>>(d_out, 16, 1.0f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_fma<MODE><<<blocks, 128>>>(d_out, iters, 1.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double fmas = (double)blocks * 128 * iters * 8 * 16;
  const double clk = ms * 1e-3 * khz * 1e3;
  printf("| %s | %.1f | %.2f |\n", name, fmas / clk / sms, ms);
}

// ------------------------------------------------------------------------------------------------ 2. UMMA pacing
template <int N>
__global__ void __launch_bounds__(128, 1) k_umma(long long* out_clk, int reps) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = sm;                         // 2 slabs [128 rows x 128 B]
  unsigned char* sB = sm + 2 * 16384;             // 2 atoms [N rows x 128 B]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 2 * N * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < (2 * 16384 + 2 * N * 128) / 16; i += 128) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *slot;
  constexpr uint32_t IDESC = make_idesc<KIND_F16>(128, N);
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      // one "step" as in the kernels: 2 slabs x 4 MMAs into one accumulator, then a commit
      for (int ka = 0; ka < 2; ++ka) {
        const uint64_t ad = make_sdesc(smem_u32(sA + ka * 16384));
        const uint64_t bd = make_sdesc(smem_u32(sB + ka * N * 128));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma<KIND_F16>(tmem + (r & 1) * N, ad + 2 * k, bd + 2 * k, IDESC, (ka | k) != 0);
      }
    }
    const long long t1 = clock64();              // all issued
    umma_commit(bar);
    mbar_wait(bar, 0, 900);
    const long long t2 = clock64();              // all executed
    out_clk[blockIdx.x * 2] = t1 - t0;
    out_clk[blockIdx.x * 2 + 1] = t2 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int N>
static void run_umma(long long* d_clk, int sms) {
  const int reps = 64;
  const size_t smem = 1024 + 2 * 16384 + 2 * N * 128 + 64;
  cudaFuncSetAttribute(k_umma<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int pass = 0; pass < 2; ++pass) k_umma<N><<<sms, 128, smem>>>(d_clk, reps);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("| %d | error: %s |\n", N, cudaGetErrorString(e)); return; }
  std::vector<long long> h(2 * sms);
  cudaMemcpy(h.data(), d_clk, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost);
  double issue = 0, exec = 0;
  for (int i = 0; i < sms; ++i) { issue += h[2 * i]; exec += h[2 * i + 1]; }
  const double n = (double)sms * reps * 8;
  printf("| %d | %.1f | %.1f | %d |\n", N, issue / n, exec / n, N / 2);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* d_out; long long* d_clk;
  cudaMalloc(&d_out, 64); cudaMalloc(&d_clk, sizeof(long long) * 2 * 1024);
  printf("## FP32 FMA forms (16 accumulators per thread, 4 warps per SM sub-partition), %d SMs\n\n", sms);
  printf("| form | FMA / clk / SM | ms |\n|---|---:|---:|\n");
  run_fma<0>("three distinct registers: `fma(a[i], acc[i], b[i])`", sms, d_out);
  run_fma<1>("shared multiplicand: `fma(a, acc[i], b[i])`", sms, d_out);
  run_fma<2>("packed `fma.rn.f32x2`", sms, d_out);
  printf("\n## tcgen05.mma kind::f16, M = 128, K = 16, A and B in shared memory (SWIZZLE_128B), one CTA per SM\n\n");
  printf("| N | clk per MMA, issue loop | clk per MMA, until commit | floor N/2 |\n|---:|---:|---:|---:|\n");
  run_umma<64>(d_clk, sms);
  run_umma<96>(d_clk, sms);
  run_umma<128>(d_clk, sms);
  run_umma<160>(d_clk, sms);
  run_umma<192>(d_clk, sms);
  run_umma<256>(d_clk, sms);
  return 0;
}
