// k_cla_front: the first half of CLA (reference network.py:174-180) in ONE kernel, F = 128, FP16 operands:
//
//      d = dwconv65( GLU( W1' . norm(x) + b1' ) )             x [rows, T, F] fp32  ->  d [rows, T, F] FP16
//
// Round 1 ran this as two kernels (cla_a: LayerNorm + GEMM + GLU -> u in HBM; k_dwconv65: u -> d in HBM) and the second
// half (cla_b) read d back: three kernels and two extra [tok, F] fp32 round trips per CLA (profiles/r2_forward_traffic.md:
// the forward moves 1.75 x its algorithmic bytes, mostly through such intermediates).  Here the gated tile never leaves
// the SM: a tile is 128 frames plus the conv's 32-frame halo on each side (192 rows: LayerNorm, GEMM1 and GLU are
// recomputed for the halo, 1.5 x - they are 2F^2 MAC per frame against the stencil's 65 F); GEMM1 runs on tcgen05 with
// the whole W1' (64 KB as FP16) resident in shared memory; the eight epilogue warps turn the accumulators into u (FP16,
// zero outside the utterance = the conv's 'same' padding) in a [192 x 128] shared tile and then run the 65-tap stencil
// over it exactly like k_dwconv65_occ (thread = one channel x 64 frames, taps in registers); d is written as FP16 -
// numerically free, because cla_b rounds d to FP16 as its MMA operand anyway (same round-to-nearest, saturating conversion).
// While the stencil of tile i runs, the MMA of tile i+1 fills the (by then drained) accumulators and the LayerNorm
// producers prepare tile i+2.
#pragma once
#include "kernels_tc.cuh"

namespace sepref {
namespace tc {

struct ClaFrontTraits {
  static constexpr int F = 128, TV = 128, HALO = 32, NT = TV + 2 * HALO;     // 192 rows per tile
  static constexpr int A_BYTES = 128 * 128, ATOM_B = NT * 128;
  static constexpr int B1_BYTES = 2 * ATOM_B;                                // [NT x 128 ch] FP16
  // gated tile, FP16, CHANNEL-major [channel][ULD frames]: a thread (= channel) writes 16 consecutive frames as two 16-byte
  // stores in the GLU pass and reads its 80-frame stencil window as ten 16-byte loads; a row stride of 200 halves
  // (100 words = 4 mod 32) keeps the 16-byte accesses of every quarter-warp on distinct banks
  static constexpr int ULD = 200;
  static constexpr int U_BYTES = F * ULD * 2;
  static constexpr int SMEM_BYTES = 1024 + 4 * A_BYTES + 2 * B1_BYTES + U_BYTES + 512;
  static constexpr int THREADS = 14 * 32;
  static constexpr int KW = 65, OB = 16;                                     // taps, outputs per register block
  static_assert(SMEM_BYTES <= 232448, "shared memory");
};

struct ClaFrontParams {
  const float* x;          // [rows, T, F]
  uint16_t* d;             // [rows, T, F] FP16
  const float *b1, *s1inv; // GEMM1 bias / inverse row scale, packed (value tile, gate tile) order [2F]
  const float *dw, *dwb;   // depthwise taps tap-major [65][F], bias [F]
  int rows, T, tiles_per_row, num_tiles;
};

__global__ void __launch_bounds__(ClaFrontTraits::THREADS, 1)
k_cla_front(const __grid_constant__ CUtensorMap map_w1, const ClaFrontParams p) {
  using TR = ClaFrontTraits;
  constexpr int F = TR::F, TV = TR::TV, HALO = TR::HALO, NT = TR::NT, ATOM_B = TR::ATOM_B, A_BYTES = TR::A_BYTES;
  constexpr int KW = TR::KW, OB = TR::OB;
  constexpr uint32_t IDESC = make_idesc<KIND_F16>(128, NT);

  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sW = sm;                                   // 4 slabs: value tile k0, k1; gate tile k0, k1
  unsigned char* sB1 = sW + 4 * A_BYTES;
  uint16_t* sU = reinterpret_cast<uint16_t*>(sB1 + 2 * TR::B1_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(sU) + TR::U_BYTES);
  uint64_t* w_full = bars;                 // [1]
  uint64_t* b1_full = w_full + 1;          // [2]
  uint64_t* b1_empty = b1_full + 2;        // [2]
  uint64_t* tm_full = b1_empty + 2;        // [1]
  uint64_t* tm_empty = tm_full + 1;        // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tm_empty + 1);

  const int pwarp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = 13 - pwarp;             // 0 weights (once), 1 MMA issue, 2-5 LayerNorm, 6-13 GLU + stencil

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&b1_full[i], 128); mbar_init(&b1_empty[i], 1); }
    mbar_init(tm_full, 1); mbar_init(tm_empty, 256);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&map_w1);
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  if (warp != 0) pdl_wait();
  const int my_iters = ((int)blockIdx.x < p.num_tiles) ? (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, 4 * A_BYTES);
      for (int half = 0; half < 2; ++half)
        for (int ka = 0; ka < 2; ++ka) tma_load_2d(&map_w1, w_full, sW + (half * 2 + ka) * A_BYTES, ka * 64, half * 128);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(w_full, 0, 1000);
      for (int it = 0; it < my_iters; ++it) {
        const uint32_t b = (uint32_t)it & 1, use = (uint32_t)it >> 1;
        mbar_wait_relaxed(&b1_full[b], use & 1, 1001);
        mbar_wait_relaxed(tm_empty, ((uint32_t)it & 1) ^ 1u, 1002);  // the GLU pass of the previous tile has drained TMEM
        tcgen05_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t dcol = tmem_base + half * NT;
#pragma unroll
          for (int ka = 0; ka < 2; ++ka) {
            const uint64_t ad = make_sdesc(smem_u32(sW + (half * 2 + ka) * A_BYTES));
            const uint64_t bd = make_sdesc(smem_u32(sB1 + b * TR::B1_BYTES + ka * ATOM_B));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma<KIND_F16>(dcol, ad + 2 * kk, bd + 2 * kk, IDESC, (ka | kk) != 0);
          }
        }
        umma_commit(tm_full);
        umma_commit(&b1_empty[b]);
      }
    }
  } else if (warp < 6) {
    // ---- LayerNorm of the 192 rows of a tile (frames t0 - 32 .. t0 + 159 of one utterance row; zero rows outside it)
    const int pw = warp - 2;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int n = tile / p.tiles_per_row, t0 = (tile % p.tiles_per_row) * TV;
      const uint32_t b = (uint32_t)it & 1, use = (uint32_t)it >> 1;
      mbar_wait_relaxed(&b1_empty[b], (use & 1) ^ 1u, 1010);
      const float4* x4 = reinterpret_cast<const float4*>(p.x) + (size_t)n * p.T * (F / 4);
      const int T = p.T;
      produce_rows<KIND_F16, F, NT, true>(sB1 + b * TR::B1_BYTES, ATOM_B, pw, lane, 1, [&](int r, int c4, int) {
        const int t = t0 - HALO + r;
        return (t >= 0 && t < T) ? __ldg(x4 + (size_t)t * (F / 4) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      });
      fence_proxy_async();
      mbar_arrive(&b1_full[b]);
    }
  } else {
    // ---- GLU pass (accumulators -> u tile in shared memory), then the 65-tap stencil over the tile
    const int ew = warp - 6;                      // 0..7
    const int eg = ew >> 2, q = pwarp & 3;
    const int ch = q * 32 + lane;                 // GLU pass: channel == TMEM lane
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    const float bv = __ldg(p.b1 + ch), bg = __ldg(p.b1 + 128 + ch), sv = __ldg(p.s1inv + ch), sg = __ldg(p.s1inv + 128 + ch);
    const uint32_t sU32 = smem_u32(sU);           // 32-bit shared addresses: LDS / STS [R + imm] instead of generic 64-bit accesses
    // (A packed variant of the stencil - tap PAIRS against the half2 words of the window, thread = channel x output parity,
    // 528 fma.rn.f32x2 instead of 1040 FFMA per 16 outputs - was measured: 2.58 ms against 2.44 ms.  The FP32 pipe, not the
    // issue slots, bounds this loop; the packed form executes at the scalar rate per lane and adds two-partial-sum overhead.)
    const int etid = ew * 32 + lane;              // stencil pass: thread = channel x half of the tile's frames
    const int cc = etid & 127, part = etid >> 7;
    float wk[KW];
#pragma unroll
    for (int j = 0; j < KW; ++j) wk[j] = __ldg(p.dw + (size_t)j * F + cc);
    const float cbias = __ldg(p.dwb + cc);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int n = tile / p.tiles_per_row, t0 = (tile % p.tiles_per_row) * TV;
      mbar_wait(tm_full, (uint32_t)it & 1, 1020);
      tcgen05_fence_after();
      // GLU pass: group eg owns columns [96 eg, 96 eg + 96); u = (hv/2)(1 + tanh(hg/2)), zero outside the utterance
      const uint32_t tv = tmem_base + tlane + eg * 96, tg = tv + NT;
#pragma unroll 1
      for (int cb = 0; cb < 96; cb += 16) {
        uint32_t rv[16], rg[16];
        tmem_ld16(tv + cb, rv);
        tmem_ld16(tg + cb, rg);
        tmem_wait_ld();
        if (cb + 16 == 96) { tcgen05_fence_before(); mbar_arrive(tm_empty); }
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float u2[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int t = t0 - HALO + eg * 96 + cb + i + e;
            const float hv = 0.5f * fmaf(__uint_as_float(rv[i + e]), sv, bv), hg = fmaf(__uint_as_float(rg[i + e]), sg, bg);
            u2[e] = ((unsigned)t < (unsigned)p.T) ? fmaf(hv, tanh_approx(0.5f * hg), hv) : 0.f;
          }
          pk[i >> 1] = pack_f16x2_sat(u2[0], u2[1]);
        }
        const uint32_t ua = sU32 + (uint32_t)((ch * TR::ULD + eg * 96 + cb) * 2);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ua), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ua + 16u), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");          // the whole u tile is in shared memory
      // stencil pass: outputs [64 part, 64 part + 64) of the tile's 128 frames, 16 at a time
      uint16_t* dst = p.d + ((size_t)n * p.T + t0) * F + cc;
#pragma unroll 1
      for (int o0 = part * 64; o0 < part * 64 + 64; o0 += OB) {
        float acc[OB];
#pragma unroll
        for (int o = 0; o < OB; ++o) acc[o] = cbias;
        const uint32_t win = sU32 + (uint32_t)((cc * TR::ULD + o0) * 2);               // frames t0 + o0 - 32 .. + 47
#pragma unroll
        for (int s8 = 0; s8 < (OB + KW - 1) / 8; ++s8) {
          uint32_t hw[4];
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(hw[0]), "=r"(hw[1]), "=r"(hw[2]), "=r"(hw[3]) : "r"(win + 16u * s8) : "memory");
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 v2 = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int s2 = s8 * 8 + e * 2 + h;
              const float v = h == 0 ? v2.x : v2.y;
#pragma unroll
              for (int o = 0; o < OB; ++o) {
                const int j = s2 - o;
                if (j >= 0 && j < KW) acc[o] = fmaf(wk[j], v, acc[o]);
              }
            }
          }
        }
#pragma unroll
        for (int o = 0; o < OB; ++o)
          if (t0 + o0 + o < p.T) dst[(size_t)(o0 + o) * F] = f16_sat(acc[o]);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");          // the tile may be overwritten by the next GLU pass
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

inline int launch_cla_front(const TcLin& l1, const float* dw, const float* dwb, const float* x, uint16_t* d, int rows, int T,
                            int sm_count, cudaStream_t st) {
  using TR = ClaFrontTraits;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaError_t e;
  static bool attr_set[16] = {};
  if (!attr_set[dev & 15]) {
    e = cudaFuncSetAttribute(k_cla_front, cudaFuncAttributeMaxDynamicSharedMemorySize, TR::SMEM_BYTES);
    if (e != cudaSuccess) { snprintf(g_tc_err, sizeof(g_tc_err), "cudaFuncSetAttribute(k_cla_front): %s", cudaGetErrorString(e)); return -1; }
    attr_set[dev & 15] = true;
  }
  ClaFrontParams p{};
  p.x = x; p.d = d; p.b1 = l1.b; p.s1inv = l1.sinv[KIND_F16]; p.dw = dw; p.dwb = dwb;
  p.rows = rows; p.T = T;
  p.tiles_per_row = (T + TR::TV - 1) / TR::TV;
  p.num_tiles = rows * p.tiles_per_row;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.num_tiles < sm_count ? p.num_tiles : sm_count);
  cfg.blockDim = dim3(TR::THREADS);
  cfg.dynamicSmemBytes = TR::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, k_cla_front, l1.map[KIND_F16], p);
  if (e != cudaSuccess) { snprintf(g_tc_err, sizeof(g_tc_err), "k_cla_front launch: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

}  // namespace tc
}  // namespace sepref
