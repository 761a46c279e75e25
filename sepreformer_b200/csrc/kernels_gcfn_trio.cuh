// k_gcfn_trio: the GCFN block (reference network.py:60-66, F = 128, FP16 operands) with its weights RESIDENT in the shared
// memory of a cluster of THREE CTAs - one 128-channel (value tile, gate tile) chunk of the 3F GLU channels per CTA.
//
// Why three: the streaming kernel k_gcfn is at the per-SM ingest ceiling (37.8 B/clk: 295 KB of weights per 94-frame
// tile), so the only way down is fewer weight bytes per frame.  A CTA PAIR (kernels_gcfn_pair.cuh) must split one chunk
// into a values-in-lanes-0-63 / gates-in-lanes-64-127 tile, and the gate -> value hand-off between warps doubled the
// epilogue's warp time.  3F = 384 = 3 x 128 splits cleanly over three CTAs: each keeps its chunk's W1 rows (2 x 128 x 128
// FP16 = 64 KB) and W2 columns (128 x 128 = 32 KB) for the whole launch, every thread again owns the value AND the gate of
// one channel (the streaming kernel's epilogue, unchanged), and what is left to exchange is the K-split of GEMM2:
//
//   every CTA   LayerNorms the same 96-frame tile (x is read three times: 144 KB per tile instead of 295 KB of weights),
//               runs GEMM1 for its chunk, the gated conv, and the K = 128 partial of GEMM2  ->  partial Y [128 ch x 96 frames];
//   CTA c       finalises frames [32c, 32c + 32) of the tile: it writes the other two thirds of its partial to an
//               L2-resident scratch (distributed shared memory moves ~10 B/clk per direction - a third of a tile period
//               for 32 KB; the L2 path takes the same 32 KB at full store/load bandwidth), signals both peers through
//               their mbarriers (release / acquire at cluster scope) and adds their partials for its own frames.
//
// Zero padding of the time convolution as in k_gcfn_pair: frames outside the utterance are zero operand rows, the folded
// interior formula sees h = b1 there, and the two frames next to an utterance boundary are corrected by w0*b1/2, w2*b1/2.
#pragma once
#include "kernels_gcfn_pair.cuh"

namespace sepref {
namespace tc {

struct TrioTraits {
  static constexpr int F = 128, NTOK = 96, NV = NTOK - 2, THIRD = NTOK / 3;
  static constexpr int A_BYTES = 128 * 128, ATOM_B = NTOK * 128;
  static constexpr int W1_SLABS = 4, W2_SLABS = 2;          // (value, gate) x 2 k slabs; 2 k slabs of this chunk's 128 channels
  static constexpr int B1_BYTES = 2 * ATOM_B, B2_BYTES = 2 * ATOM_B;
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_BYTES = 1024 + (W1_SLABS + W2_SLABS) * A_BYTES + 2 * B1_BYTES + 2 * B2_BYTES + BAR_BYTES;
  static constexpr int THREADS = 16 * 32;
  static constexpr int SCRATCH_FLOATS_PER_CLUSTER = 2 * 3 * 2 * THIRD * F;     // [parity][src][dst slot][frame][channel]
  __host__ __device__ static constexpr int tm_pair(int buf, int half) { return (buf * 2 + half) * NTOK; }
  __host__ __device__ static constexpr int tm_y() { return 4 * NTOK; }
  static_assert(SMEM_BYTES <= 232448, "shared memory");
};

struct GcfnTrioParams {
  const float* x;
  float* y;
  const float *b1, *dw, *cb, *dwf, *b2, *s2inv;      // GcfnPack's per-row constants (packed (value tile, gate tile) order)
  float* scratch;                                    // [clusters][SCRATCH_FLOATS_PER_CLUSTER], L2-resident exchange buffer
  int rows, T, tiles_per_row, num_tiles, iters;
  long long* dbg_clk;
};

__device__ __forceinline__ float ld_cg(const float* p) {       // L2 only: the line was written by another SM
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __cluster_dims__(3, 1, 1) __launch_bounds__(TrioTraits::THREADS, 1)
k_gcfn_trio(const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2, const GcfnTrioParams p) {
  using TR = TrioTraits;
  constexpr int F = TR::F, NTOK = TR::NTOK, NV = TR::NV, THIRD = TR::THIRD, ATOM_B = TR::ATOM_B, A_BYTES = TR::A_BYTES;
  constexpr uint32_t IDESC = make_idesc<KIND_F16>(128, NTOK);

  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sW1 = sm;
  unsigned char* sW2 = sW1 + TR::W1_SLABS * A_BYTES;
  unsigned char* sB1 = sW2 + TR::W2_SLABS * A_BYTES;
  unsigned char* sB2 = sB1 + 2 * TR::B1_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB2 + 2 * TR::B2_BYTES);
  uint64_t* w_full = bars;                 // [1]
  uint64_t* b1_full = w_full + 1;          // [2]
  uint64_t* b1_empty = b1_full + 2;        // [2]
  uint64_t* tm_full = b1_empty + 2;        // [2]
  uint64_t* tm_empty = tm_full + 2;        // [2]
  uint64_t* b2_full = tm_empty + 2;        // [2]
  uint64_t* b2_empty = b2_full + 2;        // [2]
  uint64_t* y_full = b2_empty + 2;         // [1]
  uint64_t* y_empty = y_full + 1;          // [1]
  uint64_t* recv_full = y_empty + 1;       // [2]  one per SOURCE peer (a peer may run a tile ahead of the other), arrived REMOTELY by
                                           //      that peer's drain warps: its partial for my frames is in the scratch
  uint64_t* recv_free = recv_full + 2;     // [2]  per scratch parity, arrived REMOTELY: both peers have read what I wrote there
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(recv_free + 2);

  const int pwarp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // roles: 0 weights (once), 1 GEMM1 issue, 15 GEMM2 issue, 14 idle, 2-5 LayerNorm, 6-13 epilogue / drain
  const int warp = pwarp >= 12 ? (pwarp == 15 ? 0 : pwarp == 14 ? 1 : pwarp == 13 ? 14 : 15) : 13 - pwarp;
  const uint32_t crank = cluster_ctarank();
  const int cid = (int)blockIdx.x / 3, ncl = (int)gridDim.x / 3;
#define TSTAMP3(itv, slot) do { if (p.dbg_clk != nullptr && blockIdx.x == 0 && (itv) < 8) p.dbg_clk[(itv) * 64 + (slot)] = clock64(); } while (0)

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&b1_full[i], 128); mbar_init(&b1_empty[i], 1);
      mbar_init(&tm_full[i], 1); mbar_init(&tm_empty[i], 128);
      mbar_init(&b2_full[i], 128); mbar_init(&b2_empty[i], 1);
      mbar_init(&recv_free[i], 8);
    }
    mbar_init(y_full, 1); mbar_init(y_empty, 128);
    mbar_init(&recv_full[0], 4); mbar_init(&recv_full[1], 4);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_w1); tma_prefetch_desc(&map_w2); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < (2 * TR::B2_BYTES) / 16; i += TR::THREADS) reinterpret_cast<uint4*>(sB2)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  if (warp != 0) pdl_wait();

  auto tile_of = [&](int it) { return cid + it * ncl; };

  // =============================================================================== warp 0: this chunk's weights, once
  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, (TR::W1_SLABS + TR::W2_SLABS) * A_BYTES);
      for (int half = 0; half < 2; ++half)
        for (int ka = 0; ka < 2; ++ka)
          tma_load_2d(&map_w1, w_full, sW1 + (half * 2 + ka) * A_BYTES, ka * 64, (2 * (int)crank + half) * 128);
      for (int ka = 0; ka < 2; ++ka) tma_load_2d(&map_w2, w_full, sW2 + ka * A_BYTES, (int)crank * 128 + ka * 64, 0);
    }
  }
  // =============================================================================== warp 1: GEMM1 issue
  else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(w_full, 0, 900);
      for (int it = 0; it < p.iters; ++it) {
        const uint32_t b = (uint32_t)it & 1, use = (uint32_t)it >> 1;
        mbar_wait(&b1_full[b], use & 1, 901);
        mbar_wait(&tm_empty[b], (use & 1) ^ 1u, 902);
        tcgen05_fence_after();
        TSTAMP3(it, 0);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t d = tmem_base + TR::tm_pair((int)b, half);
#pragma unroll
          for (int ka = 0; ka < 2; ++ka) {
            const uint64_t ad = make_sdesc(smem_u32(sW1 + (half * 2 + ka) * A_BYTES));
            const uint64_t bd = make_sdesc(smem_u32(sB1 + b * TR::B1_BYTES + ka * ATOM_B));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma<KIND_F16>(d, ad + 2 * kk, bd + 2 * kk, IDESC, (ka | kk) != 0);
          }
        }
        umma_commit(&tm_full[b]);
        umma_commit(&b1_empty[b]);
        TSTAMP3(it, 1);
      }
    }
  }
  // =============================================================================== warp 15: GEMM2 issue (K = this chunk)
  else if (warp == 15) {
    if (lane == 0) {
      mbar_wait(w_full, 0, 905);
      for (int it = 0; it < p.iters; ++it) {
        const uint32_t b = (uint32_t)it & 1, use = (uint32_t)it >> 1;
        mbar_wait(&b2_full[b], use & 1, 906);
        mbar_wait(y_empty, ((uint32_t)it & 1) ^ 1u, 907);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + TR::tm_y();
#pragma unroll
        for (int ka = 0; ka < 2; ++ka) {
          const uint64_t ad = make_sdesc(smem_u32(sW2 + ka * A_BYTES));
          const uint64_t bd = make_sdesc(smem_u32(sB2 + b * TR::B2_BYTES + ka * ATOM_B));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma<KIND_F16>(d, ad + 2 * kk, bd + 2 * kk, IDESC, (ka | kk) != 0);
        }
        umma_commit(&b2_empty[b]);
        umma_commit(y_full);
        TSTAMP3(it, 8);
      }
    }
  }
  // =============================================================================== warps 2-5: LayerNorm -> stage-1 operand
  else if (warp >= 2 && warp < 6) {
    const int pw = warp - 2;
    const int sub = lane >> 3, j = lane & 7;
    for (int it = 0; it < p.iters; ++it) {
      const int tile = tile_of(it);
      const bool live = tile < p.num_tiles;
      const int n = live ? tile / p.tiles_per_row : 0, t0 = live ? (tile % p.tiles_per_row) * NV : 0;
      const float4* x4 = reinterpret_cast<const float4*>(p.x) + (size_t)n * p.T * (F / 4);
      const uint32_t b = (uint32_t)it & 1, use = (uint32_t)it >> 1;
      unsigned char* b1buf = sB1 + b * TR::B1_BYTES;
      float4 v[6][4];                        // all 24 rows of this warp are requested before the buffer wait
#pragma unroll
      for (int g = 0; g < 6; ++g) {
        const int r = 4 * (pw + 4 * g) + sub, t = t0 - 1 + r;
        const bool ok = live && t >= 0 && t < p.T;
#pragma unroll
        for (int k = 0; k < 4; ++k) v[g][k] = ok ? __ldg(x4 + (size_t)t * (F / 4) + j + 8 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      mbar_wait(&b1_empty[b], (use & 1) ^ 1u, 910);
      if (warp == 2 && lane == 0) TSTAMP3(it, 16);
#pragma unroll
      for (int g = 0; g < 6; ++g) {
        const int r = 4 * (pw + 4 * g) + sub;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) s += v[g][k].x + v[g][k].y + v[g][k].z + v[g][k].w;
#pragma unroll
        for (int o = 1; o <= 4; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / F);
        float qq = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v[g][k].x -= mean; v[g][k].y -= mean; v[g][k].z -= mean; v[g][k].w -= mean;
          qq += v[g][k].x * v[g][k].x + v[g][k].y * v[g][k].y + v[g][k].z * v[g][k].z + v[g][k].w * v[g][k].w;
        }
#pragma unroll
        for (int o = 1; o <= 4; o <<= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
        const float rstd = rsqrtf(qq * (1.0f / F) + kLnEps);
#pragma unroll
        for (int k = 0; k < 4; ++k) store_c4<KIND_F16>(b1buf, ATOM_B, r, j + 8 * k, v[g][k], rstd);
      }
      fence_proxy_async();
      mbar_arrive(&b1_full[b]);
      if (warp == 2 && lane == 0) TSTAMP3(it, 17);
    }
  }
  // =============================================================================== warps 6-13: gated conv / drain + exchange
  else if (warp >= 6 && warp < 14) {
    const int eg = (warp - 6) >> 2;               // group eg runs the chunk of tiles with (it & 1) == eg and drains them next iteration
    const int q = pwarp & 3;
    const int ch = q * 32 + lane;                 // channel of the chunk (epilogue) / output channel (drain) == TMEM lane
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    unsigned char* sbase[8];
    make_sbase<KIND_F16>(sbase, sB2 + eg * TR::B2_BYTES, ATOM_B, q, lane);
    uint32_t sb0[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) sb0[m] = smem_u32(sbase[m]);
    // per-channel constants of this CTA's chunk (packed rows: value tile 2c, gate tile 2c + 1)
    const int rv = (2 * (int)crank) * 128 + ch, rg = rv + 128;
    const float cv = __ldg(p.cb + rv), cg = __ldg(p.cb + rg);
    const float wv0 = __ldg(p.dwf + rv), wv1 = __ldg(p.dwf + 6 * F + rv), wv2 = __ldg(p.dwf + 12 * F + rv);
    const float wg0 = __ldg(p.dwf + rg), wg1 = __ldg(p.dwf + 6 * F + rg), wg2 = __ldg(p.dwf + 12 * F + rg);
    const float b1v = __ldg(p.b1 + rv), b1g = __ldg(p.b1 + rg);
    const float flv = __ldg(p.dw + rv) * b1v, frv = __ldg(p.dw + 12 * F + rv) * b1v;      // w0*b1/2, w2*b1/2 (taps pre-scaled by 1/2)
    const float flg = __ldg(p.dw + rg) * b1g, frg = __ldg(p.dw + 12 * F + rg) * b1g;
    const float bias2 = __ldg(p.b2 + ch), s2i = __ldg(p.s2inv + ch);
    const uint32_t peerA = (crank + 1) % 3, peerB = (crank + 2) % 3;
    // in a destination d, the barrier / scratch slot of source s has index (s - d - 1) mod 3; I am source for peerA and peerB
    const uint32_t rA_full = mapa_u32(smem_u32(&recv_full[(crank + 2 - peerA) % 3]), peerA);
    const uint32_t rB_full = mapa_u32(smem_u32(&recv_full[(crank + 2 - peerB) % 3]), peerB);
    float* const scr = p.scratch + (size_t)cid * TR::SCRATCH_FLOATS_PER_CLUSTER;
    // scratch slot of (parity, source CTA s, destination d != s): dst slot k = (d - s - 1) mod 3 in {0, 1}
    auto slot = [&](uint32_t par, uint32_t s, uint32_t d) { return scr + (((size_t)par * 3 + s) * 2 + ((d + 2 - s) % 3)) * (THIRD * F); };

    auto drain = [&](int it) {
      const int tile = tile_of(it);
      const bool live = tile < p.num_tiles;
      const int n = live ? tile / p.tiles_per_row : 0, t0 = live ? (tile % p.tiles_per_row) * NV : 0;
      const int cmax = live ? min(NV, p.T - t0) : 0;
      const uint32_t par = (uint32_t)it & 1;
      const int own0 = (int)crank * THIRD;
      const float* xcol = p.x + (((long long)n * p.T + t0 - 1) * F + ch);
      float* ycol = p.y + (((long long)n * p.T + t0 - 1) * F + ch);
      float xin[THIRD];
#pragma unroll
      for (int i = 0; i < THIRD; ++i) xin[i] = (own0 + i >= 1 && own0 + i <= cmax) ? ldg_now(xcol + (own0 + i) * F) : 0.f;
      if (q == 0 && lane == 0) TSTAMP3(it, 48);
      // the scratch slots of this parity were last written for tile it-2: both peers must have read them
      mbar_wait_cl(&recv_free[par], (((uint32_t)it >> 1) & 1) ^ 1u, 920);
      mbar_wait(y_full, (uint32_t)it & 1, 921);
      tcgen05_fence_after();
      if (q == 0 && lane == 0) TSTAMP3(it, 49);
      const uint32_t ty = tmem_base + tlane + TR::tm_y();
      // ---- the peers' frames: my partial -> L2 scratch (a warp stores 128 contiguous bytes per frame)
#pragma unroll
      for (int pe = 0; pe < 2; ++pe) {
        const uint32_t d = pe == 0 ? peerA : peerB;
        float* dst = slot(par, crank, d) + ch;
#pragma unroll
        for (int cb = 0; cb < THIRD; cb += 16) {
          uint32_t r[16];
          tmem_ld16(ty + d * THIRD + cb, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) dst[(cb + i) * F] = __uint_as_float(r[i]);
        }
      }
      __syncwarp();
      if (lane == 0) { mbar_arrive_remote(rA_full); mbar_arrive_remote(rB_full); }
      if (q == 0 && lane == 0) TSTAMP3(it, 50);
      // ---- my frames: own partial from TMEM ...
      uint32_t own[THIRD];
      {
        uint32_t r[16];
        tmem_ld16(ty + own0, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) own[i] = r[i];
        tmem_ld16(ty + own0 + 16, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) own[16 + i] = r[i];
      }
      tcgen05_fence_before();
      mbar_arrive(y_empty);
      // ---- ... plus the two partials the peers wrote for me
      mbar_wait_cl(&recv_full[0], (uint32_t)it & 1, 922);       // source peerA = crank + 1: index (crank + 1 - crank - 1) = 0
      mbar_wait_cl(&recv_full[1], (uint32_t)it & 1, 923);       // source peerB = crank + 2: index 1
      if (q == 0 && lane == 0) TSTAMP3(it, 51);
      const float* sa = slot(par, peerA, crank) + ch;
      const float* sbp = slot(par, peerB, crank) + ch;
#pragma unroll
      for (int cb = 0; cb < THIRD; cb += 16) {
        float pa[16], pb[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { pa[i] = ld_cg(sa + (cb + i) * F); pb[i] = ld_cg(sbp + (cb + i) * F); }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = own0 + cb + i;
          if (c >= 1 && c <= cmax) ycol[c * F] = fmaf(__uint_as_float(own[cb + i]) + pa[i] + pb[i], s2i, xin[cb + i] + bias2);
        }
      }
      __syncwarp();
      if (lane == 0) {                              // both sources may overwrite this parity's slots (two tiles from now)
        mbar_arrive_remote(mapa_u32(smem_u32(&recv_free[par]), peerA));
        mbar_arrive_remote(mapa_u32(smem_u32(&recv_free[par]), peerB));
      }
      if (q == 0 && lane == 0) TSTAMP3(it, 52);
    };

    for (int it = 0; it <= p.iters; ++it) {
      if (it > 0 && eg == ((it - 1) & 1)) drain(it - 1);
      if (it == p.iters) break;
      if ((it & 1) != eg) continue;
      const int tile = tile_of(it);
      const bool live = tile < p.num_tiles;
      const int t0 = live ? (tile % p.tiles_per_row) * NV : 0;
      const int cL = (live && t0 == 0) ? 1 : -64;
      const int cR = (live && p.T - t0 >= 1 && p.T - t0 <= NV) ? p.T - t0 : -64;
      const bool edge = cL > 0 || cR > 0;
      const uint32_t use = (uint32_t)it >> 1;
      mbar_wait(&tm_full[eg], use & 1, 930);
      if (q == 0 && lane == 0) TSTAMP3(it, 24);
      mbar_wait(&b2_empty[eg], (use & 1) ^ 1u, 931);
      tcgen05_fence_after();
      const uint32_t tv = tmem_base + tlane + TR::tm_pair(eg, 0), tg = tmem_base + tlane + TR::tm_pair(eg, 1);
      const float2 wv0p = make_float2(wv0, wv0), wv2p = make_float2(wv2, wv2), wg0p = make_float2(wg0, wg0), wg2p = make_float2(wg2, wg2);
      const float2 cvp = make_float2(cv, cv), cgp = make_float2(cg, cg);
      float pv0 = 0.f, pv1 = 0.f, pg0 = 0.f, pg1 = 0.f;
      uint32_t sb[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) sb[k] = sb0[k];
      int cb0 = 0;                                  // first column of the batch being computed
      auto batch = [&](const uint32_t (&ev)[16], const uint32_t (&eg_)[16]) {
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const float2 av = m == 0 ? make_float2(pv0, pv1) : make_float2(__uint_as_float(ev[2 * m - 2]), __uint_as_float(ev[2 * m - 1]));
          const float2 ag = m == 0 ? make_float2(pg0, pg1) : make_float2(__uint_as_float(eg_[2 * m - 2]), __uint_as_float(eg_[2 * m - 1]));
          const float2 cvv = make_float2(__uint_as_float(ev[2 * m]), __uint_as_float(ev[2 * m + 1]));
          const float2 cgg = make_float2(__uint_as_float(eg_[2 * m]), __uint_as_float(eg_[2 * m + 1]));
          float2 dv = __ffma2_rn(wv2p, cvv, __ffma2_rn(wv0p, av, cvp));
          float2 dg = __ffma2_rn(wg2p, cgg, __ffma2_rn(wg0p, ag, cgp));
          dv.x = fmaf(wv1, av.y, dv.x); dv.y = fmaf(wv1, cvv.x, dv.y);
          dg.x = fmaf(wg1, ag.y, dg.x); dg.y = fmaf(wg1, cgg.x, dg.y);
          if (edge) {                                // columns cb0-1+2m (x) and cb0+2m (y) next to an utterance boundary
            const int c0 = cb0 - 1 + 2 * m;
            dv.x -= (c0 == cL ? flv : 0.f) + (c0 == cR ? frv : 0.f); dv.y -= (c0 + 1 == cL ? flv : 0.f) + (c0 + 1 == cR ? frv : 0.f);
            dg.x -= (c0 == cL ? flg : 0.f) + (c0 == cR ? frg : 0.f); dg.y -= (c0 + 1 == cL ? flg : 0.f) + (c0 + 1 == cR ? frg : 0.f);
          }
          const float2 th = make_float2(tanh_approx(dg.x), tanh_approx(dg.y));
          const float2 u = __ffma2_rn(dv, th, dv);
          if (m > 0 || cb0 > 0) sts_elem<KIND_F16>(sb[(2 * m + 7) & 7] + (uint32_t)(((2 * m - 1) >> 3) * 1024), u.x);
          sts_elem<KIND_F16>(sb[(2 * m) & 7] + (uint32_t)(((2 * m) >> 3) * 1024), u.y);
        }
        pv0 = __uint_as_float(ev[14]); pv1 = __uint_as_float(ev[15]); pg0 = __uint_as_float(eg_[14]); pg1 = __uint_as_float(eg_[15]);
#pragma unroll
        for (int k = 0; k < 8; ++k) sb[k] += 2048u;
        cb0 += 16;
      };
      uint32_t av[16], ag[16], bv[16], bg[16];
      tmem_ld16(tv, av); tmem_ld16(tg, ag);
      tmem_wait_ld();
#pragma unroll 1
      for (int cb = 0; cb < NTOK; cb += 32) {
        tmem_ld16(tv + cb + 16, bv); tmem_ld16(tg + cb + 16, bg);
        batch(av, ag);
        tmem_wait_ld();
        if (cb + 32 < NTOK) { tmem_ld16(tv + cb + 32, av); tmem_ld16(tg + cb + 32, ag); }
        else { tcgen05_fence_before(); mbar_arrive(&tm_empty[eg]); if (q == 0 && lane == 0) TSTAMP3(it, 26); }
        batch(bv, bg);
        if (cb + 32 < NTOK) tmem_wait_ld();
      }
      fence_proxy_async();
      mbar_arrive(&b2_full[eg]);
      if (q == 0 && lane == 0) TSTAMP3(it, 27);
    }
  }

  // ---- teardown
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
#undef TSTAMP3
}

// ------------------------------------------------------------------------------------------------ host side
struct TrioState {          // per device: the exchange scratch and the number of co-resident clusters
  float* scratch = nullptr;
  int clusters = 0;
};

// once per handle (sepref_finalize): opt-in shared memory, how many clusters of three fit, and the exchange scratch
inline int prepare_gcfn_trio(TrioState& ts, int sm_count) {
  using TR = TrioTraits;
  if (ts.clusters > 0) return 0;
  cudaError_t e = cudaFuncSetAttribute(k_gcfn_trio, cudaFuncAttributeMaxDynamicSharedMemorySize, TR::SMEM_BYTES);
  if (e != cudaSuccess) { snprintf(g_tc_err, sizeof(g_tc_err), "cudaFuncSetAttribute(k_gcfn_trio): %s", cudaGetErrorString(e)); return -1; }
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(TR::THREADS);
  cfg.dynamicSmemBytes = TR::SMEM_BYTES;
  cfg.gridDim = dim3((sm_count / 3) * 3);
  int n = 0;
  e = cudaOccupancyMaxActiveClusters(&n, k_gcfn_trio, &cfg);
  if (e != cudaSuccess || n <= 0) { cudaGetLastError(); n = sm_count / 3 - 4; }
  e = cudaMalloc(&ts.scratch, (size_t)n * TR::SCRATCH_FLOATS_PER_CLUSTER * sizeof(float));
  if (e != cudaSuccess) { snprintf(g_tc_err, sizeof(g_tc_err), "trio scratch: %s", cudaGetErrorString(e)); return -1; }
  ts.clusters = n;
  return 0;
}

inline int launch_gcfn_trio(TrioState& ts, const GcfnPack& g, const float* x, float* y, int rows, int T, int sm_count, cudaStream_t st,
                            long long* dbg_clk = nullptr) {
  using TR = TrioTraits;
  (void)sm_count;
  if (ts.clusters <= 0) { snprintf(g_tc_err, sizeof(g_tc_err), "k_gcfn_trio not prepared"); return -1; }
  cudaError_t e;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(TR::THREADS);
  cfg.dynamicSmemBytes = TR::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GcfnTrioParams p{};
  p.x = x; p.y = y; p.b1 = g.b1; p.dw = g.dw; p.cb = g.cb; p.dwf = g.dwf[KIND_F16]; p.b2 = g.b2; p.s2inv = g.s2inv[KIND_F16];
  p.scratch = ts.scratch; p.rows = rows; p.T = T; p.dbg_clk = dbg_clk;
  p.tiles_per_row = (T + TR::NV - 1) / TR::NV;
  p.num_tiles = rows * p.tiles_per_row;
  int clusters = ts.clusters < p.num_tiles ? ts.clusters : p.num_tiles;
  p.iters = (p.num_tiles + clusters - 1) / clusters;
  cfg.gridDim = dim3(3 * clusters);
  e = cudaLaunchKernelEx(&cfg, k_gcfn_trio, g.map_w1[KIND_F16][0], g.map_w2[KIND_F16][0], p);
  if (e != cudaSuccess) { snprintf(g_tc_err, sizeof(g_tc_err), "k_gcfn_trio launch: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

}  // namespace tc
}  // namespace sepref
