// k_gcfn_pair: the GCFN block (reference network.py:60-66) with its weights RESIDENT in the shared memory of a CTA PAIR.
//
// Why: round 1's k_gcfn streams the whole weight set (295 KB as FP16) from L2 for every 94-frame tile, because one CTA
// cannot hold it (227 KB) and TMEM caps a tile at 96 columns.  The in-kernel timeline showed the tensor pipe waiting on
// slab arrival for most of a 14.5 k clk tile (72 MMAs at ~200 clk each against a ~50 clk floor; the eight-slab ring
// sustains ~20 B/clk/SM of weight ingest).  Here the hidden dimension of the block - the 3F GLU channels - is split
// across the two CTAs of a cluster (tensor parallelism inside a CTA pair):
//
//   CTA c keeps   W1 rows of its 192 GLU channels (values and gates: 384 rows x 128, 96 KB)  and
//                 W2 columns of the same channels  (128 x 192, 48 KB)                         for the whole launch,
//   both CTAs     LayerNorm the same 96-frame tile (x is read twice; 1/6 of the bytes the weights were),
//   each CTA      runs GEMM1 for its channels, the gated conv, and a K = 192 partial of GEMM2,
//   the pair      adds the two partial Y tiles: each CTA finalises half of the frames and receives the other CTA's
//                 partial for those frames through distributed shared memory (24 KB per tile and direction).
//
// 192 channels are one and a half 128-row MMA tiles, and a thread can only read its own TMEM lane, so the value and the
// gate of a channel must share a lane index.  Every MMA tile ("chunk") is therefore MIXED: rows 0-63 are the values of
// 64 channels, rows 64-127 their gates.  In the epilogue the two warps that own the gate lanes compute
// t = tanh(conv3(g)/2) and hand it to the two warps that own the value lanes through an 8 KB shared scratch (eight
// columns at a time, double-buffered, mbarrier handshake); the value warps compute u = (conv3(v)/2)(1 + t), round to
// FP16 and write the stage-2 operand.  All tcgen05 instructions are cta_group::1; the pair cooperates only through the
// DSMEM exchange of Y partials.
//
// Zero padding of the time convolution without a second code path: frames outside the utterance are zero rows of the
// stage-1 operand, so their accumulator is exactly 0 and the folded interior formula yields h = b1 there instead of 0.
// Only the two frames next to an utterance boundary see such a neighbour; their conv result is corrected by the
// per-channel constants kl = w0*b1/2 (frame 0) and kr = w2*b1/2 (frame T-1).
#pragma once
#include "kernels_tc.cuh"

namespace sepref {
namespace tc {

struct PairTraits {      // F = 128, FP16 operands
  static constexpr int F = 128, NTOK = 96, NV = NTOK - 2, HALF = NTOK / 2;
  static constexpr int A_BYTES = 128 * 128;                 // one weight slab [128 rows x 128 B]
  static constexpr int ATOM_B = NTOK * 128;                 // one operand slab [NTOK rows x 128 B]
  static constexpr int W1_SLABS = 6, W2_SLABS = 3;          // per CTA: 3 chunks x 2 k slabs; 3 k slabs
  static constexpr int B1_BYTES = 2 * ATOM_B;               // stage-1 operand: NTOK x 128 channels
  static constexpr int B2_BYTES = ATOM_B;                   // stage-2 operand slab: NTOK x 64 channels (one chunk)
  static constexpr int RECV_BYTES = HALF * F * 4;           // the peer's partial Y for my half of the frames, fp32
  static constexpr int XG_BYTES = 4 * 2 * 1024;             // gate -> value scratch: 4 warp pairs x 2 halves x [4][32] float2
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_BYTES = 1024 + (W1_SLABS + W2_SLABS) * A_BYTES + B1_BYTES + 2 * B2_BYTES + RECV_BYTES + XG_BYTES + BAR_BYTES;
  static constexpr int THREADS = 14 * 32;
  static constexpr int ROWS = 6 * F;                        // packed GEMM1 rows (both CTAs)
  __host__ __device__ static constexpr int tm_c(int k) { return k * NTOK; }
  __host__ __device__ static constexpr int tm_y(int b) { return 3 * NTOK + b * NTOK; }
  static_assert(SMEM_BYTES <= 232448, "shared memory");
  static_assert(5 * NTOK <= 512, "TMEM columns");
};

struct GcfnPairPack {
  // GEMM1 rows in pair order: row (c*384 + k*128 + l) is, for l < 64, the VALUE row of GLU channel c*192 + k*64 + l and,
  // for l >= 64, the GATE row of channel c*192 + k*64 + (l - 64); FP16, per-row power-of-two scaled
  const void* w1 = nullptr;
  const float* cb = nullptr;      // [768] interior conv constant (dwb + b1 * (w0+w1+w2)) / 2, pair order
  const float* dwf = nullptr;     // [3][768] taps * s1inv / 2, pair order
  const float* kl = nullptr;      // [768] w0 * b1 / 2   (correction at the first frame of an utterance)
  const float* kr = nullptr;      // [768] w2 * b1 / 2   (correction at the last frame)
  alignas(64) CUtensorMap map_w1;
  bool ready = false;
};

struct GcfnPairParams {
  const float* x;
  float* y;
  const float *cb, *dwf, *kl, *kr, *b2, *s2inv;
  int rows, T, tiles_per_row, num_tiles, iters;
  long long* dbg_clk;   // optional [8][64] clock64 stamps of block 0's first 8 tiles (tools/gcfn_timeline.py)
};

// ---- cluster / DSMEM helpers -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t raddr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}
// arrive on a barrier in the peer's shared memory; release at cluster scope orders this thread's earlier DSMEM stores
// (and, through a preceding __syncwarp, its warp's) before the arrival
__device__ __forceinline__ void mbar_arrive_remote(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cl(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __noinline__ void mbar_wait_cl_slow(uint64_t* bar, uint32_t parity, int tag) {
  const long long t0 = clock64();
  for (;;) {
#pragma unroll 1
    for (int n = 0; n < 64; ++n)
      if (mbar_try_wait_cl(bar, parity)) return;
    if (clock64() - t0 > 4000000000LL) {
      printf("sepref: cluster mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, (int)blockIdx.x, (int)threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity, int tag) {
  if (mbar_try_wait_cl(bar, parity)) return;
  mbar_wait_cl_slow(bar, parity, tag);
}
__device__ __forceinline__ void sts_f32x2(uint32_t addr, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PairTraits::THREADS, 1)
k_gcfn_pair(const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2, const GcfnPairParams p) {
  using TR = PairTraits;
  constexpr int F = TR::F, NTOK = TR::NTOK, NV = TR::NV, HALF = TR::HALF, ATOM_B = TR::ATOM_B, A_BYTES = TR::A_BYTES;
  constexpr uint32_t IDESC = make_idesc<KIND_F16>(128, NTOK);

  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sW1 = sm;
  unsigned char* sW2 = sW1 + TR::W1_SLABS * A_BYTES;
  unsigned char* sB1 = sW2 + TR::W2_SLABS * A_BYTES;
  unsigned char* sB2 = sB1 + TR::B1_BYTES;
  unsigned char* sRecv = sB2 + 2 * TR::B2_BYTES;
  unsigned char* sXg = sRecv + TR::RECV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sXg + TR::XG_BYTES);
  uint64_t* w_full = bars;                 // [1]  resident weights have landed
  uint64_t* b1_full = w_full + 1;          // [1]
  uint64_t* b1_empty = b1_full + 1;        // [1]
  uint64_t* c_full = b1_empty + 1;         // [3]  chunk accumulators
  uint64_t* c_empty = c_full + 3;          // [3]
  uint64_t* b2_full = c_empty + 3;         // [2]  stage-2 operand slabs (one per epilogue group)
  uint64_t* b2_empty = b2_full + 2;        // [2]
  uint64_t* y_full = b2_empty + 2;         // [2]
  uint64_t* y_empty = y_full + 2;          // [2]
  uint64_t* recv_full = y_empty + 2;       // [1]  arrived REMOTELY: the peer's partial for my frames is in sRecv
  uint64_t* recv_free = recv_full + 1;     // [1]  arrived REMOTELY: the peer has consumed what I sent it
  uint64_t* xg_full = recv_free + 1;       // [8]  gate -> value scratch, [warp pair][half]
  uint64_t* xg_empty = xg_full + 8;        // [8]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xg_empty + 8);

  const int pwarp = threadIdx.x >> 5, lane = threadIdx.x & 31;     // physical warp: fixes the TMEM lane quarter
  const int warp = 13 - pwarp;                                     // role index (critical single-lane roles get top warp ids)
  const uint32_t crank = cluster_ctarank(), peer = crank ^ 1u;
  const int cid = (int)blockIdx.x >> 1, ncl = (int)gridDim.x >> 1;
#define PSTAMP(itv, slot) do { if (p.dbg_clk != nullptr && blockIdx.x == 0 && (itv) < 8) p.dbg_clk[(itv) * 64 + (slot)] = clock64(); } while (0)

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    mbar_init(b1_full, 128); mbar_init(b1_empty, 1);
    for (int i = 0; i < 3; ++i) { mbar_init(&c_full[i], 1); mbar_init(&c_empty[i], 128); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&b2_full[i], 64); mbar_init(&b2_empty[i], 1);
      mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], 128);
    }
    mbar_init(recv_full, 4); mbar_init(recv_free, 4);
    for (int i = 0; i < 8; ++i) { mbar_init(&xg_full[i], 32); mbar_init(&xg_empty[i], 32); }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_w1); tma_prefetch_desc(&map_w2); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // stage-2 operand slabs: halo rows are written with whatever the halo columns produce; start them finite
  for (int i = threadIdx.x; i < (2 * TR::B2_BYTES) / 16; i += TR::THREADS) reinterpret_cast<uint4*>(sB2)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                        // the peer's barriers exist before anyone signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  if (warp != 0) pdl_wait();                 // warp 0 only fetches weights, which no kernel writes

  auto tile_of = [&](int it) { return cid + it * ncl; };

  // =============================================================================== warp 0: resident weights, once
  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, (TR::W1_SLABS + TR::W2_SLABS) * A_BYTES);
      for (int k = 0; k < 3; ++k)
        for (int ka = 0; ka < 2; ++ka)
          tma_load_2d(&map_w1, w_full, sW1 + (k * 2 + ka) * A_BYTES, ka * 64, (int)crank * 384 + k * 128);
      for (int k = 0; k < 3; ++k) tma_load_2d(&map_w2, w_full, sW2 + k * A_BYTES, (int)crank * 192 + k * 64, 0);
    }
  }
  // =============================================================================== warp 1: MMA issue
  else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(w_full, 0, 900);
      auto g1 = [&](int it, int k) {           // C_k = W1[chunk k] . norm(x)^T
        mbar_wait(&c_empty[k], (uint32_t)(it & 1) ^ 1u, 901);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + TR::tm_c(k);
#pragma unroll
        for (int ka = 0; ka < 2; ++ka) {
          const uint64_t ad = make_sdesc(smem_u32(sW1 + (k * 2 + ka) * A_BYTES));
          const uint64_t bd = make_sdesc(smem_u32(sB1 + ka * ATOM_B));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma<KIND_F16>(d, ad + 2 * kk, bd + 2 * kk, IDESC, (ka | kk) != 0);
        }
        umma_commit(&c_full[k]);
        PSTAMP(it, 1 + k);
      };
      auto g2 = [&](int it, int k) {           // Y += W2[:, chunk k] . u_k^T
        const uint32_t gidx = (uint32_t)it * 3 + k, slot = gidx & 1, yb = (uint32_t)it & 1;
        mbar_wait(&b2_full[slot], (gidx >> 1) & 1, 902);
        if (k == 0) mbar_wait(&y_empty[yb], (((uint32_t)it >> 1) & 1) ^ 1u, 903);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + TR::tm_y(yb);
        const uint64_t ad = make_sdesc(smem_u32(sW2 + k * A_BYTES));
        const uint64_t bd = make_sdesc(smem_u32(sB2 + slot * TR::B2_BYTES));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma<KIND_F16>(d, ad + 2 * kk, bd + 2 * kk, IDESC, (k | kk) != 0);
        umma_commit(&b2_empty[slot]);
        if (k == 2) umma_commit(&y_full[yb]);
        PSTAMP(it, 8 + k);
      };
      if (p.iters > 0) {
        mbar_wait(b1_full, 0, 904);
        tcgen05_fence_after();
        PSTAMP(0, 0);
        for (int k = 0; k < 3; ++k) g1(0, k);
        umma_commit(b1_empty);
      }
      // steady state: GEMM2 of chunk (it, k) is followed at once by GEMM1 of chunk (it+1, k), whose accumulator the same
      // epilogue pass has just released - the tensor pipe never waits at a tile boundary
      for (int it = 0; it < p.iters; ++it) {
        for (int k = 0; k < 3; ++k) {
          g2(it, k);
          if (it + 1 < p.iters) {
            if (k == 0) { mbar_wait(b1_full, (uint32_t)(it + 1) & 1, 905); tcgen05_fence_after(); PSTAMP(it + 1, 0); }
            g1(it + 1, k);
            if (k == 2) umma_commit(b1_empty);
          }
        }
      }
    }
  }
  // =============================================================================== warps 2-5: LayerNorm -> stage-1 operand
  else if (warp < 6) {
    const int pw = warp - 2;
    const int sub = lane >> 3, j = lane & 7;       // 8 lanes share a row: lane j holds float4 j, j+8, j+16, j+24 of it
    for (int it = 0; it < p.iters; ++it) {
      const int tile = tile_of(it);
      const bool live = tile < p.num_tiles;
      const int n = live ? tile / p.tiles_per_row : 0, t0 = live ? (tile % p.tiles_per_row) * NV : 0;
      const float4* x4 = reinterpret_cast<const float4*>(p.x) + (size_t)n * p.T * (F / 4);
      // the single operand buffer is busy until GEMM1 of the previous tile has been issued and completed: request all 24
      // rows of this warp BEFORE waiting for it, so the wait hides the global latency (96 registers of loads in flight)
      float4 v[6][4];
#pragma unroll
      for (int g = 0; g < 6; ++g) {
        const int r = 4 * (pw + 4 * g) + sub, t = t0 - 1 + r;
        const bool ok = live && t >= 0 && t < p.T;
#pragma unroll
        for (int k = 0; k < 4; ++k) v[g][k] = ok ? __ldg(x4 + (size_t)t * (F / 4) + j + 8 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      mbar_wait(b1_empty, (uint32_t)(it & 1) ^ 1u, 910);
      if (warp == 2 && lane == 0) PSTAMP(it, 16);
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
        for (int k = 0; k < 4; ++k) store_c4<KIND_F16>(sB1, ATOM_B, r, j + 8 * k, v[g][k], rstd);
      }
      fence_proxy_async();
      mbar_arrive(b1_full);
      if (warp == 2 && lane == 0) PSTAMP(it, 17);
    }
  }
  // =============================================================================== warps 6-13: gated conv, drain, exchange
  else {
    const int eg = (warp - 6) >> 2;               // epilogue group: owns stage-2 slab eg and every second chunk
    const int q = pwarp & 3;                      // TMEM lane quarter
    const bool gate_warp = q >= 2;                // lanes 64-127 of a mixed chunk are gates
    const int pi = eg * 2 + (q & 1);              // warp pair (value warp q <-> gate warp q + 2)
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    const uint32_t xg0 = smem_u32(sXg) + (uint32_t)pi * 2048u + (uint32_t)lane * 8u;     // + m*256: [8 column pairs][32 lanes] float2
    // Hand-off of t = tanh(conv(g)) from the gate warp to the value warp of a pair: one 16-column buffer and two named
    // barriers (64 threads each) - FULL: gate arrives after its stores, value syncs before its loads; EMPTY: value
    // arrives right after its loads (before its own arithmetic), gate syncs before overwriting.  Named barriers resolve
    // in ~50 clk; the mbarrier poll round trip this replaced cost ~400 clk per hand-off and bounded the whole kernel.
    const int bar_full = 1 + pi, bar_empty = 5 + pi;
    if (!gate_warp) asm volatile("bar.arrive %0, 64;" ::"r"(bar_empty) : "memory");   // the buffer starts empty
    // value warps: store bases into this group's stage-2 slab for the 8 possible (column & 7)
    uint32_t sb0[8];
    {
      const int jch = (q & 1) * 32 + lane;        // channel within the chunk's 64
#pragma unroll
      for (int m = 0; m < 8; ++m)
        sb0[m] = smem_u32(sB2) + (uint32_t)(eg * TR::B2_BYTES + m * 128 + ((((jch >> 3)) ^ m) << 4) + (jch & 7) * 2);
    }
    const int ch = q * 32 + lane;                 // drain: output channel == TMEM lane
    const float bias2 = __ldg(p.b2 + ch), s2i = __ldg(p.s2inv + ch);
    const uint32_t recv_local = smem_u32(sRecv) + (uint32_t)ch * 4u;                   // + col * 512
    const uint32_t recv_remote = mapa_u32(smem_u32(sRecv), peer) + (uint32_t)ch * 4u;
    const uint32_t r_recv_full = mapa_u32(smem_u32(recv_full), peer), r_recv_free = mapa_u32(smem_u32(recv_free), peer);

    // y = x + (Y_mine + Y_peer) * s2inv + b2 for my half of the frames of tile `it`; the other half of my partial goes
    // to the peer.  Runs at the start of iteration it+1 in the group that then owns the smaller share of the chunks.
    auto drain = [&](int it) {
      const int tile = tile_of(it);
      const bool live = tile < p.num_tiles;
      const int n = live ? tile / p.tiles_per_row : 0, t0 = live ? (tile % p.tiles_per_row) * NV : 0;
      const int cmax = live ? min(NV, p.T - t0) : 0;               // columns 1 .. cmax are frames of the utterance
      const uint32_t yb = (uint32_t)it & 1;
      const int own0 = (int)crank * HALF, peer0 = (int)peer * HALF;
      const float* xcol = p.x + (((long long)n * p.T + t0 - 1) * F + ch);
      float* ycol = p.y + (((long long)n * p.T + t0 - 1) * F + ch);
      float xin[HALF];                         // residual values of my half, requested before any wait
#pragma unroll
      for (int i = 0; i < HALF; ++i) xin[i] = (own0 + i >= 1 && own0 + i <= cmax) ? ldg_now(xcol + (own0 + i) * F) : 0.f;
      if (q == 0 && lane == 0) PSTAMP(it, 48);
      mbar_wait_cl(recv_free, ((uint32_t)it & 1) ^ 1u, 920);       // the peer has read what I sent for the previous tile
      if (q == 0 && lane == 0) PSTAMP(it, 49);
      mbar_wait(&y_full[yb], ((uint32_t)it >> 1) & 1, 921);
      tcgen05_fence_after();
      if (q == 0 && lane == 0) PSTAMP(it, 50);
      const uint32_t ty = tmem_base + tlane + TR::tm_y(yb);
#pragma unroll 1
      for (int cb = 0; cb < HALF; cb += 16) {                       // ---- the peer's frames: ship my partial
        uint32_t r[16];
        tmem_ld16(ty + peer0 + cb, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) st_cluster_f32(recv_remote + (uint32_t)(cb + i) * 512u, __uint_as_float(r[i]));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(r_recv_full);
      if (q == 0 && lane == 0) PSTAMP(it, 51);
      mbar_wait_cl(recv_full, (uint32_t)it & 1, 922);               // the peer's partial for my frames has landed
      if (q == 0 && lane == 0) PSTAMP(it, 52);
#pragma unroll
      for (int cb = 0; cb < HALF; cb += 16) {                       // ---- my frames: add, residual, store
        uint32_t r[16];
        tmem_ld16(ty + own0 + cb, r);
        float pr[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pr[i] = lds_f32(recv_local + (uint32_t)(cb + i) * 512u);
        tmem_wait_ld();
        if (cb + 16 == HALF) { tcgen05_fence_before(); mbar_arrive(&y_empty[yb]); }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = own0 + cb + i;
          if (c >= 1 && c <= cmax) ycol[c * F] = fmaf(__uint_as_float(r[i]) + pr[i], s2i, xin[cb + i] + bias2);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(r_recv_free);               // my receive buffer may be overwritten
      if (q == 0 && lane == 0) PSTAMP(it, 53);
    };

    for (int it = 0; it <= p.iters; ++it) {
      if (it > 0 && eg == ((it - 1) & 1)) drain(it - 1);
      if (it == p.iters) break;
      const int tile = tile_of(it);
      const bool live = tile < p.num_tiles;
      const int t0 = live ? (tile % p.tiles_per_row) * NV : 0;
      // columns holding the first / last frame of the utterance, if this tile has them (else out of range)
      const int cL = (live && t0 == 0) ? 1 : -64;
      const int cR = (live && p.T - t0 >= 1 && p.T - t0 <= NV) ? p.T - t0 : -64;
      const bool edge = cL > 0 || cR > 0;
#pragma unroll 1
      for (int k = 0; k < 3; ++k) {
        const uint32_t gidx = (uint32_t)it * 3 + k;
        if ((int)(gidx & 1) != eg) continue;
        const int row = (int)crank * 384 + k * 128 + q * 32 + lane;          // packed GEMM1 row of this TMEM lane
        const float cst = ldg_now(p.cb + row);
        const float w0 = ldg_now(p.dwf + row), w1 = ldg_now(p.dwf + TR::ROWS + row), w2 = ldg_now(p.dwf + 2 * TR::ROWS + row);
        const float fl = edge ? __ldg(p.kl + row) : 0.f, fr = edge ? __ldg(p.kr + row) : 0.f;
        mbar_wait(&c_full[k], (uint32_t)it & 1, 930);
        if ((q & 1) == 0 && lane == 0) PSTAMP(it, 20 + 8 * k + (gate_warp ? 4 : 0));
        if (!gate_warp) mbar_wait(&b2_empty[eg], ((gidx >> 1) & 1) ^ 1u, 931);
        tcgen05_fence_after();
        if ((q & 1) == 0 && lane == 0) PSTAMP(it, 21 + 8 * k + (gate_warp ? 4 : 0));
        const uint32_t tc0 = tmem_base + tlane + TR::tm_c(k);
        const float2 cp = make_float2(cst, cst), w0p = make_float2(w0, w0), w1p = make_float2(w1, w1), w2p = make_float2(w2, w2);
        float2 e_p = make_float2(0.f, 0.f), o_p = e_p;       // carried column pairs (cb-2, cb-1) and (cb-1, cb)
        uint32_t sb[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) sb[m] = sb0[m];
#pragma unroll 1
        for (int cb = 0; cb < NTOK; cb += 16) {
          uint32_t e[16], o[16];
          tmem_ld16(tc0 + cb, e);
          tmem_ld16(tc0 + cb + 1, o);              // last batch: column NTOK belongs to the next region; only carried, never used
          tmem_wait_ld();
          if (cb + 16 == NTOK) { tcgen05_fence_before(); mbar_arrive(&c_empty[k]); if ((q & 1) == 0 && lane == 0) PSTAMP(it, 22 + 8 * k + (gate_warp ? 4 : 0)); }
          // conv over column pairs (cb-1+2m, cb+2m), m = 0..7 (see the packed epilogue of k_gcfn)
          float2 d[8];
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            const float2 a = m == 0 ? e_p : make_float2(__uint_as_float(e[2 * m - 2]), __uint_as_float(e[2 * m - 1]));
            const float2 b = m == 0 ? o_p : make_float2(__uint_as_float(o[2 * m - 2]), __uint_as_float(o[2 * m - 1]));
            const float2 c = make_float2(__uint_as_float(e[2 * m]), __uint_as_float(e[2 * m + 1]));
            d[m] = __ffma2_rn(w2p, c, __ffma2_rn(w1p, b, __ffma2_rn(w0p, a, cp)));
          }
          e_p = make_float2(__uint_as_float(e[14]), __uint_as_float(e[15]));
          o_p = make_float2(__uint_as_float(o[14]), __uint_as_float(o[15]));
          if (edge) {                               // the frame next to an utterance boundary saw h = b1 instead of 0 there
#pragma unroll
            for (int m = 0; m < 8; ++m) {
              const int c0 = cb - 1 + 2 * m;
              d[m].x -= (c0 == cL ? fl : 0.f) + (c0 == cR ? fr : 0.f);
              d[m].y -= (c0 + 1 == cL ? fl : 0.f) + (c0 + 1 == cR ? fr : 0.f);
            }
          }
          if (gate_warp) {
            float2 th[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) th[m] = make_float2(tanh_approx(d[m].x), tanh_approx(d[m].y));
            asm volatile("bar.sync %0, 64;" ::"r"(bar_empty) : "memory");
#pragma unroll
            for (int m = 0; m < 8; ++m) sts_f32x2(xg0 + (uint32_t)(m * 256), th[m]);
            asm volatile("bar.arrive %0, 64;" ::"r"(bar_full) : "memory");
          } else {
            float2 th[8];
            asm volatile("bar.sync %0, 64;" ::"r"(bar_full) : "memory");
#pragma unroll
            for (int m = 0; m < 8; ++m) th[m] = lds_f32x2(xg0 + (uint32_t)(m * 256));
            asm volatile("bar.arrive %0, 64;" ::"r"(bar_empty) : "memory");
#pragma unroll
            for (int m = 0; m < 8; ++m) {
              const float2 u = __ffma2_rn(d[m], th[m], d[m]);
              // columns cb-1+2m (x) and cb+2m (y): (column & 7) is (2m+7)&7 resp. (2m)&7 because cb % 16 == 0
              if (m > 0 || cb > 0) sts_elem<KIND_F16>(sb[(2 * m + 7) & 7] + (uint32_t)(((2 * m - 1) >> 3) * 1024), u.x);
              sts_elem<KIND_F16>(sb[(2 * m) & 7] + (uint32_t)(((2 * m) >> 3) * 1024), u.y);
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) sb[m] += 2048u;
          }
        }
        if (!gate_warp) { fence_proxy_async(); mbar_arrive(&b2_full[eg]); }
        if ((q & 1) == 0 && lane == 0) PSTAMP(it, 23 + 8 * k + (gate_warp ? 4 : 0));
      }
    }
  }

  // ---- teardown
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                        // nobody leaves while the peer may still write into / signal this CTA
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
#undef PSTAMP
}

// ------------------------------------------------------------------------------------------------ host side
inline int prepare_gcfn_pair(GcfnPairPack& g) {
  if (make_weight_map(&g.map_w1, g.w1, KIND_F16, PairTraits::ROWS, PairTraits::F, 128)) return -1;
  g.ready = true;
  return 0;
}

inline int launch_gcfn_pair(const GcfnPairPack& gp, const GcfnPack& g, const float* x, float* y, int rows, int T, int sm_count,
                            cudaStream_t st, long long* dbg_clk = nullptr) {
  using TR = PairTraits;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaError_t e;
  static bool attr_set[16] = {};
  if (!attr_set[dev & 15]) {
    e = cudaFuncSetAttribute(k_gcfn_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, TR::SMEM_BYTES);
    if (e != cudaSuccess) { snprintf(g_tc_err, sizeof(g_tc_err), "cudaFuncSetAttribute(k_gcfn_pair): %s", cudaGetErrorString(e)); return -1; }
    attr_set[dev & 15] = true;
  }
  GcfnPairParams p{};
  p.x = x; p.y = y; p.cb = gp.cb; p.dwf = gp.dwf; p.kl = gp.kl; p.kr = gp.kr; p.b2 = g.b2; p.s2inv = g.s2inv[KIND_F16];
  p.rows = rows; p.T = T; p.dbg_clk = dbg_clk;
  p.tiles_per_row = (T + TR::NV - 1) / TR::NV;
  p.num_tiles = rows * p.tiles_per_row;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(TR::THREADS);
  cfg.dynamicSmemBytes = TR::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static int max_pairs[16] = {0};          // CTA pairs that can be co-resident (a pair sits on one TPC)
  if (max_pairs[dev & 15] == 0) {
    cfg.gridDim = dim3((sm_count / 2) * 2);
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, k_gcfn_pair, &cfg);
    if (e != cudaSuccess || n <= 0) { cudaGetLastError(); n = sm_count / 2; }
    max_pairs[dev & 15] = n;
  }
  int pairs = max_pairs[dev & 15];
  if (pairs > p.num_tiles) pairs = p.num_tiles;
  p.iters = (p.num_tiles + pairs - 1) / pairs;
  cfg.gridDim = dim3(2 * pairs);
  e = cudaLaunchKernelEx(&cfg, k_gcfn_pair, gp.map_w1, g.map_w2[KIND_F16][0], p);
  if (e != cudaSuccess) { snprintf(g_tc_err, sizeof(g_tc_err), "k_gcfn_pair launch: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

}  // namespace tc
}  // namespace sepref
