// k_gcfn_tm: the GCFN block (reference network.py:60-66) with FRAMES as the MMA M dimension ("token-major") and the
// weights as the N operand, shared by a CTA pair through tcgen05.mma.cta_group::2.
//
// Why: k_gcfn (channels = M, frames = N) is bound by per-SM operand ingest - TMEM caps its tile at 96 frames, and every
// tile pulls the block's whole weight set (295 KB as FP16) through an L2->SM port that tops out at 37.8 B/clk
// (profiles/r1_ingest_microbench.md; k_gcfn runs at 76 % of it).  Two earlier attempts to keep the weights resident by
// splitting the hidden dimension over a cluster lost to structure.  Here the roles of the operands are
// swapped instead:
//
//   D[frames, channels] = A[frames, K] . B[channels, K]^T      A = activations produced on chip, B = weight slabs by TMA
//
// * a CTA owns 128 accumulator rows = 4 warp segments of 32 frames (30 produced + one halo frame each side), so a tile
//   is 120 frames instead of 94, and TMEM holds 3 x (64 values | 64 gates) + 128 output columns = 512;
// * with cta_group::2 the two CTAs of a pair run ONE M = 256 instruction stream: each CTA supplies its own 128 frame
//   rows and HALF of every weight slab (the N operand is split across the pair and read by both tensor cores), so a CTA
//   ingests 147 KB of weights per 120 frames where k_gcfn ingests 295 KB per 94: 2.5x fewer weight bytes per frame;
// * the k = 3 time convolution now runs ACROSS accumulator rows.  The epilogue reads TMEM with the 16x256b shape, whose
//   fragment gives a thread rows (q, q+8) of a 16-row block and two neighbouring columns; with the row <-> frame map
//   row = 32w + q + 8m  <->  frame 4q + m of warp segment w, a thread holds a RUN of four consecutive frames of two
//   channels, so only the two frames next to a run come from other threads (two shuffles by 4 lanes per run instead of
//   two per element), and the channel pair is a packed fma.rn.f32x2 operand.  Per-channel taps are per-thread data
//   again (a thread sees 16 + 16 of a chunk's 128 columns): one 16-byte shared load per channel pair and tap set.
//
// Zero padding of the convolution as in k_gcfn_pair: frames outside the utterance are zero operand rows (accumulator
// exactly 0), the interior formula folds b1 into the conv constant, and the first / last frame of an utterance subtract
// kl = w0*b1/2 / kr = w2*b1/2 (warp-uniform slow path for the segments that touch an utterance boundary).
//
// PAIR = false is the same kernel on one CTA (cta_group::1, whole weight slabs): the bring-up and fallback variant.
#pragma once
#include "kernels_gcfn_pair.cuh"

namespace sepref {
namespace tc {

template <bool PAIR>
struct TmTraits {      // F = 128, FP16 operands
  static constexpr int F = 128;
  static constexpr int NCTA = PAIR ? 2 : 1;
  static constexpr int ROWS = 128;                       // accumulator rows (frames incl. halos) per CTA
  static constexpr int SEG = 30;                         // frames a 32-row warp segment produces
  static constexpr int NCH = 6;                          // chunks of 64 GLU channels (64 values | 64 gates = 128 columns)
  static constexpr int A_SLAB = ROWS * 128;              // [128 rows x 128 B] swizzled slab
  static constexpr int A1_BYTES = 2 * A_SLAB;            // LayerNorm'd frames, K = 128
  static constexpr int H_BYTES = A_SLAB;                 // gated hidden chunk, K = 64
  static constexpr int WROWS = 128 / NCTA;               // weight rows per slab held by one CTA
  static constexpr int W_SLOT = WROWS * 128;
  static constexpr int NG = 3;                           // epilogue groups == accumulator buffers == hidden-chunk buffers
  static constexpr int NS1 = PAIR ? 8 : 3;               // GEMM1 weight ring (slots of one k slab)
  static constexpr int NS2 = PAIR ? 3 : 2;               // GEMM2 weight ring
  // per-thread constants as float4 entries, the four column pairs c of a k block 16 B apart (conflict-free LDS.128):
  static constexpr int TAP_FLOATS = NCH * 8 * 2 * 2 * 4 * 4;   // [chunk][k block][value|gate][half][c]: (w0a w0b w1a w1b) | (w2a w2b ca cb)
  static constexpr int EDGE_FLOATS = NCH * 8 * 2 * 4 * 4;      // [chunk][k block][value|gate][c]: (kla klb kra krb)
  static constexpr int TAB_FLOATS = TAP_FLOATS + EDGE_FLOATS + 2 * F;   // + s2inv[F], b2[F]
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_BYTES = 1024 + 2 * A1_BYTES + NG * H_BYTES + (NS1 + NS2) * W_SLOT + TAB_FLOATS * 4 + BAR_BYTES;
  static constexpr int THREADS = (4 * NG + 8) * 32;      // NG epilogue groups of 4 warps, 4 producer warps, 4 single-lane roles
  static constexpr int TM_Y = NG * 128;                  // TMEM columns: NG accumulators of 128, then 128 of Y
  static_assert(NCH % NG == 0 && TM_Y + 128 <= 512, "chunk -> group map, TMEM columns");
  static constexpr int ROWS_W1 = NCH * 128;              // packed GEMM1 rows
  static_assert(SMEM_BYTES <= 232448, "shared memory");
};

struct GcfnTmPack {
  // GEMM1 rows in chunk order: row (j*128 + vg*64 + i) is the value (vg = 0) / gate (vg = 1) row of GLU channel j*64 + i;
  // FP16, per-row power-of-two scaled
  const void* w1 = nullptr;
  const float* tab = nullptr;     // [TAB_FLOATS]: taps | edge constants | s2inv | b2 (layout in TmTraits)
  alignas(64) CUtensorMap map_w1[2];      // box rows 128 (single CTA) / 64 (pair)
  bool ready = false;
};

struct GcfnTmParams {
  const float* x;
  float* y;
  const float* tab;
  int rows, T, segs_per_row, num_segs, num_tiles, iters;
  int flags;            // bring-up switches: 1 = the pair's weight halves swapped, 2 = leader barrier address by bit mask instead of mapa
  long long* dbg_clk;   // optional [8][64] clock64 stamps of block 0's first 8 tiles
};

__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_16x256b_x1(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_b32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
template <bool PAIR>
__device__ __forceinline__ void umma_f16_tm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if (PAIR) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
  } else {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
  }
}
// completion of every MMA this thread has issued so far -> one arrival on the barrier at this offset in every CTA of the pair
template <bool PAIR>
__device__ __forceinline__ void umma_commit_tm(uint64_t* bar) {
  if (PAIR) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
  } else {
    umma_commit(bar);
  }
}
// Arrival on a barrier of the leader CTA (shared::cluster address).  Default semantics (release at CTA scope), as in
// CUTLASS's ClusterBarrier::arrive: the .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR (measured: it turned
// every hand-off into a > 1 k clk stall).  What the arrival orders here is local to the arriving CTA - its TMEM reads
// (tcgen05.fence::before_thread_sync) and its own shared-memory operand writes (fence.proxy.async), both consumed by that
// CTA's tensor core under an instruction the leader issues after it has observed the arrival.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t caddr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ float2 lds_f32x2_tm(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
// one half slab into this CTA's ring slot; the bytes are counted on the LEADER CTA's barrier (address from mapa)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t lead_bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(lead_bar), "r"(c0), "r"(c1)
      : "memory");
}

// ------------------------------------------------------------------------------------------------ the kernel
template <bool PAIR>
__global__ void __launch_bounds__(TmTraits<PAIR>::THREADS, 1)
k_gcfn_tm(const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2, const GcfnTmParams p) {
  using TR = TmTraits<PAIR>;
  constexpr int F = TR::F, NCTA = TR::NCTA, NCH = TR::NCH, SEG = TR::SEG, A_SLAB = TR::A_SLAB, A1_BYTES = TR::A1_BYTES, H_BYTES = TR::H_BYTES;
  constexpr int W_SLOT = TR::W_SLOT, WROWS = TR::WROWS, NS1 = TR::NS1, NS2 = TR::NS2, NG = TR::NG;
  constexpr uint32_t IDESC = make_idesc<KIND_F16>(128 * NCTA, 128);

  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA1 = sm;                                   // [2] LayerNorm'd frame tiles
  unsigned char* sH = sA1 + 2 * A1_BYTES;                    // [NG] gated hidden chunks (one per epilogue group)
  unsigned char* sW1 = sH + NG * H_BYTES;                    // [NS1] GEMM1 weight ring
  unsigned char* sW2 = sW1 + NS1 * W_SLOT;                   // [NS2] GEMM2 weight ring
  float* sTab = reinterpret_cast<float*>(sW2 + NS2 * W_SLOT);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(sTab) + TR::TAB_FLOATS * 4);
  // waited on by the LEADER's issuing threads, arrived on by both CTAs:
  uint64_t* w1_full = bars;                  // [NS1]
  uint64_t* w2_full = w1_full + NS1;         // [NS2]
  uint64_t* a_full = w2_full + NS2;          // [2]  frame tile written (4 producer warps per CTA)
  uint64_t* acc_empty = a_full + 2;          // [NG] accumulator read out (4 epilogue warps per CTA)
  uint64_t* h_full = acc_empty + NG;         // [NG] hidden chunk written
  uint64_t* y_empty = h_full + NG;           // [1]  Y read out
  // arrived on by the leader's tcgen05.commit in BOTH CTAs, waited on locally:
  uint64_t* w1_empty = y_empty + 1;          // [NS1]
  uint64_t* w2_empty = w1_empty + NS1;       // [NS2]
  uint64_t* a_empty = w2_empty + NS2;        // [2]
  uint64_t* acc_full = a_empty + 2;          // [NG]
  uint64_t* h_empty = acc_full + NG;         // [NG]
  uint64_t* y_full = h_empty + NG;           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y_full + 1);

  const int pwarp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // roles by physical warp: 0-11 epilogue (group = pwarp >> 2, TMEM lane quarter = pwarp & 3), 12-15 producers,
  // 16 / 17 TMA of the GEMM1 / GEMM2 ring, 18 / 19 GEMM1 / GEMM2 issue (highest warp ids: highest arbitration priority)
  constexpr int W_PRO = 4 * NG, W_TMA1 = W_PRO + 4, W_TMA2 = W_PRO + 5, W_MMA1 = W_PRO + 6, W_MMA2 = W_PRO + 7;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = crank == 0;
  const int cid = (int)blockIdx.x / NCTA, ncl = (int)gridDim.x / NCTA;
#define TSTAMP(itv, slot) do { if (p.dbg_clk != nullptr && blockIdx.x == 0 && (itv) < 8) p.dbg_clk[(itv) * 64 + (slot)] = clock64(); } while (0)

  // arrive on a leader-side barrier (release at cluster scope in a pair)
  auto arrive_lead = [&](uint64_t* bar) {
    if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(bar), 0));
    else mbar_arrive(bar);
  };
  auto wait_lead = [&](uint64_t* bar, uint32_t parity, int tag) {
    mbar_wait(bar, parity, tag);
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS1; ++i) { mbar_init(&w1_full[i], 1); mbar_init(&w1_empty[i], 1); }
    for (int i = 0; i < NS2; ++i) { mbar_init(&w2_full[i], 1); mbar_init(&w2_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 4 * NCTA); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NG; ++i) {
      mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4 * NCTA);
      mbar_init(&h_full[i], 4 * NCTA); mbar_init(&h_empty[i], 1);
    }
    mbar_init(y_full, 1); mbar_init(y_empty, 4 * NG * NCTA);
    fence_barrier_init();
  }
  if (pwarp == W_TMA1 && lane == 0) { tma_prefetch_desc(&map_w1); tma_prefetch_desc(&map_w2); }
  if (pwarp == W_MMA1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  // per-channel constants of this block (weights: no kernel writes them, so they may be read ahead of pdl_wait)
  for (int i = threadIdx.x; i < TR::TAB_FLOATS / 4; i += TR::THREADS)
    reinterpret_cast<float4*>(sTab)[i] = __ldg(reinterpret_cast<const float4*>(p.tab) + i);
  // hidden chunks start finite (every row is rewritten per chunk; this only covers the first use)
  for (int i = threadIdx.x; i < (NG * H_BYTES) / 16; i += TR::THREADS) reinterpret_cast<uint4*>(sH)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();              // the peer's barriers exist before anyone signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  if (pwarp != W_TMA1 && pwarp != W_TMA2) pdl_wait();     // the TMA threads only stream weights

  // iteration it of this CTA works on tile (cid + it * ncl) * NCTA + crank; tiles past the last one are dummies
  auto tile_of = [&](int it) { return (cid + it * ncl) * NCTA + (int)crank; };
  const int total = p.iters * NCH;

  // =============================================================================== weight slabs via TMA (two single-lane roles)
  if (pwarp == W_TMA1 || pwarp == W_TMA2) {
    if (lane == 0) {
      const bool g1 = pwarp == W_TMA1;
      uint64_t* full = g1 ? w1_full : w2_full;
      uint64_t* empty = g1 ? w1_empty : w2_empty;
      unsigned char* ring = g1 ? sW1 : sW2;
      const int rn = g1 ? NS1 : NS2;
      const CUtensorMap* map = g1 ? &map_w1 : &map_w2;
      int st = 0; uint32_t ph = 0;
      auto load = [&](int c0, int c1) {
        mbar_wait(&empty[st], ph ^ 1, 100);
        if (leader) mbar_arrive_expect_tx(&full[st], NCTA * W_SLOT);
        if (PAIR) tma_load_2d_pair(map, (p.flags & 2) ? (smem_u32(&full[st]) & 0xFEFFFFFFu) : mapa_u32(smem_u32(&full[st]), 0), ring + st * W_SLOT, c0, c1);
        else tma_load_2d(map, &full[st], ring + st * W_SLOT, c0, c1);
        if (++st == rn) { st = 0; ph ^= 1; }
      };
      const int half = (PAIR && (p.flags & 1)) ? 1 - (int)crank : (int)crank;     // which half of every slab this CTA supplies
      for (int g = 0; g < total; ++g) {
        const int j = g % NCH;
        if (g1) { load(0, j * 128 + half * WROWS); load(64, j * 128 + half * WROWS); }
        else load(j * 64, half * WROWS);
      }
    }
  }
  // =============================================================================== GEMM1 issue (leader CTA)
  else if (pwarp == W_MMA1) {
    if (lane == 0 && leader) {
      int st = 0; uint32_t ph = 0;
      for (int g = 0; g < total; ++g) {
        const int it = g / NCH, j = g % NCH;
        const int ab = it & 1;
        const uint32_t e = (uint32_t)g % NG, n = (uint32_t)g / NG;         // chunk g belongs to epilogue group / accumulator g % NG
        if (j == 0) { wait_lead(&a_full[ab], (uint32_t)(it >> 1) & 1, 200); TSTAMP(it, 0); }
        wait_lead(&acc_empty[e], (n & 1) ^ 1, 201);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + e * 128;
        for (int ka = 0; ka < 2; ++ka) {
          wait_lead(&w1_full[st], ph, 202);
          tcgen05_fence_after();
          const uint64_t ad = make_sdesc(smem_u32(sA1 + ab * A1_BYTES + ka * A_SLAB));
          const uint64_t bd = make_sdesc(smem_u32(sW1 + st * W_SLOT));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_tm<PAIR>(d, ad + 2 * k, bd + 2 * k, IDESC, (ka | k) != 0);
          umma_commit_tm<PAIR>(&w1_empty[st]);
          if (++st == NS1) { st = 0; ph ^= 1; }
        }
        umma_commit_tm<PAIR>(&acc_full[e]);
        TSTAMP(it, 1 + j);
        if (j == NCH - 1) umma_commit_tm<PAIR>(&a_empty[ab]);     // every GEMM1 MMA of this tile has been issued
      }
    }
  }
  // =============================================================================== GEMM2 issue (leader CTA)
  else if (pwarp == W_MMA2) {
    if (lane == 0 && leader) {
      int st = 0; uint32_t ph = 0;
      const uint32_t d = tmem_base + TR::TM_Y;
      for (int g = 0; g < total; ++g) {
        const int it = g / NCH, j = g % NCH;
        const uint32_t e = (uint32_t)g % NG, n = (uint32_t)g / NG;
        wait_lead(&h_full[e], n & 1, 210);
        if (j == 0) wait_lead(y_empty, (uint32_t)(it & 1) ^ 1, 211);
        wait_lead(&w2_full[st], ph, 212);
        tcgen05_fence_after();
        const uint64_t ad = make_sdesc(smem_u32(sH + e * H_BYTES));
        const uint64_t bd = make_sdesc(smem_u32(sW2 + st * W_SLOT));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_tm<PAIR>(d, ad + 2 * k, bd + 2 * k, IDESC, (j | k) != 0);
        umma_commit_tm<PAIR>(&w2_empty[st]);
        if (++st == NS2) { st = 0; ph ^= 1; }
        umma_commit_tm<PAIR>(&h_empty[e]);
        TSTAMP(it, 8 + j);
        if (j == NCH - 1) umma_commit_tm<PAIR>(y_full);
      }
    }
  }
  // =============================================================================== frame-tile producer (4 warps)
  else if (pwarp >= W_PRO) {
    const int pw = pwarp - W_PRO;
    const float4* x4 = reinterpret_cast<const float4*>(p.x);
    for (int it = 0; it < p.iters; ++it) {
      const int tile = tile_of(it);
      const int ab = it & 1;
      mbar_wait(&a_empty[ab], ((uint32_t)(it >> 1) & 1) ^ 1, 300);
      if (pw == 0 && lane == 0) TSTAMP(it, 16);
      const int T = p.T, spr = p.segs_per_row, nseg = p.num_segs;
      produce_rows<KIND_F16, F, 128, true>(sA1 + ab * A1_BYTES, A_SLAB, pw, lane, 1, [&](int r, int c4, int) {
        // row r = 32w + q + 8m  <->  frame 4q + m - 1 of warp segment w (frames -1 and 30 are the halos)
        const int seg = tile * 4 + (r >> 5);
        const int n = seg / spr, t = (seg - n * spr) * SEG + 4 * (r & 7) + ((r >> 3) & 3) - 1;
        return (seg < nseg && t >= 0 && t < T) ? __ldg(x4 + ((size_t)n * T + t) * (F / 4) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      });
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) arrive_lead(&a_full[ab]);
      if (pw == 0 && lane == 0) TSTAMP(it, 17);
    }
  }
  // =============================================================================== gated-conv epilogue and drain (NG groups of 4 warps)
  else {
    const int eg = pwarp >> 2, q4 = pwarp & 3;
    const int c = lane & 3, q = lane >> 2;
    const uint32_t tq = (uint32_t)(q4 * 32) << 16;
    const uint32_t sTab32 = smem_u32(sTab);
    const uint32_t hb = smem_u32(sH + eg * H_BYTES) + (uint32_t)(4 * q4) * 1024u + (uint32_t)q * 128u + (uint32_t)c * 4u;

    // y = x + Y * s2inv + b2' of the tile whose GEMM2 finished, shared by ALL epilogue groups: group e owns 48 / 48 / 32 of the
    // 128 output channels (units of 16 columns), reads them from TMEM into registers in one go and releases Y at once -
    // GEMM2 of the next tile waits for nothing but these reads; the global loads and stores follow at leisure.  (A single
    // draining group serialised the whole pipeline: GEMM2(it+1) waited ~7 k clk for it, the hidden buffers behind GEMM2.)
    // Y is read with the 16x256b shape as well: the four threads that share a row hold 8 consecutive channels of it per
    // k block, so every 8-byte global access of a warp fills whole 32-byte sectors.
    auto drain = [&](int tile, int it) {
      const int seg = tile * 4 + q4;
      const int n = seg / p.segs_per_row, ts0 = (seg - n * p.segs_per_row) * SEG;
      bool ok[4];
      size_t off[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int l = 4 * q + m, t = ts0 + l - 1;                         // frame index in the segment (0 and 31: halos)
        ok[m] = seg < p.num_segs && l >= 1 && l <= SEG && t < p.T;
        off[m] = ok[m] ? ((size_t)n * p.T + t) * F + 2 * c : (size_t)(2 * c);
      }
      const int u0 = eg * 3, nu = eg == NG - 1 ? 2 : 3;                   // this group's 16-column units [u0, u0 + nu)
      const uint32_t cst = sTab32 + (uint32_t)(TR::TAP_FLOATS + TR::EDGE_FLOATS + 2 * c) * 4u;
      const uint32_t ty = tmem_base + tq + TR::TM_Y + (uint32_t)(16 * u0);
      // x of the first unit is requested ahead of the wait for Y, x of unit u+1 ahead of the arithmetic of unit u
      auto load_x = [&](int u, float2 (&xin)[4][2]) {
        const int ch0 = 16 * (u0 + u);
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
          for (int k = 0; k < 2; ++k)
            xin[m][k] = ok[m] ? __ldg(reinterpret_cast<const float2*>(p.x + off[m] + ch0 + 8 * k)) : make_float2(0.f, 0.f);
      };
      float2 xa[4][2], xb[4][2];
      load_x(0, xa);
      mbar_wait(y_full, (uint32_t)it & 1, 400);
      tcgen05_fence_after();
      if (q4 == 0 && lane == 0 && eg == 0) TSTAMP(it + 1, 19);
      uint32_t r[3][2][8];                                                // [unit][row half][k block, row, column]
#pragma unroll
      for (int u = 0; u < 3; ++u)
        if (u < nu) {
          tmem_ld_16x256b_x2(ty + (uint32_t)(16 * u), r[u][0]);
          tmem_ld_16x256b_x2(ty + (16u << 16) + (uint32_t)(16 * u), r[u][1]);
        }
      tmem_wait_ld();
      tcgen05_fence_before(); __syncwarp();
      if (lane == 0) arrive_lead(y_empty);
      auto finish = [&](int u, const float2 (&xin)[4][2]) {
        const int ch0 = 16 * (u0 + u);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float2 s2 = lds_f32x2_tm(cst + (uint32_t)(ch0 + 8 * k) * 4u), b2 = lds_f32x2_tm(cst + (uint32_t)(F + ch0 + 8 * k) * 4u);
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            float2 o;
            o.x = fmaf(__uint_as_float(r[u][m >> 1][4 * k + 2 * (m & 1) + 0]), s2.x, xin[m][k].x + b2.x);
            o.y = fmaf(__uint_as_float(r[u][m >> 1][4 * k + 2 * (m & 1) + 1]), s2.y, xin[m][k].y + b2.y);
            if (ok[m]) *reinterpret_cast<float2*>(p.y + off[m] + ch0 + 8 * k) = o;
          }
        }
      };
      load_x(1, xb);
      finish(0, xa);
      if (nu > 2) load_x(2, xa);
      finish(1, xb);
      if (nu > 2) finish(2, xa);
      if (q4 == 0 && lane == 0 && eg == 0) TSTAMP(it + 1, 18);
    };

    for (int it = 0; it <= p.iters; ++it) {
      if (it == p.iters) { if (it > 0) drain(tile_of(it - 1), it - 1); break; }
      const int tile = tile_of(it);
      const int seg = tile * 4 + q4;
      const int n = seg / p.segs_per_row, ts = (seg - n * p.segs_per_row) * SEG;
      // frames ts-1 .. ts+30 of utterance n; an utterance boundary inside (or next to) the produced frames takes the slow path
      const bool edge = ts == 0 || ts + SEG >= p.T;
      float f0[4], f1[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int t = ts + 4 * q + m - 1;
        f0[m] = t == 0 ? 1.f : 0.f;
        f1[m] = t == p.T - 1 ? 1.f : 0.f;
      }
#pragma unroll 1
      for (int j = eg; j < NCH; j += NG) {
        // the previous tile's output is drained between this group's two chunks: its Y is complete by then, and GEMM2 of
        // this tile (which waits for the drain's TMEM reads) is not yet holding anybody up
        if (j >= NG && it > 0) drain(tile_of(it - 1), it - 1);
        const uint32_t g = (uint32_t)it * NCH + j, nuse = g / NG;
        mbar_wait(&acc_full[eg], nuse & 1, 410);
        if (q4 == 0 && lane == 0) TSTAMP(it, 24 + j * 4);
        mbar_wait(&h_empty[eg], (nuse & 1) ^ 1, 411);
        if (q4 == 0 && lane == 0) TSTAMP(it, 25 + j * 4);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + tq + (uint32_t)eg * 128u;
        const uint32_t tapj = sTab32 + (uint32_t)(j * 8 * 64 + c * 4) * 4u;                         // + k*256 B + vg*128 B + half*64 B
        const uint32_t edgj = sTab32 + (uint32_t)(TR::TAP_FLOATS + j * 8 * 32 + c * 4) * 4u;        // + k*128 B + vg*64 B
        // one k block (8 value and 8 gate columns) per step, software-pipelined: the TMEM reads of block k+1 are in flight
        // while block k is computed (two register sets of 16; the 96-register budget has no room for wider steps)
        auto issue = [&](int k, uint32_t (&v0)[4], uint32_t (&v1)[4], uint32_t (&g0)[4], uint32_t (&g1)[4]) {
          tmem_ld_16x256b_x1(tacc + (uint32_t)(8 * k), v0);                         // rows q, q+8      (m = 0, 1)
          tmem_ld_16x256b_x1(tacc + (16u << 16) + (uint32_t)(8 * k), v1);           // rows 16+q, 24+q  (m = 2, 3)
          tmem_ld_16x256b_x1(tacc + 64u + (uint32_t)(8 * k), g0);
          tmem_ld_16x256b_x1(tacc + (16u << 16) + 64u + (uint32_t)(8 * k), g1);
        };
        auto step = [&](int k, const uint32_t (&v0)[4], const uint32_t (&v1)[4], const uint32_t (&g0)[4], const uint32_t (&g1)[4]) {
          {
            constexpr int kk = 0;
            float2 av[4], ag[4];
            av[0] = make_float2(__uint_as_float(v0[4 * kk + 0]), __uint_as_float(v0[4 * kk + 1]));
            av[1] = make_float2(__uint_as_float(v0[4 * kk + 2]), __uint_as_float(v0[4 * kk + 3]));
            av[2] = make_float2(__uint_as_float(v1[4 * kk + 0]), __uint_as_float(v1[4 * kk + 1]));
            av[3] = make_float2(__uint_as_float(v1[4 * kk + 2]), __uint_as_float(v1[4 * kk + 3]));
            ag[0] = make_float2(__uint_as_float(g0[4 * kk + 0]), __uint_as_float(g0[4 * kk + 1]));
            ag[1] = make_float2(__uint_as_float(g0[4 * kk + 2]), __uint_as_float(g0[4 * kk + 3]));
            ag[2] = make_float2(__uint_as_float(g1[4 * kk + 0]), __uint_as_float(g1[4 * kk + 1]));
            ag[3] = make_float2(__uint_as_float(g1[4 * kk + 2]), __uint_as_float(g1[4 * kk + 3]));
            // the frames next to this thread's run of four live four lanes away (same column pair, neighbouring q)
            float2 pv, nv, pg, ng;
            pv.x = __shfl_up_sync(0xffffffffu, av[3].x, 4); pv.y = __shfl_up_sync(0xffffffffu, av[3].y, 4);
            nv.x = __shfl_down_sync(0xffffffffu, av[0].x, 4); nv.y = __shfl_down_sync(0xffffffffu, av[0].y, 4);
            pg.x = __shfl_up_sync(0xffffffffu, ag[3].x, 4); pg.y = __shfl_up_sync(0xffffffffu, ag[3].y, 4);
            ng.x = __shfl_down_sync(0xffffffffu, ag[0].x, 4); ng.y = __shfl_down_sync(0xffffffffu, ag[0].y, 4);
            const float4 tv0 = lds_f32x4(tapj + (uint32_t)k * 256u), tv1 = lds_f32x4(tapj + (uint32_t)k * 256u + 64u);
            const float4 tg0 = lds_f32x4(tapj + (uint32_t)k * 256u + 128u), tg1 = lds_f32x4(tapj + (uint32_t)k * 256u + 192u);
            const float2 wv0 = make_float2(tv0.x, tv0.y), wv1 = make_float2(tv0.z, tv0.w), wv2 = make_float2(tv1.x, tv1.y), cv = make_float2(tv1.z, tv1.w);
            const float2 wg0 = make_float2(tg0.x, tg0.y), wg1 = make_float2(tg0.z, tg0.w), wg2 = make_float2(tg1.x, tg1.y), cg = make_float2(tg1.z, tg1.w);
            float2 dv[4], dg[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const float2 lv = m == 0 ? pv : av[m - 1], rv = m == 3 ? nv : av[m + 1];
              const float2 lg = m == 0 ? pg : ag[m - 1], rg = m == 3 ? ng : ag[m + 1];
              dv[m] = __ffma2_rn(wv2, rv, __ffma2_rn(wv1, av[m], __ffma2_rn(wv0, lv, cv)));
              dg[m] = __ffma2_rn(wg2, rg, __ffma2_rn(wg1, ag[m], __ffma2_rn(wg0, lg, cg)));
            }
            if (edge) {
              const float4 ev = lds_f32x4(edgj + (uint32_t)k * 128u), eg4 = lds_f32x4(edgj + (uint32_t)k * 128u + 64u);
#pragma unroll
              for (int m = 0; m < 4; ++m) {
                dv[m].x -= f0[m] * ev.x + f1[m] * ev.z; dv[m].y -= f0[m] * ev.y + f1[m] * ev.w;
                dg[m].x -= f0[m] * eg4.x + f1[m] * eg4.z; dg[m].y -= f0[m] * eg4.y + f1[m] * eg4.w;
              }
            }
            const uint32_t kx = (uint32_t)(k ^ q) << 4;
            // phase by phase (eight independent MUFU results before their first use) rather than frame by frame
            float2 th[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) th[m] = make_float2(tanh_approx(dg[m].x), tanh_approx(dg[m].y));
            uint32_t uh[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const float2 u = __ffma2_rn(dv[m], th[m], dv[m]);
              uh[m] = pack_f16x2_sat(u.x, u.y);
            }
#pragma unroll
            for (int m = 0; m < 4; ++m) sts_b32(hb + (uint32_t)m * 1024u + kx, uh[m]);
          }
        };
        auto release_acc = [&]() {
          tcgen05_fence_before(); __syncwarp();
          if (lane == 0) arrive_lead(&acc_empty[eg]);
          if (q4 == 0 && lane == 0) TSTAMP(it, 26 + j * 4);
        };
        {
          uint32_t a0[4], a1[4], a2[4], a3[4], b0[4], b1[4], b2[4], b3[4];
          issue(0, a0, a1, a2, a3);
          tmem_wait_ld();
#pragma unroll 1
          for (int k = 0; k < 8; k += 2) {
            issue(k + 1, b0, b1, b2, b3);
            step(k, a0, a1, a2, a3);
            tmem_wait_ld();
            if (k + 2 < 8) issue(k + 2, a0, a1, a2, a3);
            else release_acc();                           // every column of this accumulator has been read
            step(k + 1, b0, b1, b2, b3);
            if (k + 2 < 8) tmem_wait_ld();
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) arrive_lead(&h_full[eg]);
        if (q4 == 0 && lane == 0) TSTAMP(it, 27 + j * 4);
      }
    }
  }

  // ---- teardown
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();              // nobody leaves while the peer may still signal this CTA or read its operands
  if (pwarp == W_MMA1) {
    __syncwarp();
    tcgen05_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
#undef TSTAMP
}

// ------------------------------------------------------------------------------------------------ host side
inline int prepare_gcfn_tm(GcfnTmPack& g) {
  if (g.w1 == nullptr) return 0;
  if (make_weight_map(&g.map_w1[0], g.w1, KIND_F16, TmTraits<false>::ROWS_W1, 128, 128)) return -1;
  if (make_weight_map(&g.map_w1[1], g.w1, KIND_F16, TmTraits<false>::ROWS_W1, 128, 64)) return -1;
  g.ready = true;
  return 0;
}

template <bool PAIR>
inline int launch_gcfn_tm_t(const GcfnTmPack& g, const GcfnPack& base, const float* x, float* y, int rows, int T, int sm_count,
                            cudaStream_t st, long long* dbg_clk, int flags) {
  using TR = TmTraits<PAIR>;
  static bool attr_set[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  cudaError_t e;
  if (dev < 0 || dev >= 16 || !attr_set[dev]) {
    e = cudaFuncSetAttribute(k_gcfn_tm<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, TR::SMEM_BYTES);
    if (e != cudaSuccess) { snprintf(g_tc_err, sizeof(g_tc_err), "cudaFuncSetAttribute(k_gcfn_tm): %s", cudaGetErrorString(e)); return -1; }
    if (dev >= 0 && dev < 16) attr_set[dev] = true;
  }
  GcfnTmParams p{};
  p.x = x; p.y = y; p.tab = g.tab; p.rows = rows; p.T = T; p.dbg_clk = dbg_clk; p.flags = flags;
  p.segs_per_row = (T + TR::SEG - 1) / TR::SEG;
  p.num_segs = rows * p.segs_per_row;
  p.num_tiles = (p.num_segs + 3) / 4;
  const int units = (p.num_tiles + TR::NCTA - 1) / TR::NCTA;            // tile pairs (or tiles)
  int clusters = sm_count / TR::NCTA;
  if (clusters > units) clusters = units;
  p.iters = (units + clusters - 1) / clusters;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * TR::NCTA);
  cfg.blockDim = dim3(TR::THREADS);
  cfg.dynamicSmemBytes = TR::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  if (PAIR) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  e = cudaLaunchKernelEx(&cfg, k_gcfn_tm<PAIR>, g.map_w1[PAIR ? 1 : 0], base.map_w2[KIND_F16][PAIR ? 1 : 0], p);
  if (e != cudaSuccess) { snprintf(g_tc_err, sizeof(g_tc_err), "k_gcfn_tm launch: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

// mode 1: CTA pair (cta_group::2), mode 2: single CTA; modes 3-5: the pair with bring-up flags 1-3
inline int launch_gcfn_tm(const GcfnTmPack& g, const GcfnPack& base, const float* x, float* y, int rows, int T, int sm_count,
                          cudaStream_t st, int mode, long long* dbg_clk = nullptr) {
  if (!g.ready) { snprintf(g_tc_err, sizeof(g_tc_err), "k_gcfn_tm not prepared"); return -1; }
  return mode == 2 ? launch_gcfn_tm_t<false>(g, base, x, y, rows, T, sm_count, st, dbg_clk, 0)
                   : launch_gcfn_tm_t<true>(g, base, x, y, rows, T, sm_count, st, dbg_clk, mode >= 3 ? mode - 2 : 0);
}

}  // namespace tc
}  // namespace sepref
