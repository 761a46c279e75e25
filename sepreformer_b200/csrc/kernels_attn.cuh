// Pooled multi-head attention with Shaw-style relative-position key bias
// (reference network.py:103-122 as called from EGA, network.py:145-149; table from module.py:42-57,196-198).
//
//   S[i,j] = q_i . k_j + q_i . E[clamp(i-j, -maxlen, maxlen-1) + maxlen]      (q pre-scaled by 1/sqrt(dk))
//   O      = softmax_j(S) . V
//
// Flash-style: a CTA owns 64 query rows of one (row, head); keys stream through shared memory in tiles of 64 with an
// online softmax, so neither the [Td,Td] scores nor the reference's [Td,Td,dk] gathered table ever exist.  The
// relative term uses the "skew" identity: per (q-tile, k-tile) only the 127 table rows i-j in
// [q0-k0-63, q0-k0+63] are needed; R = Q . E_slice^T is one more small MMA whose result is read back along
// diagonals (c = r - jj + 63).  Matrix products use mma.sync m16n8k8 TF32 with operands rounded to nearest;
// this is <5% of the separator's FLOPs - the tcgen05 path is reserved for the token GEMMs.
#pragma once
#include "common.cuh"

namespace sepref {
namespace attn {

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int DK>
struct AttnSmem {
  static constexpr int LD = DK + 4;       // row stride (floats) of the Q/K/V/E tiles: conflict-free fragment loads
  static constexpr int RLD = 84;          // row stride of the per-warp R tile (80 columns used)
  uint32_t q[64 * LD];
  uint32_t k[64 * LD];
  uint32_t v[64 * LD];
  uint32_t e[128 * LD];
  float r[4][16 * RLD];
};

// qkv: [N, Td, 3F] (q | k | v), table: [2*maxlen, DK], out: [N, Td, F].  grid (ceil(Td/64), H, N), block 128.
template <int DK>
__global__ void __launch_bounds__(128) k_attn_relpos(const float* __restrict__ qkv, const float* __restrict__ table,
                                                     float* __restrict__ out, int Td, int F, int maxlen) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AttnSmem<DK>& sm = *reinterpret_cast<AttnSmem<DK>*>(smem_raw);
  constexpr int LD = AttnSmem<DK>::LD, RLD = AttnSmem<DK>::RLD;
  constexpr int KS = DK / 8;              // k-steps of the Q.K^T / Q.E^T products
  constexpr int ON = DK / 8;              // n-tiles of the output accumulator
  constexpr int V4 = DK / 4;              // float4 per tile row

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * 64, h = blockIdx.y, n = blockIdx.z;
  const float* base = qkv + (size_t)n * Td * 3 * F + h * DK;

  // ---- Q tile (rows beyond Td are zero)
  for (int idx = tid; idx < 64 * V4; idx += 128) {
    const int r = idx / V4, c = (idx % V4) * 4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < Td) x = __ldg(reinterpret_cast<const float4*>(base + (size_t)(q0 + r) * 3 * F + c));
    uint32_t* d = sm.q + r * LD + c;
    d[0] = f32_to_tf32_rna(x.x); d[1] = f32_to_tf32_rna(x.y); d[2] = f32_to_tf32_rna(x.z); d[3] = f32_to_tf32_rna(x.w);
  }
  __syncthreads();
  uint32_t qa[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const uint32_t* p = sm.q + (warp * 16) * LD + ks * 8;
    qa[ks][0] = p[g * LD + t];
    qa[ks][1] = p[(g + 8) * LD + t];
    qa[ks][2] = p[g * LD + t + 4];
    qa[ks][3] = p[(g + 8) * LD + t + 4];
  }

  float oacc[ON][4];
#pragma unroll
  for (int i = 0; i < ON; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) oacc[i][j] = 0.f;
  float row_max[2] = {-INFINITY, -INFINITY}, row_sum[2] = {0.f, 0.f};

  // K/V/table tiles are fetched one key tile ahead into registers, so the global-memory latency of tile k+1 is hidden
  // behind the MMAs and softmax of tile k; only the conversion + shared-memory store sits between the two barriers.
  constexpr int KN = 64 * V4 / 128;       // float4 per thread for the K (and V) tile
  constexpr int EN = 128 * V4 / 128;      // float4 per thread for the table slice
  float4 kreg[KN], vreg[KN], ereg[EN];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < KN; ++i) {
      const int idx = tid + 128 * i, r = idx / V4, c = (idx % V4) * 4;
      kreg[i] = make_float4(0.f, 0.f, 0.f, 0.f); vreg[i] = kreg[i];
      if (k0 + r < Td) {
        const float* p = base + (size_t)(k0 + r) * 3 * F + c;
        kreg[i] = __ldg(reinterpret_cast<const float4*>(p + F));
        vreg[i] = __ldg(reinterpret_cast<const float4*>(p + 2 * F));
      }
    }
    const int delta0 = q0 - k0 - 63;      // relative offset of table-slice row 0
#pragma unroll
    for (int i = 0; i < EN; ++i) {
      const int idx = tid + 128 * i, r = idx / V4, c = (idx % V4) * 4;
      int rel = delta0 + r;
      rel = max(-maxlen, min(maxlen - 1, rel)) + maxlen;
      ereg[i] = __ldg(reinterpret_cast<const float4*>(table + (size_t)rel * DK + c));
    }
  };
  auto put = [&](uint32_t* d, const float4& x) {
    d[0] = f32_to_tf32_rna(x.x); d[1] = f32_to_tf32_rna(x.y); d[2] = f32_to_tf32_rna(x.z); d[3] = f32_to_tf32_rna(x.w);
  };
  fetch(0);
  for (int k0 = 0; k0 < Td; k0 += 64) {
    __syncthreads();   // previous tile fully consumed
#pragma unroll
    for (int i = 0; i < KN; ++i) {
      const int idx = tid + 128 * i, r = idx / V4, c = (idx % V4) * 4;
      put(sm.k + r * LD + c, kreg[i]);
      put(sm.v + r * LD + c, vreg[i]);
    }
#pragma unroll
    for (int i = 0; i < EN; ++i) {
      const int idx = tid + 128 * i, r = idx / V4, c = (idx % V4) * 4;
      put(sm.e + r * LD + c, ereg[i]);
    }
    __syncthreads();
    if (k0 + 64 < Td) fetch(k0 + 64);

    // ---- R = Q_warp . E_slice^T : [16 x 80], slice rows 16*warp + c
    float* rw = sm.r[warp];
#pragma unroll
    for (int nt = 0; nt < 10; ++nt) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      // column 79 of warp 3 maps to slice row 127, which only ever feeds unused positions
      const uint32_t* p = sm.e + (warp * 16 + nt * 8 + g) * LD;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) mma_tf32(acc, qa[ks], p[ks * 8 + t], p[ks * 8 + t + 4]);
      rw[g * RLD + nt * 8 + 2 * t] = acc[0];
      rw[g * RLD + nt * 8 + 2 * t + 1] = acc[1];
      rw[(g + 8) * RLD + nt * 8 + 2 * t] = acc[2];
      rw[(g + 8) * RLD + nt * 8 + 2 * t + 1] = acc[3];
    }
    __syncwarp();

    // ---- S = Q . K^T + skewed R
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const uint32_t* p = sm.k + (nt * 8 + g) * LD;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) mma_tf32(s[nt], qa[ks], p[ks * 8 + t], p[ks * 8 + t + 4]);
      const int jj = nt * 8 + 2 * t;
      s[nt][0] += rw[g * RLD + (g - jj + 63)];
      s[nt][1] += rw[g * RLD + (g - jj + 62)];
      s[nt][2] += rw[(g + 8) * RLD + (g + 8 - jj + 63)];
      s[nt][3] += rw[(g + 8) * RLD + (g + 8 - jj + 62)];
      if (k0 + jj >= Td) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
      if (k0 + jj + 1 >= Td) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
    }
    __syncwarp();

    // ---- online softmax (rows g and g+8 of this warp's 16)
    float mx[2] = {row_max[0], row_max[1]};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      mx[0] = fmaxf(mx[0], __shfl_xor_sync(0xffffffffu, mx[0], o));
      mx[1] = fmaxf(mx[1], __shfl_xor_sync(0xffffffffu, mx[1], o));
    }
    const float corr0 = __expf(row_max[0] - mx[0]), corr1 = __expf(row_max[1] - mx[1]);   // first tile: exp(-inf) = 0
    row_max[0] = mx[0]; row_max[1] = mx[1];
    float ps[2] = {0.f, 0.f};
    uint32_t pa[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = __expf(s[nt][0] - mx[0]), p1 = __expf(s[nt][1] - mx[0]);
      const float p2 = __expf(s[nt][2] - mx[1]), p3 = __expf(s[nt][3] - mx[1]);
      ps[0] += p0 + p1; ps[1] += p2 + p3;
      // A fragment of P for k-step nt: slot t <- key 2t, slot t+4 <- key 2t+1 (V rows are read with the same permutation)
      pa[nt][0] = f32_to_tf32_rna(p0); pa[nt][1] = f32_to_tf32_rna(p2);
      pa[nt][2] = f32_to_tf32_rna(p1); pa[nt][3] = f32_to_tf32_rna(p3);
    }
    row_sum[0] = row_sum[0] * corr0 + ps[0];
    row_sum[1] = row_sum[1] * corr1 + ps[1];
#pragma unroll
    for (int i = 0; i < ON; ++i) { oacc[i][0] *= corr0; oacc[i][1] *= corr0; oacc[i][2] *= corr1; oacc[i][3] *= corr1; }

    // ---- O += P . V
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const uint32_t* p0 = sm.v + (nt * 8 + 2 * t) * LD;
#pragma unroll
      for (int i = 0; i < ON; ++i) mma_tf32(oacc[i], pa[nt], p0[i * 8 + g], p0[LD + i * 8 + g]);
    }
  }

  // ---- normalise and store: quad-reduce the row sums first
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    row_sum[0] += __shfl_xor_sync(0xffffffffu, row_sum[0], o);
    row_sum[1] += __shfl_xor_sync(0xffffffffu, row_sum[1], o);
  }
  const float inv0 = 1.0f / row_sum[0], inv1 = 1.0f / row_sum[1];
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  float* ob = out + (size_t)n * Td * F + h * DK;
#pragma unroll
  for (int i = 0; i < ON; ++i) {
    if (r0 < Td) *reinterpret_cast<float2*>(ob + (size_t)r0 * F + i * 8 + 2 * t) = make_float2(oacc[i][0] * inv0, oacc[i][1] * inv0);
    if (r1 < Td) *reinterpret_cast<float2*>(ob + (size_t)r1 * F + i * 8 + 2 * t) = make_float2(oacc[i][2] * inv1, oacc[i][3] * inv1);
  }
}

}  // namespace attn
}  // namespace sepref
