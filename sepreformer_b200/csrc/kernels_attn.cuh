// Pooled multi-head attention with Shaw-style relative-position key bias
// (reference network.py:103-122 as called from EGA, network.py:145-149; table from module.py:42-57,196-198).
//
//   S[i,j] = q_i . k_j + q_i . E[clamp(i-j, -maxlen, maxlen-1) + maxlen]      (q pre-scaled by log2(e)/sqrt(dk))
//   O      = softmax_j(S) . V,  evaluated as 2^(S - max S): one MUFU.EX2 per score
//
// Flash-style: a CTA owns 64 query rows of one (row, head); keys stream through shared memory in tiles of 64 with an
// online softmax, so neither the [Td,Td] scores nor the reference's [Td,Td,dk] gathered table ever exist.  The
// relative term uses the "skew" identity: per (q-tile, k-tile) only the 127 table rows i-j in
// [q0-k0-63, q0-k0+63] are needed; R = Q . E_slice^T is one more small MMA whose result is read back along
// diagonals (c = r - jj + 63).  Matrix products are mma.sync m16n8k16 with FP16 operands (round-to-nearest,
// saturating; the same 11-bit significand as TF32, half the shared-memory bytes and MMA count) and fp32
// accumulation; fragments come from ldmatrix (x4, .trans for V).  This is <5% of the separator's FLOPs - the tcgen05
// path is reserved for the token GEMMs.
#pragma once
#include "common.cuh"

namespace sepref {
namespace attn {

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// four 8x8 b16 matrices; lane l supplies the address of row (l & 7) of matrix (l >> 3)
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}

template <int DK>
struct AttnSmem {
  static constexpr int LD = DK + 8;       // row stride (halves) of the Q/K/V/E tiles: 16-byte skew, conflict-free ldmatrix
  static constexpr int RLD = 88;          // row stride of the per-warp R tile (80 columns used); = 8 mod 32 so the
                                          // float2 accumulator stores of a half-warp (4 rows x 8 words) are conflict-free
  uint16_t q[64 * LD];
  uint16_t k[64 * LD];
  uint16_t v[64 * LD];
  uint16_t e[128 * LD];
  float r[4][16 * RLD];
};

// qkv: [N, Td, 3F] (q | k | v; fp32, or FP16 rows when IN16 - what the kind::f16 projection kernel writes), table:
// [2*maxlen, DK] as FP16 (rounded once at pack time - the same round-to-nearest, saturating conversion this kernel used
// to apply to every slice it staged; half the bytes per key tile), out: [N, Td, F].  grid (ceil(Td/64), H, N), block 128.
template <int DK, bool IN16>
__global__ void __launch_bounds__(128, 5) k_attn_relpos(const void* __restrict__ qkv_, const uint16_t* __restrict__ table,
                                                     float* __restrict__ out, int Td, int F, int maxlen) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AttnSmem<DK>& sm = *reinterpret_cast<AttnSmem<DK>*>(smem_raw);
  constexpr int LD = AttnSmem<DK>::LD, RLD = AttnSmem<DK>::RLD;
  constexpr int KS = DK / 16;             // k-steps of the Q.K^T / Q.E^T products
  constexpr int ON = DK / 8;              // n-tiles of the output accumulator
  constexpr int V4 = DK / 4;              // float4 per tile row

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * 64, h = blockIdx.y, n = blockIdx.z;
  // four channels of a q|k|v row as packed halves (part 0/1/2 = q/k/v)
  const size_t base_el = (size_t)n * Td * 3 * F + h * DK;
  auto ld_qkv = [&](int row, int part, int c) -> uint2 {
    const size_t off = base_el + (size_t)row * 3 * F + part * F + c;
    if (IN16) return __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(qkv_) + off));
    const float4 x = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(qkv_) + off));
    return make_uint2(pack_f16x2_sat(x.x, x.y), pack_f16x2_sat(x.z, x.w));
  };
  // tile staging: float4 number idx -> (row, first channel).  For DK = 16 a half-warp's 8-byte stores cover four rows
  // two apart (48-byte row stride: rows r, r+2, r+4, r+6 start 96 bytes apart mod 128 -> all 32 banks, no conflict).
  auto rc = [&](int idx, int& r, int& c) {
    if (DK == 16) { r = 8 * (idx >> 5) + 2 * ((idx & 15) >> 2) + ((idx >> 4) & 1); c = (idx & 3) * 4; }
    else { r = idx / V4; c = (idx % V4) * 4; }
  };

  // ---- Q tile (rows beyond Td are zero)
  for (int idx = tid; idx < 64 * V4; idx += 128) {
    int r, c; rc(idx, r, c);
    uint2 x = make_uint2(0u, 0u);
    if (q0 + r < Td) x = ld_qkv(q0 + r, 0, c);
    *reinterpret_cast<uint2*>(sm.q + r * LD + c) = x;
  }
  __syncthreads();
  // ldmatrix lane roles: A operand (rows x k) and B operand stored [n][k] share one pattern, V ([k][n]) uses .trans
  const int l7 = lane & 7, lb3 = (lane >> 3) & 1, lb4 = lane >> 4;
  uint32_t qa[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) ldsm_x4(qa[ks], sm.q + (warp * 16 + l7 + lb3 * 8) * LD + ks * 16 + lb4 * 8);

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
  uint2 kreg[KN], vreg[KN];
  uint2 ereg[EN];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < KN; ++i) {
      const int idx = tid + 128 * i; int r, c; rc(idx, r, c);
      kreg[i] = make_uint2(0u, 0u); vreg[i] = kreg[i];
      if (k0 + r < Td) { kreg[i] = ld_qkv(k0 + r, 1, c); vreg[i] = ld_qkv(k0 + r, 2, c); }
    }
    const int delta0 = q0 - k0 - 63;      // relative offset of table-slice row 0
#pragma unroll
    for (int i = 0; i < EN; ++i) {
      const int idx = tid + 128 * i; int r, c; rc(idx, r, c);
      int rel = delta0 + r;
      rel = max(-maxlen, min(maxlen - 1, rel)) + maxlen;
      ereg[i] = __ldg(reinterpret_cast<const uint2*>(table + (size_t)rel * DK + c));
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < Td; k0 += 64) {
    __syncthreads();   // previous tile fully consumed
#pragma unroll
    for (int i = 0; i < KN; ++i) {
      const int idx = tid + 128 * i; int r, c; rc(idx, r, c);
      *reinterpret_cast<uint2*>(sm.k + r * LD + c) = kreg[i];
      *reinterpret_cast<uint2*>(sm.v + r * LD + c) = vreg[i];
    }
#pragma unroll
    for (int i = 0; i < EN; ++i) {
      const int idx = tid + 128 * i; int r, c; rc(idx, r, c);
      *reinterpret_cast<uint2*>(sm.e + r * LD + c) = ereg[i];
    }
    __syncthreads();
    if (k0 + 64 < Td) fetch(k0 + 64);

    // ---- R = Q_warp . E_slice^T : [16 x 80], slice rows 16*warp + c (column 79 of warp 3 = slice row 127, which
    //      only ever feeds unused positions)
    float* rw = sm.r[warp];
#pragma unroll
    for (int np = 0; np < 5; ++np) {
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t b[4];
        ldsm_x4(b, sm.e + (warp * 16 + np * 16 + l7 + lb4 * 8) * LD + ks * 16 + lb3 * 8);
        mma_f16(acc[0], qa[ks], b[0], b[1]);
        mma_f16(acc[1], qa[ks], b[2], b[3]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int col = (2 * np + u) * 8 + 2 * t;
        *reinterpret_cast<float2*>(rw + g * RLD + col) = make_float2(acc[u][0], acc[u][1]);
        *reinterpret_cast<float2*>(rw + (g + 8) * RLD + col) = make_float2(acc[u][2], acc[u][3]);
      }
    }
    __syncwarp();

    // ---- S = Q . K^T + skewed R
    float s[8][4];
#pragma unroll
    for (int np = 0; np < 4; ++np) {
#pragma unroll
      for (int u = 0; u < 2; ++u) s[2 * np + u][0] = s[2 * np + u][1] = s[2 * np + u][2] = s[2 * np + u][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t b[4];
        ldsm_x4(b, sm.k + (np * 16 + l7 + lb4 * 8) * LD + ks * 16 + lb3 * 8);
        mma_f16(s[2 * np], qa[ks], b[0], b[1]);
        mma_f16(s[2 * np + 1], qa[ks], b[2], b[3]);
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int jj = nt * 8 + 2 * t;
      s[nt][0] += rw[g * RLD + (g - jj + 63)];
      s[nt][1] += rw[g * RLD + (g - jj + 62)];
      s[nt][2] += rw[(g + 8) * RLD + (g + 8 - jj + 63)];
      s[nt][3] += rw[(g + 8) * RLD + (g + 8 - jj + 62)];
    }
    if (k0 + 64 > Td) {        // only the last key tile has columns past the sequence end (block-uniform branch)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int jj = nt * 8 + 2 * t;
        if (k0 + jj >= Td) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
        if (k0 + jj + 1 >= Td) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
      }
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
    const float corr0 = ex2_approx(row_max[0] - mx[0]), corr1 = ex2_approx(row_max[1] - mx[1]);   // first tile: 2^-inf = 0
    row_max[0] = mx[0]; row_max[1] = mx[1];
    float ps[2] = {0.f, 0.f};
    uint32_t pa[4][4];      // A fragments of P, one per 16-key step
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = ex2_approx(s[nt][0] - mx[0]), p1 = ex2_approx(s[nt][1] - mx[0]);
      const float p2 = ex2_approx(s[nt][2] - mx[1]), p3 = ex2_approx(s[nt][3] - mx[1]);
      ps[0] += p0 + p1; ps[1] += p2 + p3;
      pa[nt >> 1][(nt & 1) * 2] = pack_f16x2_sat(p0, p1);          // row g,   keys 8*nt + 2t, +1
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_f16x2_sat(p2, p3);      // row g+8
    }
    row_sum[0] = row_sum[0] * corr0 + ps[0];
    row_sum[1] = row_sum[1] * corr1 + ps[1];
#pragma unroll
    for (int i = 0; i < ON; ++i) { oacc[i][0] *= corr0; oacc[i][1] *= corr0; oacc[i][2] *= corr1; oacc[i][3] *= corr1; }

    // ---- O += P . V   (V tile is [key][d]: the B fragments are its transposed 8x8 blocks)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int ip = 0; ip < ON / 2; ++ip) {
        uint32_t b[4];
        ldsm_x4_trans(b, sm.v + (kk * 16 + l7 + lb3 * 8) * LD + (ip * 2 + lb4) * 8);
        mma_f16(oacc[2 * ip], pa[kk], b[0], b[1]);
        mma_f16(oacc[2 * ip + 1], pa[kk], b[2], b[3]);
      }
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
