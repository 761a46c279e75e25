// Exact-fp32 CUDA-core kernels of the separator path (GEMM_PATH 0) plus the stencil / layout / normalisation
// kernels that both paths share.  Activations are channels-last [rows, T, F] fp32.
#pragma once
#include "common.cuh"

namespace sepref {
namespace simt {

// ------------------------------------------------------------------------------------------------ layout
// [B, F, Tin] -> [B, Tpad, F] with zero right-padding (Separator.pad_signal, module.py:220-234, fused with the
// channels-first -> channels-last change).  grid (ceil(Tpad/32), F/32, B), block (32, 8).
__global__ void k_nct_to_ntc_pad(const float* __restrict__ in, float* __restrict__ out, int F, int Tin, int Tpad) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
  const float* src = in + (size_t)b * F * Tin;
  float* dst = out + (size_t)b * Tpad * F;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int f = f0 + threadIdx.y + 8 * k, t = t0 + threadIdx.x;
    tile[threadIdx.y + 8 * k][threadIdx.x] = (t < Tin) ? src[(size_t)f * Tin + t] : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int t = t0 + threadIdx.y + 8 * k, f = f0 + threadIdx.x;
    if (t < Tpad) dst[(size_t)t * F + f] = tile[threadIdx.x][threadIdx.y + 8 * k];
  }
}

// [N, T, F] -> [N, F, T].  grid (ceil(T/32), F/32, N), block (32, 8).
__global__ void k_ntc_to_nct(const float* __restrict__ in, float* __restrict__ out, int F, int T) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, t0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
  const float* src = in + (size_t)n * T * F;
  float* dst = out + (size_t)n * F * T;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int t = t0 + threadIdx.y + 8 * k, f = f0 + threadIdx.x;
    tile[threadIdx.y + 8 * k][threadIdx.x] = (t < T) ? src[(size_t)t * F + f] : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int f = f0 + threadIdx.y + 8 * k, t = t0 + threadIdx.x;
    if (t < T) dst[(size_t)f * T + t] = tile[threadIdx.x][threadIdx.y + 8 * k];
  }
}

// ------------------------------------------------------------------------------------------------ layer norm
// out[o, :] = normalise(mean_{j<r} x[o*r + j, :]) - no affine (gamma/beta are folded into the next linear map).
// One warp per output row; r = 1 is a plain LayerNorm, r > 1 fuses EGA's adaptive_avg_pool1d (network.py:146).
template <int F>
__global__ void __launch_bounds__(256) k_pool_layernorm(const float* __restrict__ x, float* __restrict__ out,
                                                        int out_rows, int r) {
  constexpr int V = F / 128;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= out_rows) return;
  float4 v[V];
#pragma unroll
  for (int i = 0; i < V; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* src = reinterpret_cast<const float4*>(x + (size_t)row * r * F);
  for (int j = 0; j < r; ++j) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float4 a = __ldg(src + (size_t)j * (F / 4) + lane + 32 * i);
      v[i].x += a.x; v[i].y += a.y; v[i].z += a.z; v[i].w += a.w;
    }
  }
  const float inv_r = 1.0f / (float)r;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i].x *= inv_r; v[i].y *= inv_r; v[i].z *= inv_r; v[i].w *= inv_r;
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) * (1.0f / F);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / F) + kLnEps);
  float4* dst = reinterpret_cast<float4*>(out + (size_t)row * F);
#pragma unroll
  for (int i = 0; i < V; ++i)
    dst[lane + 32 * i] = make_float4(v[i].x * rstd, v[i].y * rstd, v[i].z * rstd, v[i].w * rstd);
}

// ------------------------------------------------------------------------------------------------ fp32 GEMM
enum Epi { EPI_BIAS = 0, EPI_GELU = 1, EPI_RES = 2, EPI_GATE = 3 };

struct GemmArgs {
  const float* A; int lda;       // [M, K] activations
  const float* W;                // [N, K] weights (row-major, as torch.nn.Linear stores them)
  const float* bias;             // [N]
  float* C; int ldc;             // [M, N]
  int M, N, K;
  const float* res; int ldres;   // EPI_RES / EPI_GATE: residual rows [M, N]
  const float* up; int up_div;   // EPI_GATE: pooled attention output rows [M / up_div, N] (nearest upsample)
};

// C = epi(A . W^T + bias).  128x128 tile, 16-deep k slices, 256 threads x (8x8) accumulators.  N % 128 == 0, K % 16 == 0.
template <int EPI>
__global__ void __launch_bounds__(256) k_gemm_f32(GemmArgs a) {
  __shared__ __align__(16) float As[16][128 + 4];
  __shared__ __align__(16) float Ws[16][128 + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * 128, n0 = blockIdx.x * 128;
  const int lr = tid >> 2, lc = (tid & 3) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < a.K; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = lr + 64 * i;
      const int m = m0 + r;
      float4 v = (m < a.M) ? __ldg(reinterpret_cast<const float4*>(a.A + (size_t)m * a.lda + k0 + lc))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      As[lc + 0][r] = v.x; As[lc + 1][r] = v.y; As[lc + 2][r] = v.z; As[lc + 3][r] = v.w;
      float4 w = __ldg(reinterpret_cast<const float4*>(a.W + (size_t)(n0 + r) * a.K + k0 + lc));
      Ws[lc + 0][r] = w.x; Ws[lc + 1][r] = w.y; Ws[lc + 2][r] = w.z; Ws[lc + 3][r] = w.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float av[8], bv[8];
      *reinterpret_cast<float4*>(av) = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      *reinterpret_cast<float4*>(av + 4) = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      *reinterpret_cast<float4*>(bv) = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      *reinterpret_cast<float4*>(bv + 4) = *reinterpret_cast<const float4*>(&Ws[k][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= a.M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + h * 64 + tx * 4;
      const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + n));
      float4 v = make_float4(acc[i][h * 4 + 0] + b.x, acc[i][h * 4 + 1] + b.y, acc[i][h * 4 + 2] + b.z,
                             acc[i][h * 4 + 3] + b.w);
      if (EPI == EPI_GELU) {
        v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
      } else if (EPI == EPI_RES) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(a.res + (size_t)m * a.ldres + n));
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      } else if (EPI == EPI_GATE) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(a.res + (size_t)m * a.ldres + n));
        const float4 u = __ldg(reinterpret_cast<const float4*>(a.up + (size_t)(m / a.up_div) * a.N + n));
        v.x = r.x + sigmoid_acc(v.x) * u.x; v.y = r.y + sigmoid_acc(v.y) * u.y;
        v.z = r.z + sigmoid_acc(v.z) * u.z; v.w = r.w + sigmoid_acc(v.w) * u.w;
      }
      *reinterpret_cast<float4*>(a.C + (size_t)m * a.ldc + n) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ gated stencils
// GCFN middle (network.py:62-65): u[t, c] = d[t, c] * sigmoid(d[t, C + c]),  d = depthwise k=3 (zero pad 1) of h [N, T, 2C].
// w is tap-major [3][2C].  One thread per (token, 4 channels).
__global__ void __launch_bounds__(256) k_dw3_glu(const float* __restrict__ h, const float* __restrict__ w,
                                                 const float* __restrict__ wb, float* __restrict__ u, int rows, int T,
                                                 int C) {
  const int c4 = C / 4;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)rows * c4) return;
  const int row = (int)(idx / c4), c = (int)(idx % c4) * 4;
  const int t = row % T;
  float4 acc[2];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int ch = c + g * C;
    float4 s = __ldg(reinterpret_cast<const float4*>(wb + ch));
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int tt = t + j - 1;
      if (tt < 0 || tt >= T) continue;
      const float4 x = __ldg(reinterpret_cast<const float4*>(h + (size_t)(row + j - 1) * 2 * C + ch));
      const float4 k = __ldg(reinterpret_cast<const float4*>(w + (size_t)j * 2 * C + ch));
      s.x = fmaf(k.x, x.x, s.x); s.y = fmaf(k.y, x.y, s.y); s.z = fmaf(k.z, x.z, s.z); s.w = fmaf(k.w, x.w, s.w);
    }
    acc[g] = s;
  }
  float4 o = make_float4(acc[0].x * sigmoid_acc(acc[1].x), acc[0].y * sigmoid_acc(acc[1].y),
                         acc[0].z * sigmoid_acc(acc[1].z), acc[0].w * sigmoid_acc(acc[1].w));
  *reinterpret_cast<float4*>(u + (size_t)row * C + c) = o;
}

// GLU over the last dim: out[r, c] = h[r, c] * sigmoid(h[r, C + c]).
__global__ void __launch_bounds__(256) k_glu(const float* __restrict__ h, float* __restrict__ out, size_t rows, int C) {
  const int c4 = C / 4;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * c4) return;
  const size_t row = idx / c4;
  const int c = (int)(idx % c4) * 4;
  const float4 a = __ldg(reinterpret_cast<const float4*>(h + row * 2 * C + c));
  const float4 g = __ldg(reinterpret_cast<const float4*>(h + row * 2 * C + C + c));
  *reinterpret_cast<float4*>(out + row * C + c) =
      make_float4(a.x * sigmoid_acc(g.x), a.y * sigmoid_acc(g.y), a.z * sigmoid_acc(g.z), a.w * sigmoid_acc(g.w));
}

// Depthwise 'same' convolution along time, odd K (CLA's k=65, network.py:166,180).  w is tap-major [K][C].
// One thread = one channel x TB consecutive frames; grid (ceil(T/TB), N, C/128), block 128.
template <int K, int TB>
__global__ void __launch_bounds__(128) k_dwconv_same(const float* __restrict__ u, const float* __restrict__ w,
                                                     const float* __restrict__ wb, float* __restrict__ out, int T,
                                                     int C) {
  constexpr int P = (K - 1) / 2;
  const int c = blockIdx.z * 128 + threadIdx.x;
  const int n = blockIdx.y, t0 = blockIdx.x * TB;
  const float* src = u + (size_t)n * T * C + c;
  float wk[K];
#pragma unroll
  for (int j = 0; j < K; ++j) wk[j] = __ldg(w + (size_t)j * C + c);
  float acc[TB];
  const float b = __ldg(wb + c);
#pragma unroll
  for (int o = 0; o < TB; ++o) acc[o] = b;
#pragma unroll
  for (int s = 0; s < TB + K - 1; ++s) {
    const int t = t0 + s - P;
    const float v = (t >= 0 && t < T) ? __ldg(src + (size_t)t * C) : 0.f;
#pragma unroll
    for (int o = 0; o < TB; ++o) {
      const int j = s - o;
      if (j >= 0 && j < K) acc[o] = fmaf(wk[j], v, acc[o]);
    }
  }
  float* dst = out + (size_t)n * T * C + c;
#pragma unroll
  for (int o = 0; o < TB; ++o)
    if (t0 + o < T) dst[(size_t)(t0 + o) * C] = acc[o];
}

// Shared-memory tiled variant of the k=65 depthwise convolution (the one that runs in the forward pass): a block
// stages frames [t0-32, t0+TB+32) x C once (coalesced float4), then thread = channel slides its 65-tap window along
// time in registers (80 conflict-free LDS.64 feed 16 outputs x 65 packed FMAs), so the kernel runs at the FP32 FMA rate
// instead of re-reading every input 5x through L1.  grid (ceil(T/TB), N, C/CB), block CB threads (CB channels
// per block keep the tile small enough for several resident blocks), smem (TB+64)*CB*4 bytes.
template <int C, int TB, int CB>
__global__ void __launch_bounds__(CB) k_dwconv65_tiled(const float* __restrict__ u, const float* __restrict__ w,
                                                        const float* __restrict__ wb, float* __restrict__ out, int T) {
  constexpr int K = 65, P = 32, ROWS = TB + K - 1;
  extern __shared__ __align__(16) float tile[];           // [ROWS][CB]
  const int n = blockIdx.y, t0 = blockIdx.x * TB, cb = blockIdx.z * CB;
  const float* src = u + (size_t)n * T * C + cb;
  for (int idx = threadIdx.x; idx < ROWS * (CB / 4); idx += CB) {
    const int r = idx / (CB / 4), c4 = idx % (CB / 4);
    const int t = t0 - P + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= 0 && t < T) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)t * C) + c4);
    reinterpret_cast<float4*>(tile + (size_t)r * CB)[c4] = v;
  }
  // thread = two adjacent channels x half of the block's frames: every multiply-add is one packed fma.rn.f32x2
  // (half the instructions; the FMA rate itself is the same as the scalar form's, profiles/r1_pipe_rates.md)
  const int cl = 2 * (threadIdx.x % (CB / 2)), part = threadIdx.x / (CB / 2), cp = cb + cl;
  float2 wk[K];
#pragma unroll
  for (int j = 0; j < K; ++j) wk[j] = __ldg(reinterpret_cast<const float2*>(w + (size_t)j * C + cp));
  const float2 b = __ldg(reinterpret_cast<const float2*>(wb + cp));
  __syncthreads();
  float* dst = out + ((size_t)n * T + t0) * C + cp;
  constexpr int OB = 16, HALF = TB / 2;                    // outputs per register block: 80 LDS.64 feed 16 x 65 FFMA2
  static_assert(HALF % OB == 0, "dwconv65 tiling");
#pragma unroll 1
  for (int o0 = part * HALF; o0 < (part + 1) * HALF; o0 += OB) {
    float2 acc[OB];
#pragma unroll
    for (int o = 0; o < OB; ++o) acc[o] = b;
#pragma unroll
    for (int s2 = 0; s2 < OB + K - 1; ++s2) {
      const float2 v = *reinterpret_cast<const float2*>(tile + (size_t)(o0 + s2) * CB + cl);
#pragma unroll
      for (int o = 0; o < OB; ++o) {
        const int j = s2 - o;
        if (j >= 0 && j < K) acc[o] = __ffma2_rn(wk[j], v, acc[o]);
      }
    }
#pragma unroll
    for (int o = 0; o < OB; ++o)
      if (t0 + o0 + o < T) *reinterpret_cast<float2*>(dst + (size_t)(o0 + o) * C) = acc[o];
  }
}

// Higher-occupancy form of the same kernel: thread = ONE channel x 1/PARTS of the block's frames (65 tap registers
// instead of 130), CB * PARTS threads per block, four blocks per SM -> 16 warps per SM instead of 8.  The FP32 pipe
// issues ~120 FMA/clk/SM for this instruction form (profiles/r1_pipe_rates.md); the packed variant above reached 61.
template <int C, int TB, int CB, int PARTS>
__global__ void __launch_bounds__(CB * PARTS, 4) k_dwconv65_occ(const float* __restrict__ u, const float* __restrict__ w,
                                                                const float* __restrict__ wb, float* __restrict__ out, int T) {
  constexpr int K = 65, P = 32, ROWS = TB + K - 1, NT = CB * PARTS;
  extern __shared__ __align__(16) float tile[];           // [ROWS][CB]
  const int n = blockIdx.y, t0 = blockIdx.x * TB, cb = blockIdx.z * CB;
  const float* src = u + (size_t)n * T * C + cb;
  for (int idx = threadIdx.x; idx < ROWS * (CB / 4); idx += NT) {
    const int r = idx / (CB / 4), c4 = idx % (CB / 4);
    const int t = t0 - P + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= 0 && t < T) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)t * C) + c4);
    reinterpret_cast<float4*>(tile + (size_t)r * CB)[c4] = v;
  }
  const int cl = threadIdx.x % CB, part = threadIdx.x / CB, cp = cb + cl;
  float wk[K];
#pragma unroll
  for (int j = 0; j < K; ++j) wk[j] = __ldg(w + (size_t)j * C + cp);
  const float b = __ldg(wb + cp);
  __syncthreads();
  float* dst = out + ((size_t)n * T + t0) * C + cp;
  constexpr int OB = 16, SPAN = TB / PARTS;               // outputs per register block: 80 LDS feed 16 x 65 FMAs
  static_assert(SPAN % OB == 0, "dwconv65 tiling");
#pragma unroll 1
  for (int o0 = part * SPAN; o0 < (part + 1) * SPAN; o0 += OB) {
    float acc[OB];
#pragma unroll
    for (int o = 0; o < OB; ++o) acc[o] = b;
#pragma unroll
    for (int s2 = 0; s2 < OB + K - 1; ++s2) {
      const float v = tile[(size_t)(o0 + s2) * CB + cl];
#pragma unroll
      for (int o = 0; o < OB; ++o) {
        const int j = s2 - o;
        if (j >= 0 && j < K) acc[o] = fmaf(wk[j], v, acc[o]);
      }
    }
#pragma unroll
    for (int o = 0; o < OB; ++o)
      if (t0 + o0 + o < T) dst[(size_t)(o0 + o) * C] = acc[o];
  }
}

// DownConvLayer (module.py:72-78): y[t] = GELU(b' + sum_j w'[j] x[2t + j - P]) with BatchNorm folded into w', b'.
// w tap-major [K][C]; one thread per (out frame, 4 channels).
__global__ void __launch_bounds__(256) k_downconv_gelu(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ wb, float* __restrict__ y, int N, int T,
                                                       int C, int K) {
  const int c4 = C / 4, To = T / 2, P = (K - 1) / 2;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * To * c4) return;
  const int c = (int)(idx % c4) * 4;
  const size_t orow = idx / c4;
  const int n = (int)(orow / To), to = (int)(orow % To);
  float4 s = __ldg(reinterpret_cast<const float4*>(wb + c));
  for (int j = 0; j < K; ++j) {
    const int t = 2 * to + j - P;
    if (t < 0 || t >= T) continue;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + ((size_t)n * T + t) * C + c));
    const float4 k = __ldg(reinterpret_cast<const float4*>(w + (size_t)j * C + c));
    s.x = fmaf(k.x, v.x, s.x); s.y = fmaf(k.y, v.y, s.y); s.z = fmaf(k.z, v.z, s.z); s.w = fmaf(k.w, v.w, s.w);
  }
  *reinterpret_cast<float4*>(y + orow * C + c) = make_float4(gelu_erf(s.x), gelu_erf(s.y), gelu_erf(s.z), gelu_erf(s.w));
}

// [up2(low) || skip] rows for the fusion 1x1 conv (module.py:212-213): out [rows, 2F].
__global__ void __launch_bounds__(256) k_concat_up(const float* __restrict__ low, const float* __restrict__ skip,
                                                   float* __restrict__ out, int N, int T, int F) {
  const int f4 = F / 4;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * T * 2 * f4) return;
  const int c = (int)(idx % (2 * f4));
  const size_t row = idx / (2 * f4);
  const int n = (int)(row / T), t = (int)(row % T);
  float4 v;
  if (c < f4) v = __ldg(reinterpret_cast<const float4*>(low + ((size_t)n * (T / 2) + t / 2) * F) + c);
  else v = __ldg(reinterpret_cast<const float4*>(skip + row * F) + (c - f4));
  reinterpret_cast<float4*>(out + row * 2 * F)[c] = v;
}

// ------------------------------------------------------------------------------------------------ GroupNorm(1 group)
// SpkSplitStage tail (module.py:122-124).  h2 is [B, T, S*F]; speaker s of utterance b owns channels [sF, (s+1)F).
// stats[(b*S+s)*2 + {0,1}] accumulate sum / sum of squares in double.  grid (chunks, B*S), block 256.
__global__ void __launch_bounds__(256) k_gn_stats(const float* __restrict__ h2, double* __restrict__ stats, int T, int F,
                                                  int S, int rows_per_block) {
  const int bs = blockIdx.y, b = bs / S, s = bs % S;
  const int t0 = blockIdx.x * rows_per_block;
  const int t1 = min(T, t0 + rows_per_block);
  const int f4 = F / 4;
  const int count = (t1 - t0) * f4;
  double sum = 0.0, sq = 0.0;
  for (int e = threadIdx.x; e < count; e += blockDim.x) {
    const int t = t0 + e / f4, c = (e % f4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(h2 + ((size_t)b * T + t) * S * F + s * F + c));
    sum += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
    sq += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  __shared__ double sh[2][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[0][warp] = sum; sh[1][warp] = sq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { ts += sh[0][w]; tq += sh[1][w]; }
    atomicAdd(stats + (size_t)bs * 2 + 0, ts);
    atomicAdd(stats + (size_t)bs * 2 + 1, tq);
  }
}

// out[(b*S+s), t, c] = (h2[b, t, sF+c] - mu) * rstd * gamma[c] + beta[c]
__global__ void __launch_bounds__(256) k_gn_apply_split(const float* __restrict__ h2, const double* __restrict__ stats,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float* __restrict__ out, int B, int T, int F, int S) {
  const int f4 = F / 4;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * S * T * f4) return;
  const int c = (int)(idx % f4) * 4;
  const size_t orow = idx / f4;                 // (b*S + s)*T + t
  const int t = (int)(orow % T);
  const int bs = (int)(orow / T);
  const int b = bs / S, s = bs % S;
  const double cnt = (double)T * F;
  const double mu = stats[bs * 2] / cnt;
  const double var = fmax(stats[bs * 2 + 1] / cnt - mu * mu, 0.0);
  const float rstd = (float)(1.0 / sqrt(var + (double)kGnEps));
  const float m = (float)mu;
  const float4 v = __ldg(reinterpret_cast<const float4*>(h2 + ((size_t)b * T + t) * S * F + s * F + c));
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
  const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
  *reinterpret_cast<float4*>(out + orow * F + c) =
      make_float4((v.x - m) * rstd * g.x + be.x, (v.y - m) * rstd * g.y + be.y, (v.z - m) * rstd * g.z + be.z,
                  (v.w - m) * rstd * g.w + be.w);
}

// ------------------------------------------------------------------------------------------------ speaker attention
// SpkAttention's 2-token attention (network.py:240-246 with MultiHeadAttention 106-122, pos_k=None), S = 2.
// qkv is [2B, T, 3F] (row 2b+s), q already scaled by log2(e)/sqrt(dk) (softmax in base 2).  One thread per (b, t, head).
template <int DK>
__global__ void __launch_bounds__(256) k_spk_attn2(const float* __restrict__ qkv, float* __restrict__ o, int B, int T,
                                                   int F) {
  const int H = F / DK;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * T * H) return;
  const int h = (int)(idx % H);
  const size_t bt = idx / H;
  const int b = (int)(bt / T), t = (int)(bt % T);
  const float* r0 = qkv + (((size_t)(2 * b) * T + t) * 3 * F) + h * DK;
  const float* r1 = qkv + (((size_t)(2 * b + 1) * T + t) * 3 * F) + h * DK;
  float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
#pragma unroll
  for (int d = 0; d < DK; d += 4) {
    const float4 q0 = __ldg(reinterpret_cast<const float4*>(r0 + d));
    const float4 q1 = __ldg(reinterpret_cast<const float4*>(r1 + d));
    const float4 k0 = __ldg(reinterpret_cast<const float4*>(r0 + F + d));
    const float4 k1 = __ldg(reinterpret_cast<const float4*>(r1 + F + d));
    s00 += q0.x * k0.x + q0.y * k0.y + q0.z * k0.z + q0.w * k0.w;
    s01 += q0.x * k1.x + q0.y * k1.y + q0.z * k1.z + q0.w * k1.w;
    s10 += q1.x * k0.x + q1.y * k0.y + q1.z * k0.z + q1.w * k0.w;
    s11 += q1.x * k1.x + q1.y * k1.y + q1.z * k1.z + q1.w * k1.w;
  }
  const float m0 = fmaxf(s00, s01), m1 = fmaxf(s10, s11);
  float p00 = exp2f(s00 - m0), p01 = exp2f(s01 - m0), p10 = exp2f(s10 - m1), p11 = exp2f(s11 - m1);   // q carries log2(e) (pack_mha)
  const float i0 = 1.0f / (p00 + p01), i1 = 1.0f / (p10 + p11);
  p00 *= i0; p01 *= i0; p10 *= i1; p11 *= i1;
  float* o0 = o + ((size_t)(2 * b) * T + t) * F + h * DK;
  float* o1 = o + ((size_t)(2 * b + 1) * T + t) * F + h * DK;
#pragma unroll
  for (int d = 0; d < DK; d += 4) {
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(r0 + 2 * F + d));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(r1 + 2 * F + d));
    *reinterpret_cast<float4*>(o0 + d) = make_float4(p00 * v0.x + p01 * v1.x, p00 * v0.y + p01 * v1.y,
                                                     p00 * v0.z + p01 * v1.z, p00 * v0.w + p01 * v1.w);
    *reinterpret_cast<float4*>(o1 + d) = make_float4(p10 * v0.x + p11 * v1.x, p10 * v0.y + p11 * v1.y,
                                                     p10 * v0.z + p11 * v1.z, p10 * v0.w + p11 * v1.w);
  }
}

}  // namespace simt
}  // namespace sepref
