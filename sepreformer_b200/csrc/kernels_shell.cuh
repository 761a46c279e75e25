// Kernels for the layers either side of the separator, so that a waveform enters the GPU and
// waveforms (or the metric) leave it - the 393 MB of feature traffic per B=32 step of the separator-only boundary
// shrinks to 4 MB in and 8 MB out:
//
//   n2  AudioEncoder + FeatureProjector  (reference modules/module.py:12-35)
//         e = GELU(Conv1d(1 -> 256, k = 16, stride 4, no bias)(mix))        -> k_enc_stats / k_enc_norm (recomputed, never stored)
//         z = GroupNorm(1 group, eps 1e-8)(e) over (256, T) per utterance   -> two passes: statistics, then normalise
//         x = Conv1d(256 -> F, k = 1, no bias)(z)                           -> tcgen05 token GEMM (k_tok), output channels-last
//                                                                              and already zero-padded: the separator's layout
//   n1  OutputLayer (module.py:237-265, masking = False): Linear(F -> 4F), GLU, Linear(2F -> 256)  -> two-stage k_tok
//   n3  AudioDecoder (module.py:268-283): ConvTranspose1d(256 -> 1, k = 16, stride 4, no bias).  Being linear, it is folded
//       into the output layer's second matrix at pack time (W' = w_dec^T . W2: [16, 2F]), so the GEMM emits the 16-sample
//       frame of every encoder step directly; k_overlap_add sums the four frames that cover a sample.
//   n4  PIT SI-SNRi (utils/implements/criterions.py:221-260), batched, one block per utterance -> k_pit_sisnri
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace sepref {
namespace shell {

constexpr int kEncC = 256, kEncK = 16, kEncS = 4;

// e[c] for one frame from its 16 samples; w tap-major [16][256]
__device__ __forceinline__ float enc_value(const float (&win)[kEncK], const float* __restrict__ w, int c) {
  float a = 0.f;
#pragma unroll
  for (int j = 0; j < kEncK; ++j) a = fmaf(__ldg(w + j * kEncC + c), win[j], a);
  return gelu_erf(a);
}

// Pass 1: per-utterance sum and sum of squares of e over (256 channels, T frames).  grid (ceil(T / FR), B), block 256
// (thread = channel).  stats[b*2 + {0,1}] in double.
template <int FR>
__global__ void __launch_bounds__(kEncC) k_enc_stats(const float* __restrict__ mix, const float* __restrict__ w,
                                                     double* __restrict__ stats, int n, int T) {
  __shared__ float s[FR * kEncS + kEncK];
  const int b = blockIdx.y, t0 = blockIdx.x * FR, c = threadIdx.x;
  const int nt = min(FR, T - t0);
  const float* src = mix + (size_t)b * n + (size_t)t0 * kEncS;
  const int ns = (nt - 1) * kEncS + kEncK;
  for (int i = threadIdx.x; i < ns; i += kEncC) s[i] = __ldg(src + i);
  __syncthreads();
  float wr[kEncK];
#pragma unroll
  for (int j = 0; j < kEncK; ++j) wr[j] = __ldg(w + j * kEncC + c);
  float sum = 0.f, sq = 0.f;
  for (int t = 0; t < nt; ++t) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < kEncK; ++j) a = fmaf(wr[j], s[t * kEncS + j], a);
    const float e = gelu_erf(a);
    sum += e; sq = fmaf(e, e, sq);
  }
  double ds = sum, dq = sq;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { ds += __shfl_xor_sync(0xffffffffu, ds, o); dq += __shfl_xor_sync(0xffffffffu, dq, o); }
  __shared__ double sh[2][8];
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = ds; sh[1][threadIdx.x >> 5] = dq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, q = 0.0;
    for (int i = 0; i < 8; ++i) { a += sh[0][i]; q += sh[1][i]; }
    atomicAdd(stats + 2 * b, a);
    atomicAdd(stats + 2 * b + 1, q);
  }
}

// Pass 2: z[b, t, c] = (e - mu_b) * rstd_b * gamma[c] + beta[c] for t < T, zero rows for T <= t < Tp (the separator's
// right padding, module.py:220-234: zero FEATURES, i.e. zero projector input because the projector has no bias).
template <int FR>
__global__ void __launch_bounds__(kEncC) k_enc_norm(const float* __restrict__ mix, const float* __restrict__ w,
                                                    const double* __restrict__ stats, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, float* __restrict__ z, int n, int T, int Tp) {
  __shared__ float s[FR * kEncS + kEncK];
  const int b = blockIdx.y, t0 = blockIdx.x * FR, c = threadIdx.x;
  const int nt = max(0, min(FR, T - t0)), ntp = min(FR, Tp - t0);
  if (nt > 0) {
    const float* src = mix + (size_t)b * n + (size_t)t0 * kEncS;
    const int ns = (nt - 1) * kEncS + kEncK;
    for (int i = threadIdx.x; i < ns; i += kEncC) s[i] = __ldg(src + i);
  }
  __syncthreads();
  const double cnt = (double)T * kEncC;
  const double mu = stats[2 * b] / cnt;
  const double var = fmax(stats[2 * b + 1] / cnt - mu * mu, 0.0);
  const float mean = (float)mu, rstd = (float)(1.0 / sqrt(var + 1e-8));
  const float g = __ldg(gamma + c) * rstd, be = __ldg(beta + c);
  float wr[kEncK];
#pragma unroll
  for (int j = 0; j < kEncK; ++j) wr[j] = __ldg(w + j * kEncC + c);
  float* dst = z + ((size_t)b * Tp + t0) * kEncC + c;
  for (int t = 0; t < ntp; ++t) {
    float v = 0.f;
    if (t < nt) {
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < kEncK; ++j) a = fmaf(wr[j], s[t * kEncS + j], a);
      v = fmaf(gelu_erf(a) - mean, g, be);
    }
    dst[(size_t)t * kEncC] = v;
  }
}

// audio[s][b][m] = sum over the (up to four) frames t with 0 <= m - 4t < 16 of frames[(b*S + s)*Tp + t][m - 4t];
// frames rows are `ld` floats apart (the GEMM's padded row), n_out = (T - 1) * 4 + 16.
__global__ void __launch_bounds__(256) k_overlap_add(const float* __restrict__ frames, float* __restrict__ audio, int B, int S,
                                                     int T, int Tp, int ld, int n_out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * S * n_out) return;
  const int m = (int)(idx % n_out);
  const int bs = (int)(idx / n_out), b = bs / S, s = bs % S;
  const float* fr = frames + (size_t)bs * Tp * ld;
  float a = 0.f;
  const int tq = m / kEncS;
#pragma unroll
  for (int k = 0; k < kEncK / kEncS; ++k) {
    const int t = tq - k, j = m - kEncS * t;
    if (t >= 0 && t < T && j < kEncK) a += __ldg(fr + (size_t)t * ld + j);
  }
  audio[((size_t)s * B + b) * n_out + m] = a;
}

// ---- n4: batched PIT SI-SNR improvement, two speakers (criterions.py:232-260; eps as passed by engine.py:131) --------
// est: [S, B, ld_est] (the model's output), tgt: [S, B, n], mix: [B, n]; out[b*3 + {0,1,2}] = best-permutation sum over
// speakers of (SI-SNR(est, tgt) - SI-SNR(mix, tgt)) in dB, and the two per-speaker terms of that permutation.
// One block per utterance.  Everything the metric needs is a second-order statistic of zero-mean signals, so one pass
// accumulates the 5 sums, 5 sums of squares and 6 cross products in double; the projections follow in closed form:
//   |proj| = |<a,t>| |t| / (|t|^2 + eps),   |a - proj|^2 = |a|^2 - 2 k <a,t> + k^2 |t|^2,  k = <a,t> / (|t|^2 + eps).
__global__ void __launch_bounds__(256) k_pit_sisnri(const float* __restrict__ est, const float* __restrict__ tgt,
                                                    const float* __restrict__ mix, float* __restrict__ out, int B, int n,
                                                    int ld_est, double eps) {
  const int b = blockIdx.x;
  const float* e0 = est + (size_t)b * ld_est;
  const float* e1 = est + ((size_t)B + b) * ld_est;
  const float* t0 = tgt + (size_t)b * n;
  const float* t1 = tgt + ((size_t)B + b) * n;
  const float* mx = mix + (size_t)b * n;
  // signals: 0 e0, 1 e1, 2 mix, 3 t0, 4 t1
  double s[5] = {0, 0, 0, 0, 0}, q[5] = {0, 0, 0, 0, 0}, x[6] = {0, 0, 0, 0, 0, 0};   // x: e0t0 e0t1 e1t0 e1t1 mt0 mt1
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v[5] = {(double)__ldg(e0 + i), (double)__ldg(e1 + i), (double)__ldg(mx + i), (double)__ldg(t0 + i), (double)__ldg(t1 + i)};
#pragma unroll
    for (int k = 0; k < 5; ++k) { s[k] += v[k]; q[k] += v[k] * v[k]; }
    x[0] += v[0] * v[3]; x[1] += v[0] * v[4]; x[2] += v[1] * v[3]; x[3] += v[1] * v[4]; x[4] += v[2] * v[3]; x[5] += v[2] * v[4];
  }
  __shared__ double sh[16][8];
  double all[16];
#pragma unroll
  for (int k = 0; k < 5; ++k) { all[k] = s[k]; all[5 + k] = q[k]; }
#pragma unroll
  for (int k = 0; k < 6; ++k) all[10 + k] = x[k];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) all[k] += __shfl_xor_sync(0xffffffffu, all[k], o);
    if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = all[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double v[16];
    for (int k = 0; k < 16; ++k) { v[k] = 0.0; for (int w = 0; w < 8; ++w) v[k] += sh[k][w]; }
    const double N = (double)n;
    auto var = [&](int a) { return v[5 + a] - v[a] * v[a] / N; };                       // |a_zm|^2
    auto cov = [&](int a, int t, int xi) { return v[10 + xi] - v[a] * v[t] / N; };    // <a_zm, t_zm>
    auto sisnr = [&](int a, int t, int xi) {
      const double tt = fmax(var(t), 0.0), aa = fmax(var(a), 0.0), at = cov(a, t, xi);
      const double k = at / (tt + eps);
      const double proj = fabs(k) * sqrt(tt);
      const double res = sqrt(fmax(aa - 2.0 * k * at + k * k * tt, 0.0));
      return 20.0 * log10(eps + proj / (res + eps));
    };
    const double m0 = sisnr(2, 3, 4), m1 = sisnr(2, 4, 5);
    const double a00 = sisnr(0, 3, 0) - m0, a01 = sisnr(0, 4, 1) - m1, a10 = sisnr(1, 3, 2) - m0, a11 = sisnr(1, 4, 3) - m1;
    const double p0 = a00 + a11, p1 = a01 + a10;       // identity permutation / swapped
    if (p0 >= p1) { out[3 * b] = (float)p0; out[3 * b + 1] = (float)a00; out[3 * b + 2] = (float)a11; }
    else { out[3 * b] = (float)p1; out[3 * b + 1] = (float)a01; out[3 * b + 2] = (float)a10; }
  }
}

}  // namespace shell
}  // namespace sepref
