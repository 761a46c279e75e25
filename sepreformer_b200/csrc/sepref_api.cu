// libsepref_b200.so - C ABI, weight packing and the launch schedule of the B200 separator.
// The ABI is declared in include/sepref.h; each entry point there names the reference code it replaces.
#include "../../include/sepref.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <array>
#include <functional>
#include <map>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "kernels_attn.cuh"
#include "kernels_simt.cuh"
#include "kernels_tc.cuh"
#include "kernels_gcfn_pair.cuh"
#include "kernels_gcfn_trio.cuh"
#include "kernels_gcfn_tm.cuh"
#include "kernels_cla_front.cuh"
#include "kernels_shell.cuh"

namespace sepref {

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CU_TRY(expr)                                                                               \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess) return fail(SEPREF_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// ------------------------------------------------------------------------------------------------ weights
struct HostT {
  std::vector<int64_t> shape;
  std::vector<float> v;
  bool set = false;
  bool optional = false;       // model-shell tensors ("@..." keys): only the model-level entry points need them
};

struct GcfnW {   // network.py:46-66 after folding: LN affine -> w1/b1, LayerScale -> w2/b2
  const float *w1, *b1;        // [6F, F], [6F]
  const float *dw, *dwb;       // tap-major [3][6F], [6F]
  const float *w2, *b2;        // [F, 3F], [F]
  tc::GcfnPack tc;             // tensor-core operand copies (TF32-rounded, re-tiled)
  tc::GcfnPairPack pair;       // F = 128: FP16 operands in CTA-pair order (kernels_gcfn_pair.cuh)
  tc::GcfnTmPack tm;           // F = 128: FP16 operands in chunk order for the frames-as-M kernel (kernels_gcfn_tm.cuh)
  bool f16_ok = true;          // pack-time range bound of the FP16 stage-2 operand (see range_bound_*)
};
struct MhaW {    // network.py:76-88: q|k|v stacked, LN affine and 1/sqrt(dk) folded in, LayerScale folded into out
  const float *wqkv, *bqkv;    // [3F, F], [3F]
  const float *wo, *bo;        // [F, F], [F]
  tc::TcLin tqkv, to;          // TF32 copies + TMA maps
};
struct EgaW {    // network.py:126-136
  MhaW att;
  const float *wg, *bg;        // gate linear with its LayerNorm affine folded: [F, F], [F]
  tc::TcLin tg;
};
struct ClaW {    // network.py:159-172: LN -> w1, BN -> w2, LayerScale -> w3
  const float *w1, *b1;        // [2F, F]
  const float *dw, *dwb;       // tap-major [K][F], [F]
  const float *w2, *b2;        // [2F, F]
  const float *w3, *b3;        // [F, 2F]
  tc::TcLin t1, t2, t3;        // t1 rows re-ordered into (value tile, gate tile) pairs
  bool f16_ok_b = true;        // cla_b: conv output d and GELU output both provably inside the FP16 range
};
struct DownW { const float *dw, *b; };                     // module.py:63-70, BN folded; tap-major [K][F]
struct SplitW { const float *wa, *ba, *wb, *bb, *gamma, *beta; tc::TcLin ta, tb; };   // module.py:110-118 (ta pair-ordered)
struct FuseW { const float *w, *b; tc::TcLin t; };         // module.py:187: [F, 2F]
struct SpkW { MhaW att; const GcfnW* ff = nullptr; };                       // network.py:227-231
struct ShellW {    // the layers around the separator (module.py:12-35, 237-283), see kernels_shell.cuh
  const float* enc_w = nullptr;        // AudioEncoder filter, tap-major [16][256]
  const float *gn_g = nullptr, *gn_b = nullptr;   // FeatureProjector GroupNorm affine [256]
  tc::TcLin proj;                      // FeatureProjector 1x1 conv [F, 256], zero bias
  tc::TcLin out1;                      // OutputLayer Linear(F -> 4F), (value tile, gate tile) pair order
  tc::TcLin out2;                      // w_dec^T . Linear(2F -> 256): [16 of 128, 2F] + folded bias
  bool ready = false;
};

}  // namespace sepref

using namespace sepref;

struct sepref_handle {
  sepref_config cfg;
  int device = 0;
  int gemm_path = 1;
  int debug_sync = 0;
  int cluster = 2;          // CTAs per cluster sharing multicast weight slabs (1, 2 or 4)
  int launches = 0;
  int sm_count = 148;
  bool finalized = false;
  std::map<std::string, HostT> params;          // expected keys (ordered for stable "first missing")
  std::string missing_key;
  float* slab = nullptr;                         // all packed weights
  size_t slab_floats = 0;
  std::unordered_map<std::string, GcfnW> gcfn;
  std::unordered_map<std::string, EgaW> ega;
  std::unordered_map<std::string, ClaW> cla;
  std::unordered_map<std::string, SpkW> spk;
  std::unordered_map<std::string, DownW> down;
  std::unordered_map<std::string, SplitW> split;
  std::unordered_map<std::string, FuseW> fuse;
  const float* pe_k = nullptr;                   // [2*maxlen, dk]
  const void* pe_k_h = nullptr;                  // the same table as FP16 (what k_attn_relpos multiplies)
  ShellW shell;
  // optional per-launch timing (SEPREF_OPT_PROFILE): one event after every launch, names alongside
  int profile = 0;
  std::vector<cudaEvent_t> prof_events;
  std::vector<const char*> prof_names;
  size_t prof_used = 0;
  // host-buffer entry point: staging arena, copy streams and per-sub-batch events
  char* arena = nullptr;
  size_t arena_bytes = 0;
  long long* dbg_clk = nullptr;          // tools: timeline buffer for one k_tok configuration (dbg_which)
  char dbg_name[32] = "";
  int dbg_flags = 0;
  int gcfn_wide = 0;                     // SEPREF_OPT_GCFN_WIDE: 160-frame GCFN tiles (fp16, F = 128)
  int gcfn_pair = 0;                     // SEPREF_OPT_GCFN_PAIR: weights resident in a CTA pair (FP16 operands, F = 128)
  int cla_fused = 1;                     // SEPREF_OPT_CLA_FUSED: CLA's first half as one kernel (k_cla_front), FP16 d (FP16 operands, F = 128)
  int gcfn_tm = 0;                       // SEPREF_OPT_GCFN_TM: frames-as-M GCFN kernel, 1 = CTA pair (cta_group::2), 2 = single CTA
  int gcfn_trio = 0;                     // SEPREF_OPT_GCFN_TRIO: weights resident in a cluster of three CTAs (FP16 operands, F = 128)
  tc::TrioState trio;
  int raw_f16 = 0;                       // SEPREF_OPT_RAW_F16: FP16 operands also for GEMMs fed by the raw residual stream
  int* range_flags = nullptr;            // device: kRangeSites flags + 1 re-run counter (TOK_LAUNCH_RAW)
  int f16_fallbacks = 0;                 // GEMM groups whose pack-time range bound forces TF32 operands on gemm_path 2
  double attn_bound = 0.0;               // largest pack-time bound of a q/k/v element (attention runs on FP16 operands)
  std::map<std::tuple<int, int, int>, size_t> ws_cache;   // (batch, t_enc, tensor-core path?) -> workspace bytes
  // SEPREF_OPT_CUDA_GRAPH: the launch sequence of a forward is captured once per (shapes, options, buffer addresses)
  // and replayed with one cudaGraphLaunch - the ~260 launches of a forward cost ~2 ms of host time otherwise
  struct GraphEntry {
    std::array<uintptr_t, 20> key{};
    cudaGraphExec_t exec = nullptr;
    int launches = 0;
    bool bad = false;          // capture failed once: this key stays on the eager path
    unsigned long long stamp = 0;
  };
  int use_graphs = 0;
  int graph_replays = 0;                 // forwards served by a graph launch since finalize (tests)
  unsigned long long graph_clock = 0;
  std::vector<GraphEntry> graphs;
  cudaStream_t s_cap = nullptr;
  int host_chunk = 16;                   // utterances per sub-batch of sepref_separator_forward_host
  cudaStream_t s_in = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_done;
  // pipelined host entry (submit / wait): two staging slots share the copy streams and one compute stream
  struct HostSlot {
    char* arena = nullptr;
    size_t bytes = 0;
    cudaEvent_t ev_in = nullptr, ev_cmp = nullptr, ev_out = nullptr;
    bool pending = false;
  } slots[2];
  cudaStream_t s_cmp = nullptr;
};

namespace sepref {

// ------------------------------------------------------------------------------------------------ expected keys
static void expect(sepref_handle* h, const std::string& key, std::vector<int64_t> shape) {
  HostT t;
  t.shape = std::move(shape);
  h->params[key] = std::move(t);
}
static void expect_linear(sepref_handle* h, const std::string& p, int64_t out, int64_t in) {
  expect(h, p + "weight", {out, in});
  expect(h, p + "bias", {out});
}
static void expect_norm(sepref_handle* h, const std::string& p, int64_t c) {
  expect(h, p + "weight", {c});
  expect(h, p + "bias", {c});
}
static void expect_bn(sepref_handle* h, const std::string& p, int64_t c) {
  expect_norm(h, p, c);
  expect(h, p + "running_mean", {c});
  expect(h, p + "running_var", {c});
}
static void expect_gcfn(sepref_handle* h, const std::string& p) {
  const int64_t F = h->cfg.feat;
  expect_norm(h, p + "net1.0.", F);
  expect_linear(h, p + "net1.1.", 6 * F, F);
  expect(h, p + "depthwise.weight", {6 * F, 1, 3});
  expect(h, p + "depthwise.bias", {6 * F});
  expect_linear(h, p + "net2.2.", F, 3 * F);
  expect(h, p + "Layer_scale.layer_scale", {1, 1, F});
}
static void expect_mha(sepref_handle* h, const std::string& p) {
  const int64_t F = h->cfg.feat;
  expect_norm(h, p + "layer_norm.", F);
  for (const char* n : {"linear_q.", "linear_k.", "linear_v.", "linear_out."}) expect_linear(h, p + n, F, F);
  expect(h, p + "Layer_scale.layer_scale", {1, 1, F});
}
static void expect_global(sepref_handle* h, const std::string& p) {
  const int64_t F = h->cfg.feat;
  expect_mha(h, p + "block.ega.block.self_attn.");
  expect_norm(h, p + "block.ega.block.linear.0.", F);
  expect_linear(h, p + "block.ega.block.linear.1.", F, F);
  expect_gcfn(h, p + "block.gcfn.");
}
static void expect_local(sepref_handle* h, const std::string& p) {
  const int64_t F = h->cfg.feat;
  const std::string c = p + "block.cla.";
  expect_norm(h, c + "layer_norm.", F);
  expect_linear(h, c + "linear1.", 2 * F, F);
  expect(h, c + "dw_conv_1d.weight", {F, 1, h->cfg.cla_kernel});
  expect(h, c + "dw_conv_1d.bias", {F});
  expect_linear(h, c + "linear2.", 2 * F, F);
  expect_bn(h, c + "BN.", 2 * F);
  expect_linear(h, c + "linear3.1.", F, 2 * F);
  expect(h, c + "Layer_scale.layer_scale", {1, 1, F});
  expect_gcfn(h, p + "block.gcfn.");
}
static void expect_enc_stage(sepref_handle* h, const std::string& p, bool down) {
  const int64_t F = h->cfg.feat;
  expect_global(h, p + "g_block_1.");
  expect_local(h, p + "l_block_1.");
  expect_global(h, p + "g_block_2.");
  expect_local(h, p + "l_block_2.");
  if (down) {
    expect(h, p + "downconv.down_conv.weight", {F, 1, h->cfg.down_kernel});
    expect(h, p + "downconv.down_conv.bias", {F});
    expect_bn(h, p + "downconv.BN.", F);
  }
}
static void expect_split(sepref_handle* h, const std::string& p) {
  const int64_t F = h->cfg.feat, S = h->cfg.num_spks;
  expect(h, p + "linear.0.weight", {4 * F * S, F, 1});
  expect(h, p + "linear.0.bias", {4 * F * S});
  expect(h, p + "linear.2.weight", {F * S, 2 * F * S, 1});
  expect(h, p + "linear.2.bias", {F * S});
  expect_norm(h, p + "norm.", F);
}
static const char* const kShellKeys[] = {
    "@audio_encoder.conv1d.weight", "@feature_projector.norm.weight", "@feature_projector.norm.bias",
    "@feature_projector.conv1d.weight", "@out_layer.end_conv1x1.0.weight", "@out_layer.end_conv1x1.0.bias",
    "@out_layer.end_conv1x1.2.weight", "@out_layer.end_conv1x1.2.bias", "@audio_decoder.weight"};
static void build_expected(sepref_handle* h) {
  const sepref_config& c = h->cfg;
  const int64_t F = c.feat;
  // model shell, keyed by "@" + the key in Model.state_dict() (model.py:24-29); optional
  expect(h, "@audio_encoder.conv1d.weight", {shell::kEncC, 1, shell::kEncK});
  expect_norm(h, "@feature_projector.norm.", shell::kEncC);
  expect(h, "@feature_projector.conv1d.weight", {F, shell::kEncC, 1});
  expect_linear(h, "@out_layer.end_conv1x1.0.", 4 * F, F);
  expect_linear(h, "@out_layer.end_conv1x1.2.", shell::kEncC, 2 * F);
  expect(h, "@audio_decoder.weight", {shell::kEncC, 1, shell::kEncK});
  for (const char* k : kShellKeys) h->params[k].optional = true;
  expect(h, "pos_emb.pe_k.weight", {2 * (int64_t)c.maxlen, F / c.heads});
  for (int s = 0; s < c.num_stages; ++s) expect_enc_stage(h, "enc_stages." + std::to_string(s) + ".", true);
  expect_enc_stage(h, "bottleneck_G.", false);
  if (c.per_stage_split)
    for (int s = 0; s <= c.num_stages; ++s) expect_split(h, "spk_split_blocks." + std::to_string(s) + ".");
  else
    expect_split(h, "spk_split_block.");
  for (int s = 0; s < c.num_stages; ++s) {
    expect(h, "simple_fusion." + std::to_string(s) + ".weight", {F, 2 * F, 1});
    expect(h, "simple_fusion." + std::to_string(s) + ".bias", {F});
    const std::string p = "dec_stages." + std::to_string(s) + ".";
    for (int n = 1; n <= 3; ++n) {
      expect_global(h, p + "g_block_" + std::to_string(n) + ".");
      expect_local(h, p + "l_block_" + std::to_string(n) + ".");
      expect_mha(h, p + "spk_attn_" + std::to_string(n) + ".self_attn.");
      expect_gcfn(h, p + "spk_attn_" + std::to_string(n) + ".feed_forward.");
    }
  }
}

// ------------------------------------------------------------------------------------------------ packing
static float tf32_rna_host(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;
  u = (u + 0x1000u) & 0xffffe000u;
  float y;
  memcpy(&y, &u, 4);
  return y;
}

struct Packer {
  sepref_handle* h;
  std::vector<float> host;
  std::vector<std::pair<const float**, size_t>> fix;
  const std::vector<float>& P(const std::string& k) const { return h->params.at(k).v; }
  void put(const float** field, const std::vector<float>& v) {
    size_t off = (host.size() + 63) & ~size_t(63);     // 256-byte alignment of every tensor
    host.resize(off);
    host.insert(host.end(), v.begin(), v.end());
    fix.emplace_back(field, off);
  }
  void put_half(const void** field, const std::vector<uint16_t>& v) {   // fp16 payload stored inside the float slab
    size_t off = (host.size() + 63) & ~size_t(63);
    host.resize(off + (v.size() + 1) / 2);
    memcpy(host.data() + off, v.data(), v.size() * sizeof(uint16_t));
    fix.emplace_back(reinterpret_cast<const float**>(field), off);
  }
  void put_void(const void** field, const std::vector<float>& v) { put(reinterpret_cast<const float**>(field), v); }
};

// Operand copies of a [rows, cols] weight matrix for both tensor-core kinds:
//   TF32: rounded to nearest, inverse scale 1
//   FP16: each row multiplied by 2^-floor(log2(max|w|)) (row maximum lands in [1,2)) and rounded to nearest; the
//         epilogue multiplies the accumulator by the exact inverse 2^e, so the scaling costs no precision.
static void put_operands(Packer& pk, const std::vector<float>& w, int rows, int cols, const void* (&wdst)[2],
                         const float* (&sdst)[2], std::vector<float>* sinv_f16_out = nullptr) {
  std::vector<float> wt(w.size()), ones(rows, 1.0f), sinv(rows);
  std::vector<uint16_t> wh(w.size());
  for (int r = 0; r < rows; ++r) {
    float m = 0.f;
    for (int i = 0; i < cols; ++i) m = std::fmax(m, std::fabs(w[(size_t)r * cols + i]));
    int e = 0;
    if (m > 0.f && std::isfinite(m)) e = (int)std::floor(std::log2((double)m));
    const float sc = (float)std::ldexp(1.0, -e);
    sinv[r] = (float)std::ldexp(1.0, e);
    for (int i = 0; i < cols; ++i) {
      const float v = w[(size_t)r * cols + i];
      wt[(size_t)r * cols + i] = tf32_rna_host(v);
      const __half hv = __float2half_rn(v * sc);
      memcpy(&wh[(size_t)r * cols + i], &hv, 2);
    }
  }
  pk.put_void(&wdst[0], wt);
  pk.put_half(&wdst[1], wh);
  pk.put(&sdst[0], ones);
  pk.put(&sdst[1], sinv);
  if (sinv_f16_out) *sinv_f16_out = sinv;
}

// y = W . (gamma * n + beta) + b  ==  (W diag(gamma)) . n + (b + W . beta)
static void fold_ln_in(std::vector<float>& w, std::vector<float>& b, int out, int in, const std::vector<float>& gamma,
                       const std::vector<float>& beta) {
  for (int o = 0; o < out; ++o) {
    double acc = b[o];
    for (int i = 0; i < in; ++i) {
      acc += (double)w[(size_t)o * in + i] * beta[i];
      w[(size_t)o * in + i] *= gamma[i];
    }
    b[o] = (float)acc;
  }
}
// row scaling: y = s * (W x + b)
static void scale_rows(std::vector<float>& w, std::vector<float>& b, int out, int in, const std::vector<float>& s) {
  for (int o = 0; o < out; ++o) {
    for (int i = 0; i < in; ++i) w[(size_t)o * in + i] *= s[o];
    b[o] *= s[o];
  }
}
// [C,1,K] -> tap-major [K][C]
static std::vector<float> tap_major(const std::vector<float>& w, int C, int K) {
  std::vector<float> o((size_t)C * K);
  for (int c = 0; c < C; ++c)
    for (int k = 0; k < K; ++k) o[(size_t)k * C + c] = w[(size_t)c * K + k];
  return o;
}

// ---- FP16 range bounds (gemm_path 2) ----------------------------------------------------------------------------------
// FP16 operands carry TF32's significand but only a 5-bit exponent, and the producers convert with saturation, so an
// out-of-range activation would be clipped silently.  Every activation that is written as an FP16 operand is therefore
// bounded at pack time from the weights alone; a GEMM group whose bound is not safely inside the FP16 range runs with
// TF32 operands (fp32 range) instead.  The bounds start from |LayerNorm(x)|_2 <= sqrt(F) (Cauchy-Schwarz per row):
//   |W' . LN(x) + b'|_c <= |W'_c|_2 * sqrt(F) + |b'_c|
// (W', b' with the LayerNorm affine folded in).  GLU and GELU do not increase magnitudes; depthwise filters multiply
// the bound by the sum of their absolute taps.  Activations that are NOT normalised first (the raw residual stream fed
// to SpkSplit, the fusion conv and the output layer) have no such bound: those GEMMs check the range at run time and are
// re-computed with TF32 operands when it was exceeded (TOK_LAUNCH_RAW below).
static constexpr double kF16Safe = 3.0e4;      // half of the FP16 maximum (65504)
static std::vector<double> ln_fed_row_bounds(const std::vector<float>& w, const std::vector<float>& b, int out, int in) {
  std::vector<double> r(out);
  for (int o = 0; o < out; ++o) {
    double ss = 0.0;
    for (int i = 0; i < in; ++i) ss += (double)w[(size_t)o * in + i] * w[(size_t)o * in + i];
    r[o] = std::sqrt(ss) * std::sqrt((double)in) + std::fabs((double)b[o]);
  }
  return r;
}

// Operand copies for the tensor-core path.  pair_c > 0: rows [0,pair_c) are GLU values and [pair_c, 2*pair_c) the
// matching gates; they are interleaved in tiles of 128 (value tile j, gate tile j) so that one MMA step yields a pair.
static void pack_tc_lin(Packer& pk, tc::TcLin& l, const std::vector<float>& w, const std::vector<float>& b, int rows,
                        int cols, int pair_c) {
  std::vector<float> wr((size_t)rows * cols), bt(rows);
  for (int dst = 0; dst < rows; ++dst) {
    int src = dst;
    if (pair_c > 0) {
      const int tile = dst / 128, r = dst % 128;
      src = (tile & 1) * pair_c + (tile >> 1) * 128 + r;
    }
    std::copy(w.begin() + (size_t)src * cols, w.begin() + (size_t)(src + 1) * cols, wr.begin() + (size_t)dst * cols);
    bt[dst] = b[src];
  }
  l.rows = rows; l.cols = cols;
  put_operands(pk, wr, rows, cols, l.w, l.sinv);
  pk.put(&l.b, bt);
}

static void pack_gcfn(Packer& pk, const std::string& p, GcfnW& g) {
  const int F = pk.h->cfg.feat;
  std::vector<float> w1 = pk.P(p + "net1.1.weight"), b1 = pk.P(p + "net1.1.bias");
  fold_ln_in(w1, b1, 6 * F, F, pk.P(p + "net1.0.weight"), pk.P(p + "net1.0.bias"));
  std::vector<float> w2 = pk.P(p + "net2.2.weight"), b2 = pk.P(p + "net2.2.bias");
  scale_rows(w2, b2, F, 3 * F, pk.P(p + "Layer_scale.layer_scale"));
  std::vector<float> dw = tap_major(pk.P(p + "depthwise.weight"), 6 * F, 3);
  {   // FP16 stage-2 operand u = conv3(h)_value * sigmoid(.): |u_c| <= sum|taps_c| * max|h_c| + |bias_c|
    const std::vector<double> hb = ln_fed_row_bounds(w1, b1, 6 * F, F);
    const auto& db = pk.P(p + "depthwise.bias");
    double worst = 0.0;
    for (int c = 0; c < 3 * F; ++c) {
      const double taps = std::fabs((double)dw[c]) + std::fabs((double)dw[(size_t)6 * F + c]) + std::fabs((double)dw[(size_t)12 * F + c]);
      worst = std::fmax(worst, taps * hb[c] + std::fabs((double)db[c]));
    }
    g.f16_ok = std::isfinite(worst) && worst < kF16Safe;
    if (!g.f16_ok) ++pk.h->f16_fallbacks;
  }
  pk.put(&g.w1, w1); pk.put(&g.b1, b1);
  pk.put(&g.dw, dw); pk.put(&g.dwb, pk.P(p + "depthwise.bias"));
  pk.put(&g.w2, w2); pk.put(&g.b2, b2);
  // tensor-core copies: GEMM1 rows re-ordered into (value tile, gate tile) pairs of 128 channels.
  // Depthwise taps / bias are stored pre-scaled by 1/2 for the tanh form of the gate, u = (dv/2) * (1 + tanh(dg/2));
  // cb is the conv constant of interior columns, where h = D + b1 everywhere: (dwb + b1 * (w0+w1+w2)) / 2
  std::vector<float> w1r((size_t)6 * F * F), b1t(6 * F), dwt((size_t)3 * 6 * F), dwbt(6 * F), cbt(6 * F);
  const int C = 3 * F, nchunk = C / 128;
  for (int j = 0; j < nchunk; ++j)
    for (int half = 0; half < 2; ++half)
      for (int r = 0; r < 128; ++r) {
        const int src = half * C + j * 128 + r;            // original output channel
        const int dst = (2 * j + half) * 128 + r;          // packed row
        for (int i = 0; i < F; ++i) w1r[(size_t)dst * F + i] = w1[(size_t)src * F + i];
        b1t[dst] = b1[src];
        double wsum = 0.0;
        for (int k = 0; k < 3; ++k) {
          dwt[(size_t)k * 6 * F + dst] = 0.5f * dw[(size_t)k * 6 * F + src];
          wsum += dw[(size_t)k * 6 * F + src];
        }
        const float db = pk.P(p + "depthwise.bias")[src];
        dwbt[dst] = 0.5f * db;
        cbt[dst] = (float)(0.5 * ((double)db + (double)b1[src] * wsum));
      }
  std::vector<float> s1h;
  put_operands(pk, w1r, 6 * F, F, g.tc.w1, g.tc.s1inv, &s1h);
  put_operands(pk, w2, F, 3 * F, g.tc.w2, g.tc.s2inv);
  std::vector<float> dwf16(dwt.size());
  for (int k = 0; k < 3; ++k)
    for (int r = 0; r < 6 * F; ++r) dwf16[(size_t)k * 6 * F + r] = dwt[(size_t)k * 6 * F + r] * s1h[r];
  pk.put(&g.tc.dwf[0], dwt);        // TF32: inverse scale is 1
  pk.put(&g.tc.dwf[1], dwf16);
  pk.put(&g.tc.b1, b1t); pk.put(&g.tc.dw, dwt); pk.put(&g.tc.dwb, dwbt); pk.put(&g.tc.cb, cbt); pk.put(&g.tc.b2, b2);
  if (F == tc::PairTraits::F) {
    // CTA-pair order (kernels_gcfn_pair.cuh): CTA c owns GLU channels [192c, 192c + 192) as three mixed chunks of
    // 64 value rows followed by the 64 matching gate rows; FP16 operands with per-row power-of-two scaling
    constexpr int R = tc::PairTraits::ROWS;
    std::vector<uint16_t> wh((size_t)R * F);
    std::vector<float> cbp(R), dwfp((size_t)3 * R), klp(R), krp(R);
    const auto& dbias = pk.P(p + "depthwise.bias");
    for (int c = 0; c < 2; ++c)
      for (int k = 0; k < 3; ++k)
        for (int l = 0; l < 128; ++l) {
          const int dst = c * 384 + k * 128 + l;
          const int chn = c * 192 + k * 64 + (l & 63);
          const int src = (l < 64 ? 0 : 3 * F) + chn;                 // row of the folded [6F, F] matrix
          float m = 0.f;
          for (int i = 0; i < F; ++i) m = std::fmax(m, std::fabs(w1[(size_t)src * F + i]));
          int e = 0;
          if (m > 0.f && std::isfinite(m)) e = (int)std::floor(std::log2((double)m));
          const float sc = (float)std::ldexp(1.0, -e), sinv = (float)std::ldexp(1.0, e);
          for (int i = 0; i < F; ++i) {
            const __half hv = __float2half_rn(w1[(size_t)src * F + i] * sc);
            memcpy(&wh[(size_t)dst * F + i], &hv, 2);
          }
          const double t0 = dw[src], t1 = dw[(size_t)6 * F + src], t2 = dw[(size_t)12 * F + src];
          cbp[dst] = (float)(0.5 * ((double)dbias[src] + (double)b1[src] * (t0 + t1 + t2)));
          dwfp[dst] = (float)(0.5 * t0) * sinv; dwfp[(size_t)R + dst] = (float)(0.5 * t1) * sinv; dwfp[(size_t)2 * R + dst] = (float)(0.5 * t2) * sinv;
          klp[dst] = (float)(0.5 * t0 * (double)b1[src]);
          krp[dst] = (float)(0.5 * t2 * (double)b1[src]);
        }
    pk.put_half(&g.pair.w1, wh);
    pk.put(&g.pair.cb, cbp); pk.put(&g.pair.dwf, dwfp); pk.put(&g.pair.kl, klp); pk.put(&g.pair.kr, krp);
  }
  if (F == tc::TmTraits<true>::F) {
    // frames-as-M order (kernels_gcfn_tm.cuh): chunk j = GLU channels [64j, 64j + 64): 64 value rows, then the 64 gate rows
    using TT = tc::TmTraits<true>;
    std::vector<uint16_t> wh((size_t)TT::ROWS_W1 * F);
    std::vector<float> tab(TT::TAB_FLOATS, 0.f);
    const auto& dbias = pk.P(p + "depthwise.bias");
    for (int j = 0; j < TT::NCH; ++j)
      for (int vg = 0; vg < 2; ++vg)
        for (int i = 0; i < 64; ++i) {
          const int dst = j * 128 + vg * 64 + i;
          const int src = vg * 3 * F + j * 64 + i;                    // row of the folded [6F, F] matrix
          float m = 0.f;
          for (int k = 0; k < F; ++k) m = std::fmax(m, std::fabs(w1[(size_t)src * F + k]));
          int e = 0;
          if (m > 0.f && std::isfinite(m)) e = (int)std::floor(std::log2((double)m));
          const float sc = (float)std::ldexp(1.0, -e), sinv = (float)std::ldexp(1.0, e);
          for (int k = 0; k < F; ++k) {
            const __half hv = __float2half_rn(w1[(size_t)src * F + k] * sc);
            memcpy(&wh[(size_t)dst * F + k], &hv, 2);
          }
          const double t0 = dw[src], t1 = dw[(size_t)6 * F + src], t2 = dw[(size_t)12 * F + src];
          // accumulator column i of the chunk = 8*kb + 2*c + e2: float4 entries [j][kb][vg][half][c] / [j][kb][vg][c]
          const int kb = i >> 3, c = (i & 7) >> 1, e2 = i & 1;
          float* tp = &tab[((((size_t)(j * 8 + kb) * 2 + vg) * 2 + 0) * 4 + c) * 4];
          tp[0 + e2] = (float)(0.5 * t0) * sinv;
          tp[2 + e2] = (float)(0.5 * t1) * sinv;
          tp[16 + 0 + e2] = (float)(0.5 * t2) * sinv;                  // half 1 is 4 float4 = 16 floats further
          tp[16 + 2 + e2] = (float)(0.5 * ((double)dbias[src] + (double)b1[src] * (t0 + t1 + t2)));
          float* ep = &tab[TT::TAP_FLOATS + (((size_t)(j * 8 + kb) * 2 + vg) * 4 + c) * 4];
          ep[0 + e2] = (float)(0.5 * t0 * (double)b1[src]);
          ep[2 + e2] = (float)(0.5 * t2 * (double)b1[src]);
        }
    {   // s2inv | b2 of the FP16 second matrix: the same per-row scaling put_operands applied to w2
      for (int r = 0; r < F; ++r) {
        float m = 0.f;
        for (int k = 0; k < 3 * F; ++k) m = std::fmax(m, std::fabs(w2[(size_t)r * 3 * F + k]));
        int e = 0;
        if (m > 0.f && std::isfinite(m)) e = (int)std::floor(std::log2((double)m));
        tab[TT::TAP_FLOATS + TT::EDGE_FLOATS + r] = (float)std::ldexp(1.0, e);
        tab[TT::TAP_FLOATS + TT::EDGE_FLOATS + F + r] = b2[r];
      }
    }
    pk.put_half(&g.tm.w1, wh);
    pk.put(&g.tm.tab, tab);
  }
}

static void pack_mha(Packer& pk, const std::string& p, MhaW& m) {
  const int F = pk.h->cfg.feat, dk = F / pk.h->cfg.heads;
  std::vector<float> w((size_t)3 * F * F), b(3 * F);
  const char* names[3] = {"linear_q.", "linear_k.", "linear_v."};
  for (int j = 0; j < 3; ++j) {
    const auto& wj = pk.P(p + names[j] + "weight");
    const auto& bj = pk.P(p + names[j] + "bias");
    std::copy(wj.begin(), wj.end(), w.begin() + (size_t)j * F * F);
    std::copy(bj.begin(), bj.end(), b.begin() + (size_t)j * F);
  }
  fold_ln_in(w, b, 3 * F, F, pk.P(p + "layer_norm.weight"), pk.P(p + "layer_norm.bias"));
  // scores / sqrt(dk) (network.py:110,112), times log2(e): the attention kernels evaluate softmax as 2^(s' - max s'), one
  // MUFU.EX2 per score instead of expf()'s scale / range-check / rescale sequence (exactly the same softmax)
  const float qs = (float)(1.4426950408889634 / std::sqrt((double)dk));
  for (size_t i = 0; i < (size_t)F * F; ++i) w[i] *= qs;
  for (int i = 0; i < F; ++i) b[i] *= qs;
  for (double v : ln_fed_row_bounds(w, b, 3 * F, F)) pk.h->attn_bound = std::fmax(pk.h->attn_bound, std::isfinite(v) ? v : 1e300);
  std::vector<float> wo = pk.P(p + "linear_out.weight"), bo = pk.P(p + "linear_out.bias");
  scale_rows(wo, bo, F, F, pk.P(p + "Layer_scale.layer_scale"));
  pk.put(&m.wqkv, w); pk.put(&m.bqkv, b); pk.put(&m.wo, wo); pk.put(&m.bo, bo);
  pack_tc_lin(pk, m.tqkv, w, b, 3 * F, F, 0);
  pack_tc_lin(pk, m.to, wo, bo, F, F, 0);
}

static void pack_ega(Packer& pk, const std::string& p, EgaW& e) {
  const int F = pk.h->cfg.feat;
  pack_mha(pk, p + "block.self_attn.", e.att);
  std::vector<float> w = pk.P(p + "block.linear.1.weight"), b = pk.P(p + "block.linear.1.bias");
  fold_ln_in(w, b, F, F, pk.P(p + "block.linear.0.weight"), pk.P(p + "block.linear.0.bias"));
  pk.put(&e.wg, w); pk.put(&e.bg, b);
  pack_tc_lin(pk, e.tg, w, b, F, F, 0);
}

static void bn_scale_shift(const Packer& pk, const std::string& p, int C, std::vector<float>& s, std::vector<float>& sh) {
  const auto &g = pk.P(p + "weight"), &b = pk.P(p + "bias"), &m = pk.P(p + "running_mean"), &v = pk.P(p + "running_var");
  s.resize(C); sh.resize(C);
  for (int c = 0; c < C; ++c) {
    s[c] = (float)((double)g[c] / std::sqrt((double)v[c] + (double)kBnEps));
    sh[c] = b[c] - m[c] * s[c];
  }
}

static void pack_cla(Packer& pk, const std::string& p, ClaW& c) {
  const int F = pk.h->cfg.feat, K = pk.h->cfg.cla_kernel;
  std::vector<float> w1 = pk.P(p + "linear1.weight"), b1 = pk.P(p + "linear1.bias");
  fold_ln_in(w1, b1, 2 * F, F, pk.P(p + "layer_norm.weight"), pk.P(p + "layer_norm.bias"));
  std::vector<float> w2 = pk.P(p + "linear2.weight"), b2 = pk.P(p + "linear2.bias"), s, sh;
  bn_scale_shift(pk, p + "BN.", 2 * F, s, sh);
  scale_rows(w2, b2, 2 * F, F, s);
  for (int i = 0; i < 2 * F; ++i) b2[i] += sh[i];
  std::vector<float> w3 = pk.P(p + "linear3.1.weight"), b3 = pk.P(p + "linear3.1.bias");
  scale_rows(w3, b3, F, 2 * F, pk.P(p + "Layer_scale.layer_scale"));
  {   // cla_b writes two FP16 operands: d = conv65(GLU(h1)) and g = GELU(W2'.d + b2')
    const std::vector<double> hb = ln_fed_row_bounds(w1, b1, 2 * F, F);       // rows [0,F) are the GLU values
    const auto &cw = pk.P(p + "dw_conv_1d.weight"), &cb = pk.P(p + "dw_conv_1d.bias");
    std::vector<double> dbound(F);
    double worst = 0.0;
    for (int ch = 0; ch < F; ++ch) {
      double taps = 0.0;
      for (int k = 0; k < K; ++k) taps += std::fabs((double)cw[(size_t)ch * K + k]);
      dbound[ch] = taps * hb[ch] + std::fabs((double)cb[ch]);
      worst = std::fmax(worst, dbound[ch]);
    }
    for (int j = 0; j < 2 * F; ++j) {
      double acc = std::fabs((double)b2[j]);
      for (int ch = 0; ch < F; ++ch) acc += std::fabs((double)w2[(size_t)j * F + ch]) * dbound[ch];
      worst = std::fmax(worst, acc);
    }
    c.f16_ok_b = std::isfinite(worst) && worst < kF16Safe;
    if (!c.f16_ok_b) ++pk.h->f16_fallbacks;
  }
  pk.put(&c.w1, w1); pk.put(&c.b1, b1);
  pk.put(&c.dw, tap_major(pk.P(p + "dw_conv_1d.weight"), F, K)); pk.put(&c.dwb, pk.P(p + "dw_conv_1d.bias"));
  pk.put(&c.w2, w2); pk.put(&c.b2, b2); pk.put(&c.w3, w3); pk.put(&c.b3, b3);
  pack_tc_lin(pk, c.t1, w1, b1, 2 * F, F, F);
  pack_tc_lin(pk, c.t2, w2, b2, 2 * F, F, 0);
  pack_tc_lin(pk, c.t3, w3, b3, F, 2 * F, 0);
}

static void pack_down(Packer& pk, const std::string& p, DownW& d) {
  const int F = pk.h->cfg.feat, K = pk.h->cfg.down_kernel;
  std::vector<float> w = pk.P(p + "down_conv.weight"), b = pk.P(p + "down_conv.bias"), s, sh;
  bn_scale_shift(pk, p + "BN.", F, s, sh);
  for (int c = 0; c < F; ++c) {
    for (int k = 0; k < K; ++k) w[(size_t)c * K + k] *= s[c];
    b[c] = b[c] * s[c] + sh[c];
  }
  pk.put(&d.dw, tap_major(w, F, K)); pk.put(&d.b, b);
}

static void pack_split(Packer& pk, const std::string& p, SplitW& s) {
  pk.put(&s.wa, pk.P(p + "linear.0.weight")); pk.put(&s.ba, pk.P(p + "linear.0.bias"));
  pk.put(&s.wb, pk.P(p + "linear.2.weight")); pk.put(&s.bb, pk.P(p + "linear.2.bias"));
  pk.put(&s.gamma, pk.P(p + "norm.weight")); pk.put(&s.beta, pk.P(p + "norm.bias"));
  const int F = pk.h->cfg.feat, S = pk.h->cfg.num_spks;
  pack_tc_lin(pk, s.ta, pk.P(p + "linear.0.weight"), pk.P(p + "linear.0.bias"), 4 * F * S, F, 2 * F * S);
  pack_tc_lin(pk, s.tb, pk.P(p + "linear.2.weight"), pk.P(p + "linear.2.bias"), F * S, 2 * F * S, 0);
}

// Model shell (kernels_shell.cuh).  The decoder is a linear map of the output layer's result, so it is folded into the
// second matrix:  frame[j] = sum_c w_dec[c, j] * (W2[c, :] . g + b2[c])  =  (w_dec^T W2)[j, :] . g + (w_dec^T b2)[j].
static void pack_shell(Packer& pk, ShellW& sh) {
  const int F = pk.h->cfg.feat, C = shell::kEncC, K = shell::kEncK;
  pk.put(&sh.enc_w, tap_major(pk.P("@audio_encoder.conv1d.weight"), C, K));
  pk.put(&sh.gn_g, pk.P("@feature_projector.norm.weight"));
  pk.put(&sh.gn_b, pk.P("@feature_projector.norm.bias"));
  pack_tc_lin(pk, sh.proj, pk.P("@feature_projector.conv1d.weight"), std::vector<float>(F, 0.f), F, C, 0);
  pack_tc_lin(pk, sh.out1, pk.P("@out_layer.end_conv1x1.0.weight"), pk.P("@out_layer.end_conv1x1.0.bias"), 4 * F, F, 2 * F);
  const auto &w2 = pk.P("@out_layer.end_conv1x1.2.weight"), &b2 = pk.P("@out_layer.end_conv1x1.2.bias"), &wd = pk.P("@audio_decoder.weight");
  std::vector<float> wf((size_t)128 * 2 * F, 0.f), bf(128, 0.f);
  for (int j = 0; j < K; ++j) {
    double bacc = 0.0;
    for (int c = 0; c < C; ++c) bacc += (double)wd[(size_t)c * K + j] * b2[c];
    bf[j] = (float)bacc;
    for (int i = 0; i < 2 * F; ++i) {
      double acc = 0.0;
      for (int c = 0; c < C; ++c) acc += (double)wd[(size_t)c * K + j] * w2[(size_t)c * 2 * F + i];
      wf[(size_t)j * 2 * F + i] = (float)acc;
    }
  }
  pack_tc_lin(pk, sh.out2, wf, bf, 128, 2 * F, 0);
}

// ------------------------------------------------------------------------------------------------ launch context
struct Arena {          // bump allocator over caller memory; measure == true is a dry run that only sizes
  char* base = nullptr;
  size_t off = 0, cap = 0, peak = 0;
  bool measure = false;
  bool overflow = false;      // a real run asked for more than the caller's workspace holds
  bool dry() const { return measure; }
  float* f32(size_t n) { return reinterpret_cast<float*>(raw(n * sizeof(float))); }
  void* raw(size_t bytes) {
    off = (off + 255) & ~size_t(255);
    void* p = base ? base + off : nullptr;
    off += bytes;
    if (off > peak) peak = off;
    if (base && off > cap) { overflow = true; return base; }     // never hand out memory past the workspace
    return p;
  }
};

struct Ctx {
  sepref_handle* h;
  cudaStream_t st;
  Arena ws;
  int rc = 0;
  int raw_site = 0;      // raw-stream GEMM launches of this call so far (index of the next range flag)
  bool dry() const { return ws.dry(); }
  bool ok() const { return rc == 0 && !ws.overflow; }
  cudaError_t prof_mark(const char* what) {
    if (h->prof_used == h->prof_events.size()) {
      cudaEvent_t ev;
      cudaError_t e = cudaEventCreate(&ev);
      if (e != cudaSuccess) return e;
      h->prof_events.push_back(ev);
      h->prof_names.push_back(what);
    }
    h->prof_names[h->prof_used] = what;
    return cudaEventRecord(h->prof_events[h->prof_used++], st);
  }
  void after(const char* what) {
    if (dry() || rc) return;
    ++h->launches;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && h->profile) e = prof_mark(what);
    if (e == cudaSuccess && h->debug_sync) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fail(SEPREF_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  }
};

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// ---- thin launchers --------------------------------------------------------------------------------------------
static void pool_layernorm(Ctx& c, const float* x, float* out, size_t out_rows, int r) {
  if (c.dry() || !c.ok()) return;
  const int F = c.h->cfg.feat;
  if (F == 128) simt::k_pool_layernorm<128><<<cdiv(out_rows, 8), 256, 0, c.st>>>(x, out, (int)out_rows, r);
  else simt::k_pool_layernorm<256><<<cdiv(out_rows, 8), 256, 0, c.st>>>(x, out, (int)out_rows, r);
  c.after("k_pool_layernorm");
}

static void gemm(Ctx& c, int epi, const float* A, int lda, const float* W, const float* bias, float* C, int ldc, size_t M,
                 int N, int K, const float* res = nullptr, int ldres = 0, const float* up = nullptr, int up_div = 1) {
  if (c.dry() || !c.ok()) return;
  simt::GemmArgs a{A, lda, W, bias, C, ldc, (int)M, N, K, res, ldres, up, up_div};
  dim3 grid(N / 128, cdiv(M, 128));
  switch (epi) {
    case simt::EPI_BIAS: simt::k_gemm_f32<simt::EPI_BIAS><<<grid, 256, 0, c.st>>>(a); break;
    case simt::EPI_GELU: simt::k_gemm_f32<simt::EPI_GELU><<<grid, 256, 0, c.st>>>(a); break;
    case simt::EPI_RES: simt::k_gemm_f32<simt::EPI_RES><<<grid, 256, 0, c.st>>>(a); break;
    default: simt::k_gemm_f32<simt::EPI_GATE><<<grid, 256, 0, c.st>>>(a); break;
  }
  c.after("k_gemm_f32");
}

static tc::TokParams tok_params(const float* a0, float* out, int ld_out, const tc::TcLin& l1, size_t M) {
  tc::TokParams p{};
  p.a0 = a0; p.out = out; p.ld_out = ld_out; p.M = (long long)M; p.pool_r = 1;
  (void)l1;
  return p;
}
#define TOK_LAUNCH_K(FAMILY, kind, l1, l2, params, what)                                                      \
  do {                                                                                                        \
    if (!c.dry() && c.ok()) {                                                                                 \
      params.dbg_clk = (c.h->dbg_clk && strstr(what, c.h->dbg_name)) ? c.h->dbg_clk : nullptr;                \
      params.dbg_flags = c.h->dbg_flags;                                                                      \
      if (SEPREF_TOK_DISPATCH(FAMILY, c.h->cfg.feat, (kind), l1, l2, params, c.h->sm_count, c.st)) {          \
        c.rc = fail(SEPREF_ERR_CUDA, "%s: %s", what, tc::last_error());                                       \
      } else {                                                                                                \
        c.after(what);                                                                                        \
      }                                                                                                       \
    }                                                                                                         \
  } while (0)
#define TOK_LAUNCH(FAMILY, l1, l2, params, what) TOK_LAUNCH_K(FAMILY, c.h->gemm_path - 1, l1, l2, params, what)
// GEMMs fed by the un-normalised residual stream (SpkSplit, fusion conv, output layer): no pack-time bound covers their
// operands.  On gemm_path 2 they run with FP16 operands and report a range excess at run time; the TF32 launch behind
// them re-computes the same output only if that happened (it exits before any setup otherwise, ~3 us).
constexpr int kRangeSites = 32;
#define TOK_LAUNCH_RAW(FAMILY, l1, l2, params, what)                                                          \
  do {                                                                                                        \
    if (c.h->gemm_path == 2 && !c.h->raw_f16 && c.h->range_flags != nullptr) {                                \
      if (!c.dry() && c.ok() && c.raw_site == 0) {                                                            \
        cudaError_t e__ = cudaMemsetAsync(c.h->range_flags, 0, kRangeSites * sizeof(int), c.st);              \
        if (e__ != cudaSuccess) c.rc = fail(SEPREF_ERR_CUDA, "range flags: %s", cudaGetErrorString(e__));     \
      }                                                                                                       \
      int* flag__ = c.h->range_flags + (c.raw_site++ % kRangeSites);                                          \
      params.range_flag = flag__; params.only_if = nullptr; params.rerun_count = nullptr;                     \
      TOK_LAUNCH_K(FAMILY, tc::KIND_F16, l1, l2, params, what);                                               \
      params.range_flag = nullptr; params.only_if = flag__; params.rerun_count = c.h->range_flags + kRangeSites; \
      TOK_LAUNCH_K(FAMILY, tc::KIND_TF32, l1, l2, params, what);                                              \
    } else {                                                                                                  \
      TOK_LAUNCH_K(FAMILY, kind_of(c, c.h->raw_f16 != 0), l1, l2, params, what);                              \
    }                                                                                                         \
  } while (0)
static int log2i(int v) { int s = 0; while ((1 << s) < v) ++s; return s; }
// operand kind of a GEMM group on the tensor-core paths: FP16 on gemm_path 2 unless its pack-time range bound failed
static int kind_of(const Ctx& c, bool f16_ok) { return (c.h->gemm_path == 2 && f16_ok) ? tc::KIND_F16 : tc::KIND_TF32; }

// CLA's depthwise k=65 'same' convolution (network.py:166,180): u [N,T,F] -> d
static void dwconv65(Ctx& c, const float* u, const float* w, const float* wb, float* d, int N, int T) {
  if (c.dry() || !c.ok()) return;
  const int F = c.h->cfg.feat;
  constexpr int TB = 128, CB = 64, PARTS = 2;
  const size_t smem = (size_t)(TB + 64) * CB * sizeof(float);
  cudaError_t e = cudaSuccess;
  const dim3 grid(cdiv(T, TB), N, F / CB);
  static bool attr_sets[2][16] = {};  // per-device function attribute of the two instantiations: set once
  bool* attr_set = attr_sets[F == 128 ? 0 : 1];
  const int di = c.h->device & 15;
  if (F == 128) {
    if (!attr_set[di]) e = cudaFuncSetAttribute(simt::k_dwconv65_occ<128, TB, CB, PARTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) simt::k_dwconv65_occ<128, TB, CB, PARTS><<<grid, CB * PARTS, smem, c.st>>>(u, w, wb, d, T);
  } else {
    if (!attr_set[di]) e = cudaFuncSetAttribute(simt::k_dwconv65_occ<256, TB, CB, PARTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) simt::k_dwconv65_occ<256, TB, CB, PARTS><<<grid, CB * PARTS, smem, c.st>>>(u, w, wb, d, T);
  }
  if (e == cudaSuccess) attr_set[di] = true;
  if (e != cudaSuccess) { c.rc = fail(SEPREF_ERR_CUDA, "dwconv65 setup: %s", cudaGetErrorString(e)); return; }
  c.after("k_dwconv65_occ");
}

// ---- blocks ----------------------------------------------------------------------------------------------------
// GCFN.forward (network.py:60-66).  x, y: [rows = N*T, F]; y must not alias x.
static void run_gcfn(Ctx& c, const GcfnW& g, const float* x, float* y, int N, int T) {
  const int F = c.h->cfg.feat;
  const size_t rows = (size_t)N * T;
  if (c.h->gemm_path >= 1) {
    if (!c.dry() && c.ok()) {
      const int kind = kind_of(c, g.f16_ok);
      if (kind == tc::KIND_F16 && c.h->gcfn_trio && F == tc::TrioTraits::F && c.h->trio.clusters > 0) {
        if (tc::launch_gcfn_trio(c.h->trio, g.tc, x, y, N, T, c.h->sm_count, c.st)) { c.rc = fail(SEPREF_ERR_CUDA, "tc::launch_gcfn_trio failed: %s", tc::last_error()); return; }
        c.after("tc::k_gcfn");
        return;
      }
      if (kind == tc::KIND_F16 && c.h->gcfn_pair && g.pair.ready) {
        if (tc::launch_gcfn_pair(g.pair, g.tc, x, y, N, T, c.h->sm_count, c.st)) { c.rc = fail(SEPREF_ERR_CUDA, "tc::launch_gcfn_pair failed: %s", tc::last_error()); return; }
        c.after("tc::k_gcfn");
        return;
      }
      if (kind == tc::KIND_F16 && c.h->gcfn_tm && g.tm.ready) {
        if (tc::launch_gcfn_tm(g.tm, g.tc, x, y, N, T, c.h->sm_count, c.st, c.h->gcfn_tm)) { c.rc = fail(SEPREF_ERR_CUDA, "tc::launch_gcfn_tm failed: %s", tc::last_error()); return; }
        c.after("tc::k_gcfn");
        return;
      }
      int rc = tc::launch_gcfn(g.tc, x, y, N, T, F, c.h->sm_count, c.st, nullptr, nullptr, c.h->cluster, kind, c.h->gcfn_wide != 0);
      if (rc) { c.rc = fail(SEPREF_ERR_CUDA, "tc::launch_gcfn failed: %s", tc::last_error()); return; }
      c.after("tc::k_gcfn");
    }
    return;
  }
  const size_t mark = c.ws.off;
  float* ln = c.ws.f32(rows * F);
  float* hbuf = c.ws.f32(rows * 6 * F);
  float* u = c.ws.f32(rows * 3 * F);
  pool_layernorm(c, x, ln, rows, 1);
  gemm(c, simt::EPI_BIAS, ln, F, g.w1, g.b1, hbuf, 6 * F, rows, 6 * F, F);
  if (!c.dry() && c.ok()) {
    simt::k_dw3_glu<<<cdiv(rows * (3 * F / 4), 256), 256, 0, c.st>>>(hbuf, g.dw, g.dwb, u, (int)rows, T, 3 * F);
    c.after("k_dw3_glu");
  }
  gemm(c, simt::EPI_RES, u, 3 * F, g.w2, g.b2, y, F, rows, F, 3 * F, x, F);
  c.ws.off = mark;
}

// CLA.forward (network.py:174-187)
static void run_cla(Ctx& c, const ClaW& w, const float* x, float* y, int N, int T) {
  const int F = c.h->cfg.feat;
  const size_t rows = (size_t)N * T;
  const size_t mark = c.ws.off;
  if (c.h->gemm_path == 2 && c.h->cla_fused && F == tc::ClaFrontTraits::F && w.f16_ok_b && c.h->cfg.cla_kernel == tc::ClaFrontTraits::KW) {
    // LN + GEMM1 + GLU + depthwise k=65 in one kernel (d as FP16); GEMM2 + GELU + GEMM3 + residual -> y
    uint16_t* d16 = reinterpret_cast<uint16_t*>(c.ws.raw(rows * F * sizeof(uint16_t)));
    if (!c.dry() && c.ok()) {
      if (tc::launch_cla_front(w.t1, w.dw, w.dwb, x, d16, N, T, c.h->sm_count, c.st)) { c.rc = fail(SEPREF_ERR_CUDA, "tc::launch_cla_front failed: %s", tc::last_error()); return; }
      c.after("tc::k_cla_front");
    }
    tc::TokParams pb = tok_params(reinterpret_cast<const float*>(d16), y, F, w.t2, rows);
    pb.res = x;
    TOK_LAUNCH_K(tc::CfgClaB16, tc::KIND_F16, w.t2, &w.t3, pb, "tc::k_tok<cla_b>");
    c.ws.off = mark;
    return;
  }
  if (c.h->gemm_path >= 1) {     // LN+GEMM1+GLU -> u ; depthwise k=65 -> d ; GEMM2+GELU+GEMM3+residual -> y
    float* u = c.ws.f32(rows * F);
    float* d = c.ws.f32(rows * F);
    tc::TokParams pa = tok_params(x, u, F, w.t1, rows);
    TOK_LAUNCH(tc::CfgClaA, w.t1, nullptr, pa, "tc::k_tok<cla_a>");
    dwconv65(c, u, w.dw, w.dwb, d, N, T);
    tc::TokParams pb = tok_params(d, y, F, w.t2, rows);
    pb.res = x;
    TOK_LAUNCH_K(tc::CfgClaB, kind_of(c, w.f16_ok_b), w.t2, &w.t3, pb, "tc::k_tok<cla_b>");
    c.ws.off = mark;
    return;
  }
  float* ln = c.ws.f32(rows * F);
  float* hbuf = c.ws.f32(rows * 2 * F);
  float* u = c.ws.f32(rows * F);
  pool_layernorm(c, x, ln, rows, 1);
  gemm(c, simt::EPI_BIAS, ln, F, w.w1, w.b1, hbuf, 2 * F, rows, 2 * F, F);
  if (!c.dry() && c.ok()) {
    simt::k_glu<<<cdiv(rows * (F / 4), 256), 256, 0, c.st>>>(hbuf, u, rows, F);
    c.after("k_glu");
  }
  float* d = ln;   // LayerNorm output is dead: reuse it for the convolution result
  if (!c.dry() && c.ok()) {
    dim3 grid(cdiv(T, 16), N, F / 128);
    if (c.h->cfg.cla_kernel == 65) simt::k_dwconv_same<65, 16><<<grid, 128, 0, c.st>>>(u, w.dw, w.dwb, d, T, F);
    else { c.rc = fail(SEPREF_ERR_ARG, "CLA kernel size %d not built (only 65)", c.h->cfg.cla_kernel); return; }
    c.after("k_dwconv_same");
  }
  gemm(c, simt::EPI_GELU, d, F, w.w2, w.b2, hbuf, 2 * F, rows, 2 * F, F);
  gemm(c, simt::EPI_RES, hbuf, 2 * F, w.w3, w.b3, y, F, rows, F, 2 * F, x, F);
  c.ws.off = mark;
}

// EGA.forward (network.py:138-155) incl. the pooled MultiHeadAttention (network.py:90-124)
static void run_ega(Ctx& c, const EgaW& w, const float* x, float* y, int N, int T, int Td) {
  const int F = c.h->cfg.feat, H = c.h->cfg.heads, dk = F / H;
  const int r = T / Td;
  const size_t rows = (size_t)N * T, prow = (size_t)N * Td;
  const size_t mark = c.ws.off;
  float* qkv = c.ws.f32(prow * 3 * F);
  float* o = c.ws.f32(prow * F);
  float* a = c.ws.f32(prow * F);
  const bool tcp = c.h->gemm_path >= 1;
  float* z = tcp ? nullptr : c.ws.f32(prow * F);
  float* ln = tcp ? nullptr : c.ws.f32(rows * F);
  const bool h16 = c.h->gemm_path == 2;     // q|k|v rows are stored as FP16 between the projection and the attention
  if (tcp) {
    tc::TokParams pq = tok_params(x, qkv, 3 * F, w.att.tqkv, prow);
    pq.pool_r = r;
    if (r <= 64 && 64 % r == 0) TOK_LAUNCH(tc::CfgQkvPool16R, w.att.tqkv, nullptr, pq, "tc::k_tok<qkv_pool>");
    else TOK_LAUNCH(tc::CfgQkvPool16, w.att.tqkv, nullptr, pq, "tc::k_tok<qkv_pool>");
  } else {
    pool_layernorm(c, x, z, prow, r);
    gemm(c, simt::EPI_BIAS, z, F, w.att.wqkv, w.att.bqkv, qkv, 3 * F, prow, 3 * F, F);
  }
  if (!c.dry() && c.ok()) {
    dim3 grid(cdiv(Td, 64), H, N);
    if (dk == 16) {
      if (h16) attn::k_attn_relpos<16, true><<<grid, 128, sizeof(attn::AttnSmem<16>), c.st>>>(qkv, reinterpret_cast<const uint16_t*>(c.h->pe_k_h), o, Td, F, c.h->cfg.maxlen);
      else attn::k_attn_relpos<16, false><<<grid, 128, sizeof(attn::AttnSmem<16>), c.st>>>(qkv, reinterpret_cast<const uint16_t*>(c.h->pe_k_h), o, Td, F, c.h->cfg.maxlen);
    } else {
      if (h16) attn::k_attn_relpos<32, true><<<grid, 128, sizeof(attn::AttnSmem<32>), c.st>>>(qkv, reinterpret_cast<const uint16_t*>(c.h->pe_k_h), o, Td, F, c.h->cfg.maxlen);
      else attn::k_attn_relpos<32, false><<<grid, 128, sizeof(attn::AttnSmem<32>), c.st>>>(qkv, reinterpret_cast<const uint16_t*>(c.h->pe_k_h), o, Td, F, c.h->cfg.maxlen);
    }
    c.after("k_attn_relpos");
  }
  if (tcp) {
    tc::TokParams po = tok_params(o, a, F, w.att.to, prow);
    TOK_LAUNCH(tc::CfgProj, w.att.to, nullptr, po, "tc::k_tok<proj>");
    tc::TokParams pg = tok_params(x, y, F, w.tg, rows);
    pg.res = x; pg.up = a; pg.up_shift = log2i(r);
    TOK_LAUNCH(tc::CfgGate, w.tg, nullptr, pg, "tc::k_tok<gate>");
  } else {
    gemm(c, simt::EPI_BIAS, o, F, w.att.wo, w.att.bo, a, F, prow, F, F);
    pool_layernorm(c, x, ln, rows, 1);
    gemm(c, simt::EPI_GATE, ln, F, w.wg, w.bg, y, F, rows, F, F, x, F, a, r);
  }
  c.ws.off = mark;
}

// SpkAttention.forward (network.py:233-252); rows = B*S with S = 2
static void run_spk(Ctx& c, const SpkW& w, const float* x, float* y, int N, int T) {
  const int F = c.h->cfg.feat, H = c.h->cfg.heads, dk = F / H;
  const size_t rows = (size_t)N * T;
  const size_t mark = c.ws.off;
  float* ln = c.ws.f32(rows * F);
  float* qkv = c.ws.f32(rows * 3 * F);
  float* mid = c.ws.f32(rows * F);
  const bool tcp = c.h->gemm_path >= 1;
  // attention inside the out-projection's producer (one head per lane; fp32 q|k|v rows of dk = 32 would not fit its registers)
  const bool fused = tcp && H == 8 && (c.h->gemm_path == 2 || F == 128);
  if (fused) {
    tc::TokParams pq = tok_params(x, qkv, 3 * F, w.att.tqkv, rows);
    TOK_LAUNCH(tc::CfgQkv16, w.att.tqkv, nullptr, pq, "tc::k_tok<qkv>");
  } else if (tcp) {
    tc::TokParams pq = tok_params(x, qkv, 3 * F, w.att.tqkv, rows);
    TOK_LAUNCH(tc::CfgQkv, w.att.tqkv, nullptr, pq, "tc::k_tok<qkv>");
  } else {
    pool_layernorm(c, x, ln, rows, 1);
    gemm(c, simt::EPI_BIAS, ln, F, w.att.wqkv, w.att.bqkv, qkv, 3 * F, rows, 3 * F, F);
  }
  if (fused) {
    // the 2-token attention runs inside the out-projection kernel's operand producer (one head per lane)
    tc::TokParams po = tok_params(qkv, mid, F, w.att.to, rows);
    po.res = x; po.spk_T = T;
    TOK_LAUNCH(tc::CfgSpkProj, w.att.to, nullptr, po, "tc::k_tok<spk_proj>");
  } else {
    float* o = ln;
    if (!c.dry() && c.ok()) {
      const size_t n = (size_t)(N / 2) * T * H;
      if (dk == 16) simt::k_spk_attn2<16><<<cdiv(n, 256), 256, 0, c.st>>>(qkv, o, N / 2, T, F);
      else simt::k_spk_attn2<32><<<cdiv(n, 256), 256, 0, c.st>>>(qkv, o, N / 2, T, F);
      c.after("k_spk_attn2");
    }
    if (tcp) {
      tc::TokParams po = tok_params(o, mid, F, w.att.to, rows);
      po.res = x;
      TOK_LAUNCH(tc::CfgProjRes, w.att.to, nullptr, po, "tc::k_tok<proj_res>");
    } else {
      gemm(c, simt::EPI_RES, o, F, w.att.wo, w.att.bo, mid, F, rows, F, F, x, F);
    }
  }
  run_gcfn(c, *w.ff, mid, y, N, T);   // its scratch is bumped beyond `mid`
  c.ws.off = mark;
}

// DownConvLayer.forward (module.py:72-78)
static void run_down(Ctx& c, const DownW& w, const float* x, float* y, int N, int T) {
  if (c.dry() || !c.ok()) return;
  const int F = c.h->cfg.feat;
  simt::k_downconv_gelu<<<cdiv((size_t)N * (T / 2) * (F / 4), 256), 256, 0, c.st>>>(x, w.dw, w.b, y, N, T, F,
                                                                                  c.h->cfg.down_kernel);
  c.after("k_downconv_gelu");
}

// SpkSplitStage.forward (module.py:120-125): x [N,T,F] -> y [N*S,T,F]
static void run_split(Ctx& c, const SplitW& w, const float* x, float* y, int N, int T) {
  const int F = c.h->cfg.feat, S = c.h->cfg.num_spks;
  const size_t rows = (size_t)N * T;
  const size_t mark = c.ws.off;
  float* hbuf = c.ws.f32(rows * 4 * F * S);
  float* g = c.ws.f32(rows * 2 * F * S);
  float* h2 = c.ws.f32(rows * F * S);
  double* stats = reinterpret_cast<double*>(c.ws.raw(sizeof(double) * 2 * N * S));
  if (c.h->gemm_path >= 1) {
    tc::TokParams ps = tok_params(x, h2, F * S, w.ta, rows);
    // the producer rounds the RAW residual stream (no LayerNorm in front, module.py:113): no pack-time range bound
    TOK_LAUNCH_RAW(tc::CfgSplit, w.ta, &w.tb, ps, "tc::k_tok<split>");
  } else {
    gemm(c, simt::EPI_BIAS, x, F, w.wa, w.ba, hbuf, 4 * F * S, rows, 4 * F * S, F);
    if (!c.dry() && c.ok()) {
      simt::k_glu<<<cdiv(rows * (2 * F * S / 4), 256), 256, 0, c.st>>>(hbuf, g, rows, 2 * F * S);
      c.after("k_glu");
    }
    gemm(c, simt::EPI_BIAS, g, 2 * F * S, w.wb, w.bb, h2, F * S, rows, F * S, 2 * F * S);
  }
  if (!c.dry() && c.ok()) {
    cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * 2 * N * S, c.st);
    if (e != cudaSuccess) { c.rc = fail(SEPREF_ERR_CUDA, "memset: %s", cudaGetErrorString(e)); return; }
    const int rpb = 64;
    simt::k_gn_stats<<<dim3(cdiv(T, rpb), N * S), 256, 0, c.st>>>(h2, stats, T, F, S, rpb);
    c.after("k_gn_stats");
    simt::k_gn_apply_split<<<cdiv(rows * S * (F / 4), 256), 256, 0, c.st>>>(h2, stats, w.gamma, w.beta, y, N, T, F, S);
    c.after("k_gn_apply_split");
  }
  c.ws.off = mark;
}

// upsample + cat + simple_fusion (module.py:212-214): low [N,T/2,F], skip [N,T,F] -> y [N,T,F]
static void run_fuse(Ctx& c, const FuseW& w, const float* low, const float* skip, float* y, int N, int T) {
  const int F = c.h->cfg.feat;
  const size_t rows = (size_t)N * T;
  const size_t mark = c.ws.off;
  if (c.h->gemm_path >= 1) {
    tc::TokParams pf = tok_params(low, y, F, w.t, rows);
    pf.a1 = skip;
    TOK_LAUNCH_RAW(tc::CfgFuse, w.t, nullptr, pf, "tc::k_tok<fuse>");   // raw stream, as above
    return;
  }
  float* cat = c.ws.f32(rows * 2 * F);
  if (!c.dry() && c.ok()) {
    simt::k_concat_up<<<cdiv(rows * 2 * (F / 4), 256), 256, 0, c.st>>>(low, skip, cat, N, T, F);
    c.after("k_concat_up");
  }
  gemm(c, simt::EPI_BIAS, cat, 2 * F, w.w, w.b, y, F, rows, F, 2 * F);
  c.ws.off = mark;
}

template <class M>
static const typename M::mapped_type* find_block(const M& m, const std::string& key, Ctx& c, const char* kind) {
  auto it = m.find(key);
  if (it == m.end()) {
    c.rc = fail(SEPREF_ERR_ARG, "no %s weights under prefix '%s'", kind, key.c_str());
    return nullptr;
  }
  return &it->second;
}

static void run_global(Ctx& c, const std::string& p, const float* x, float* y, int N, int T, int Td) {
  const EgaW* e = find_block(c.h->ega, p + "block.ega.", c, "EGA");
  const GcfnW* g = find_block(c.h->gcfn, p + "block.gcfn.", c, "GCFN");
  if (!e || !g) return;
  const size_t mark = c.ws.off;
  float* mid = c.ws.f32((size_t)N * T * c.h->cfg.feat);
  run_ega(c, *e, x, mid, N, T, Td);
  run_gcfn(c, *g, mid, y, N, T);
  c.ws.off = mark;
}
static void run_local(Ctx& c, const std::string& p, const float* x, float* y, int N, int T) {
  const ClaW* l = find_block(c.h->cla, p + "block.cla.", c, "CLA");
  const GcfnW* g = find_block(c.h->gcfn, p + "block.gcfn.", c, "GCFN");
  if (!l || !g) return;
  const size_t mark = c.ws.off;
  float* mid = c.ws.f32((size_t)N * T * c.h->cfg.feat);
  run_cla(c, *l, x, mid, N, T);
  run_gcfn(c, *g, mid, y, N, T);
  c.ws.off = mark;
}

// Model-level callers (run_model) feed and drain the separator in its own channels-last layout instead of through the
// [B, F, T] boundary tensors: `fill` writes the padded features [B, Tp, F] straight into the first activation buffer,
// and the last activation / a dead buffer of the same size are handed back instead of being transposed out.
struct SepHooks {
  std::function<void(float* cur)> fill;      // produce [B, Tp, F] channels-last, padding rows zero
  float* last_ntc = nullptr;                 // out: final activation [B*S, Tp, F]
  float* scratch_ntc = nullptr;              // out: a dead buffer of the same size
};

// Separator.forward (module.py:190-218)
static void run_separator(Ctx& c, const float* x_in, int B, int t_enc, float* out_last, float* const* out_stages,
                          SepHooks* hooks = nullptr) {
  const sepref_config& cf = c.h->cfg;
  const int F = cf.feat, R = cf.num_stages, S = cf.num_spks;
  const int chunk = 1 << R;
  const int Tp = (t_enc % chunk == 0) ? t_enc : (t_enc / chunk + 1) * chunk;
  const int Td = Tp >> R;

  float* cur = c.ws.f32((size_t)B * Tp * F);
  float* tmp = c.ws.f32((size_t)B * Tp * F);
  if (hooks) {
    hooks->fill(cur);
  } else if (!c.dry() && c.ok()) {
    simt::k_nct_to_ntc_pad<<<dim3(cdiv(Tp, 32), F / 32, B), dim3(32, 8), 0, c.st>>>(x_in, cur, F, t_enc, Tp);
    c.after("k_nct_to_ntc_pad");
  }
  std::vector<float*> skips(R);
  auto split_prefix = [&](int idx) {
    return cf.per_stage_split ? "spk_split_blocks." + std::to_string(idx) + "." : std::string("spk_split_block.");
  };
  auto enc_stage = [&](const std::string& p, int T) {   // G, L, G, L (module.py:88-100); result ends in `cur`
    run_global(c, p + "g_block_1.", cur, tmp, B, T, Td);
    run_local(c, p + "l_block_1.", tmp, cur, B, T);
    run_global(c, p + "g_block_2.", cur, tmp, B, T, Td);
    run_local(c, p + "l_block_2.", tmp, cur, B, T);
  };
  for (int s = 0; s < R && c.ok(); ++s) {
    const int T = Tp >> s;
    const std::string p = "enc_stages." + std::to_string(s) + ".";
    enc_stage(p, T);
    skips[s] = c.ws.f32((size_t)B * S * T * F);
    if (const SplitW* w = find_block(c.h->split, split_prefix(s), c, "SpkSplit")) run_split(c, *w, cur, skips[s], B, T);
    if (const DownW* d = find_block(c.h->down, p + "downconv.", c, "DownConv")) run_down(c, *d, cur, tmp, B, T);
    std::swap(cur, tmp);      // activations now [B, T/2, F] inside the larger buffer
  }
  if (!c.ok()) return;
  enc_stage("bottleneck_G.", Td);
  float* xd = c.ws.f32((size_t)B * S * Tp * F);
  float* xt = c.ws.f32((size_t)B * S * Tp * F);
  if (const SplitW* w = find_block(c.h->split, split_prefix(R), c, "SpkSplit")) run_split(c, *w, cur, xd, B, Td);

  const int N2 = B * S;
  for (int i = 0; i < R && c.ok(); ++i) {
    const int Tl = Td << i, T = Tl * 2;
    if (out_stages && out_stages[i] && !c.dry()) {
      simt::k_ntc_to_nct<<<dim3(cdiv(Tl, 32), F / 32, N2), dim3(32, 8), 0, c.st>>>(xd, out_stages[i], F, Tl);
      c.after("k_ntc_to_nct");
    }
    const std::string p = "dec_stages." + std::to_string(i) + ".";
    if (const FuseW* w = find_block(c.h->fuse, "simple_fusion." + std::to_string(i) + ".", c, "fusion"))
      run_fuse(c, *w, xd, skips[R - 1 - i], xt, N2, T);
    std::swap(xd, xt);
    for (int n = 1; n <= 3 && c.ok(); ++n) {       // module.py:150-166
      const std::string sn = std::to_string(n) + ".";
      run_global(c, p + "g_block_" + sn, xd, xt, N2, T, Td);
      run_local(c, p + "l_block_" + sn, xt, xd, N2, T);
      if (const SpkW* w = find_block(c.h->spk, p + "spk_attn_" + sn, c, "SpkAttention")) run_spk(c, *w, xd, xt, N2, T);
      std::swap(xd, xt);
    }
  }
  if (hooks) {
    hooks->last_ntc = xd;
    hooks->scratch_ntc = xt;
  } else if (!c.dry() && c.ok()) {
    simt::k_ntc_to_nct<<<dim3(cdiv(Tp, 32), F / 32, N2), dim3(32, 8), 0, c.st>>>(xd, out_last, F, Tp);
    c.after("k_ntc_to_nct");
  }
}

// Model.forward without the auxiliary heads (model.py:38-45): mixture [B, n] -> audio [S, B, (T-1)*4 + 16]
static void run_model(Ctx& c, const float* mix, int B, int n, float* audio, float* const* out_stages) {
  const sepref_config& cf = c.h->cfg;
  const int F = cf.feat, S = cf.num_spks;
  const ShellW& sh = c.h->shell;
  const int T = (n - shell::kEncK) / shell::kEncS + 1;
  const int chunk = 1 << cf.num_stages;
  const int Tp = (T % chunk == 0) ? T : (T / chunk + 1) * chunk;
  double* stats = reinterpret_cast<double*>(c.ws.raw(sizeof(double) * 2 * B));
  float* z = c.ws.f32((size_t)B * Tp * shell::kEncC);
  SepHooks hooks;
  hooks.fill = [&](float* cur) {
    if (!c.dry() && c.ok()) {
      constexpr int FR = 64;
      cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B, c.st);
      if (e != cudaSuccess) { c.rc = fail(SEPREF_ERR_CUDA, "memset: %s", cudaGetErrorString(e)); return; }
      shell::k_enc_stats<FR><<<dim3(cdiv(T, FR), B), shell::kEncC, 0, c.st>>>(mix, sh.enc_w, stats, n, T);
      c.after("k_enc_stats");
      shell::k_enc_norm<FR><<<dim3(cdiv(Tp, FR), B), shell::kEncC, 0, c.st>>>(mix, sh.enc_w, stats, sh.gn_g, sh.gn_b, z, n, T, Tp);
      c.after("k_enc_norm");
    }
    // |z| <= sqrt(256 T) by construction of the normalisation: FP16 operands are range-safe here
    tc::TokParams pp = tok_params(z, cur, F, sh.proj, (size_t)B * Tp);
    if (c.h->gemm_path >= 1) TOK_LAUNCH_K(tc::CfgEncProj, kind_of(c, true), sh.proj, nullptr, pp, "tc::k_tok<enc_proj>");
  };
  run_separator(c, nullptr, B, T, nullptr, out_stages, &hooks);
  if (!c.ok()) return;
  // OutputLayer + AudioDecoder on the channels-last activation (the crop to T frames, module.py:251, happens in the
  // overlap-add: padding frames are computed and ignored).  Its input is the raw residual stream: TF32 unless RAW_F16.
  float* frames = hooks.scratch_ntc;            // [B*S*Tp, 128], 16 valid floats per row
  tc::TokParams po = tok_params(hooks.last_ntc, frames, 128, sh.out1, (size_t)B * S * Tp);
  po.out_ch = shell::kEncK;
  if (c.h->gemm_path >= 1) TOK_LAUNCH_RAW(tc::CfgOutDec, sh.out1, &sh.out2, po, "tc::k_tok<out_dec>");
  if (!c.dry() && c.ok()) {
    const int n_out = (T - 1) * shell::kEncS + shell::kEncK;
    shell::k_overlap_add<<<cdiv((size_t)B * S * n_out, 256), 256, 0, c.st>>>(frames, audio, B, S, T, Tp, 128, n_out);
    c.after("k_overlap_add");
  }
}

static void drop_graphs(sepref_handle* h) {
  for (auto& g : h->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  h->graphs.clear();
}

// Run `enqueue(stream)` (which launches a whole forward) either directly or through a cached CUDA graph.  First sight
// of a key runs eagerly (lazy per-device initialisation - function attributes, occupancy queries - happens there), the
// second captures on an internal stream (the caller's may be the legacy default stream, which cannot be captured) and
// replays; later calls only replay.  Any capture problem leaves the key on the eager path.
template <class Fn>
static int run_graphed(sepref_handle* h, const std::array<uintptr_t, 20>& key, cudaStream_t st, Fn&& enqueue) {
  if (!h->use_graphs || h->profile || h->debug_sync || h->dbg_clk) return enqueue(st);
  sepref_handle::GraphEntry* ent = nullptr;
  for (auto& g : h->graphs)
    if (g.key == key) { ent = &g; break; }
  if (ent && ent->exec) {
    ent->stamp = ++h->graph_clock;
    cudaError_t e = cudaGraphLaunch(ent->exec, st);
    if (e != cudaSuccess) return fail(SEPREF_ERR_CUDA, "cudaGraphLaunch: %s", cudaGetErrorString(e));
    h->launches = ent->launches;
    ++h->graph_replays;
    return 0;
  }
  if (!ent) {
    if (h->graphs.size() >= 8) {       // evict the least recently used entry
      size_t lru = 0;
      for (size_t i = 1; i < h->graphs.size(); ++i)
        if (h->graphs[i].stamp < h->graphs[lru].stamp) lru = i;
      if (h->graphs[lru].exec) cudaGraphExecDestroy(h->graphs[lru].exec);
      h->graphs.erase(h->graphs.begin() + (long)lru);
    }
    sepref_handle::GraphEntry g;
    g.key = key;
    g.stamp = ++h->graph_clock;
    h->graphs.push_back(g);
    return enqueue(st);
  }
  if (ent->bad) return enqueue(st);
  ent->stamp = ++h->graph_clock;
  if (!h->s_cap && cudaStreamCreateWithFlags(&h->s_cap, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); ent->bad = true; return enqueue(st); }
  if (cudaStreamBeginCapture(h->s_cap, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); ent->bad = true; return enqueue(st); }
  const int rc = enqueue(h->s_cap);
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(h->s_cap, &graph);
  // (`ent` may dangle if enqueue touched h->graphs - it does not: nested calls use other entry points)
  if (rc != 0 || e != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    ent->bad = true;
    return rc != 0 ? rc : enqueue(st);
  }
  e = cudaGraphInstantiate(&ent->exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { cudaGetLastError(); ent->exec = nullptr; ent->bad = true; return enqueue(st); }
  ent->launches = h->launches;
  e = cudaGraphLaunch(ent->exec, st);
  if (e != cudaSuccess) return fail(SEPREF_ERR_CUDA, "cudaGraphLaunch: %s", cudaGetErrorString(e));
  ++h->graph_replays;
  return 0;
}

static std::array<uintptr_t, 20> graph_key(const sepref_handle* h, int mode, const void* in, const void* out, float* const* stages,
                                           const void* ws, size_t ws_bytes, int batch, int len) {
  std::array<uintptr_t, 20> k{};
  k[0] = (uintptr_t)mode; k[1] = (uintptr_t)in; k[2] = (uintptr_t)out; k[3] = (uintptr_t)ws; k[4] = (uintptr_t)ws_bytes;
  k[5] = (uintptr_t)batch; k[6] = (uintptr_t)len;
  k[7] = (uintptr_t)((h->gemm_path) | (h->cluster << 4) | (h->gcfn_wide << 8) | (h->gcfn_pair << 9) | (h->raw_f16 << 10) | (h->gcfn_trio << 11) | (h->cla_fused << 12) | (h->gcfn_tm << 13));
  for (int i = 0; i < h->cfg.num_stages && i < 8; ++i) k[8 + i] = stages ? (uintptr_t)stages[i] : 0;
  return k;
}

static int check_ready(sepref_handle* h) {
  if (!h) return fail(SEPREF_ERR_ARG, "null handle");
  if (!h->finalized) return fail(SEPREF_ERR_STATE, "sepref_finalize has not been called (or parameters changed since)");
  return 0;
}

static int check_device_ptr(const void* p, const char* name) {
  if (!p) return fail(SEPREF_ERR_ARG, "%s is NULL", name);
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(SEPREF_ERR_CUDA, "cudaPointerGetAttributes(%s): %s", name, cudaGetErrorString(e)); }
  if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged)
    return fail(SEPREF_ERR_ARG, "%s is not a CUDA device pointer (no CPU fallback exists)", name);
  return 0;
}

}  // namespace sepref

// ================================================================================================ C ABI
extern "C" {

const char* sepref_last_error(void) { return g_err; }
const char* sepref_version(void) { return "sepref-b200 0.1 (sm_100a)"; }

int sepref_create(const sepref_config* cfg, int device, sepref_handle** out) {
  if (!cfg || !out) return fail(SEPREF_ERR_ARG, "null argument");
  if (cfg->feat != 128 && cfg->feat != 256) return fail(SEPREF_ERR_ARG, "feat=%d unsupported (kernels are built for 128 and 256)", cfg->feat);
  if (cfg->heads <= 0 || cfg->feat % cfg->heads || (cfg->feat / cfg->heads != 16 && cfg->feat / cfg->heads != 32))
    return fail(SEPREF_ERR_ARG, "head width %d unsupported (16 or 32)", cfg->heads > 0 ? cfg->feat / cfg->heads : -1);
  if (cfg->num_spks != 2) return fail(SEPREF_ERR_ARG, "num_spks=%d unsupported (2)", cfg->num_spks);
  if (cfg->num_stages < 1 || cfg->num_stages > 6) return fail(SEPREF_ERR_ARG, "num_stages=%d unsupported", cfg->num_stages);
  if (cfg->cla_kernel != 65) return fail(SEPREF_ERR_ARG, "cla_kernel=%d unsupported (65)", cfg->cla_kernel);
  if (cfg->down_kernel < 1 || cfg->down_kernel % 2 == 0) return fail(SEPREF_ERR_ARG, "down_kernel must be odd");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(SEPREF_ERR_CUDA, "no CUDA device available (%s); this library has no CPU path", cudaGetErrorString(e));
  }
  if (device < 0 || device >= ndev) return fail(SEPREF_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(SEPREF_ERR_CUDA, "device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
  sepref_handle* h = new sepref_handle();
  h->cfg = *cfg;
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  build_expected(h);
  *out = h;
  return 0;
}

void sepref_destroy(sepref_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  drop_graphs(h);
  if (h->trio.scratch) cudaFree(h->trio.scratch);
  if (h->range_flags) cudaFree(h->range_flags);
  if (h->s_cap) cudaStreamDestroy(h->s_cap);
  if (h->slab) cudaFree(h->slab);
  if (h->arena) cudaFree(h->arena);
  for (cudaEvent_t ev : h->prof_events) cudaEventDestroy(ev);
  for (cudaEvent_t ev : h->ev_in) cudaEventDestroy(ev);
  for (cudaEvent_t ev : h->ev_done) cudaEventDestroy(ev);
  for (auto& sl : h->slots) {
    if (sl.pending && sl.ev_out) cudaEventSynchronize(sl.ev_out);
    if (sl.arena) cudaFree(sl.arena);
    if (sl.ev_in) cudaEventDestroy(sl.ev_in);
    if (sl.ev_cmp) cudaEventDestroy(sl.ev_cmp);
    if (sl.ev_out) cudaEventDestroy(sl.ev_out);
  }
  if (h->s_cmp) cudaStreamDestroy(h->s_cmp);
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  delete h;
}

int sepref_set_option(sepref_handle* h, int option, int value) {
  if (!h) return fail(SEPREF_ERR_ARG, "null handle");
  switch (option) {
    case SEPREF_OPT_GEMM_PATH:
      if (value < 0 || value > 2) return fail(SEPREF_ERR_ARG, "gemm path must be 0, 1 or 2");
      h->gemm_path = value;
      return 0;
    case SEPREF_OPT_DEBUG_SYNC: h->debug_sync = value ? 1 : 0; return 0;
    case SEPREF_OPT_PROFILE: h->profile = value ? 1 : 0; return 0;
    case SEPREF_OPT_GCFN_WIDE: h->gcfn_wide = value; return 0;
    case SEPREF_OPT_RAW_F16: h->raw_f16 = value ? 1 : 0; return 0;
    case SEPREF_OPT_GCFN_PAIR: h->gcfn_pair = value ? 1 : 0; return 0;
    case SEPREF_OPT_GCFN_TRIO: h->gcfn_trio = value ? 1 : 0; return 0;
    case SEPREF_OPT_GCFN_TM: h->gcfn_tm = (value >= 0 && value <= 5) ? value : 0; return 0;
    case SEPREF_OPT_CLA_FUSED: h->cla_fused = value ? 1 : 0; h->ws_cache.clear(); return 0;
    case SEPREF_OPT_CUDA_GRAPH: h->use_graphs = value ? 1 : 0; if (!value) drop_graphs(h); return 0;
    case 99: h->dbg_flags = value; return 0;      // tuning experiments (kernels_tc.cuh TokParams::dbg_flags)
    case SEPREF_OPT_HOST_CHUNK:
      if (value < 1) return fail(SEPREF_ERR_ARG, "host chunk must be >= 1");
      h->host_chunk = value;
      return 0;
    case SEPREF_OPT_CLUSTER:
      if (value != 1 && value != 2 && value != 4) return fail(SEPREF_ERR_ARG, "cluster size must be 1, 2 or 4");
      h->cluster = value;
      return 0;
    default: return fail(SEPREF_ERR_ARG, "unknown option %d", option);
  }
}

int sepref_set_param(sepref_handle* h, const char* key, const float* data, const int64_t* shape, int ndim) {
  if (!h || !key || !data || (!shape && ndim > 0)) return fail(SEPREF_ERR_ARG, "null argument");
  const std::string k(key);
  const std::string tail = "num_batches_tracked";
  if (k.size() >= tail.size() && k.compare(k.size() - tail.size(), tail.size(), tail) == 0) return 0;
  auto it = h->params.find(k);
  if (it == h->params.end()) return fail(SEPREF_ERR_ARG, "unknown parameter key '%s'", key);
  HostT& t = it->second;
  if ((int)t.shape.size() != ndim) return fail(SEPREF_ERR_ARG, "'%s': expected %zu dims, got %d", key, t.shape.size(), ndim);
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    if (shape[i] != t.shape[i]) return fail(SEPREF_ERR_ARG, "'%s': dim %d is %lld, expected %lld", key, i, (long long)shape[i], (long long)t.shape[i]);
    n *= (size_t)shape[i];
  }
  t.v.assign(data, data + n);
  t.set = true;
  h->finalized = false;
  return 0;
}

int sepref_missing_params(sepref_handle* h, const char** first_missing) {
  if (!h) return fail(SEPREF_ERR_ARG, "null handle");
  int n = 0;
  h->missing_key.clear();
  for (auto& kv : h->params)
    if (!kv.second.set && !kv.second.optional) {
      if (n == 0) h->missing_key = kv.first;
      ++n;
    }
  if (first_missing) *first_missing = n ? h->missing_key.c_str() : nullptr;
  return n;
}

int sepref_finalize(sepref_handle* h) {
  if (!h) return fail(SEPREF_ERR_ARG, "null handle");
  const char* miss = nullptr;
  int n = sepref_missing_params(h, &miss);
  if (n) return fail(SEPREF_ERR_STATE, "%d parameters not set; first missing: %s", n, miss);
  CU_TRY(cudaSetDevice(h->device));
  h->gcfn.clear(); h->ega.clear(); h->cla.clear(); h->spk.clear(); h->down.clear(); h->split.clear(); h->fuse.clear();
  h->f16_fallbacks = 0; h->attn_bound = 0.0; h->ws_cache.clear();
  drop_graphs(h);                      // captured launches hold pointers into the weight slab that is about to be replaced
  h->graph_replays = 0;
  Packer pk{h};
  // discover blocks from the key set
  for (auto& kv : h->params) {
    const std::string& k = kv.first;
    auto ends = [&](const char* s) { size_t l = strlen(s); return k.size() >= l && k.compare(k.size() - l, l, s) == 0; };
    auto pre = [&](const char* s) { return k.substr(0, k.size() - strlen(s)); };
    if (ends("net1.1.weight")) h->gcfn[pre("net1.1.weight")];
    else if (ends("block.self_attn.linear_q.weight")) h->ega[pre("block.self_attn.linear_q.weight")];
    else if (ends("linear3.1.weight")) h->cla[pre("linear3.1.weight")];
    else if (ends("self_attn.linear_q.weight") && !ends("block.self_attn.linear_q.weight")) h->spk[pre("self_attn.linear_q.weight")];
    else if (ends("down_conv.weight")) h->down[pre("down_conv.weight")];
    else if (ends("linear.2.weight")) h->split[pre("linear.2.weight")];
    else if (k.compare(0, 14, "simple_fusion.") == 0 && ends("weight")) h->fuse[pre("weight")];
  }
  for (auto& kv : h->gcfn) pack_gcfn(pk, kv.first, kv.second);
  for (auto& kv : h->ega) pack_ega(pk, kv.first, kv.second);
  for (auto& kv : h->cla) pack_cla(pk, kv.first, kv.second);
  for (auto& kv : h->spk) pack_mha(pk, kv.first + "self_attn.", kv.second.att);
  for (auto& kv : h->down) pack_down(pk, kv.first, kv.second);
  for (auto& kv : h->split) pack_split(pk, kv.first, kv.second);
  for (auto& kv : h->fuse) {
    pk.put(&kv.second.w, pk.P(kv.first + "weight")); pk.put(&kv.second.b, pk.P(kv.first + "bias"));
    pack_tc_lin(pk, kv.second.t, pk.P(kv.first + "weight"), pk.P(kv.first + "bias"), h->cfg.feat, 2 * h->cfg.feat, 0);
  }
  pk.put(&h->pe_k, pk.P("pos_emb.pe_k.weight"));
  {
    const auto& pe = pk.P("pos_emb.pe_k.weight");
    std::vector<uint16_t> ph(pe.size());
    for (size_t i = 0; i < pe.size(); ++i) {
      const float v = std::fmax(-65504.f, std::fmin(65504.f, pe[i]));      // saturating, as cvt.rn.satfinite
      const __half hv = __float2half_rn(v);
      memcpy(&ph[i], &hv, 2);
    }
    pk.put_half(&h->pe_k_h, ph);
  }
  h->shell = ShellW();
  bool have_shell = true;
  for (const char* k : kShellKeys) have_shell = have_shell && h->params.at(k).set;
  if (have_shell) pack_shell(pk, h->shell);
  {   // k_attn_relpos multiplies FP16 operands on every path (q, k, v rows and the relative-position table)
    double pe = 0.0;
    for (float v : pk.P("pos_emb.pe_k.weight")) pe = std::fmax(pe, std::isfinite(v) ? std::fabs((double)v) : 1e300);
    if (h->attn_bound >= kF16Safe || pe >= kF16Safe)
      return fail(SEPREF_ERR_RANGE, "attention operands can exceed the FP16 range (q/k/v bound %.3g, |pe_k| max %.3g, limit %.3g): "
                  "the attention kernel has FP16 operands only and would saturate", h->attn_bound, pe, kF16Safe);
  }
  if (h->slab) { cudaFree(h->slab); h->slab = nullptr; }
  h->slab_floats = pk.host.size();
  CU_TRY(cudaMalloc(&h->slab, h->slab_floats * sizeof(float)));
  CU_TRY(cudaMemcpy(h->slab, pk.host.data(), h->slab_floats * sizeof(float), cudaMemcpyHostToDevice));
  for (auto& f : pk.fix) *f.first = h->slab + f.second;
  // SpkAttention's feed-forward GCFN is packed under its own prefix; link it
  for (auto& kv : h->spk) kv.second.ff = &h->gcfn.at(kv.first + "feed_forward.");
  // opt-in shared memory sizes
  CU_TRY(cudaFuncSetAttribute(attn::k_attn_relpos<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(attn::AttnSmem<16>)));
  CU_TRY(cudaFuncSetAttribute(attn::k_attn_relpos<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(attn::AttnSmem<32>)));
  CU_TRY(cudaFuncSetAttribute(attn::k_attn_relpos<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(attn::AttnSmem<16>)));
  CU_TRY(cudaFuncSetAttribute(attn::k_attn_relpos<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(attn::AttnSmem<32>)));
  int rc = tc::init(h->cfg.feat);
  if (rc) return fail(SEPREF_ERR_CUDA, "tensor-core kernel setup failed: %s", tc::last_error());
  for (auto& kv : h->gcfn) {
    rc |= tc::prepare_gcfn(kv.second.tc, h->cfg.feat);
    if (h->cfg.feat == tc::PairTraits::F) rc |= tc::prepare_gcfn_pair(kv.second.pair);
    if (h->cfg.feat == tc::TmTraits<true>::F) rc |= tc::prepare_gcfn_tm(kv.second.tm);
  }
  for (auto& kv : h->ega) rc |= tc::prepare_lin(kv.second.att.tqkv) | tc::prepare_lin(kv.second.att.to) | tc::prepare_lin(kv.second.tg);
  for (auto& kv : h->cla) rc |= tc::prepare_lin(kv.second.t1) | tc::prepare_lin(kv.second.t2) | tc::prepare_lin(kv.second.t3);
  for (auto& kv : h->spk) rc |= tc::prepare_lin(kv.second.att.tqkv) | tc::prepare_lin(kv.second.att.to);
  for (auto& kv : h->split) rc |= tc::prepare_lin(kv.second.ta) | tc::prepare_lin(kv.second.tb);
  for (auto& kv : h->fuse) rc |= tc::prepare_lin(kv.second.t);
  if (h->cfg.feat == tc::TrioTraits::F) rc |= tc::prepare_gcfn_trio(h->trio, h->sm_count);
  if (h->range_flags == nullptr) {
    if (cudaMalloc(&h->range_flags, (kRangeSites + 1) * sizeof(int)) != cudaSuccess ||
        cudaMemset(h->range_flags, 0, (kRangeSites + 1) * sizeof(int)) != cudaSuccess)
      return fail(SEPREF_ERR_CUDA, "range flags: %s", cudaGetErrorString(cudaGetLastError()));
  }
  if (have_shell) {
    rc |= tc::prepare_lin(h->shell.proj) | tc::prepare_lin(h->shell.out1) | tc::prepare_lin(h->shell.out2);
    h->shell.ready = rc == 0;
  }
  if (rc) return fail(SEPREF_ERR_CUDA, "tensor map setup failed: %s", tc::last_error());
  h->finalized = true;
  return 0;
}

int sepref_padded_frames(const sepref_handle* h, int t_enc) {
  if (!h || t_enc <= 0) return fail(SEPREF_ERR_ARG, "bad argument");
  const int chunk = 1 << h->cfg.num_stages;
  return (t_enc % chunk == 0) ? t_enc : (t_enc / chunk + 1) * chunk;
}

size_t sepref_workspace_bytes(const sepref_handle* h, int batch, int t_enc) {
  if (!h || batch <= 0 || t_enc <= 0) { fail(SEPREF_ERR_ARG, "sepref_workspace_bytes: bad argument"); return 0; }
  if (!h->finalized) { fail(SEPREF_ERR_STATE, "sepref_workspace_bytes before sepref_finalize"); return 0; }
  sepref_handle* hm = const_cast<sepref_handle*>(h);
  const auto key = std::make_tuple(batch, t_enc, h->gemm_path >= 1 ? 1 : 0);
  auto it = hm->ws_cache.find(key);
  if (it != hm->ws_cache.end()) return it->second;
  Ctx c{hm, nullptr};
  c.ws.measure = true;                 // dry run of the launch schedule: sizes only
  run_separator(c, nullptr, batch, t_enc, nullptr, nullptr);
  if (c.rc) return 0;                  // a block is missing: the size would be truncated (g_err says which)
  const size_t need = c.ws.peak + 512;
  if (hm->ws_cache.size() > 64) hm->ws_cache.clear();
  hm->ws_cache[key] = need;
  return need;
}

int sepref_separator_forward(sepref_handle* h, const float* x, int batch, int t_enc, float* out_last,
                             float* const* out_stages, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_ready(h)) return rc;
  if (batch <= 0 || t_enc <= 0) return fail(SEPREF_ERR_ARG, "batch and t_enc must be positive");
  if (int rc = check_device_ptr(x, "x")) return rc;
  if (int rc = check_device_ptr(out_last, "out_last")) return rc;
  if (int rc = check_device_ptr(workspace, "workspace")) return rc;
  const int Tp = sepref_padded_frames(h, t_enc);
  if ((Tp >> h->cfg.num_stages) < 1) return fail(SEPREF_ERR_ARG, "sequence too short");
  const size_t need = sepref_workspace_bytes(h, batch, t_enc);
  if (need == 0) return SEPREF_ERR_STATE;
  if (workspace_bytes < need) return fail(SEPREF_ERR_WORKSPACE, "workspace %zu bytes < required %zu", workspace_bytes, need);
  CU_TRY(cudaSetDevice(h->device));
  Ctx c{h, (cudaStream_t)stream};
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  c.ws.base = base;
  c.ws.cap = workspace_bytes - (size_t)(base - reinterpret_cast<char*>(workspace));
  h->prof_used = 0;
  const size_t cap = c.ws.cap;
  return run_graphed(h, graph_key(h, 0, x, out_last, out_stages, workspace, workspace_bytes, batch, t_enc), (cudaStream_t)stream,
                     [&](cudaStream_t st) -> int {
    Ctx cc{h, st};
    cc.ws.base = base;
    cc.ws.cap = cap;
    h->launches = 0;
    if (h->profile) CU_TRY(cc.prof_mark("start"));
    run_separator(cc, x, batch, t_enc, out_last, out_stages);
    if (cc.rc == 0 && cc.ws.overflow) return fail(SEPREF_ERR_WORKSPACE, "workspace overflow: schedule needs %zu bytes", cc.ws.peak);
    return cc.rc;
  });
}

int sepref_separator_forward_host(sepref_handle* h, const float* x_host, int batch, int t_enc, float* out_last_host,
                                  float* const* out_stages_host, void* stream) {
  if (int rc = check_ready(h)) return rc;
  if (!x_host || !out_last_host || batch <= 0 || t_enc <= 0) return fail(SEPREF_ERR_ARG, "bad argument");
  CU_TRY(cudaSetDevice(h->device));
  const sepref_config& cf = h->cfg;
  const int F = cf.feat, R = cf.num_stages, S = cf.num_spks;
  const int Tp = sepref_padded_frames(h, t_enc), Td = Tp >> R;
  // Utterances are independent, so the batch is walked in sub-batches: the H2D copy of sub-batch i+1 and the D2H copy
  // of sub-batch i-1 run on two copy streams while sub-batch i computes (host buffers should be pinned for that).
  const int sub = batch < h->host_chunk ? batch : h->host_chunk;
  const int nsub = (batch + sub - 1) / sub;
  const size_t in_b = (size_t)batch * F * t_enc * 4, out_b = (size_t)batch * S * F * Tp * 4;
  size_t stage_b[8] = {0}, stage_tot = 0;
  for (int i = 0; i < R; ++i) {
    stage_b[i] = (out_stages_host && out_stages_host[i]) ? (size_t)batch * S * F * (Td << i) * 4 : 0;
    stage_tot += (stage_b[i] + 255) & ~size_t(255);
  }
  const size_t ws_b = sepref_workspace_bytes(h, sub, t_enc);
  const size_t total = ((in_b + 255) & ~size_t(255)) + ((out_b + 255) & ~size_t(255)) + stage_tot + ws_b + 1024;
  if (h->arena_bytes < total) {
    if (h->arena) cudaFree(h->arena);
    h->arena = nullptr; h->arena_bytes = 0;
    CU_TRY(cudaMalloc(&h->arena, total));
    h->arena_bytes = total;
  }
  if (!h->s_in) {
    CU_TRY(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
  }
  while ((int)h->ev_in.size() < nsub) {
    cudaEvent_t a, b;
    CU_TRY(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    h->ev_in.push_back(a); h->ev_done.push_back(b);
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* p = h->arena;
  float* d_in = reinterpret_cast<float*>(p); p += (in_b + 255) & ~size_t(255);
  float* d_out = reinterpret_cast<float*>(p); p += (out_b + 255) & ~size_t(255);
  float* d_stage[8] = {nullptr};
  for (int i = 0; i < R; ++i)
    if (stage_b[i]) { d_stage[i] = reinterpret_cast<float*>(p); p += (stage_b[i] + 255) & ~size_t(255); }
  const size_t ws_avail = h->arena_bytes - (size_t)(p - h->arena);
  int launches = 0;
  // the copy streams must not start before work already queued on the caller's stream that produced x_host's contents
  // has nothing to do with them (host memory): no dependency needed at entry.
  for (int i = 0; i < nsub; ++i) {
    const int b0 = i * sub, nb = (b0 + sub <= batch) ? sub : batch - b0;
    const size_t in_off = (size_t)b0 * F * t_enc, in_n = (size_t)nb * F * t_enc;
    CU_TRY(cudaMemcpyAsync(d_in + in_off, x_host + in_off, in_n * 4, cudaMemcpyHostToDevice, h->s_in));
    CU_TRY(cudaEventRecord(h->ev_in[i], h->s_in));
    CU_TRY(cudaStreamWaitEvent(st, h->ev_in[i], 0));
    float* stage_ptrs[8] = {nullptr};
    for (int k = 0; k < R; ++k)
      if (stage_b[k]) stage_ptrs[k] = d_stage[k] + (size_t)b0 * S * F * (Td << k);
    float* out_ptr = d_out + (size_t)b0 * S * F * Tp;
    int rc = sepref_separator_forward(h, d_in + in_off, nb, t_enc, out_ptr, stage_ptrs, p, ws_avail, st);
    if (rc) return rc;
    launches += h->launches;
    CU_TRY(cudaEventRecord(h->ev_done[i], st));
    CU_TRY(cudaStreamWaitEvent(h->s_out, h->ev_done[i], 0));
    CU_TRY(cudaMemcpyAsync(out_last_host + (size_t)b0 * S * F * Tp, out_ptr, (size_t)nb * S * F * Tp * 4, cudaMemcpyDeviceToHost, h->s_out));
    for (int k = 0; k < R; ++k)
      if (stage_b[k]) {
        const size_t off = (size_t)b0 * S * F * (Td << k), n = (size_t)nb * S * F * (Td << k);
        CU_TRY(cudaMemcpyAsync(out_stages_host[k] + off, d_stage[k] + off, n * 4, cudaMemcpyDeviceToHost, h->s_out));
      }
  }
  h->launches = launches;
  CU_TRY(cudaStreamSynchronize(h->s_out));
  CU_TRY(cudaStreamSynchronize(st));
  return 0;
}

int sepref_separator_wait_host(sepref_handle* h, int slot) {
  if (!h || slot < 0 || slot > 1) return fail(SEPREF_ERR_ARG, "bad slot");
  auto& sl = h->slots[slot];
  if (!sl.pending) return 0;
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaEventSynchronize(sl.ev_out));
  sl.pending = false;
  return 0;
}

int sepref_separator_submit_host(sepref_handle* h, int slot, const float* x_host, int batch, int t_enc,
                                 float* out_last_host, float* const* out_stages_host) {
  if (int rc = check_ready(h)) return rc;
  if (slot < 0 || slot > 1) return fail(SEPREF_ERR_ARG, "slot must be 0 or 1");
  if (!x_host || !out_last_host || batch <= 0 || t_enc <= 0) return fail(SEPREF_ERR_ARG, "bad argument");
  CU_TRY(cudaSetDevice(h->device));
  auto& sl = h->slots[slot];
  // a slot is reused only after its previous request has fully landed in host memory
  if (int rc = sepref_separator_wait_host(h, slot)) return rc;
  const sepref_config& cf = h->cfg;
  const int F = cf.feat, R = cf.num_stages, S = cf.num_spks;
  const int Tp = sepref_padded_frames(h, t_enc), Td = Tp >> R;
  auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
  const size_t in_b = (size_t)batch * F * t_enc * 4, out_b = (size_t)batch * S * F * Tp * 4;
  size_t stage_b[8] = {0}, stage_tot = 0;
  for (int i = 0; i < R; ++i) {
    stage_b[i] = (out_stages_host && out_stages_host[i]) ? (size_t)batch * S * F * (Td << i) * 4 : 0;
    stage_tot += up(stage_b[i]);
  }
  const size_t ws_b = sepref_workspace_bytes(h, batch, t_enc);
  const size_t total = up(in_b) + up(out_b) + stage_tot + ws_b + 1024;
  if (sl.bytes < total) {
    if (sl.arena) cudaFree(sl.arena);
    sl.arena = nullptr; sl.bytes = 0;
    CU_TRY(cudaMalloc(&sl.arena, total));
    sl.bytes = total;
  }
  if (!h->s_in) {
    CU_TRY(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
  }
  if (!h->s_cmp) CU_TRY(cudaStreamCreateWithFlags(&h->s_cmp, cudaStreamNonBlocking));
  if (!sl.ev_in) {
    CU_TRY(cudaEventCreateWithFlags(&sl.ev_in, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&sl.ev_cmp, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&sl.ev_out, cudaEventDisableTiming));
  }
  char* p = sl.arena;
  float* d_in = reinterpret_cast<float*>(p); p += up(in_b);
  float* d_out = reinterpret_cast<float*>(p); p += up(out_b);
  float* d_stage[8] = {nullptr};
  for (int i = 0; i < R; ++i)
    if (stage_b[i]) { d_stage[i] = reinterpret_cast<float*>(p); p += up(stage_b[i]); }
  const size_t ws_avail = sl.bytes - (size_t)(p - sl.arena);
  // H2D on the input copy stream (overlaps the other slot's kernels), kernels on the shared compute stream in
  // submission order, D2H on the output copy stream (overlaps the next request's kernels)
  CU_TRY(cudaMemcpyAsync(d_in, x_host, in_b, cudaMemcpyHostToDevice, h->s_in));
  CU_TRY(cudaEventRecord(sl.ev_in, h->s_in));
  CU_TRY(cudaStreamWaitEvent(h->s_cmp, sl.ev_in, 0));
  if (int rc = sepref_separator_forward(h, d_in, batch, t_enc, d_out, d_stage, p, ws_avail, h->s_cmp)) return rc;
  CU_TRY(cudaEventRecord(sl.ev_cmp, h->s_cmp));
  CU_TRY(cudaStreamWaitEvent(h->s_out, sl.ev_cmp, 0));
  CU_TRY(cudaMemcpyAsync(out_last_host, d_out, out_b, cudaMemcpyDeviceToHost, h->s_out));
  for (int k = 0; k < R; ++k)
    if (stage_b[k]) CU_TRY(cudaMemcpyAsync(out_stages_host[k], d_stage[k], stage_b[k], cudaMemcpyDeviceToHost, h->s_out));
  CU_TRY(cudaEventRecord(sl.ev_out, h->s_out));
  sl.pending = true;
  return 0;
}

// ---- model-level entry points (waveform in, waveforms out) -----------------------------------------------------------
static int check_model_ready(sepref_handle* h, int batch, int samples) {
  if (int rc = check_ready(h)) return rc;
  if (!h->shell.ready) return fail(SEPREF_ERR_STATE, "model-shell parameters (\"@audio_encoder...\", \"@out_layer...\", ...) were not all set before sepref_finalize");
  if (h->gemm_path < 1) return fail(SEPREF_ERR_ARG, "the model-level entry points exist on the tensor-core paths only (gemm_path 1 or 2)");
  if (batch <= 0 || samples < shell::kEncK) return fail(SEPREF_ERR_ARG, "batch must be positive and samples >= %d", shell::kEncK);
  const int T = (samples - shell::kEncK) / shell::kEncS + 1;
  if ((sepref_padded_frames(h, T) >> h->cfg.num_stages) < 1) return fail(SEPREF_ERR_ARG, "sequence too short");
  return 0;
}

int sepref_model_frames(const sepref_handle* h, int samples) {
  if (!h || samples < shell::kEncK) return fail(SEPREF_ERR_ARG, "bad argument");
  return (samples - shell::kEncK) / shell::kEncS + 1;
}
int sepref_model_output_samples(const sepref_handle* h, int samples) {
  if (!h || samples < shell::kEncK) return fail(SEPREF_ERR_ARG, "bad argument");
  return (sepref_model_frames(h, samples) - 1) * shell::kEncS + shell::kEncK;
}

size_t sepref_model_workspace_bytes(const sepref_handle* h, int batch, int samples) {
  if (!h || batch <= 0 || samples < shell::kEncK) { fail(SEPREF_ERR_ARG, "sepref_model_workspace_bytes: bad argument"); return 0; }
  if (!h->finalized || !h->shell.ready) { fail(SEPREF_ERR_STATE, "sepref_model_workspace_bytes: model shell not finalized"); return 0; }
  sepref_handle* hm = const_cast<sepref_handle*>(h);
  const auto key = std::make_tuple(batch, -samples, h->gemm_path >= 1 ? 1 : 0);
  auto it = hm->ws_cache.find(key);
  if (it != hm->ws_cache.end()) return it->second;
  Ctx c{hm, nullptr};
  c.ws.measure = true;
  run_model(c, nullptr, batch, samples, nullptr, nullptr);
  if (c.rc) return 0;
  const size_t need = c.ws.peak + 512;
  if (hm->ws_cache.size() > 64) hm->ws_cache.clear();
  hm->ws_cache[key] = need;
  return need;
}

int sepref_model_forward(sepref_handle* h, const float* mix, int batch, int samples, float* audio, float* const* out_stages,
                         void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_model_ready(h, batch, samples)) return rc;
  if (int rc = check_device_ptr(mix, "mix")) return rc;
  if (int rc = check_device_ptr(audio, "audio")) return rc;
  if (int rc = check_device_ptr(workspace, "workspace")) return rc;
  const size_t need = sepref_model_workspace_bytes(h, batch, samples);
  if (need == 0) return SEPREF_ERR_STATE;
  if (workspace_bytes < need) return fail(SEPREF_ERR_WORKSPACE, "workspace %zu bytes < required %zu", workspace_bytes, need);
  CU_TRY(cudaSetDevice(h->device));
  Ctx c{h, (cudaStream_t)stream};
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  c.ws.base = base;
  c.ws.cap = workspace_bytes - (size_t)(base - reinterpret_cast<char*>(workspace));
  h->prof_used = 0;
  const size_t cap = c.ws.cap;
  return run_graphed(h, graph_key(h, 1, mix, audio, out_stages, workspace, workspace_bytes, batch, samples), (cudaStream_t)stream,
                     [&](cudaStream_t st) -> int {
    Ctx cc{h, st};
    cc.ws.base = base;
    cc.ws.cap = cap;
    h->launches = 0;
    if (h->profile) CU_TRY(cc.prof_mark("start"));
    run_model(cc, mix, batch, samples, audio, out_stages);
    if (cc.rc == 0 && cc.ws.overflow) return fail(SEPREF_ERR_WORKSPACE, "workspace overflow: schedule needs %zu bytes", cc.ws.peak);
    return cc.rc;
  });
}

int sepref_model_submit_host(sepref_handle* h, int slot, const float* mix_host, int batch, int samples, float* audio_host) {
  if (int rc = check_model_ready(h, batch, samples)) return rc;
  if (slot < 0 || slot > 1) return fail(SEPREF_ERR_ARG, "slot must be 0 or 1");
  if (!mix_host || !audio_host) return fail(SEPREF_ERR_ARG, "null host buffer");
  CU_TRY(cudaSetDevice(h->device));
  auto& sl = h->slots[slot];
  if (int rc = sepref_separator_wait_host(h, slot)) return rc;       // the slot's previous request has fully landed
  auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
  const int n_out = sepref_model_output_samples(h, samples);
  const size_t in_b = (size_t)batch * samples * 4, out_b = (size_t)batch * h->cfg.num_spks * n_out * 4;
  const size_t ws_b = sepref_model_workspace_bytes(h, batch, samples);
  if (ws_b == 0) return SEPREF_ERR_STATE;
  const size_t total = up(in_b) + up(out_b) + ws_b + 1024;
  if (sl.bytes < total) {
    if (sl.arena) cudaFree(sl.arena);
    sl.arena = nullptr; sl.bytes = 0;
    CU_TRY(cudaMalloc(&sl.arena, total));
    sl.bytes = total;
  }
  if (!h->s_in) {
    CU_TRY(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
  }
  if (!h->s_cmp) CU_TRY(cudaStreamCreateWithFlags(&h->s_cmp, cudaStreamNonBlocking));
  if (!sl.ev_in) {
    CU_TRY(cudaEventCreateWithFlags(&sl.ev_in, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&sl.ev_cmp, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&sl.ev_out, cudaEventDisableTiming));
  }
  char* p = sl.arena;
  float* d_in = reinterpret_cast<float*>(p); p += up(in_b);
  float* d_out = reinterpret_cast<float*>(p); p += up(out_b);
  const size_t ws_avail = sl.bytes - (size_t)(p - sl.arena);
  CU_TRY(cudaMemcpyAsync(d_in, mix_host, in_b, cudaMemcpyHostToDevice, h->s_in));
  CU_TRY(cudaEventRecord(sl.ev_in, h->s_in));
  CU_TRY(cudaStreamWaitEvent(h->s_cmp, sl.ev_in, 0));
  if (int rc = sepref_model_forward(h, d_in, batch, samples, d_out, nullptr, p, ws_avail, h->s_cmp)) return rc;
  CU_TRY(cudaEventRecord(sl.ev_cmp, h->s_cmp));
  CU_TRY(cudaStreamWaitEvent(h->s_out, sl.ev_cmp, 0));
  CU_TRY(cudaMemcpyAsync(audio_host, d_out, out_b, cudaMemcpyDeviceToHost, h->s_out));
  CU_TRY(cudaEventRecord(sl.ev_out, h->s_out));
  sl.pending = true;
  return 0;
}
int sepref_model_wait_host(sepref_handle* h, int slot) { return sepref_separator_wait_host(h, slot); }

// ---- batched PIT SI-SNR improvement on the device (criterions.py:221-260) ---------------------------------------------
int sepref_pit_sisnri(sepref_handle* h, const float* est, const float* tgt, const float* mix, int batch, int n, int ld_est,
                      double eps, float* out, void* stream) {
  if (!h) return fail(SEPREF_ERR_ARG, "null handle");
  if (h->cfg.num_spks != 2) return fail(SEPREF_ERR_ARG, "two speakers only");
  if (batch <= 0 || n <= 0 || ld_est < n) return fail(SEPREF_ERR_ARG, "batch, n must be positive and ld_est >= n");
  if (int rc = check_device_ptr(est, "est")) return rc;
  if (int rc = check_device_ptr(tgt, "tgt")) return rc;
  if (int rc = check_device_ptr(mix, "mix")) return rc;
  if (int rc = check_device_ptr(out, "out")) return rc;
  CU_TRY(cudaSetDevice(h->device));
  shell::k_pit_sisnri<<<batch, 256, 0, (cudaStream_t)stream>>>(est, tgt, mix, out, batch, n, ld_est, eps);
  CU_TRY(cudaGetLastError());
  return 0;
}

int sepref_last_launch_count(const sepref_handle* h) { return h ? h->launches : 0; }
int sepref_f16_fallback_count(const sepref_handle* h) { return (h && h->finalized) ? h->f16_fallbacks : -1; }
long long sepref_range_rerun_count(sepref_handle* h) {
  if (!h || !h->finalized || !h->range_flags) return -1;
  int v = 0;
  if (cudaSetDevice(h->device) != cudaSuccess) return -1;
  if (cudaMemcpy(&v, h->range_flags + 32, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return v;
}
int sepref_graph_replay_count(const sepref_handle* h) { return h ? h->graph_replays : -1; }

int sepref_profile_report(sepref_handle* h, char* buf, size_t cap) {
  if (!h || !buf || cap == 0) return fail(SEPREF_ERR_ARG, "bad argument");
  buf[0] = 0;
  if (h->prof_used < 2) return 0;
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaEventSynchronize(h->prof_events[h->prof_used - 1]));
  std::map<std::string, std::pair<double, int>> acc;
  for (size_t i = 1; i < h->prof_used; ++i) {
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, h->prof_events[i - 1], h->prof_events[i]));
    auto& a = acc[h->prof_names[i]];
    a.first += ms;
    a.second += 1;
  }
  size_t off = 0;
  for (auto& kv : acc) {
    int n = snprintf(buf + off, cap - off, "%s %.6f %d\n", kv.first.c_str(), kv.second.first, kv.second.second);
    if (n < 0 || (size_t)n >= cap - off) break;
    off += (size_t)n;
  }
  return 0;
}

// ---- block-level entry points ------------------------------------------------------------------------------------
size_t sepref_block_workspace_bytes(const sepref_handle* h, int rows, int t) {
  if (!h || rows <= 0 || t <= 0) return 0;
  // generous: the SIMT GCFN needs 10F floats per token, SpkSplit 14F, plus a block-sized intermediate
  return ((size_t)rows * t * h->cfg.feat * 4) * 24 + (1 << 20);
}

#define BLOCK_PROLOGUE()                                                                          \
  if (int rc__ = check_ready(h)) return rc__;                                                     \
  if (!prefix) return fail(SEPREF_ERR_ARG, "null prefix");                                        \
  if (rows <= 0 || t <= 0) return fail(SEPREF_ERR_ARG, "rows and t must be positive");            \
  if (int rc__ = check_device_ptr(x, "x")) return rc__;                                           \
  if (int rc__ = check_device_ptr(y, "y")) return rc__;                                           \
  CU_TRY(cudaSetDevice(h->device));                                                               \
  Ctx c{h, (cudaStream_t)stream};                                                                 \
  h->launches = 0;

#define BLOCK_WS()                                                                                \
  if (int rc__ = check_device_ptr(workspace, "workspace")) return rc__;                           \
  if (workspace_bytes < sepref_block_workspace_bytes(h, rows, t))                                 \
    return fail(SEPREF_ERR_WORKSPACE, "block workspace too small");                               \
  c.ws.base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255)); \
  c.ws.cap = workspace_bytes - (size_t)(c.ws.base - reinterpret_cast<char*>(workspace));

#define BLOCK_RET() \
  return (c.rc == 0 && c.ws.overflow) ? fail(SEPREF_ERR_WORKSPACE, "block workspace overflow: needs %zu bytes", c.ws.peak) : c.rc

int sepref_gcfn_forward(sepref_handle* h, const char* prefix, const float* x, int rows, int t, float* y, void* workspace,
                        size_t workspace_bytes, void* stream) {
  BLOCK_PROLOGUE();
  BLOCK_WS();
  if (const GcfnW* w = find_block(h->gcfn, prefix, c, "GCFN")) run_gcfn(c, *w, x, y, rows, t);
  BLOCK_RET();
}
int sepref_debug_gcfn_h(sepref_handle* h, const char* prefix, const float* x, int rows, int t, float* y, float* h_out,
                        void* stream) {
  BLOCK_PROLOGUE();
  if (int rc = check_device_ptr(h_out, "h_out")) return rc;
  if (const GcfnW* w = find_block(h->gcfn, prefix, c, "GCFN")) {
    if (tc::launch_gcfn(w->tc, x, y, rows, t, h->cfg.feat, h->sm_count, c.st, h_out, nullptr, h->cluster, kind_of(c, w->f16_ok), h->gcfn_wide != 0)) return fail(SEPREF_ERR_CUDA, "%s", tc::last_error());
    c.after("tc::k_gcfn");
  }
  BLOCK_RET();
}
int sepref_debug_gcfn_timeline(sepref_handle* h, const char* prefix, const float* x, int rows, int t, float* y,
                               long long* clk_out, void* stream) {
  BLOCK_PROLOGUE();
  if (int rc = check_device_ptr(clk_out, "clk_out")) return rc;
  if (const GcfnW* w = find_block(h->gcfn, prefix, c, "GCFN")) {
    if (kind_of(c, w->f16_ok) == tc::KIND_F16 && h->gcfn_trio && h->trio.clusters > 0) {
      if (tc::launch_gcfn_trio(h->trio, w->tc, x, y, rows, t, h->sm_count, c.st, clk_out)) return fail(SEPREF_ERR_CUDA, "%s", tc::last_error());
      c.after("tc::k_gcfn");
      return c.rc;
    }
    if (kind_of(c, w->f16_ok) == tc::KIND_F16 && h->gcfn_pair && w->pair.ready) {
      if (tc::launch_gcfn_pair(w->pair, w->tc, x, y, rows, t, h->sm_count, c.st, clk_out)) return fail(SEPREF_ERR_CUDA, "%s", tc::last_error());
      c.after("tc::k_gcfn");
      return c.rc;
    }
    if (kind_of(c, w->f16_ok) == tc::KIND_F16 && h->gcfn_tm && w->tm.ready) {
      if (tc::launch_gcfn_tm(w->tm, w->tc, x, y, rows, t, h->sm_count, c.st, h->gcfn_tm, clk_out)) return fail(SEPREF_ERR_CUDA, "%s", tc::last_error());
      c.after("tc::k_gcfn");
      return c.rc;
    }
    if (tc::launch_gcfn(w->tc, x, y, rows, t, h->cfg.feat, h->sm_count, c.st, nullptr, clk_out, h->cluster, kind_of(c, w->f16_ok), h->gcfn_wide != 0)) return fail(SEPREF_ERR_CUDA, "%s", tc::last_error());
    c.after("tc::k_gcfn");
  }
  BLOCK_RET();
}
int sepref_debug_tok_timeline(sepref_handle* h, const char* kernel_tag, long long* clk_out) {
  if (!h) return fail(SEPREF_ERR_ARG, "null handle");
  h->dbg_clk = clk_out;
  snprintf(h->dbg_name, sizeof(h->dbg_name), "%s", kernel_tag ? kernel_tag : "");
  return 0;
}
int sepref_cla_forward(sepref_handle* h, const char* prefix, const float* x, int rows, int t, float* y, void* workspace,
                       size_t workspace_bytes, void* stream) {
  BLOCK_PROLOGUE();
  BLOCK_WS();
  if (const ClaW* w = find_block(h->cla, prefix, c, "CLA")) run_cla(c, *w, x, y, rows, t);
  BLOCK_RET();
}
int sepref_ega_forward(sepref_handle* h, const char* prefix, const float* x, int rows, int t, int td, float* y,
                       void* workspace, size_t workspace_bytes, void* stream) {
  BLOCK_PROLOGUE();
  BLOCK_WS();
  if (td <= 0 || t % td) return fail(SEPREF_ERR_ARG, "t must be a multiple of td");
  if (h->gemm_path >= 1 && ((t / td) & (t / td - 1)))
    return fail(SEPREF_ERR_ARG, "tensor-core paths need t / td to be a power of two (the gate indexes pooled rows by shift); got %d", t / td);
  if (const EgaW* w = find_block(h->ega, prefix, c, "EGA")) run_ega(c, *w, x, y, rows, t, td);
  BLOCK_RET();
}
int sepref_global_block_forward(sepref_handle* h, const char* prefix, const float* x, int rows, int t, int td, float* y,
                                void* workspace, size_t workspace_bytes, void* stream) {
  BLOCK_PROLOGUE();
  BLOCK_WS();
  if (td <= 0 || t % td) return fail(SEPREF_ERR_ARG, "t must be a multiple of td");
  if (h->gemm_path >= 1 && ((t / td) & (t / td - 1)))
    return fail(SEPREF_ERR_ARG, "tensor-core paths need t / td to be a power of two (the gate indexes pooled rows by shift); got %d", t / td);
  run_global(c, prefix, x, y, rows, t, td);
  BLOCK_RET();
}
int sepref_local_block_forward(sepref_handle* h, const char* prefix, const float* x, int rows, int t, float* y,
                               void* workspace, size_t workspace_bytes, void* stream) {
  BLOCK_PROLOGUE();
  BLOCK_WS();
  run_local(c, prefix, x, y, rows, t);
  BLOCK_RET();
}
int sepref_spk_attention_forward(sepref_handle* h, const char* prefix, const float* x, int rows, int t, float* y,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  BLOCK_PROLOGUE();
  BLOCK_WS();
  if (rows % h->cfg.num_spks) return fail(SEPREF_ERR_ARG, "rows must be a multiple of num_spks");
  if (const SpkW* w = find_block(h->spk, prefix, c, "SpkAttention")) run_spk(c, *w, x, y, rows, t);
  BLOCK_RET();
}
int sepref_down_conv_forward(sepref_handle* h, const char* prefix, const float* x, int rows, int t, float* y, void* stream) {
  BLOCK_PROLOGUE();
  if (t % 2) return fail(SEPREF_ERR_ARG, "t must be even");
  if (const DownW* w = find_block(h->down, prefix, c, "DownConv")) run_down(c, *w, x, y, rows, t);
  BLOCK_RET();
}
int sepref_spk_split_forward(sepref_handle* h, const char* prefix, const float* x, int rows, int t, float* y,
                             void* workspace, size_t workspace_bytes, void* stream) {
  BLOCK_PROLOGUE();
  BLOCK_WS();
  if (const SplitW* w = find_block(h->split, prefix, c, "SpkSplit")) run_split(c, *w, x, y, rows, t);
  BLOCK_RET();
}
int sepref_fusion_forward(sepref_handle* h, const char* prefix, const float* x_low, const float* skip, int rows, int t,
                          float* y, void* workspace, size_t workspace_bytes, void* stream) {
  const float* x = skip;
  BLOCK_PROLOGUE();
  BLOCK_WS();
  if (int rc = check_device_ptr(x_low, "x_low")) return rc;
  if (t % 2) return fail(SEPREF_ERR_ARG, "t must be even");
  if (const FuseW* w = find_block(h->fuse, prefix, c, "fusion")) run_fuse(c, *w, x_low, skip, y, rows, t);
  BLOCK_RET();
}

}  // extern "C"
