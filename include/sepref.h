/*
 * sepref.h - C ABI of the B200-native SepReformer separator (libsepref_b200.so).
 *
 * This is the drop-in boundary for the reference's separator hot path.  Every entry point names the
 * reference interface it replaces (paths relative to the reference repository,
 * models/SepReformer_Base_WSJ0/...).  Plain pointers and sizes only: no torch / C++ types cross it.
 *
 * Conventions
 *   - all tensors are fp32; "device" pointers are CUDA device pointers on the handle's device
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream)
 *   - every function returns 0 on success, <0 on error; sepref_last_error() describes the last failure
 *     of the calling thread.  Nothing throws across the ABI.  There is NO CPU fallback: a missing GPU,
 *     a host pointer where a device pointer is required, or an unsupported shape is an error.
 *   - the caller owns inputs, outputs and the workspace; the handle owns only the re-packed weights.
 *     No allocation and no host synchronisation happen inside the *_forward calls on device buffers.
 *   - one handle = one device and ONE forward in flight at a time: calls on the same handle must be ordered on one
 *     stream (or externally serialised).  The handle carries per-call device state - the range flags of the raw-stream
 *     GEMMs, the captured CUDA graph and its static buffers, the profiling events.  The host-pipelined entry points
 *     (sepref_*_submit_host) already order their two slots on one internal compute stream.  Use one handle per
 *     thread / stream for concurrent forwards (torch's data_parallel replicas get one handle per device).
 */
#ifndef SEPREF_H_
#define SEPREF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEPREF_OK 0
#define SEPREF_ERR_ARG (-1)        /* bad argument / unsupported configuration */
#define SEPREF_ERR_STATE (-2)      /* call order: parameters missing, not finalized, ... */
#define SEPREF_ERR_CUDA (-3)       /* CUDA runtime / driver error (no device, launch failure, ...) */
#define SEPREF_ERR_WORKSPACE (-4)  /* workspace too small */
#define SEPREF_ERR_RANGE (-5)      /* weights whose activations cannot be held in the operand format (see sepref_finalize) */

typedef struct sepref_handle sepref_handle;

/* The numbers Separator.__init__ receives through configs.yaml:46-83 (modules/module.py:39). */
typedef struct sepref_config {
  int32_t feat;            /* F: 128 (Base) or 256 (Large)               enc_stage.global_blocks.in_channels   */
  int32_t heads;           /* 8                                          ...num_mha_heads                      */
  int32_t num_stages;      /* R = 4                                      num_stages                            */
  int32_t num_spks;        /* 2                                          spk_split_stage.num_spks              */
  int32_t cla_kernel;      /* 65                                         local_blocks.kernel_size              */
  int32_t down_kernel;     /* 5                                          down_conv_layer.samp_kernel_size      */
  int32_t maxlen;          /* 2000                                       relative_positional_encoding.maxlen   */
  int32_t per_stage_split; /* 1 for SepReformer_Large_DM_WHAM (module.py:182-184 there), else 0                */
} sepref_config;

/* Options for sepref_set_option(). */
#define SEPREF_OPT_GEMM_PATH 1   /* 0 = fp32 CUDA-core kernels, 1 = tcgen05 kind::tf32, 2 = tcgen05 kind::f16 (fp16
                                  * operands: TF32's 11-bit significand at half the bytes; row-scaled weights)      */
#define SEPREF_OPT_DEBUG_SYNC 2  /* 1 = synchronise + check after every launch (debugging only; default 0)   */
#define SEPREF_OPT_CLUSTER 4     /* CTAs per cluster sharing TMA-multicast weight slabs: 1, 2 (default) or 4            */
#define SEPREF_OPT_GCFN_WIDE 6   /* 1: GCFN kernel with 160-frame tiles and single-buffered accumulators (f16 path, F = 128); 0 */
#define SEPREF_OPT_HOST_CHUNK 5  /* utterances per sub-batch of sepref_separator_forward_host (copy/compute overlap); 16 */
#define SEPREF_OPT_RAW_F16 7     /* gemm_path 2 only.  0 (default): the GEMMs fed by the un-normalised residual stream (SpkSplit, fusion
                                  * conv, output layer) run with FP16 operands, check the range at run time and are re-computed with
                                  * TF32 operands if it was exceeded (sepref_range_rerun_count); 1: FP16 only, no re-computation */
#define SEPREF_OPT_GCFN_PAIR 8   /* 1: GCFN blocks with FP16 operands and F = 128 run as k_gcfn_pair - weights resident in the shared
                                  * memory of a CTA pair, partial sums exchanged through DSMEM; 0 (default): the streaming kernel k_gcfn */
#define SEPREF_OPT_GCFN_TRIO 10  /* 1: GCFN blocks with FP16 operands and F = 128 run as k_gcfn_trio - weights resident in the shared memory of
                                  * a cluster of three CTAs (one 128-channel chunk each), partial sums exchanged through an L2 scratch */
#define SEPREF_OPT_CLA_FUSED 11  /* 1 (default): with FP16 operands and F = 128, CLA's LayerNorm + linear1 + GLU + depthwise k=65 conv run as
                                  * one kernel (k_cla_front) that hands d to the second half as FP16; 0: three kernels, fp32 intermediates */
#define SEPREF_OPT_GCFN_TM 12    /* 0 (default): k_gcfn (channels as the MMA M dimension).  1: k_gcfn_tm, frames as M and the weight slabs as the N
                                  * operand shared by a CTA pair (tcgen05.mma.cta_group::2); 2: the same on single CTAs.  FP16 operands, F = 128. */
#define SEPREF_OPT_CUDA_GRAPH 9  /* 1: sepref_separator_forward / sepref_model_forward capture their ~260 launches into a CUDA graph the
                                  * second time they see the same shapes, options AND buffer addresses, and replay it afterwards (one
                                  * cudaGraphLaunch instead of ~2 ms of launch calls).  Keep the buffers alive and at the same addresses
                                  * to benefit; up to 8 graphs are cached per handle, sepref_finalize drops them.  0 (default): eager */
#define SEPREF_OPT_PROFILE 3     /* 1 = record a CUDA event after every launch of sepref_separator_forward      */

const char* sepref_last_error(void);
const char* sepref_version(void);

/* Replaces Separator.__init__ (modules/module.py:39,172-188): creates an empty handle on CUDA `device`. */
int sepref_create(const sepref_config* cfg, int device, sepref_handle** out);
void sepref_destroy(sepref_handle* h);
int sepref_set_option(sepref_handle* h, int option, int value);

/* Replaces load_state_dict on the separator (utils/util_engine.py:43): hand over one tensor by its
 * state_dict key relative to the separator (e.g. "enc_stages.0.g_block_1.block.gcfn.net1.1.weight").
 * `data` is a HOST pointer to `prod(shape)` contiguous floats; it is copied.  Unknown keys are an error,
 * "....num_batches_tracked" keys are accepted and ignored. */
int sepref_set_param(sepref_handle* h, const char* key, const float* data, const int64_t* shape, int ndim);

/* Number of parameters still missing before sepref_finalize() can succeed; `first_missing` (optional)
 * receives a pointer to the first missing key (valid until the next call on this handle). */
int sepref_missing_params(sepref_handle* h, const char** first_missing);

/* Folds eval-mode BatchNorm, LayerNorm affine, LayerScale and 1/sqrt(dk) into the neighbouring linear
 * maps, rounds tensor-core operands to TF32 / FP16 (round-to-nearest), re-tiles them and uploads everything.
 * May be called again after further sepref_set_param calls.  SEPREF_ERR_RANGE: see sepref_f16_fallback_count. */
int sepref_finalize(sepref_handle* h);

/* Separator.pad_signal (modules/module.py:220-234): frames after right-padding to a multiple of 2^R
 * (unchanged when already a multiple). */
int sepref_padded_frames(const sepref_handle* h, int t_enc);

/* Bytes of device scratch sepref_separator_forward needs for a batch of `batch` utterances (cached per shape).
 * Returns 0 - and sets sepref_last_error() - before sepref_finalize or for bad arguments. */
size_t sepref_workspace_bytes(const sepref_handle* h, int batch, int t_enc);

/* Replaces Separator.forward (modules/module.py:190-218).
 *   x          device [batch, F, t_enc]                      (the FeatureProjector output, model.py:40)
 *   out_last   device [batch*num_spks, F, T_pad]             row index = b*num_spks + spk (module.py:123)
 *   out_stages R device pointers, stage i is [batch*num_spks, F, T_pad / 2^(R-i)]; NULL (or NULL entries)
 *              skips writing that auxiliary output (they only feed the training-time aux heads, model.py:47-51)
 */
int sepref_separator_forward(sepref_handle* h, const float* x, int batch, int t_enc, float* out_last,
                             float* const* out_stages, void* workspace, size_t workspace_bytes, void* stream);

/* Same call with HOST buffers (what engine.py:165-167 does through data_parallel: the mixture features
 * arrive from the host and the separated features go back): copies in, runs, copies out and synchronises
 * `stream`.  The batch is processed in sub-batches (utterances are independent) so that the H2D / D2H copies of
 * neighbouring sub-batches overlap the kernels (pin the host buffers for that).  Uses an internal device arena
 * that grows on demand (the only entry point that allocates). */
int sepref_separator_forward_host(sepref_handle* h, const float* x_host, int batch, int t_enc,
                                  float* out_last_host, float* const* out_stages_host, void* stream);

/* Pipelined form of the host-buffer call for serving loops (engine.py:165-167 iterates the test set one batch after
 * another): submit queues H2D copy -> kernels -> D2H copy for one batch in staging slot `slot` (0 or 1) and returns
 * without waiting; wait blocks until that slot's outputs are in host memory.  With two slots in flight the copies of
 * batch i+1 / i-1 overlap the kernels of batch i.  Host buffers must stay valid (and should be pinned) until wait
 * returns; submitting to a slot that is still pending waits for it first.  Kernels of all submissions run in order
 * on one internal stream. */
int sepref_separator_submit_host(sepref_handle* h, int slot, const float* x_host, int batch, int t_enc,
                                 float* out_last_host, float* const* out_stages_host);
int sepref_separator_wait_host(sepref_handle* h, int slot);

/* ---- model-level entry points: the layers either side of the separator ----------------------
 * Hand the shell tensors to sepref_set_param under "@" + their key in Model.state_dict() (model.py:24-29):
 *   "@audio_encoder.conv1d.weight" [256,1,16]           AudioEncoder        modules/module.py:12-22
 *   "@feature_projector.norm.{weight,bias}" [256]       FeatureProjector    modules/module.py:24-35
 *   "@feature_projector.conv1d.weight" [F,256,1]
 *   "@out_layer.end_conv1x1.{0,2}.{weight,bias}"        OutputLayer         modules/module.py:237-265 (masking = False)
 *   "@audio_decoder.weight" [256,1,16]                  AudioDecoder        modules/module.py:268-283
 * They are optional: sepref_finalize succeeds without them, the calls below then return SEPREF_ERR_STATE. */

/* Encoder frames for `samples` input samples ((samples - 16) / 4 + 1) and samples of one decoded waveform ((T - 1) * 4 + 16). */
int sepref_model_frames(const sepref_handle* h, int samples);
int sepref_model_output_samples(const sepref_handle* h, int samples);
size_t sepref_model_workspace_bytes(const sepref_handle* h, int batch, int samples);

/* Replaces Model.forward without its training-time auxiliary heads (model.py:38-45): audio_encoder -> feature_projector
 * -> separator -> out_layer -> audio_decoder, modules/module.py:12-35,190-218,237-283.
 *   mix         device [batch, samples]
 *   audio       device [num_spks, batch, sepref_model_output_samples()]   (the list Model.forward returns, stacked)
 *   out_stages  as in sepref_separator_forward (feeds the auxiliary heads if the caller wants them); NULL to skip */
int sepref_model_forward(sepref_handle* h, const float* mix, int batch, int samples, float* audio,
                         float* const* out_stages, void* workspace, size_t workspace_bytes, void* stream);

/* Pipelined HOST-buffer form (the loop of engine.py:165-172 / 113-149: mixture from the host, waveforms back): as
 * sepref_separator_submit_host / wait_host, but what crosses PCIe is the waveform (4 B per sample) in and
 * num_spks waveforms out instead of F-channel feature maps.  Shares the two staging slots with the separator calls. */
int sepref_model_submit_host(sepref_handle* h, int slot, const float* mix_host, int batch, int samples, float* audio_host);
int sepref_model_wait_host(sepref_handle* h, int slot);

/* PIT_SISNRi for two speakers, batched on the device (utils/implements/criterions.py:221-260 with scale_inv = True;
 * engine.py:131 passes eps = 1e-15): est device [2, batch, ld_est] (first n samples of each row are used), tgt device
 * [2, batch, n], mix device [batch, n]; out device [batch, 3] = {best-permutation sum over speakers of the SI-SNR
 * improvement in dB (engine.py:132 divides it by num_spks), and its two per-speaker terms}. */
int sepref_pit_sisnri(sepref_handle* h, const float* est, const float* tgt, const float* mix, int batch, int n, int ld_est,
                      double eps, float* out, void* stream);

/* Number of kernels the last sepref_separator_forward* call on this handle launched. */
int sepref_last_launch_count(const sepref_handle* h);

/* gemm_path 2 range safety: FP16 operands have a 5-bit exponent, so sepref_finalize bounds every activation that is
 * written as an FP16 operand from the weights alone (|LayerNorm(x)|_2 <= sqrt(F)); a GEMM group (one GCFN, one CLA
 * tail) whose bound is not below 3e4 runs with TF32 operands instead.  Returns how many groups that is (0 for sane
 * weights), -1 before sepref_finalize.  The attention kernel has FP16 operands only: if a q/k/v bound or the
 * relative-position table exceeds the limit, sepref_finalize fails with SEPREF_ERR_RANGE rather than saturate. */
int sepref_f16_fallback_count(const sepref_handle* h);
/* The three GEMMs fed by the UN-normalised residual stream (SpkSplitStage convs module.py:113-116, the fusion conv
 * module.py:212-214, OutputLayer module.py:244-247) have no such pack-time bound.  On gemm_path 2 they run with FP16
 * operands and detect a range excess at run time (a source value or a stage-2 operand beyond 65504; the conversions
 * saturate, never produce inf); the same GEMM is then re-computed with TF32 operands by a conditional launch that exits
 * at once otherwise.  Returns how many such re-computations have run on this handle so far (synchronises the device),
 * -1 before sepref_finalize.  SEPREF_OPT_RAW_F16 = 1 drops the re-computation (FP16 only). */
long long sepref_range_rerun_count(sepref_handle* h);

/* Forwards served by replaying a captured CUDA graph since the last sepref_finalize (SEPREF_OPT_CUDA_GRAPH). */
int sepref_graph_replay_count(const sepref_handle* h);

/* With SEPREF_OPT_PROFILE on: device time of the last sepref_separator_forward per kernel, measured with CUDA
 * events on the launching stream (interval between consecutive launches' completion).  Writes lines
 * "<kernel> <total ms> <launches>\n" into buf; synchronises on the last event. */
int sepref_profile_report(sepref_handle* h, char* buf, size_t cap);

/* ---- block-level entry points (unit parity).  Activations are channels-last device [rows, T, F];
 * `prefix` selects the weights by state_dict prefix, e.g. "dec_stages.1.g_block_2.block.gcfn.".
 * `workspace` must hold sepref_block_workspace_bytes(h, rows, t) bytes. */
size_t sepref_block_workspace_bytes(const sepref_handle* h, int rows, int t);
/* GCFN.forward, modules/network.py:60-66 */
int sepref_gcfn_forward(sepref_handle* h, const char* prefix, const float* x, int rows, int t, float* y,
                        void* workspace, size_t workspace_bytes, void* stream);
/* Test hook: the tensor-core GCFN kernel, additionally dumping h = W1.LN(x)+b1 as [rows*t, 6F] in the reference's
 * channel order (the tensor that modules/network.py:61 calls y before the depthwise conv), so that tests can
 * check GEMM1 separately from the gated convolution and GEMM2. */
int sepref_debug_gcfn_h(sepref_handle* h, const char* prefix, const float* x, int rows, int t, float* y, float* h_out,
                        void* stream);
/* Test/tuning hook: the tensor-core GCFN kernel with a pipeline timeline: clk_out (device, 8*64 int64) receives
 * clock64() stamps of block 0's first 8 tiles (slot meaning in csrc/kernels_tc.cuh, STAMP). */
int sepref_debug_gcfn_timeline(sepref_handle* h, const char* prefix, const float* x, int rows, int t, float* y,
                               long long* clk_out, void* stream);
/* Tuning hook: subsequent launches of the k_tok configuration whose tag contains `kernel_tag` ("gate", "cla_a",
 * "cla_b", "qkv", ...) write clock64 pipeline stamps of block 0's first 8 tiles into clk_out (device, 8*64 int64);
 * clk_out = NULL switches it off. */
int sepref_debug_tok_timeline(sepref_handle* h, const char* kernel_tag, long long* clk_out);
/* CLA.forward, modules/network.py:174-187 */
int sepref_cla_forward(sepref_handle* h, const char* prefix, const float* x, int rows, int t, float* y,
                       void* workspace, size_t workspace_bytes, void* stream);
/* EGA.forward (+ MultiHeadAttention with relative positions), modules/network.py:138-155,90-124; td = pooled length */
int sepref_ega_forward(sepref_handle* h, const char* prefix, const float* x, int rows, int t, int td, float* y,
                       void* workspace, size_t workspace_bytes, void* stream);
/* GlobalBlock.forward, modules/network.py:198-209 */
int sepref_global_block_forward(sepref_handle* h, const char* prefix, const float* x, int rows, int t, int td,
                                float* y, void* workspace, size_t workspace_bytes, void* stream);
/* LocalBlock.forward, modules/network.py:220-224 */
int sepref_local_block_forward(sepref_handle* h, const char* prefix, const float* x, int rows, int t, float* y,
                               void* workspace, size_t workspace_bytes, void* stream);
/* SpkAttention.forward, modules/network.py:233-252; rows = batch*num_spks */
int sepref_spk_attention_forward(sepref_handle* h, const char* prefix, const float* x, int rows, int t, float* y,
                                 void* workspace, size_t workspace_bytes, void* stream);
/* DownConvLayer.forward, modules/module.py:72-78; y is [rows, t/2, F] */
int sepref_down_conv_forward(sepref_handle* h, const char* prefix, const float* x, int rows, int t, float* y,
                             void* stream);
/* SpkSplitStage.forward, modules/module.py:120-125; y is [rows*num_spks, t, F] */
int sepref_spk_split_forward(sepref_handle* h, const char* prefix, const float* x, int rows, int t, float* y,
                             void* workspace, size_t workspace_bytes, void* stream);
/* upsample + cat + simple_fusion[i], modules/module.py:212-214; x_low is [rows, t/2, F], skip and y [rows, t, F] */
int sepref_fusion_forward(sepref_handle* h, const char* prefix, const float* x_low, const float* skip, int rows,
                          int t, float* y, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEPREF_H_ */
