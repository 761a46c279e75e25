"""Parity at the shapes the benchmark actually runs ("what's weak" 1-4).

* pooled attention (EGA / GlobalBlock) against the CPU oracle at Td = 500 and Td = 1250 keys (8 and 20 key tiles of
  the online softmax), head widths 16 and 32, all three kernel paths - compared on the whole output AND on the
  attention term alone (``y - x``), which the residual would otherwise dilute;
* F = 256 blocks at lengths where every persistent kernel walks several tiles per CTA (barrier phase wrap,
  single-buffered operand paths), against the oracle, and one whole Large separator against a reference golden;
* the single-utterance case C1 (``sample_WSJ.wav`` length: 18396 frames -> 18400, Td = 1150);
* FP16-operand robustness: scaled inputs, a planted outlier channel, large LayerScale, and weights whose pack-time
  range bound exceeds FP16 (the library must fall back to TF32 operands for those GEMMs, never saturate silently).
"""
import os

import pytest
import torch

from oracle import separator_oracle as O
from sepreformer_b200 import MODEL_SHAPES, Separator, separator_kwargs

from _util import check_generator_stable, load_golden, model_state, rel_l2, seeded_input

pytestmark = pytest.mark.gpu

PATHS = [int(p) for p in os.environ.get("SEPREF_TEST_PATHS", "0,1,2").split(",")]
# whole-output bound: north-star 1e-3 on the tensor-core paths; the fp32 CUDA-core path still rounds its attention
# operands to fp16 (kernels_attn.cuh), hence 5e-4 rather than 1e-4 where attention is involved
TOL = {0: 5e-4, 1: 1e-3, 2: 1e-3}
BASE, LARGE = "SepReformer_Base_WSJ0", "SepReformer_Large_DM_WSJ0"
_models = {}


def gpu_model(name, wseed, sd=None):
    key = (name, wseed)
    if sd is not None or key not in _models:
        _models.clear()
        shape = MODEL_SHAPES[name]
        m = Separator(**separator_kwargs(shape), per_stage_split=shape.per_stage_split)
        m.load_state_dict(sd if sd is not None else model_state(name, wseed), strict=True)
        m = m.cuda().eval()
        if sd is not None:
            return m
        _models[key] = m
    return _models[key]


def fparams(sd):
    return {k: v for k, v in sd.items() if v.is_floating_point()}


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("model,td,r", [(BASE, 500, 1), (BASE, 1250, 1), (LARGE, 500, 1), (LARGE, 1250, 1),
                                        (BASE, 500, 16), (LARGE, 250, 4), (BASE, 250, 2), (BASE, 250, 4), (BASE, 125, 8)])
def test_pooled_attention_matches_oracle_at_benchmark_key_counts(model, td, r, path):
    shape = MODEL_SHAPES[model]
    sd = model_state(model, 1)
    p = fparams(sd)
    m = gpu_model(model, 1)
    m.gemm_path = path
    pre = "dec_stages.2.g_block_1.block.ega."
    rows = 2 if r == 1 else 1
    x = seeded_input(700 + td + r, rows, td * r, shape.feat)
    with torch.no_grad():
        ref = O.ega(x, p, pre, shape.heads, td, p["pos_emb.pe_k.weight"], shape.maxlen)
    y = m.run_block("ega", pre, x.cuda(), td=td).cpu()
    whole, term = rel_l2(y, ref), rel_l2(y - x, ref - x)
    print(f"ega {model} Td={td} r={r} path={path}: whole {whole:.2e} attention-term {term:.2e}")
    assert whole < TOL[path]
    assert term < 4e-3        # gate * upsampled attention output alone (fp16-rounded q, k, v, E and P)


@pytest.mark.parametrize("path", PATHS)
def test_large_blocks_match_oracle_over_many_tiles(path):
    """F = 256: 12 rows x 4002 frames = 48 k tokens -> 3-4 tiles per CTA in every persistent kernel."""
    shape = MODEL_SHAPES[LARGE]
    sd = model_state(LARGE, 5)
    p = fparams(sd)
    m = gpu_model(LARGE, 5)
    m.gemm_path = path
    rows, t = 12, 4002
    x = seeded_input(811, rows, t, shape.feat)
    low = seeded_input(812, rows, t // 2, shape.feat)
    with torch.no_grad():
        want = {
            ("gcfn", "enc_stages.0.g_block_1.block.gcfn."): O.gcfn(x, p, "enc_stages.0.g_block_1.block.gcfn.", True),
            ("cla", "enc_stages.0.l_block_2.block.cla."): O.cla(x, p, "enc_stages.0.l_block_2.block.cla.", True),
            ("spk_attention", "dec_stages.3.spk_attn_1."): O.spk_attention(x, p, "dec_stages.3.spk_attn_1.", shape.heads, 2, True),
            ("spk_split", "spk_split_block."): O.spk_split(x[:6], p, "spk_split_block.", 2),
            ("fusion", "simple_fusion.3."): O.fuse(low, x, p, "simple_fusion.3."),
        }
    for (kind, prefix), ref in want.items():
        xin = x[:6] if kind == "spk_split" else x
        y = m.run_block(kind, prefix, xin.cuda(), x_low=low.cuda())
        torch.cuda.synchronize()
        err = rel_l2(y.cpu(), ref)
        print(f"large {kind} path={path}: {err:.2e}")
        assert err < (1e-4 if path == 0 else 1e-3), (kind, path)


@pytest.mark.parametrize("path", PATHS)
def test_large_separator_matches_reference_golden_medium(path):
    gold = load_golden("sep_large_medium")
    name = str(gold["model"])
    shape = MODEL_SHAPES[name]
    m = gpu_model(name, int(gold["wseed"]))
    m.gemm_path = path
    x = seeded_input(int(gold["xseed"]), int(gold["batch"]), shape.feat, int(gold["t_enc"]))
    check_generator_stable(gold, model_state(name, int(gold["wseed"])), x)
    with torch.no_grad():
        last, stages = m(x.cuda())
    st = int(gold["stride"])
    err = rel_l2(last.cpu()[..., ::st], gold["last"])
    print(f"sep_large_medium path={path}: {err:.2e}")
    assert err < (1e-4 if path == 0 else 1e-3)
    for i, s in enumerate(stages):
        assert rel_l2(s.cpu()[..., ::st], gold[f"stage{i}"]) < (1e-4 if path == 0 else 1e-3), f"stage {i}"


@pytest.mark.parametrize("path", [p for p in PATHS if p > 0])
def test_single_utterance_c1_matches_oracle(path):
    """BASELINE.json configs[0]: one utterance of sample_WSJ.wav's length (73593 samples -> 18396 frames)."""
    shape = MODEL_SHAPES[BASE]
    sd = model_state(BASE, 1)
    m = gpu_model(BASE, 1)
    m.gemm_path = path
    x = seeded_input(901, 1, shape.feat, 18396)
    with torch.no_grad():
        ref, ref_stages = O.separator_forward(x, fparams(sd), fast=True)
        got, stages = m(x.cuda())
    assert got.shape == (2, 128, 18400)
    err = rel_l2(got.cpu(), ref)
    print(f"C1 path={path}: {err:.2e}")
    assert err < 1e-3
    assert rel_l2(stages[0].cpu(), ref_stages[0]) < 1e-3      # bottleneck output: Td = 1150 keys


# ------------------------------------------------------------------------------------------------ fp16 range
def _run_vs_oracle(m, sd, x, path):
    m.gemm_path = path
    with torch.no_grad():
        ref, _ = O.separator_forward(x, fparams(sd), fast=True)
        got, _ = m(x.cuda())
    assert bool(torch.isfinite(got).all())
    return rel_l2(got.cpu(), ref)


@pytest.mark.parametrize("scale", [1e-3, 1e2])
def test_f16_path_scaled_inputs(scale):
    sd = model_state(BASE, 1)
    m = gpu_model(BASE, 1)
    x = seeded_input(911, 2, 128, 317) * scale
    err = _run_vs_oracle(m, sd, x, 2)
    print(f"f16 path, input scale {scale:g}: {err:.2e}")
    assert err < 1e-3


def test_f16_path_outlier_channel():
    sd = model_state(BASE, 1)
    m = gpu_model(BASE, 1)
    x = seeded_input(912, 2, 128, 317)
    x[:, 5, :] = 1.0e4
    x[1, 77, 100:120] = -3.0e4
    err = _run_vs_oracle(m, sd, x, 2)
    print(f"f16 path, planted outliers: {err:.2e}")
    assert err < 1e-3


def test_f16_path_large_layer_scale():
    sd = {k: v.clone() for k, v in model_state(BASE, 1).items()}
    for k in sd:
        if k.endswith("layer_scale"):
            sd[k].fill_(3.0)
    m = gpu_model(BASE, -1, sd=sd)
    x = seeded_input(913, 2, 128, 317)
    err = _run_vs_oracle(m, sd, x, 2)
    print(f"f16 path, LayerScale 3.0: {err:.2e}")
    assert err < 1e-3


def test_f16_path_falls_back_to_tf32_when_weights_exceed_fp16_range():
    """Weights whose worst-case intermediate bound passes the FP16 range: the pack-time check must route those GEMMs
    to TF32 operands (fp32 range) instead of saturating."""
    from sepreformer_b200 import _lib
    sd = {k: v.clone() for k, v in model_state(BASE, 1).items()}
    big = ["enc_stages.0.g_block_1.block.gcfn.", "dec_stages.3.l_block_2.block.gcfn."]
    f3 = 3 * 128
    for pre in big:          # the VALUE half of the gated conv grows 2e4-fold (the gates, and so the conditioning of the
        sd[pre + "depthwise.weight"][:f3] *= 2.0e4       # block, stay as they are); W2 brings the product back
        sd[pre + "depthwise.bias"][:f3] *= 2.0e4
        sd[pre + "net2.2.weight"] /= 2.0e4
    cl = "enc_stages.1.l_block_1.block.cla."
    sd[cl + "dw_conv_1d.weight"] *= 1.0e5
    sd[cl + "dw_conv_1d.bias"] *= 1.0e5
    sd[cl + "linear2.weight"] /= 1.0e5
    m = gpu_model(BASE, -2, sd=sd)
    x = seeded_input(914, 2, 128, 317)
    err = _run_vs_oracle(m, sd, x, 2)
    n = _lib.lib().sepref_f16_fallback_count(m.handle())
    print(f"f16 path, out-of-range weights: {err:.2e}, GEMM groups on TF32 operands: {n}")
    assert n >= 3
    assert err < 1e-3


def test_tensor_core_gate_rejects_non_power_of_two_upsampling():
    m = gpu_model(BASE, 1)
    m.gemm_path = 2
    x = seeded_input(915, 1, 75 * 3, 128).cuda()
    with pytest.raises(RuntimeError, match="power of two"):
        m.run_block("ega", "dec_stages.2.g_block_1.block.ega.", x, td=75)
    m.gemm_path = 0
    m.run_block("ega", "dec_stages.2.g_block_1.block.ega.", x, td=75)     # the CUDA-core path divides by r


@pytest.mark.parametrize("mode", [1, 2])
def test_frames_as_m_gcfn_kernel_matches_streaming_kernel_and_oracle(mode):
    """k_gcfn_tm (SEPREF_OPT_GCFN_TM: frames as the MMA M dimension; 1 = weight slabs shared by a CTA pair through
    tcgen05.mma.cta_group::2, 2 = single CTAs) against k_gcfn and the CPU oracle of GCFN (network.py:60-66): block level
    at segment / tile / utterance boundaries (30-frame warp segments, 120-frame tiles), then one whole forward."""
    sd = model_state(BASE, 1)
    m = gpu_model(BASE, 1)
    prefix = "dec_stages.1.g_block_2.block.gcfn."
    try:
        for rows, T in ((1, 1), (1, 2), (1, 29), (1, 30), (1, 31), (2, 61), (1, 120), (3, 121), (2, 1000), (5, 1234)):
            x = seeded_input(930 + T, rows, 128, T).transpose(1, 2).contiguous().cuda()      # [rows, T, F]
            m.gcfn_tm = 0
            y0 = m.run_block("gcfn", prefix, x)
            m.gcfn_tm = mode
            y1 = m.run_block("gcfn", prefix, x)
            with torch.no_grad():
                ref = O.gcfn(x.cpu(), fparams(sd), prefix)
            d, e = rel_l2(y1.cpu(), y0.cpu()), rel_l2((y1 - x).cpu(), ref - x.cpu())
            print(f"k_gcfn_tm mode {mode} rows={rows} T={T}: vs k_gcfn {d:.2e}, block term vs oracle {e:.2e}")
            assert bool(torch.isfinite(y1).all()) and d < 1e-5 and e < 1e-3
        x = seeded_input(931, 2, 128, 1999).cuda()
        m.gcfn_tm = 0
        y0, _ = m(x)
        m.gcfn_tm = mode
        y1, _ = m(x)
        d = rel_l2(y1.cpu(), y0.cpu())
        print(f"k_gcfn_tm mode {mode}: whole forward vs k_gcfn {d:.2e}")
        assert d < 5e-4
    finally:
        m.gcfn_tm = 0


def test_raw_stream_gemms_recompute_with_tf32_when_fp16_range_is_exceeded():
    """SpkSplit / fusion / output-layer GEMMs read the un-normalised residual stream: FP16 operands by default on the f16
    path, with a run-time range check and a conditional TF32 re-computation (include/sepref.h, sepref_range_rerun_count)."""
    from sepreformer_b200 import _lib
    sd = model_state(BASE, 1)
    m = gpu_model(BASE, 1)
    m.gemm_path = 2
    L = _lib.lib()
    x = seeded_input(941, 2, 128, 317)
    before = L.sepref_range_rerun_count(m.handle())
    err = _run_vs_oracle(m, sd, x, 2)
    mid = L.sepref_range_rerun_count(m.handle())
    print(f"unit-scale input: {err:.2e}, re-computations {mid - before}")
    assert before >= 0 and mid == before and err < 1e-3
    # Residual stream far beyond 65504.  At this scale the whole-network error is set by the 11-bit rounding of the raw
    # stream itself (1e-2 against the fp32 oracle on the TF32 path as well), so the reference here is the TF32 path:
    # FP16 + re-computation must agree with it, FP16 alone (SEPREF_OPT_RAW_F16) must not.
    xb = (x * 3.0e4).cuda()
    with torch.no_grad():
        m.gemm_path = 1
        y_tf32, _ = m(xb)
        m.gemm_path = 2
        y_f16, _ = m(xb)
        after = L.sepref_range_rerun_count(m.handle())
        m.raw_f16 = 1
        try:
            y_clamped, _ = m(xb)
        finally:
            m.raw_f16 = 0
    d_ok, d_bad = rel_l2(y_f16.cpu(), y_tf32.cpu()), rel_l2(y_clamped.cpu(), y_tf32.cpu())
    print(f"input x 3e4: re-computations {after - mid}, f16 + re-computation vs tf32 path {d_ok:.2e}, f16 alone {d_bad:.2e}")
    assert after > mid and bool(torch.isfinite(y_f16).all()) and d_ok < 2e-3 and d_bad > 4 * d_ok
