"""CPU oracle for the SepReformer separator forward pass.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The product path (``sepreformer_b200``) never does, and fails loudly without its CUDA library.

It restates, as plain tensor arithmetic on channels-last ``[batch, time, feat]`` arrays
(torch CPU tensors, fp32 or fp64), what the reference computes in
``/root/reference/models/SepReformer_Base_WSJ0/modules/{network,module}.py``.  Every
function cites the reference lines it follows.  The arithmetic lives in PyTorch (the
reference pins torch==2.1.2); nothing is copied: blocks are written as formulas over a flat
``{state_dict key: tensor}`` mapping instead of ``nn.Module`` trees.

Parity pin: ``tests/golden/make_golden.py`` imports the *reference itself* in the build
container, loads the same seeded ``state_dict`` into it and stores its outputs under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` holds this oracle to those vectors
(fp64 agreement <= 1e-9, fp32 <= 2e-5) and, when ``/root/reference`` is present, to the live
reference.  The reference ships no tests or golden vectors of its own.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]

LN_EPS = 1e-5      # torch.nn.LayerNorm default, network.py:50,81,133,162
BN_EPS = 1e-5      # torch.nn.BatchNorm1d default, network.py:167, module.py:69
GN_EPS = 1e-8      # module.py:117


# --------------------------------------------------------------------------- primitives
def layer_norm(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)          # biased, as torch
    return (x - mu) / torch.sqrt(var + LN_EPS) * w + b


def affine(x: Tensor, w: Tensor, b: Tensor | None) -> Tensor:
    y = x @ w.transpose(0, 1)
    return y if b is None else y + b


def sigmoid(x: Tensor) -> Tensor:
    return 1.0 / (1.0 + torch.exp(-x))


def gelu_erf(x: Tensor) -> Tensor:
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))   # GELU(approximate='none')


def glu_last(x: Tensor) -> Tensor:
    half = x.shape[-1] // 2
    return x[..., :half] * sigmoid(x[..., half:])


def dwconv_time(x: Tensor, w: Tensor, b: Tensor, pad: int, stride: int = 1, fast: bool = False) -> Tensor:
    """Depthwise cross-correlation along time of ``x[B,T,C]`` with ``w[C,1,k]``, zero padding ``pad`` each side.

    out[b,t,c] = b[c] + sum_j w[c,0,j] * x[b, stride*t + j - pad, c]   (torch Conv1d, groups=C).
    """
    k = w.shape[-1]
    B, T, C = x.shape
    t_out = (T + 2 * pad - k) // stride + 1
    if fast:   # same numbers through the library primitive; used only to time the CPU baseline
        y = torch.nn.functional.conv1d(x.transpose(1, 2), w, b, stride=stride, padding=pad, groups=C)
        return y.transpose(1, 2)
    xp = torch.zeros(B, T + 2 * pad, C, dtype=x.dtype)
    xp[:, pad:pad + T] = x
    out = torch.zeros(B, t_out, C, dtype=x.dtype) + b
    for j in range(k):
        out = out + xp[:, j:j + stride * (t_out - 1) + 1:stride] * w[:, 0, j]
    return out


# --------------------------------------------------------------------------- blocks
def gcfn(x: Tensor, p: Params, pre: str, fast: bool = False) -> Tensor:
    """network.py:46-66.  x + ls * W2 . GLU(dw3(W1 . LN(x)))."""
    h = affine(layer_norm(x, p[pre + "net1.0.weight"], p[pre + "net1.0.bias"]),
               p[pre + "net1.1.weight"], p[pre + "net1.1.bias"])
    d = dwconv_time(h, p[pre + "depthwise.weight"], p[pre + "depthwise.bias"], pad=1, fast=fast)
    y = affine(glu_last(d), p[pre + "net2.2.weight"], p[pre + "net2.2.bias"])
    return x + y * p[pre + "Layer_scale.layer_scale"].reshape(-1)


def rel_pos_index(td: int, maxlen: int) -> Tensor:
    """module.py:52-57,196-197: clamp(i - j, -maxlen, maxlen-1) + maxlen."""
    i = torch.arange(td)
    return (i[:, None] - i[None, :]).clamp(-maxlen, maxlen - 1) + maxlen


def mha(x: Tensor, p: Params, pre: str, heads: int, pe_k: Tensor | None, maxlen: int) -> Tensor:
    """network.py:90-124 with mask=None (the only way it is called).  Returns ls * out_proj(attn)."""
    n, t, f = x.shape
    dk = f // heads
    z = layer_norm(x, p[pre + "layer_norm.weight"], p[pre + "layer_norm.bias"])
    q = affine(z, p[pre + "linear_q.weight"], p[pre + "linear_q.bias"]).reshape(n, t, heads, dk)
    k = affine(z, p[pre + "linear_k.weight"], p[pre + "linear_k.bias"]).reshape(n, t, heads, dk)
    v = affine(z, p[pre + "linear_v.weight"], p[pre + "linear_v.bias"]).reshape(n, t, heads, dk)
    s = torch.einsum("nihd,njhd->nhij", q, k)
    if pe_k is not None:
        e = pe_k[rel_pos_index(t, maxlen)]                     # [t, t, dk], shared by heads and layers
        s = s + torch.einsum("nihd,ijd->nhij", q, e)
    s = s / math.sqrt(dk)
    s = s - s.amax(-1, keepdim=True)
    w = torch.exp(s)
    w = w / w.sum(-1, keepdim=True)
    o = torch.einsum("nhij,njhd->nihd", w, v).reshape(n, t, f)
    o = affine(o, p[pre + "linear_out.weight"], p[pre + "linear_out.bias"])
    return o * p[pre + "Layer_scale.layer_scale"].reshape(-1)


def ega(x: Tensor, p: Params, pre: str, heads: int, td: int, pe_k: Tensor, maxlen: int) -> Tensor:
    """network.py:138-155.  Pool to td, attend, nearest-upsample, gate.  T must be a multiple of td."""
    n, t, f = x.shape
    r = t // td
    assert r * td == t, "separator lengths are td * 2^s"
    xd = x.reshape(n, td, r, f).mean(2)                        # adaptive_avg_pool1d, equal windows
    a = mha(xd, p, pre + "block.self_attn.", heads, pe_k, maxlen)
    up = a.repeat_interleave(r, dim=1)                         # nearest upsample by integer factor
    gate = sigmoid(affine(layer_norm(x, p[pre + "block.linear.0.weight"], p[pre + "block.linear.0.bias"]),
                          p[pre + "block.linear.1.weight"], p[pre + "block.linear.1.bias"]))
    return x + gate * up


def fold_bn(lin_w: Tensor, lin_b: Tensor, p: Params, pre: str) -> Tuple[Tensor, Tensor]:
    """Eval-mode BatchNorm1d after a per-channel affine map collapses into that map."""
    s = p[pre + "weight"] / torch.sqrt(p[pre + "running_var"] + BN_EPS)
    return lin_w * s.reshape(-1, *([1] * (lin_w.dim() - 1))), (lin_b - p[pre + "running_mean"]) * s + p[pre + "bias"]


def cla(x: Tensor, p: Params, pre: str, fast: bool = False) -> Tensor:
    """network.py:174-187."""
    u = glu_last(affine(layer_norm(x, p[pre + "layer_norm.weight"], p[pre + "layer_norm.bias"]),
                        p[pre + "linear1.weight"], p[pre + "linear1.bias"]))
    k = p[pre + "dw_conv_1d.weight"].shape[-1]
    d = dwconv_time(u, p[pre + "dw_conv_1d.weight"], p[pre + "dw_conv_1d.bias"], pad=(k - 1) // 2, fast=fast)
    w2, b2 = fold_bn(p[pre + "linear2.weight"], p[pre + "linear2.bias"], p, pre + "BN.")
    g = gelu_erf(affine(d, w2, b2))
    y = affine(g, p[pre + "linear3.1.weight"], p[pre + "linear3.1.bias"])
    return x + y * p[pre + "Layer_scale.layer_scale"].reshape(-1)


def global_block(x, p, pre, heads, td, pe_k, maxlen, fast=False):
    """network.py:198-209 (the trailing permute is a layout change only)."""
    return gcfn(ega(x, p, pre + "block.ega.", heads, td, pe_k, maxlen), p, pre + "block.gcfn.", fast)


def local_block(x, p, pre, fast=False):
    """network.py:220-224."""
    return gcfn(cla(x, p, pre + "block.cla.", fast), p, pre + "block.gcfn.", fast)


def spk_attention(x: Tensor, p: Params, pre: str, heads: int, num_spks: int, fast: bool = False) -> Tensor:
    """network.py:233-252.  x is [B*S, T, F] with row b*S+s; attention runs over the S speakers of each (b,t)."""
    n, t, f = x.shape
    b = n // num_spks
    tok = x.reshape(b, num_spks, t, f).permute(0, 2, 1, 3).reshape(b * t, num_spks, f)
    tok = tok + mha(tok, p, pre + "self_attn.", heads, None, 0)
    x = tok.reshape(b, t, num_spks, f).permute(0, 2, 1, 3).reshape(n, t, f)
    return gcfn(x, p, pre + "feed_forward.", fast)


def down_conv(x: Tensor, p: Params, pre: str, fast: bool = False) -> Tensor:
    """module.py:72-78: depthwise k=5 stride 2 pad 2, BatchNorm(eval), GELU."""
    w, b = fold_bn(p[pre + "down_conv.weight"], p[pre + "down_conv.bias"], p, pre + "BN.")
    k = w.shape[-1]
    return gelu_erf(dwconv_time(x, w, b, pad=(k - 1) // 2, stride=2, fast=fast))


def spk_split(x: Tensor, p: Params, pre: str, num_spks: int) -> Tensor:
    """module.py:120-125.  [B,T,F] -> [B*S,T,F]; GroupNorm(1 group) statistics over all (T,F) of a row."""
    n, t, f = x.shape
    h = affine(x, p[pre + "linear.0.weight"][:, :, 0], p[pre + "linear.0.bias"])
    h = glu_last(h)                                            # GLU(dim=-2) on [B,C,T] == GLU over channels
    h = affine(h, p[pre + "linear.2.weight"][:, :, 0], p[pre + "linear.2.bias"])     # [B,T,S*F]
    y = h.reshape(n, t, num_spks, f).permute(0, 2, 1, 3).reshape(n * num_spks, t, f)
    mu = y.mean((1, 2), keepdim=True)
    var = ((y - mu) ** 2).mean((1, 2), keepdim=True)
    return (y - mu) / torch.sqrt(var + GN_EPS) * p[pre + "norm.weight"] + p[pre + "norm.bias"]


def fuse(x_low: Tensor, skip: Tensor, p: Params, pre: str) -> Tensor:
    """module.py:212-214: nearest upsample to skip's length, concat [up, skip] on channels, 1x1 conv."""
    r = skip.shape[1] // x_low.shape[1]
    cat = torch.cat([x_low.repeat_interleave(r, dim=1), skip], dim=-1)
    return affine(cat, p[pre + "weight"][:, :, 0], p[pre + "bias"])


def pad_frames(x: Tensor, chunk: int) -> Tensor:
    """module.py:220-234 on [B,T,F]: right zero-pad T to a multiple of chunk; untouched if already one."""
    t = x.shape[1]
    rest = 0 if t % chunk == 0 else (t // chunk + 1) * chunk - t
    if rest == 0:
        return x
    return torch.cat([x, torch.zeros(x.shape[0], rest, x.shape[2], dtype=x.dtype)], dim=1)


# --------------------------------------------------------------------------- the path
def separator_forward(inp: Tensor, p: Params, *, heads: int = 8, num_stages: int = 4, num_spks: int = 2,
                      maxlen: int = 2000, per_stage_split: bool = False, fast: bool = False,
                      taps: dict | None = None) -> Tuple[Tensor, List[Tensor]]:
    """module.py:190-218.  ``inp`` is ``[B, F, T_enc]`` as handed over by ``Model.forward`` (model.py:41).

    Returns ``(last [B*S, F, T_pad], [stage outputs [B*S, F, T_pad / 2^(R-i)]])`` in the reference's layout.
    ``taps``, if given, receives named intermediate tensors (channels-last) for block-level tests.
    """
    x = pad_frames(inp.transpose(1, 2), 2 ** num_stages)
    td = x.shape[1] // 2 ** num_stages
    pe_k = p["pos_emb.pe_k.weight"]

    def G(x, pre):
        return global_block(x, p, pre, heads, td, pe_k, maxlen, fast)

    def split(x, idx):
        pre = f"spk_split_blocks.{idx}." if per_stage_split else "spk_split_block."
        return spk_split(x, p, pre, num_spks)

    def enc_stage(x, pre, down):
        x = G(x, pre + "g_block_1.")
        x = local_block(x, p, pre + "l_block_1.", fast)
        x = G(x, pre + "g_block_2.")
        x = local_block(x, p, pre + "l_block_2.", fast)
        return (down_conv(x, p, pre + "downconv.", fast) if down else x), x

    skips = []
    for s in range(num_stages):
        x, sk = enc_stage(x, f"enc_stages.{s}.", True)
        skips.append(split(sk, s))
        if taps is not None:
            taps[f"enc{s}"] = x
    x, _ = enc_stage(x, "bottleneck_G.", False)
    x = split(x, num_stages)
    if taps is not None:
        taps["bottleneck"] = x

    stage_outs = []
    for i in range(num_stages):
        stage_outs.append(x)
        x = fuse(x, skips[num_stages - 1 - i], p, f"simple_fusion.{i}.")
        pre = f"dec_stages.{i}."
        for n in (1, 2, 3):
            x = G(x, pre + f"g_block_{n}.")
            x = local_block(x, p, pre + f"l_block_{n}.", fast)
            x = spk_attention(x, p, pre + f"spk_attn_{n}.", heads, num_spks, fast)
        if taps is not None:
            taps[f"dec{i}"] = x
    return x.transpose(1, 2).contiguous(), [s.transpose(1, 2).contiguous() for s in stage_outs]


# --------------------------------------------------------------------------- metric
def _l2(x: Tensor) -> Tensor:
    return torch.sqrt((x * x).sum(-1))


def si_snr_db(est: Tensor, src: Tensor, eps: float) -> Tensor:
    """One (estimate, source) pair of utils/implements/criterions.py:242-249 (scale-invariant branch)."""
    est = est - est.mean(-1, keepdim=True)
    src = src - src.mean(-1, keepdim=True)
    proj = (est * src).sum(-1, keepdim=True) / (_l2(src).unsqueeze(-1) ** 2 + eps) * src
    return 20.0 * torch.log10(eps + _l2(proj) / (_l2(est - proj) + eps))


def pit_si_snri(estims: Sequence[Tensor], targets: Sequence[Tensor], mixture: Tensor, eps: float = 1e-15) -> Tensor:
    """criterions.py:232-260 for two or more speakers: best-permutation SI-SNR improvement per utterance [B]."""
    import itertools
    n = len(estims)
    best = None
    for perm in itertools.permutations(range(n)):
        tot = sum(si_snr_db(estims[s], targets[t], eps) - si_snr_db(mixture, targets[t], eps)
                  for s, t in enumerate(perm))
        best = tot if best is None else torch.maximum(best, tot)
    return best / n


# --------------------------------------------------------------------------- model shell (for the SI-SNRi metric only)
# Restatement of the layers around the separator (reference module.py:12-35, 237-283, model.py:38-45), so that the
# north-star metric "SI-SNRi delta vs reference" can be evaluated: both sides use this same shell, only the
# separator differs.  Shell weights use the keys of Model.state_dict() (audio_encoder.*, feature_projector.*, ...).
def audio_encoder(mix: Tensor, p: Params) -> Tensor:
    """module.py:12-22: Conv1d(1 -> C, k, stride, no bias) + GELU on [B, n] -> [B, C, T]."""
    w = p["audio_encoder.conv1d.weight"]                       # [C, 1, k]
    k = w.shape[-1]
    stride = 4                                                 # configs.yaml:37
    frames = mix.unfold(-1, k, stride)                         # [B, T, k]
    return gelu_erf(torch.einsum("btk,ck->bct", frames, w[:, 0]))


def feature_projector(e: Tensor, p: Params) -> Tensor:
    """module.py:24-35: GroupNorm(1 group, eps 1e-8) over (C, T) + 1x1 conv (no bias)."""
    mu = e.mean((1, 2), keepdim=True)
    var = ((e - mu) ** 2).mean((1, 2), keepdim=True)
    z = (e - mu) / torch.sqrt(var + GN_EPS) * p["feature_projector.norm.weight"][None, :, None] \
        + p["feature_projector.norm.bias"][None, :, None]
    return torch.einsum("bct,fc->bft", z, p["feature_projector.conv1d.weight"][:, :, 0])


def output_layer(sep: Tensor, enc: Tensor, p: Params, num_spks: int, pre: str = "out_layer.") -> Tensor:
    """module.py:250-265 with masking=False: crop to the encoder length, Linear-GLU-Linear over channels.
    sep [B*S, F, T_pad] -> [S, B, C, T]."""
    x = sep[..., : enc.shape[-1]].transpose(1, 2)
    x = glu_last(affine(x, p[pre + "end_conv1x1.0.weight"], p[pre + "end_conv1x1.0.bias"]))
    x = affine(x, p[pre + "end_conv1x1.2.weight"], p[pre + "end_conv1x1.2.bias"]).transpose(1, 2)
    bs, c, t = x.shape
    return x.reshape(bs // num_spks, num_spks, c, t).transpose(0, 1)


def audio_decoder(x: Tensor, p: Params, pre: str = "audio_decoder.") -> Tensor:
    """module.py:268-283: ConvTranspose1d(C -> 1, k, stride, no bias) as overlap-add.  x [B, C, T] -> [B, (T-1)*stride + k]."""
    w = p[pre + "weight"]                                      # [C, 1, k]
    k, stride = w.shape[-1], 4
    frames = torch.einsum("bct,ck->btk", x, w[:, 0])           # [B, T, k]
    b, t, _ = frames.shape
    out = torch.zeros(b, (t - 1) * stride + k, dtype=x.dtype)
    for j in range(k):
        out[:, j: j + (t - 1) * stride + 1: stride] += frames[:, :, j]
    return out


def model_forward(mix: Tensor, shell: Params, separator_fn, num_spks: int = 2) -> List[Tensor]:
    """model.py:38-45 (the auxiliary heads of 47-51 do not feed the returned audio): mixture [B, n] -> per-speaker audio.
    ``separator_fn(features [B,F,T]) -> last [B*S,F,T_pad]`` is the separator under test (oracle or CUDA)."""
    enc = audio_encoder(mix, shell)
    feat = feature_projector(enc, shell)
    last = separator_fn(feat)
    out = output_layer(last.to(enc.dtype), enc, shell, num_spks)
    return [audio_decoder(out[s], shell) for s in range(num_spks)]
